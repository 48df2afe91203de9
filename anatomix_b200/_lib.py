"""ctypes binding of ``include/anatomix_b200.h`` (the C ABI).

Loading never compiles anything: the shared library is built in-tree by
``anatomix_b200.build`` / ``__graft_entry__.build()``.  A missing library is a
hard error for the engine path (there is no CPU or eager fallback behind it).
"""
from __future__ import annotations

import ctypes as C
import os

_VARIANT = os.environ.get("ANX_LIB_VARIANT", "")      # experiments only, see build.py
LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib",
                        "libanatomix_b200" + ("_" + _VARIANT if _VARIANT else "") + ".so")

ANX_OK = 0
STATUS_NAMES = {0: "OK", 1: "BAD_ARG", 2: "BAD_SHAPE", 3: "UNSUPPORTED", 4: "WORKSPACE",
                5: "CUDA", 6: "NOT_READY", 7: "NO_DEVICE"}
NORM = {"none": 0, "batch": 1, "instance": 2}
ACT = {"none": 0, "relu": 1, "lrelu": 2}
POOL = {"Max": 0, "Avg": 1}
INTERP = {"nearest": 0, "trilinear": 1}
FLAG_FORCE_SIMT = 1
FLAG_STORE_FP16 = 2
FLAG_STORE_BF16 = 4
FLAG_DEPTH_HALO_INPUT = 8
FLAG_NO_UPCONV = 16
FLAG_NO_ROWS = 32
FLAG_NO_WS_REUSE = 64
PAYLOAD_F32_NCDHW = 0
PAYLOAD_CL16 = 1

# every symbol include/anatomix_b200.h declares (checked by tests/test_abi.py)
EXPORTS = [
    "anx_engine_create", "anx_engine_destroy", "anx_engine_num_convs", "anx_engine_conv_info",
    "anx_engine_set_conv", "anx_engine_workspace_bytes", "anx_engine_forward",
    "anx_engine_forward_host", "anx_engine_launches_per_forward", "anx_engine_profile",
    "anx_engine_num_buffers", "anx_engine_buffer_info", "anx_status_string",
    "anx_engine_last_error", "anx_version", "anx_selftest",
    "anx_engine_num_steps", "anx_engine_step_info", "anx_engine_run_steps",
    "anx_engine_forward_allgather", "anx_engine_row_layout", "anx_engine_set_head", "anx_engine_out_channels",
    "anx_engine_num_taps", "anx_engine_tap_info", "anx_engine_export_tap", "anx_avgpool3d_scale_f32", "anx_blend_window_f32", "anx_engine_set_slab", "anx_engine_step_stats",
    "anx_engine_forward_gather", "anx_engine_forward_cl16", "anx_engine_storage_type", "anx_widen_cl16_f32",
    "anx_push_to_peers", "anx_engine_forward_slab", "anx_engine_forward_host_ex",
    "anx_engine_forward_concat", "anx_channel_normalize_f32",
    "anx_engine_tap_kind", "anx_engine_set_tap_conv", "anx_engine_export_prenorm_tap",
    "anx_engine_forward_host_pipelined", "anx_engine_host_wait",
]


class UnetDesc(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("input_nc", C.c_int32), ("output_nc", C.c_int32),
                ("num_downs", C.c_int32), ("ngf", C.c_int32), ("norm_kind", C.c_int32),
                ("norm_eps", C.c_float), ("act_kind", C.c_int32), ("act_slope", C.c_float),
                ("pool_kind", C.c_int32), ("interp_kind", C.c_int32), ("device", C.c_int32),
                ("flags", C.c_uint32)]


class SlabLinks(C.Structure):
    _fields_ = [("struct_size", C.c_uint32), ("lower_workspace", C.c_void_p), ("upper_workspace", C.c_void_p),
                ("flags", C.c_void_p), ("lower_flags", C.c_void_p), ("upper_flags", C.c_void_p)]


class EngineError(RuntimeError):
    def __init__(self, status, detail=""):
        self.status = status
        super().__init__(f"anatomix_b200 engine: {STATUS_NAMES.get(status, status)}"
                         + (f": {detail}" if detail else ""))


_lib = None


def load():
    """The loaded library (cached).  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m anatomix_b200.build` "
            "(the engine has no fallback path)")
    lib = C.CDLL(LIB_PATH)
    vp, fp, i32, sz = C.c_void_p, C.POINTER(C.c_float), C.c_int32, C.c_size_t
    lib.anx_engine_create.argtypes = [C.POINTER(UnetDesc), C.POINTER(vp)]
    lib.anx_engine_create.restype = i32
    lib.anx_engine_destroy.argtypes = [vp]
    lib.anx_engine_destroy.restype = None
    lib.anx_engine_num_convs.argtypes = [vp]
    lib.anx_engine_num_convs.restype = i32
    lib.anx_engine_conv_info.argtypes = [vp, i32] + [C.POINTER(i32)] * 4
    lib.anx_engine_conv_info.restype = i32
    lib.anx_engine_set_conv.argtypes = [vp, i32] + [vp] * 6 + [i32]
    lib.anx_engine_set_conv.restype = i32
    lib.anx_engine_workspace_bytes.argtypes = [vp, i32, i32, i32, i32]
    lib.anx_engine_workspace_bytes.restype = sz
    lib.anx_engine_forward.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, sz, vp]
    lib.anx_engine_forward.restype = i32
    lib.anx_engine_forward_host.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, vp, vp, sz, vp]
    lib.anx_engine_forward_host.restype = i32
    lib.anx_engine_launches_per_forward.argtypes = [vp, i32, i32, i32, i32]
    lib.anx_engine_launches_per_forward.restype = i32
    lib.anx_engine_profile.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, sz, vp, fp, vp, i32,
                                       C.POINTER(i32)]
    lib.anx_engine_profile.restype = i32
    lib.anx_engine_num_buffers.argtypes = [vp]
    lib.anx_engine_num_buffers.restype = i32
    lib.anx_engine_buffer_info.argtypes = [vp, i32, i32, i32, i32, i32, C.POINTER(sz), C.POINTER(sz),
                                           C.POINTER(i32), C.POINTER(i32)]
    lib.anx_engine_buffer_info.restype = i32
    lib.anx_engine_num_steps.argtypes = [vp]
    lib.anx_engine_num_steps.restype = i32
    lib.anx_engine_step_info.argtypes = [vp, i32] + [C.POINTER(i32)] * 4 + [C.c_char_p]
    lib.anx_engine_step_info.restype = i32
    lib.anx_engine_run_steps.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, sz, vp, i32, i32]
    lib.anx_engine_run_steps.restype = i32
    lib.anx_engine_forward_allgather.argtypes = [vp, vp, C.POINTER(vp), i32, i32, i32, i32, i32, i32, vp, sz, vp]
    lib.anx_engine_forward_allgather.restype = i32
    lib.anx_engine_forward_gather.argtypes = [vp, vp, C.POINTER(vp), i32, i32, i32, i32, i32, i32, i32, vp, sz, vp]
    lib.anx_engine_forward_gather.restype = i32
    lib.anx_engine_forward_cl16.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, sz, vp]
    lib.anx_engine_forward_cl16.restype = i32
    lib.anx_engine_storage_type.argtypes = [vp]
    lib.anx_engine_storage_type.restype = i32
    lib.anx_widen_cl16_f32.argtypes = [vp, vp, C.c_int64, i32, i32, i32, i32, i32, vp]
    lib.anx_widen_cl16_f32.restype = i32
    lib.anx_push_to_peers.argtypes = [vp, vp, C.POINTER(vp), i32, i32, sz, vp]
    lib.anx_push_to_peers.restype = i32
    lib.anx_engine_forward_slab.argtypes = [vp, vp, vp, i32, i32, i32, i32, vp, sz, C.POINTER(SlabLinks), vp]
    lib.anx_engine_forward_slab.restype = i32
    lib.anx_engine_forward_host_ex.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, sz, vp]
    lib.anx_engine_forward_host_ex.restype = i32
    lib.anx_engine_forward_concat.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp, sz, vp]
    lib.anx_engine_forward_concat.restype = i32
    lib.anx_channel_normalize_f32.argtypes = [vp, vp, C.c_int64, i32, i32, i32, i32, i32, C.c_float, vp]
    lib.anx_channel_normalize_f32.restype = i32
    lib.anx_engine_tap_kind.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32)]
    lib.anx_engine_tap_kind.restype = i32
    lib.anx_engine_set_tap_conv.argtypes = [vp, i32, vp, vp, i32]
    lib.anx_engine_set_tap_conv.restype = i32
    lib.anx_engine_export_prenorm_tap.argtypes = [vp, i32, vp, i32, i32, i32, i32, vp, sz, vp, vp]
    lib.anx_engine_export_prenorm_tap.restype = i32
    lib.anx_engine_forward_host_pipelined.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, vp, vp, vp, sz, vp]
    lib.anx_engine_forward_host_pipelined.restype = i32
    lib.anx_engine_host_wait.argtypes = [vp, vp]
    lib.anx_engine_host_wait.restype = i32
    lib.anx_engine_row_layout.argtypes = [vp, i32, C.POINTER(i32), C.POINTER(i32)]
    lib.anx_engine_row_layout.restype = i32
    lib.anx_engine_set_head.argtypes = [vp, i32, vp, vp, i32]
    lib.anx_engine_set_head.restype = i32
    lib.anx_engine_out_channels.argtypes = [vp]
    lib.anx_engine_out_channels.restype = i32
    lib.anx_engine_num_taps.argtypes = [vp]
    lib.anx_engine_num_taps.restype = i32
    lib.anx_engine_tap_info.argtypes = [vp, i32] + [C.POINTER(i32)] * 5
    lib.anx_engine_tap_info.restype = i32
    lib.anx_engine_export_tap.argtypes = [vp, i32, i32, i32, i32, i32, vp, sz, vp, vp]
    lib.anx_engine_export_tap.restype = i32
    lib.anx_avgpool3d_scale_f32.argtypes = [vp, vp, C.c_int64, i32, i32, i32, i32, C.c_float, vp]
    lib.anx_avgpool3d_scale_f32.restype = i32
    lib.anx_blend_window_f32.argtypes = [vp, vp, vp, vp] + [i32] * 10 + [vp]
    lib.anx_blend_window_f32.restype = i32
    lib.anx_engine_set_slab.argtypes = [vp, i32, i32, i32]
    lib.anx_engine_set_slab.restype = i32
    lib.anx_engine_step_stats.argtypes = [vp, i32, i32, i32, i32, i32, C.POINTER(sz), C.POINTER(sz)]
    lib.anx_engine_step_stats.restype = i32
    lib.anx_status_string.argtypes = [i32]
    lib.anx_status_string.restype = C.c_char_p
    lib.anx_engine_last_error.argtypes = [vp]
    lib.anx_engine_last_error.restype = C.c_char_p
    lib.anx_version.restype = i32
    lib.anx_selftest.argtypes = [i32, C.c_char_p, sz]
    lib.anx_selftest.restype = i32
    _lib = lib
    return lib


def selftest(device=0):
    """(ok, report) of the tcgen05/TMA primitive probes."""
    buf = C.create_string_buffer(16384)
    st = load().anx_selftest(device, buf, len(buf))
    return st == ANX_OK, buf.value.decode()
