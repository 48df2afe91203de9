"""ctypes front-end of the plain-C oracle (oracle/unet_ref.c).  Test
infrastructure only; see that file's header."""
import ctypes

import numpy as np

from .build_oracle import build
from .unet_oracle import DEFAULTS, layer_program

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.ref_unet_forward.restype = ctypes.c_int
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def unet_forward_c(cfg, state, x):
    """Standard forward through the C restatement.  ``x``: float32 ndarray
    [N,C,D,H,W]; ``state``: dict name -> array-like."""
    c = dict(DEFAULTS)
    c.update(cfg)
    assert c["activation"] == "relu" and c["doubleconv"] and c["use_skip_connection"]
    assert c["norm"] in ("batch", "instance") and c["final_act"] == "none"
    prog, _, _ = layer_program(c["input_nc"], c["output_nc"], c["num_downs"], c["ngf"],
                               c["norm"], c["activation"], c["final_act"], True, True)
    f32 = lambda k: np.ascontiguousarray(np.asarray(state[k], dtype=np.float32))
    ws, bs, bns, keep = [], [], [], []
    for pos, (op, idx, ci, co) in enumerate(prog):
        if op != "conv":
            continue
        w = f32(f"model.{idx}.weight"); keep.append(w); ws.append(_fp(w))
        if c["norm"] == "instance":
            b = f32(f"model.{idx}.bias"); keep.append(b); bs.append(_fp(b))
        if c["norm"] == "batch" and pos + 1 < len(prog) and prog[pos + 1][0] == "norm":
            pack = np.concatenate([f32(f"model.{idx+1}.{k}") for k in
                                   ("weight", "bias", "running_mean", "running_var")])
            keep.append(pack); bns.append(_fp(pack))
        else:
            bns.append(ctypes.POINTER(ctypes.c_float)())
    n = len(ws)
    PF = ctypes.POINTER(ctypes.c_float)
    W = (PF * n)(*ws)
    B = (PF * n)(*bs) if bs else None
    BN = (PF * n)(*bns)
    x = np.ascontiguousarray(x, dtype=np.float32)
    N, _, D, H, Wd = x.shape
    out = np.empty((N, c["output_nc"], D, H, Wd), dtype=np.float32)
    rc = lib().ref_unet_forward(
        _fp(x), _fp(out), N, D, H, Wd, c["input_nc"], c["output_nc"], c["num_downs"], c["ngf"],
        0 if c["norm"] == "batch" else 1, ctypes.c_float(c["norm_eps"]),
        0 if c["pooling"] == "Max" else 1, 0 if c["interp"] == "nearest" else 1, W, B, BN)
    if rc != 0:
        raise ValueError("unsupported volume shape for this U-Net depth")
    return out
