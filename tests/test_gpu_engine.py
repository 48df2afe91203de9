"""Parity of the CUDA engine (through the C ABI) against the CPU oracle and the
golden vectors minted from the reference.  Everything here needs a B200.

Tolerances (floating point); measured values of round 2 on B200 in brackets
(gpurun_out/parity_values.txt of the round, summarised in profiles/r2_summary.md):
  * LOOSE  engine (bf16 operands, fp32 accumulate) vs the fp32 oracle:
           rel-L2 <= 2e-2 and per-voxel cosine >= 0.998 [worst: 1.35e-2 / 0.99951; 7.9e-3 at 128^3].
           For scale: the reference itself under CPU bf16 autocast sits at rel-L2 1.8e-2 from its fp32 self.
  * TIGHT  engine vs the oracle rounding to bf16 at the same points (weights after
           BN folding, stored activations), fp32 accumulate: rel-L2 <= 8e-3 [worst: 6.2e-3; 3.7e-3 at
           128^3].  The two differ only by fp32 summation order, which flips an occasional bf16
           rounding (measured: 3 of 16384 activations after the third conv, one ulp
           each); later layers spread such a flip, which is why this gate is not
           tighter.  The first two layers are additionally held to 1e-3.
  * Degenerate inputs (all zeros, single impulses) make every voxel round the same
           way, so bf16 storage itself sits up to 0.18 rel-L2 from fp32 there (CPU
           emulation); for those the engine must be no worse than that emulation.
"""
import contextlib
import io

import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import CFG_6M, golden, min_cosine, rand_input, rel_l2
from oracle import unet_oracle as O

pytestmark = pytest.mark.gpu

LOOSE_REL, LOOSE_COS, TIGHT_REL = 2e-2, 0.998, 8e-3


def make_engine(cfg, state, flags=0):
    from anatomix_b200.engine import Engine
    eng = Engine(cfg, "cuda:0", flags=flags)
    eng.load_state(state)
    return eng


def check_against_oracle(cfg, state, x, y_gpu, tight=True, upconv=True):
    want = O.unet_forward(cfg, state, x)
    got = y_gpu.float().cpu()
    assert torch.isfinite(got).all()
    r, c = rel_l2(got, want), min_cosine(got, want)
    assert r <= LOOSE_REL and c >= LOOSE_COS, f"loose gate: rel-L2 {r:.3e}, min cosine {c:.5f}"
    rt = float("nan")
    if tight:
        emu = O.unet_forward(cfg, state, x, engine_rounding=True, emulate_upconv=upconv)
        rt = rel_l2(got, emu)
        assert rt <= TIGHT_REL, f"tight gate: rel-L2 {rt:.3e} vs bf16-emulating oracle"
    print(f"PARITY shape {tuple(x.shape)} loose rel-L2 {r:.3e} cos {c:.5f} tight rel-L2 {rt:.3e}")
    return r, c


def localize(eng_a, eng_b, shape):
    """Per-buffer comparison of two engines' workspaces after a forward."""
    n, _, d, h, w = shape
    wa, wb = eng_a.workspace(n, d, h, w), eng_b.workspace(n, d, h, w)
    sdt = O.engine_storage_dtype(eng_a.cfg)
    lines = []
    for i, (off, nb, lvl, grp) in enumerate(eng_a.buffer_table(n, d, h, w)):
        a = wa[off:off + nb].view(sdt).float()
        b = wb[off:off + nb].view(sdt).float()
        finite = torch.isfinite(a) & torch.isfinite(b)
        diff = (a - b)[finite].abs().max().item() if finite.any() else float("nan")
        lines.append(f"buffer {i:2d} level {lvl} groups {grp:3d}: max|diff| {diff:.4g} "
                     f"nonfinite {int((~finite).sum())}")
    return "\n".join(lines)


def test_primitive_selftest():
    from anatomix_b200 import _lib
    ok, report = _lib.selftest(0)
    print(report)
    assert ok, report


def small_cfg(**kw):
    cfg = dict(dimension=3, input_nc=1, output_nc=16, num_downs=2, ngf=16)
    cfg.update(kw)
    return cfg


@pytest.mark.parametrize("shape", [(1, 1, 8, 16, 8), (2, 1, 16, 32, 24)])
def test_simt_path_matches_oracle_small(shape):
    """Layout, stem conv, pooling, upsample/concat and weight packing, with the
    tensor cores out of the picture (debug conv kernel on the same packed weights)."""
    from anatomix_b200 import _lib
    cfg = small_cfg()
    state = O.random_state(cfg, seed=11)
    x = rand_input(shape, 3)
    eng = make_engine(cfg, state, flags=_lib.FLAG_FORCE_SIMT)
    y = eng.forward(x.cuda())
    torch.cuda.synchronize()
    check_against_oracle(cfg, state, x, y)


@pytest.mark.parametrize("shape,cfgkw", [
    ((1, 1, 8, 16, 8), {}),
    ((2, 1, 16, 32, 24), {}),
    ((1, 1, 16, 16, 16), dict(num_downs=3, ngf=32, output_nc=24)),
    ((1, 1, 8, 8, 8), dict(num_downs=1, ngf=64, output_nc=5, pooling="Avg", interp="trilinear",
                           activation="lrelu")),
])
def test_tensor_core_path_matches_simt_and_oracle_small(shape, cfgkw):
    from anatomix_b200 import _lib
    cfg = small_cfg(**cfgkw)
    state = O.random_state(cfg, seed=5)
    x = rand_input(shape, 9)
    ref_eng = make_engine(cfg, state, flags=_lib.FLAG_FORCE_SIMT)
    # same launch structure as the CUDA-core path; every intermediate tensor kept (compared layer by layer below)
    eng = make_engine(cfg, state, flags=_lib.FLAG_NO_UPCONV | _lib.FLAG_NO_WS_REUSE)
    xs = x.cuda()
    y_ref = ref_eng.forward(xs)
    y = eng.forward(xs)
    torch.cuda.synchronize()
    r = rel_l2(y.cpu(), y_ref.cpu())
    assert r < 5e-3, f"tensor-core vs CUDA-core conv: rel-L2 {r:.3e}\n" + localize(eng, ref_eng, shape)
    check_against_oracle(cfg, state, x, y, upconv=False)
    # default engine: the decoder's level-0 conv runs as low-resolution + skip launches
    y_up = make_engine(cfg, state).forward(xs)
    torch.cuda.synchronize()
    check_against_oracle(cfg, state, x, y_up)
    assert rel_l2(y_up.cpu(), y.cpu()) < 1e-2
    # the first layers have had no chance to spread a rounding flip yet
    from gpu_debug import compare
    full = dict(O.DEFAULTS); full.update(cfg)
    report = compare(eng, full, state, x)
    first = [float(l.split("interior")[1].split()[0]) for l in report.splitlines()[:2]]
    assert max(first) < 1e-3, report


@pytest.mark.parametrize("shape,cfgkw", [
    ((1, 1, 16, 16, 8), dict(norm="instance", pooling="Avg", interp="trilinear", norm_eps=1e-2)),
    ((2, 1, 8, 16, 24), dict(norm="instance", num_downs=1, ngf=32, output_nc=32, norm_eps=1e-2)),
    ((1, 1, 16, 16, 16), dict(norm="instance", num_downs=3, ngf=64, output_nc=8, pooling="Avg",
                              interp="trilinear", norm_eps=1e-2)),     # 512-wide bottleneck: channel splits
])
def test_instance_norm_path_small(shape, cfgkw):
    """InstanceNorm networks: statistics from the fp32 accumulators, fp16 storage,
    normalise + ReLU pass, AvgPool / trilinear, Cout > 256 split across CTAs."""
    from anatomix_b200 import _lib
    cfg = small_cfg(**cfgkw)
    state = O.random_state(cfg, seed=21)
    x = rand_input(shape, 13)
    ref_eng = make_engine(cfg, state, flags=_lib.FLAG_FORCE_SIMT | _lib.FLAG_NO_WS_REUSE)
    eng = make_engine(cfg, state)
    xs = x.cuda()
    y_ref = ref_eng.forward(xs)
    y = eng.forward(xs)
    torch.cuda.synchronize()
    r = rel_l2(y.cpu(), y_ref.cpu())
    assert r < 5e-3, f"tensor-core vs CUDA-core conv: rel-L2 {r:.3e}\n" + localize(eng, ref_eng, shape)
    check_against_oracle(cfg, state, x, y)
    y2 = eng.forward(xs)                       # statistics buffers are re-zeroed every forward
    torch.cuda.synchronize()
    assert rel_l2(y2.cpu(), y.cpu()) < 1e-3


def test_g4_dev_variant_94m_seeded_init():
    """anatomix-dev (94M) config with the reference constructor's seeded default init
    (its released weights are Hub-only): golden G4 from the unmodified reference."""
    from conftest import CFG_94M
    from anatomix_b200 import Unet
    g = golden("g4_94m_64.npz")
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = Unet(**CFG_94M)
    sd = m.state_dict()
    x = rand_input((1, 1, 64, 64, 64), 0)
    eng = make_engine(CFG_94M, sd)
    y = eng.forward(x.cuda()).cpu()
    want = torch.from_numpy(g["out_s2"])
    got = y[:, :, ::2, ::2, ::2]
    r, c = rel_l2(got, want), min_cosine(got, want)
    assert r <= LOOSE_REL and c >= LOOSE_COS, f"G4: rel-L2 {r:.3e}, min cosine {c:.5f}"
    emu = O.unet_forward(CFG_94M, sd, x, engine_rounding=True)
    # 24 re-normalised layers spread single rounding flips further than the 6M net does
    assert rel_l2(y, emu) <= 2e-2, f"G4 tight: {rel_l2(y, emu):.3e}"
    m = m.cuda().eval()
    with torch.no_grad():
        assert m.engine_ineligible_reason(x.cuda()) is None
        ym = m(x.cuda()).cpu()
    assert rel_l2(ym, y) < 1e-3


def test_g1_smallest_legal_size(state_6m):
    g = golden("g1_6m_32.npz")
    x = rand_input((1, 1, 32, 32, 32), 0)
    eng = make_engine(CFG_6M, state_6m)
    y = eng.forward(x.cuda())
    check_against_oracle(CFG_6M, state_6m, x, y)
    got, want = y.cpu(), torch.from_numpy(g["out"])
    assert rel_l2(got, want) <= LOOSE_REL and min_cosine(got, want) >= LOOSE_COS


def test_g2_batch_noncubic(state_6m):
    g = golden("g2_6m_2x32x48x32.npz")
    x = rand_input((2, 1, 32, 48, 32), 1)
    eng = make_engine(CFG_6M, state_6m)
    y = eng.forward(x.cuda())
    check_against_oracle(CFG_6M, state_6m, x, y)
    assert rel_l2(y.cpu()[:, :, ::2, ::2, ::2], torch.from_numpy(g["out_s2"])) <= LOOSE_REL


def test_odd_bottleneck_size_48(state_6m):
    x = rand_input((1, 1, 48, 32, 48), 4)      # bottleneck 3 x 2 x 3: both mirrors hit voxel 1
    eng = make_engine(CFG_6M, state_6m)
    check_against_oracle(CFG_6M, state_6m, x, eng.forward(x.cuda()))


def test_g3_headline_shape_against_golden(state_6m):
    g = golden("g3_6m_128.npz")
    x = rand_input((1, 1, 128, 128, 128), 0)
    eng = make_engine(CFG_6M, state_6m)
    y = eng.forward(x.cuda()).cpu()
    check_against_oracle(CFG_6M, state_6m, x, y)
    assert rel_l2(y[:, :, ::8, ::8, ::8], torch.from_numpy(g["out_s8"])) <= LOOSE_REL
    assert abs(y.double().mean().item() - g["mom"][0]) < 2e-2
    assert abs(y.double().std().item() - g["mom"][1]) < 2e-2


def test_g6_structured_inputs(state_6m):
    import importlib.util, os
    spec = importlib.util.spec_from_file_location(
        "make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = golden("g6_6m_structured.npz")
    eng = make_engine(CFG_6M, state_6m)
    for name, x in mg.structured_inputs().items():
        y = eng.forward(x.cuda()).cpu()
        want = torch.from_numpy(g[name])
        emu = O.unet_forward(CFG_6M, state_6m, x, engine_rounding=True)
        r = rel_l2(y[:, :, ::2, ::2, ::2], want)
        r_emu = rel_l2(emu[:, :, ::2, ::2, ::2], want)
        # correlated rounding on constant fields: one flipped ulp shifts a whole plateau
        assert rel_l2(y, emu) <= 2.5e-2, f"{name}: rel-L2 {rel_l2(y, emu):.3e} vs bf16-emulating oracle"
        assert r <= max(LOOSE_REL, 1.2 * r_emu + 1e-3), f"{name}: rel-L2 {r:.3e} (bf16 emulation {r_emu:.3e})"


def test_full_batch_properties(state_6m):
    """BASELINE config 2 size (8 x 128^3): determinism, per-sample independence
    (eval-BN has no cross-sample coupling) and oracle parity of the first and
    last sample."""
    xs = [rand_input((1, 1, 128, 128, 128), s) for s in (0, 1)]
    x = torch.cat([xs[0], xs[1], xs[0], xs[1], xs[1], xs[0], xs[0], xs[1]]).cuda()
    eng = make_engine(CFG_6M, state_6m)
    y1 = eng.forward(x).clone()
    y2 = eng.forward(x)
    torch.cuda.synchronize()
    assert torch.equal(y1, y2), "two runs on the same input differ"
    assert torch.equal(y1[0], y1[2]) and torch.equal(y1[0], y1[5]) and torch.equal(y1[1], y1[7])
    single = eng.forward(xs[1].cuda())
    assert torch.equal(single[0], y1[1]), "batched and single-volume results differ"
    check_against_oracle(CFG_6M, state_6m, xs[0], y1[0:1], tight=True)
    check_against_oracle(CFG_6M, state_6m, xs[1], y1[7:8], tight=True)


def test_module_routes_to_engine_and_back(state_6m):
    from anatomix_b200 import Unet
    with contextlib.redirect_stdout(io.StringIO()):
        m = Unet(**CFG_6M)
    m.load_state_dict(state_6m)
    m = m.cuda()
    x = rand_input((2, 1, 32, 32, 32), 2)
    xc = x.cuda()
    # train mode (the state load_from_hf returns): BatchNorm batch statistics -> stock torch path
    with torch.no_grad():
        assert "train mode" in m.engine_ineligible_reason(xc)
    m.eval()
    assert "autograd" in m.engine_ineligible_reason(xc)
    with torch.no_grad():
        assert m.engine_ineligible_reason(xc) is None
        y = m(xc)
        check_against_oracle(CFG_6M, state_6m, x, y)
        y_taps, taps = m(xc, layers=[8])          # a stored tensor: served by the engine too
        assert rel_l2(y.cpu(), y_taps.cpu()) <= LOOSE_REL
        # finetuning mutates parameters in place: the packed weights must follow
        m.model[65].weight.mul_(2.0)
        y2 = m(xc)
    assert rel_l2(y2.cpu(), 2 * y.cpu()) < 1e-2
    assert list(m._engine_binding().engines) == [(torch.device("cuda", 0), 0)]


@pytest.mark.parametrize("input_nc", [2, 3, 4])
def test_multichannel_stem(input_nc):
    """Stem with several input channels (K = 9*Cin taps spread over 2-3 tensor-core K chunks)."""
    cfg = small_cfg(input_nc=input_nc, num_downs=1)
    state = O.random_state(cfg, seed=30 + input_nc)
    x = rand_input((2, input_nc, 8, 16, 16), 31)
    eng = make_engine(cfg, state)
    check_against_oracle(cfg, state, x, eng.forward(x.cuda()))


def test_storage_type_override(state_6m):
    """ANX_FLAG_STORE_FP16 on a BatchNorm network: 8x smaller rounding error than bf16
    (inputs in [0,1] keep the activations far inside fp16 range)."""
    from anatomix_b200 import _lib
    x = rand_input((1, 1, 32, 32, 32), 0)
    want = O.unet_forward(CFG_6M, state_6m, x)
    y16 = make_engine(CFG_6M, state_6m, flags=_lib.FLAG_STORE_FP16).forward(x.cuda()).cpu()
    ybf = make_engine(CFG_6M, state_6m).forward(x.cuda()).cpu()
    r16, rbf = rel_l2(y16, want), rel_l2(ybf, want)
    assert r16 < 4e-3 and r16 < 0.4 * rbf, (r16, rbf)


def test_sliding_window_on_engine(state_6m):
    """Whole-scan extraction: engine as the window predictor vs the
    oracle as the predictor, same window grid and gaussian blend."""
    from anatomix_b200.sliding import sliding_window_features
    x = rand_input((1, 1, 48, 32, 64), 8)
    eng = make_engine(CFG_6M, state_6m)
    got = sliding_window_features(x.cuda(), (32, 32, 32), 6, eng.forward, overlap=0.5, mode="gaussian",
                                  sigma_scale=0.25).cpu()
    want = sliding_window_features(x, (32, 32, 32), 6, lambda t: O.unet_forward(CFG_6M, state_6m, t), overlap=0.5,
                                   mode="gaussian", sigma_scale=0.25)
    assert got.shape == (1, 16, 48, 32, 64)
    assert rel_l2(got, want) <= LOOSE_REL and min_cosine(got, want) >= LOOSE_COS


def test_host_buffer_entry_point(state_6m):
    """anx_engine_forward_host: pinned host buffers in and out, chunked upload / compute /
    download pipeline; must equal the device-buffer forward bit for bit."""
    eng = make_engine(CFG_6M, state_6m)
    for n in (1, 3, 8):
        x = rand_input((n, 1, 32, 32, 32), 20 + n)
        xh = x.pin_memory()
        yh = torch.empty((n, 16, 32, 32, 32), dtype=torch.float32).pin_memory()
        dev_in = torch.empty((n, 1, 32, 32, 32), device="cuda")
        dev_out = torch.empty((n, 16, 32, 32, 32), device="cuda")
        eng.forward_host(xh, yh, dev_in, dev_out)
        torch.cuda.synchronize()
        want = eng.forward(x.cuda()).cpu()
        assert torch.equal(yh, want), f"host-buffer forward differs at batch {n}"


def test_engine_reuse_across_shapes_and_streams(state_6m):
    """One engine, several shapes back to back (plan cache), and two CUDA streams with their own
    workspaces running concurrently (the engine is immutable after packing)."""
    import ctypes as C
    eng = make_engine(CFG_6M, state_6m)
    shapes = [(1, 1, 32, 32, 32), (2, 1, 32, 48, 32), (1, 1, 64, 32, 32), (1, 1, 32, 32, 32), (3, 1, 32, 32, 64)]
    outs = []
    for i, sh in enumerate(shapes):
        x = rand_input(sh, 40 + i)
        outs.append((x, eng.forward(x.cuda())))
    torch.cuda.synchronize()
    for x, y in outs:
        assert rel_l2(y.cpu(), O.unet_forward(CFG_6M, state_6m, x)) <= LOOSE_REL
    # two streams, distinct workspaces and outputs
    xa, xb = rand_input((1, 1, 32, 32, 32), 50).cuda(), rand_input((1, 1, 32, 32, 32), 51).cuda()
    want_a, want_b = eng.forward(xa).clone(), eng.forward(xb).clone()
    need = eng.workspace_bytes(1, 32, 32, 32)
    ws = [torch.empty(need, dtype=torch.uint8, device="cuda") for _ in range(2)]
    ys = [torch.empty_like(want_a) for _ in range(2)]
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for rep in range(3):
        for k, (s, x) in enumerate(zip(streams, (xa, xb))):
            st = eng.lib.anx_engine_forward(eng._h, x.data_ptr(), ys[k].data_ptr(), 1, 32, 32, 32, ws[k].data_ptr(), need,
                                            s.cuda_stream)
            assert st == 0
    torch.cuda.synchronize()
    assert torch.equal(ys[0], want_a) and torch.equal(ys[1], want_b)


def test_empty_batch_and_cpu_inputs_stay_on_torch(state_6m):
    from anatomix_b200 import Unet
    with contextlib.redirect_stdout(io.StringIO()):
        m = Unet(**CFG_6M)
    m.load_state_dict(state_6m)
    m = m.cuda().eval()
    with torch.no_grad():
        assert m.engine_ineligible_reason(torch.zeros(0, 1, 32, 32, 32, device="cuda")) == "empty batch"
        assert m(torch.zeros(0, 1, 32, 32, 32, device="cuda")).shape == (0, 16, 32, 32, 32)
        assert "autocast" in m.engine_ineligible_reason(torch.zeros(1, 1, 32, 32, 32, device="cuda")) \
            if torch.is_autocast_enabled() else True
        assert m.engine_ineligible_reason(torch.zeros(1, 1, 32, 32, 32, device="cuda", dtype=torch.float64)) == \
            "input is not fp32"


def test_errors_mirror_the_reference(state_6m):
    from anatomix_b200.engine import Engine, EngineError
    eng = make_engine(CFG_6M, state_6m)
    for bad in [(1, 1, 16, 32, 32), (1, 1, 72, 32, 32)]:
        with pytest.raises(ValueError):
            eng.forward(torch.zeros(*bad, device="cuda"))
    with pytest.raises(ValueError):
        eng.forward(torch.zeros(1, 2, 32, 32, 32, device="cuda"))
    fresh = Engine(CFG_6M, "cuda:0")
    with pytest.raises(EngineError) as ei:
        fresh.forward(torch.zeros(1, 1, 32, 32, 32, device="cuda"))
    assert "NOT_READY" in str(ei.value)
    with pytest.raises(EngineError):
        Engine(dict(CFG_6M, ngf=24), "cuda:0")   # widths the tensor-core tiles cannot take


# ------------------------------------------------------------------ rows next to the hot path
def test_feature_taps_on_the_engine_g7(state_6m):
    """`forward(x, layers=[...])` (network.py:475-529) served from the tensors the engine stores: against
    golden G7 minted from the reference (loose gate: 16-bit storage) and the 16-bit-emulating oracle."""
    from anatomix_b200 import Unet
    g = golden("g7_6m_taps.npz")
    ids = [int(i) for i in g["tap_ids"]]
    with contextlib.redirect_stdout(io.StringIO()):
        m = Unet(**CFG_6M)
    m.load_state_dict(state_6m)
    m = m.cuda().eval()
    x = rand_input((1, 1, 32, 32, 32), 3)
    with torch.no_grad():
        assert m.engine_ineligible_reason(x.cuda(), ids) is None
        y, taps = m(x.cuda(), layers=ids)
        assert m.engine_ineligible_reason(x.cuda(), [8, 59]) is None     # slot 59: a pre-norm conv output, engine too
        y_t, taps_t = m(x.cuda(), layers=[8, 59])
    torch.cuda.synchronize()
    assert len(taps) == len(ids) and taps[-1] is y                       # slot 65 is the output itself
    _, emu = O.unet_forward(CFG_6M, state_6m, x, layers=ids, engine_rounding=True)
    for i, t, e in zip(ids, taps, emu):
        got = t.float().cpu()
        want = torch.from_numpy(g[f"tap{i}"])
        sub = got[:, :, ::2, ::2, ::2] if got.shape[-1] > 4 else got
        assert sub.shape == want.shape, (i, sub.shape, want.shape)
        r = rel_l2(sub, want)
        assert r <= LOOSE_REL, f"tap {i}: rel-L2 {r:.3e} vs reference golden"
        rt = rel_l2(got, e)
        assert rt <= TIGHT_REL, f"tap {i}: rel-L2 {rt:.3e} vs 16-bit-emulating oracle"
    assert rel_l2(y_t.cpu(), y.cpu()) <= LOOSE_REL and len(taps_t) == 2
    # encode_only stops at layers[-1] (= 15 here) and returns the taps alone
    enc = [int(i) for i in g["enc_ids"]]
    with torch.no_grad():
        only = m(x.cuda(), layers=enc, encode_only=True)
    assert isinstance(only, list) and len(only) == int(g["enc_count"])
    for k, t in enumerate(only):
        r = rel_l2(t.float().cpu()[:, :, ::2, ::2, ::2], torch.from_numpy(g[f"enc{k}"]))
        assert r <= LOOSE_REL, f"encode_only tap {k}: {r:.3e}"


def test_feature_taps_instance_norm_network():
    cfg = small_cfg(norm="instance", pooling="Avg", interp="trilinear", norm_eps=1e-2, num_downs=2)
    state = O.random_state(cfg, seed=4)
    x = rand_input((2, 1, 16, 16, 8), 6)
    eng = make_engine(cfg, state)
    table = eng.tap_table()
    ids = sorted(table)
    y, taps = eng.forward_taps(x.cuda(), ids)
    torch.cuda.synchronize()
    _, emu = O.unet_forward(cfg, state, x, layers=ids, engine_rounding=True)
    _, ref = O.unet_forward(cfg, state, x, layers=ids)
    for i, t, e, f in zip(ids, taps, emu, ref):
        assert t.shape == e.shape, (i, t.shape, e.shape)
        # pre-norm conv slots (conv + bias, re-evaluated by an un-folded clone) have no counterpart among the tensors the
        # emulation models: they are held to the fp32 oracle; stored tensors to the 16-bit emulation
        want = f if table[i][5] >= 0 else e
        r = rel_l2(t.float().cpu(), want)
        assert r <= 2e-2, f"tap {i}: rel-L2 {r:.3e}"


def test_fused_output_head_matches_pointwise_conv(state_6m):
    """nn.Sequential(Unet, UnetOutBlock) (segmentation_utils.py:114-115) with the 1x1x1 conv evaluated in the
    last conv's epilogue: identical (fp32 rounding) to applying the conv to the engine's own features."""
    import torch.nn.functional as F
    from anatomix_b200 import Unet
    from anatomix_b200.heads import UnetOutBlock, fuse_output_head, scaled_features
    with contextlib.redirect_stdout(io.StringIO()):
        m = Unet(**CFG_6M)
    m.load_state_dict(state_6m)
    torch.manual_seed(3)
    head = UnetOutBlock(3, 16, 5, False)
    seq = fuse_output_head(m, head).cuda().eval()
    assert list(seq.state_dict())[-2:] == ["1.conv.conv.weight", "1.conv.conv.bias"]
    x = rand_input((2, 1, 32, 48, 32), 8)
    with torch.no_grad():
        assert seq.fused_ineligible_reason(x.cuda()) is None
        got = seq(x.cuda())
        feats = m(x.cuda())
        # fp64 on the CPU (cuDNN would use TF32 for this conv)
        want = F.conv3d(feats.cpu().double(), head.conv.conv.weight.cpu().double(), head.conv.conv.bias.cpu().double())
        assert got.shape == (2, 5, 32, 48, 32)
        assert rel_l2(got.cpu(), want.cpu()) < 1e-5
        # against the fp32 oracle end to end
        ref = F.conv3d(O.unet_forward(CFG_6M, state_6m, x), head.conv.conv.weight.cpu(), head.conv.conv.bias.cpu())
        assert rel_l2(got.cpu(), ref) <= LOOSE_REL
        # finetuning the head in place is picked up
        head.conv.conv.bias.add_(1.0)
        assert rel_l2(seq(x.cuda()).cpu(), (want + 1.0).cpu()) < 1e-5
        # registration: pred * downscale_feat_scalar folded into the epilogue (a diagonal head)
        scaled = scaled_features(m, 0.1)(x.cuda())
        assert rel_l2(scaled.cpu(), (feats * 0.1).cpu()) < 1e-6
    seq.train()
    with torch.no_grad():
        assert seq.fused_ineligible_reason(x.cuda()) is not None      # BatchNorm batch statistics: stock torch


@pytest.mark.parametrize("k,shape", [(2, (1, 16, 32, 32, 32)), (3, (2, 5, 20, 17, 31)), (4, (1, 3, 16, 24, 36))])
def test_scaled_average_pooling_kernel(k, shape):
    """scale * F.avg_pool3d(x, k, stride=k) (run_convex_adam_with_network_feats.py:166-167, 198-205)."""
    import torch.nn.functional as F
    from anatomix_b200.heads import avg_pool3d_scaled
    x = rand_input(shape, 21).cuda()
    got = avg_pool3d_scaled(x, k, 0.1)
    want = F.avg_pool3d(x, k, stride=k) * 0.1
    assert got.shape == want.shape
    assert torch.allclose(got, want, atol=1e-6, rtol=1e-5)


def test_native_blend_kernel_matches_torch_accumulate():
    """anx_blend_window_f32 (the accumulate step of the sliding-window inferer) against the same window
    grid accumulated with torch ops on the CPU; ragged volume, overlapping windows, gaussian weights."""
    from anatomix_b200.sliding import sliding_window_features
    pred = lambda p: torch.cat([p * 1.0, p * p, 1.0 - p], dim=1)
    x = rand_input((2, 1, 40, 37, 52), 17)
    want = sliding_window_features(x, (16, 16, 24), 4, pred, overlap=0.6, mode="gaussian", sigma_scale=0.25)
    got = sliding_window_features(x.cuda(), (16, 16, 24), 4, pred, overlap=0.6, mode="gaussian", sigma_scale=0.25)
    assert got.shape == want.shape == (2, 3, 40, 37, 52)
    assert torch.allclose(got.cpu(), want, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("shape", [(1, 1, 16, 16, 128), (2, 1, 32, 24, 256), (3, 1, 16, 8, 128), (1, 1, 48, 40, 128),
                                   (1, 1, 16, 16, 224)])      # 224 = one whole x tile + one pulled back to the border
def test_row_kernel_matches_generic_kernel(shape):
    """conv3_rows_kernel (dy and dz folded into N, lanes = one 128-voxel row) against the generic tile kernel
    on the same packed 16-bit operands: only the fp32 summation order differs."""
    cfg = small_cfg()
    state = O.random_state(cfg, seed=13)
    x = rand_input(shape, 19)
    from anatomix_b200 import _lib
    ref = make_engine(cfg, state, flags=_lib.FLAG_NO_ROWS).forward(x.cuda())
    eng = make_engine(cfg, state)
    got = eng.forward(x.cuda())
    torch.cuda.synchronize()
    r = rel_l2(got.cpu(), ref.cpu())
    # the two engines also run different stem kernels (row form / tile form): both fp32-accurate, but ~4e-5 of the
    # stored stem values round the other way (test_row_form_stem_matches_tile_form_stem) and later layers spread that
    assert r < 6e-3, f"row kernels vs tile kernels: rel-L2 {r:.3e}"
    check_against_oracle(cfg, state, x, got)
    # fused head on the row kernel's fp32 epilogue
    torch.manual_seed(1)
    hw, hb = torch.randn(5, 16), torch.randn(5)
    eng.set_head(hw, hb)
    got_h = eng.forward(x.cuda())
    want_h = torch.einsum("kc,ncdhw->nkdhw", hw.double(), got.cpu().double()) + hb.double().view(1, -1, 1, 1, 1)
    assert rel_l2(got_h.cpu(), want_h) < 1e-5


@pytest.mark.parametrize("cfgkw,flags", [
    (dict(norm="none", activation="lrelu"), 0),            # no norm (zero shift), leaky ReLU in the row epilogue
    (dict(), 2),                                           # ANX_FLAG_STORE_FP16: fp16 operands through the row kernel
    (dict(num_downs=1, output_nc=5), 0),                   # 5 output channels: masked fp32 stores of the last conv
])
def test_row_kernel_variants(cfgkw, flags):
    cfg = small_cfg(**cfgkw)
    state = O.random_state(cfg, seed=21)
    x = rand_input((1, 1, 16, 24, 128), 23)
    got = make_engine(cfg, state, flags=flags).forward(x.cuda())
    torch.cuda.synchronize()
    want = O.unet_forward(cfg, state, x)
    assert torch.isfinite(got).all()
    r, c = rel_l2(got.cpu(), want), min_cosine(got.cpu(), want)
    assert r <= LOOSE_REL and c >= LOOSE_COS, f"rel-L2 {r:.3e}, min cosine {c:.5f}"


def _seeded_94m():
    from conftest import CFG_94M
    from anatomix_b200 import Unet
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = Unet(**CFG_94M)
    return CFG_94M, m.state_dict()


def test_g8_94m_at_the_baseline_volume_size():
    """BASELINE configs[2] network at its volume size (1 x 128^3) against golden G8 minted from the unmodified
    reference (seeded default init: the released 94M weights are Hub-only)."""
    cfg, sd = _seeded_94m()
    g = golden("g8_94m_128.npz")
    x = rand_input((1, 1, 128, 128, 128), 0)
    y = make_engine(cfg, sd).forward(x.cuda()).cpu()
    assert torch.isfinite(y).all()
    got, want = y[:, :, ::8, ::8, ::8], torch.from_numpy(g["out_s8"])
    r, c = rel_l2(got, want), min_cosine(got, want)
    print(f"G8 94M 1x128^3: rel-L2 {r:.3e}, min cosine {c:.5f}")
    assert r <= LOOSE_REL and c >= LOOSE_COS, f"G8: rel-L2 {r:.3e}, min cosine {c:.5f}"
    assert abs(y.double().mean().item() - g["mom"][0]) < 2e-2 and abs(y.double().std().item() - g["mom"][1]) < 2e-2
    # the batch of BASELINE configs[2] (4 volumes): per-sample independence, determinism
    xb = torch.cat([x, rand_input((1, 1, 128, 128, 128), 1), x, x]).cuda()
    eng = make_engine(cfg, sd)
    yb = eng.forward(xb)
    torch.cuda.synchronize()
    assert torch.equal(yb[0], yb[2]) and torch.equal(yb[0], yb[3]) and torch.equal(yb[0].cpu(), y[0])


def test_g9_raw_prenorm_store_survives_large_conv_outputs():
    """The InstanceNorm path stores the RAW conv output in fp16 and normalises it in place afterwards.  With every
    conv weight scaled x30 the raw tensors are 30x larger (the normalisation undoes the scale): neither overflow
    nor lost resolution may show -- golden G9 from the reference with the same scaled weights."""
    cfg, sd = _seeded_94m()
    sd = {k: (v * 30.0 if k.endswith(".weight") else v) for k, v in sd.items()}
    g = golden("g9_94m_64_w30.npz")
    x = rand_input((1, 1, 64, 64, 64), 0)
    y = make_engine(cfg, sd).forward(x.cuda()).cpu()
    assert torch.isfinite(y).all()
    got, want = y[:, :, ::4, ::4, ::4], torch.from_numpy(g["out_s4"])
    r, c = rel_l2(got, want), min_cosine(got, want)
    print(f"G9 94M x30 weights: rel-L2 {r:.3e}, min cosine {c:.5f}")
    assert r <= LOOSE_REL and c >= LOOSE_COS, f"G9: rel-L2 {r:.3e}, min cosine {c:.5f}"


@pytest.mark.parametrize("shape", [(2, 1, 32, 32, 32), (1, 1, 32, 32, 128)])    # generic tile kernel / row kernel
def test_channels_last_16bit_output_and_widen(state_6m, shape):
    """ANX_PAYLOAD_CL16: the features as 16-bit channels-last [N, D, H, W, C] are the fp32 output rounded once."""
    x = rand_input(shape, 12)
    eng = make_engine(CFG_6M, state_6m)
    y = eng.forward(x.cuda())
    y16 = eng.forward_cl16(x.cuda())
    torch.cuda.synchronize()
    assert y16.dtype == torch.bfloat16 and tuple(y16.shape) == (shape[0],) + shape[2:] + (16,)
    assert torch.equal(y16.permute(0, 4, 1, 2, 3), y.to(torch.bfloat16))
    wide = eng.widen(y16)
    torch.cuda.synchronize()
    assert torch.equal(wide, y.to(torch.bfloat16).float())
    # host-buffer entry point with the 16-bit payload
    n, _, d, h, w = shape
    xh, oh = x.pin_memory(), torch.empty((n, d, h, w, 16), dtype=torch.bfloat16).pin_memory()
    dev_in, dev_out = torch.empty(shape, device="cuda"), torch.empty((n, d, h, w, 16), dtype=torch.bfloat16, device="cuda")
    eng.forward_host_cl16(xh, oh, dev_in, dev_out)
    torch.cuda.synchronize()
    assert torch.equal(oh, y16.cpu())


def test_channels_last_output_of_a_32_channel_instance_norm_network():
    cfg = small_cfg(norm="instance", num_downs=1, ngf=32, output_nc=32, norm_eps=1e-2)
    state = O.random_state(cfg, seed=41)
    x = rand_input((1, 1, 8, 16, 24), 42)
    eng = make_engine(cfg, state)
    y, y16 = eng.forward(x.cuda()), eng.forward_cl16(x.cuda())
    torch.cuda.synchronize()
    assert y16.dtype == torch.float16 and torch.equal(y16.permute(0, 4, 1, 2, 3), y.to(torch.float16))


def test_forward_argument_validation(state_6m):
    eng = make_engine(CFG_6M, state_6m)
    x = rand_input((1, 1, 32, 32, 32), 0).cuda()
    with pytest.raises(ValueError):
        eng.forward(x, out=torch.empty((1, 16, 32, 32, 16), device="cuda"))
    with pytest.raises(ValueError):
        eng.forward(x, out=torch.empty((1, 16, 32, 32, 32), device="cuda", dtype=torch.float16))
    with pytest.raises(ValueError):
        eng.forward(x, out=torch.empty((1, 16, 32, 32, 64), device="cuda")[..., ::2])


def test_data_edits_reach_the_engine(state_6m, monkeypatch):
    """In-place edits through `.data` bump no version counter (reference pretraining_networks.py:695-713 initialises
    weights that way): `invalidate_engine()` re-packs at once, the periodic content check catches it otherwise."""
    from anatomix_b200 import Unet
    from anatomix_b200 import engine as E
    with contextlib.redirect_stdout(io.StringIO()):
        m = Unet(**CFG_6M)
    m.load_state_dict(state_6m)
    m = m.cuda().eval()
    x = rand_input((1, 1, 32, 32, 32), 2).cuda()
    with torch.no_grad():
        y = m(x)
        m.model[65].weight.data.mul_(2.0)
        m.invalidate_engine()
        y2 = m(x)
        assert rel_l2(y2.cpu(), 2 * y.cpu()) < 1e-2
        monkeypatch.setattr(E, "VERIFY_EVERY", 1)
        m.model[65].weight.data.mul_(0.5)
        y3 = m(x)
        assert rel_l2(y3.cpu(), y.cpu()) < 1e-2


def test_hooks_send_the_call_down_the_stock_loop(state_6m):
    from anatomix_b200 import Unet
    with contextlib.redirect_stdout(io.StringIO()):
        m = Unet(**CFG_6M)
    m.load_state_dict(state_6m)
    m = m.cuda().eval()
    seen = []
    h = m.model[3].register_forward_hook(lambda mod, i, o: seen.append(tuple(o.shape)))
    x = rand_input((1, 1, 32, 32, 32), 2).cuda()
    with torch.no_grad():
        assert "hooks" in m.engine_ineligible_reason(x)
        m(x)
        assert seen == [(1, 16, 32, 32, 32)]
        h.remove()
        assert m.engine_ineligible_reason(x) is None


@pytest.mark.parametrize("cfgkw,shape", [({}, (1, 1, 32, 32, 128)),                       # 6M: row kernel, fused pooling, upconv
                                         (dict(num_downs=2, pooling="Avg", interp="trilinear"), (2, 1, 16, 16, 24))])
def test_open_slab_faces_are_left_to_the_neighbour(state_6m, cfgkw, shape):
    """Depth-slab mode: a shell plane at a face that borders another slab belongs to the NEIGHBOUR, whose boundary
    plane may arrive before this slab's producer of that tensor has even started (anx_engine_forward_slab), so no
    kernel of this slab may write there.  Sentinel-fill the workspace, run a forward with both faces open, and
    check that every such plane is untouched; with the faces closed the same planes hold the reflect copies."""
    from anatomix_b200 import _lib
    from anatomix_b200.engine import Engine
    if cfgkw:
        cfg = small_cfg(**cfgkw)
        state = O.random_state(cfg, seed=3)
    else:
        cfg, state = CFG_6M, state_6m
    n, _, d, h, w = shape
    eng = Engine(cfg, "cuda:0", flags=_lib.FLAG_DEPTH_HALO_INPUT)
    eng.load_state(state)
    x = torch.rand(n, 1, d + 2, h, w, device="cuda")
    out = torch.empty((n, cfg["output_nc"], d, h, w), device="cuda")
    ws = eng.workspace(n, d, h, w)
    steps = len(eng.step_table())
    for open_faces in (True, False):
        eng.set_slab(open_faces, open_faces, 3 * d if open_faces else 0)
        ws.fill_(0x7B)
        eng.run_steps(x, out, 0, steps)
        torch.cuda.synchronize()
        touched = 0
        for off, nbytes, level, groups in eng.buffer_table(n, d, h, w):
            dl, hl, wl = d >> level, h >> level, w >> level
            plane = (hl + 2) * eng.row_layout(wl)[1] * 16
            v = ws[off:off + n * groups * (dl + 2) * plane].view(n, groups, dl + 2, plane)
            touched += int((v[:, :, 0] != 0x7B).sum().item()) + int((v[:, :, dl + 1] != 0x7B).sum().item())
        if open_faces:
            assert touched == 0, f"{touched} bytes of neighbour-owned shell planes were written"
        else:
            assert touched > 0
    eng.set_slab(False, False, 0)


@pytest.mark.parametrize("shape,c", [((2, 16, 8, 12, 20), 16), ((1, 32, 4, 6, 10), 32), ((1, 5, 3, 3, 7), 5)])
def test_voxelwise_channel_normalisation(shape, c):
    """README.md:13,49 of the reference: unit norm or zero mean / unit std across channels, per voxel."""
    import torch.nn.functional as F
    from anatomix_b200.heads import normalize_features
    x = torch.randn(*shape, generator=torch.Generator().manual_seed(8)) * 3 + 1
    got = normalize_features(x.cuda(), "unit").cpu()
    assert torch.allclose(got, F.normalize(x, dim=1), atol=1e-6, rtol=1e-5)
    got = normalize_features(x.cuda(), "zscore").cpu()
    want = (x - x.mean(1, keepdim=True)) / x.std(1, keepdim=True)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-4)
    y = x.cuda().clone()
    assert normalize_features(y, "unit", inplace=True) is y and torch.allclose(y.cpu(), F.normalize(x, dim=1), atol=1e-6, rtol=1e-5)


@pytest.mark.parametrize("shape", [(2, 1, 32, 32, 32), (1, 1, 32, 32, 128)])
def test_zero_copy_concat_behind_other_features(state_6m, shape):
    """instance_optimization.py:16-119: MIND-SSC descriptors (12 channels) in front of the network features."""
    from anatomix_b200 import Unet
    from anatomix_b200.heads import features_behind
    with contextlib.redirect_stdout(io.StringIO()):
        m = Unet(**CFG_6M)
    m.load_state_dict(state_6m)
    m = m.cuda().eval()
    x = rand_input(shape, 5).cuda()
    front = torch.randn((shape[0], 12) + shape[2:], device="cuda")
    with torch.no_grad():
        got = features_behind(m, x, front)
        want = torch.cat([front, m(x)], dim=1)
    assert got.shape == want.shape and torch.equal(got, want)


@pytest.mark.parametrize("shape,halo", [((2, 1, 32, 32, 128), False), ((1, 1, 16, 8, 256), False), ((1, 1, 32, 16, 128), True),
                                        ((1, 1, 16, 16, 240), False)])     # last x tile pulled back to the border
def test_row_form_stem_matches_tile_form_stem(shape, halo):
    """The first conv in row form (conv3_rows_kernel<.., STEM>: one N = 144 MMA per input row, K = the three dx taps of
    the hi / lo split) against the tile-form stem kernel and the fp32 oracle conv: both keep fp32-level accuracy, so the
    stored 16-bit tensors agree except where a value sits on a rounding boundary."""
    import torch.nn.functional as F
    from anatomix_b200 import _lib
    from anatomix_b200.engine import Engine
    cfg = small_cfg(num_downs=1)
    state = O.random_state(cfg, seed=17)
    n, _, d, h, w = shape
    x = rand_input((n, 1, d + (2 if halo else 0), h, w), 18).cuda()
    flags = _lib.FLAG_DEPTH_HALO_INPUT if halo else 0
    bufs = []
    for extra in (0, _lib.FLAG_NO_ROWS):
        eng = Engine(cfg, "cuda:0", flags=flags | extra)
        eng.load_state(state)
        if halo:
            eng.set_slab(True, True, 3 * d)
        out = torch.empty((n, 16, d, h, w), device="cuda")
        ws = eng.workspace(n, d, h, w)
        ws.zero_()
        eng.run_steps(x, out, 0, 1)
        torch.cuda.synchronize()
        kind, buf, goff, groups, name = eng.step_table()[0]
        off, nb, lvl, grp = eng.buffer_table(n, d, h, w)[buf]
        pitch = eng.row_layout(w)[1]
        t = ws[off:off + nb].view(torch.bfloat16)[: n * grp * (d + 2) * (h + 2) * pitch * 8]
        bufs.append(t.view(n, grp, d + 2, h + 2, pitch, 8).float().clone())
    a, b = bufs
    if halo:                                  # neighbour-owned shell planes are not written by either kernel
        a, b = a[:, :, 1:-1], b[:, :, 1:-1]
    differ = (a != b).float().mean().item()
    rel = ((a - b).abs() / b.abs().clamp_min(1e-3)).max().item()
    print(f"row-form vs tile-form stem: {differ:.2e} of the stored values differ, worst relative difference {rel:.2e}")
    assert differ < 2e-3 and rel <= 2 ** -7, (differ, rel)
    # and against the fp32 reference conv + folded BatchNorm + ReLU (network.py:309-326), interior voxels
    xi = x[:, :, 1:-1] if halo else x
    if halo:
        xp = F.pad(x, (1, 1, 1, 1, 0, 0), mode="reflect")
    else:
        xp = F.pad(xi, (1, 1, 1, 1, 1, 1), mode="reflect")
    wgt, g, bta, mu, var = (state[k].cuda() for k in ("model.0.weight", "model.1.weight", "model.1.bias",
                                                      "model.1.running_mean", "model.1.running_var"))
    want = F.relu(F.batch_norm(F.conv3d(xp, wgt), mu, var, g, bta, False, 0.0, 1e-5))
    got = bufs[0][:, :, 1:-1, 1:-1, 1:w + 1].permute(0, 1, 5, 2, 3, 4).reshape(n, 16, d, h, w)
    r = rel_l2(got.cpu(), want.cpu())
    assert r < 4e-3, f"row-form stem vs fp32 conv: rel-L2 {r:.3e} (bf16 storage alone is ~2e-3)"


def test_prenorm_conv_taps_g10(state_6m):
    """Taps at conv slots that a norm follows (network.py:504-515; the pretraining defaults): the reference returns the
    conv output BEFORE the norm.  The engine re-evaluates it with an un-folded clone of the conv right after that
    conv's launch (anx_engine_set_tap_conv / anx_engine_export_prenorm_tap).  Golden G10 from the unmodified reference."""
    from anatomix_b200 import Unet
    g = golden("g10_prenorm_taps.npz")
    ids = [int(i) for i in g["tap_ids"]]
    with contextlib.redirect_stdout(io.StringIO()):
        m = Unet(**CFG_6M)
    m.load_state_dict(state_6m)
    m = m.cuda().eval()
    x = rand_input((1, 1, 32, 32, 32), 5)
    with torch.no_grad():
        assert m.engine_ineligible_reason(x.cuda(), ids) is None
        y, taps = m(x.cuda(), layers=ids)
        # mixed with stored tensors and with encode_only
        y2, taps2 = m(x.cuda(), layers=[3, 8, 9, 10])
        only = m(x.cuda(), layers=[0, 10], encode_only=True)
    torch.cuda.synchronize()
    assert rel_l2(y.cpu()[:, :, ::2, ::2, ::2], torch.from_numpy(g["out_s2"])) <= LOOSE_REL
    for i, t in zip(ids, taps):
        got = t.float().cpu()
        want = torch.from_numpy(g[f"tap{i}"])
        sub = got[:, :, ::2, ::2, ::2] if got.shape[-1] > 4 else got
        assert sub.shape == want.shape, (i, sub.shape, want.shape)
        r = rel_l2(sub, want)
        print(f"pre-norm tap {i}: rel-L2 {r:.3e}")
        assert r <= LOOSE_REL, f"pre-norm tap {i}: rel-L2 {r:.3e} vs reference golden"
    assert torch.equal(taps2[0], taps[1]) and torch.equal(taps2[3], taps[3]) and len(taps2) == 4
    assert len(only) == 2 and torch.equal(only[0], taps[0]) and torch.equal(only[1], taps[3])
    # weights edited in place: the clone follows the re-pack
    with torch.no_grad():
        m.model[3].weight.mul_(2.0)
        _, t2 = m(x.cuda(), layers=[3])
    assert rel_l2(t2[0].cpu(), 2 * taps[1].cpu()) < 1e-2
    # row-kernel shape (128 wide): the clones of the stem and of the 16 -> 16 convs run on the row kernels
    xw = rand_input((1, 1, 32, 32, 128), 6)
    with torch.no_grad():
        m.model[3].weight.mul_(0.5)
        _, tw = m(xw.cuda(), layers=[0, 3, 62])
    _, want = O.unet_forward(CFG_6M, state_6m, xw, layers=[0, 3, 62])
    for i, a, b in zip([0, 3, 62], tw, want):
        assert rel_l2(a.cpu(), b) <= LOOSE_REL, f"pre-norm tap {i} at 128 wide: {rel_l2(a.cpu(), b):.3e}"


def test_prenorm_conv_taps_instance_norm_94m():
    cfg, sd = _seeded_94m()
    g = golden("g10_prenorm_taps.npz")
    ids = [int(i) for i in g["tap_ids_94m"]]
    from anatomix_b200 import Unet
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        m = Unet(**cfg)
    m = m.cuda().eval()
    x = rand_input((1, 1, 64, 64, 64), 0)
    with torch.no_grad():
        assert m.engine_ineligible_reason(x.cuda(), ids) is None
        _, taps = m(x.cuda(), layers=ids)
    for i, t in zip(ids, taps):
        got = t.float().cpu()
        want = torch.from_numpy(g[f"m94_tap{i}"])
        sub = got[:, :, ::4, ::4, ::4] if got.shape[-1] > 4 else got
        r = rel_l2(sub, want)
        print(f"94M pre-norm tap {i}: rel-L2 {r:.3e}")
        assert r <= LOOSE_REL, f"94M pre-norm tap {i}: rel-L2 {r:.3e}"


def test_pipelined_host_calls(state_6m):
    """anx_engine_forward_host_pipelined: back-to-back host-buffer forwards whose downloads overlap the next call;
    two {device output, host output} sets alternate, a third and fourth call reuse them (the convs wait for the
    download that still reads the buffer).  Every result equals the device-resident forward."""
    eng = make_engine(CFG_6M, state_6m)
    shape = (2, 1, 32, 32, 128)
    xs = [rand_input(shape, 40 + i).pin_memory() for i in range(4)]
    want = [eng.forward(x.cuda()).cpu() for x in xs]
    dev_in = torch.empty(shape, device="cuda")
    outs = [torch.empty((2, 16, 32, 32, 128), device="cuda") for _ in range(2)]
    hosts = [torch.empty((2, 16, 32, 32, 128)).pin_memory() for _ in range(4)]
    for i in range(4):
        eng.forward_host_pipelined(xs[i], hosts[i], dev_in, outs[i % 2])
    eng.host_wait()
    torch.cuda.synchronize()
    for i in range(4):
        assert torch.equal(hosts[i], want[i]), f"pipelined call {i} differs"
    # a plain call afterwards still completes on the stream by itself
    y = torch.empty((2, 16, 32, 32, 128)).pin_memory()
    eng.forward_host(xs[0], y, dev_in, outs[0])
    torch.cuda.synchronize()
    assert torch.equal(y, want[0])
