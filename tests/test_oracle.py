"""Pins the oracle (torch-functional restatement and plain-C restatement) to the
golden vectors minted from the unmodified reference (tests/golden/make_golden.py)."""
import numpy as np
import pytest
import torch

from conftest import CFG_6M, CFG_94M, golden, rand_input
from oracle import unet_oracle as O
from oracle.cref import unet_forward_c

TOL = 2e-4   # fp32 reorder noise between two CPU evaluation orders (3.5e-5 observed)


def sub(t, s):
    return t[:, :, ::s, ::s, ::s].numpy()


def test_layer_program_matches_reference_indices():
    prog, enc, dec = O.layer_program(1, 16, 4, 16, "batch", "relu", "none", True, True)
    assert len(prog) == 66 and enc == [8, 15, 22, 29] and dec == [37, 44, 51, 58]
    assert [p[1] for p in prog if p[0] == "conv"][-1] == 65
    prog, enc, dec = O.layer_program(1, 32, 5, 32, "instance", "relu", "none", True, True)
    assert len(prog) == 80 and enc == [8, 15, 22, 29, 36] and dec == [44, 51, 58, 65, 72]


def test_g1_full_output_and_taps(state_6m):
    g = golden("g1_6m_32.npz")
    taps_ids = [int(i) for i in g["tap_ids"]]
    y, taps = O.unet_forward(CFG_6M, state_6m, rand_input((1, 1, 32, 32, 32), 0), layers=taps_ids)
    np.testing.assert_allclose(y.numpy(), g["out"], atol=TOL, rtol=1e-4)
    for i, t in zip(taps_ids, taps):
        want = g[f"tap{i}"]
        got = sub(t, 2) if t.shape[-1] > 4 else t.numpy()
        np.testing.assert_allclose(got, want, atol=TOL, rtol=1e-4, err_msg=f"tap {i}")


def test_g2_batch_noncubic(state_6m):
    g = golden("g2_6m_2x32x48x32.npz")
    y = O.unet_forward(CFG_6M, state_6m, rand_input((2, 1, 32, 48, 32), 1))
    np.testing.assert_allclose(sub(y, 2), g["out_s2"], atol=TOL, rtol=1e-4)


def test_g3_headline_shape(state_6m):
    g = golden("g3_6m_128.npz")
    y = O.unet_forward(CFG_6M, state_6m, rand_input((1, 1, 128, 128, 128), 0))
    np.testing.assert_allclose(sub(y, 8), g["out_s8"], atol=TOL, rtol=1e-4)
    np.testing.assert_allclose(y[0, :4, 64, 64, 64].numpy(), g["probe"], atol=TOL)
    # survey-time fingerprint of the reference
    np.testing.assert_allclose(g["probe"], [-1.573407, 4.128320, 1.144190, 2.786164], atol=2e-4)
    assert abs(y.double().mean().item() - g["mom"][0]) < 1e-5


def test_g4_94m_instance_avg_trilinear():
    from anatomix_b200.unet import Unet
    g = golden("g4_94m_64.npz")
    torch.manual_seed(0)
    m = Unet(**CFG_94M)          # same RNG consumption as the reference constructor
    sd = m.state_dict()
    for name, sums in zip(g["param_names"], g["param_sums"]):
        t = sd[str(name)].double()
        assert abs(t.sum().item() - sums[0]) < 1e-6 * max(1.0, abs(sums[0])), name
    ids = [int(i) for i in g["tap_ids"]]
    y, taps = O.unet_forward(CFG_94M, sd, rand_input((1, 1, 64, 64, 64), 0), layers=ids)
    np.testing.assert_allclose(sub(y, 2), g["out_s2"], atol=5e-4, rtol=1e-3)
    for i, t in zip(ids, taps):
        got = sub(t, 4) if t.shape[-1] > 4 else t.numpy()
        np.testing.assert_allclose(got, g[f"tap{i}"], atol=5e-4, rtol=1e-3, err_msg=f"tap {i}")


def test_g9_94m_with_conv_weights_scaled_by_30():
    """Golden G9 (the reference with every conv weight x30: InstanceNorm undoes the scale up to its eps)."""
    import contextlib, io
    from anatomix_b200 import Unet
    g = golden("g9_94m_64_w30.npz")
    torch.manual_seed(0)
    with contextlib.redirect_stdout(io.StringIO()):
        sd = Unet(**CFG_94M).state_dict()
    sd = {k: (v * 30.0 if k.endswith(".weight") else v) for k, v in sd.items()}
    y = O.unet_forward(CFG_94M, sd, rand_input((1, 1, 64, 64, 64), 0))
    np.testing.assert_allclose(y[:, :, ::4, ::4, ::4].numpy(), g["out_s4"], atol=5e-4, rtol=1e-3)


def test_g10_prenorm_conv_taps(state_6m):
    """Golden G10: taps at conv slots that a norm follows hold the PRE-norm conv output (network.py:504-515)."""
    g = golden("g10_prenorm_taps.npz")
    ids = [int(i) for i in g["tap_ids"]]
    y, taps = O.unet_forward(CFG_6M, state_6m, rand_input((1, 1, 32, 32, 32), 5), layers=ids)
    np.testing.assert_allclose(sub(y, 2), g["out_s2"], atol=TOL, rtol=1e-4)
    for i, t in zip(ids, taps):
        got = sub(t, 2) if t.shape[-1] > 4 else t.numpy()
        np.testing.assert_allclose(got, g[f"tap{i}"], atol=TOL, rtol=1e-4, err_msg=f"tap {i}")


def test_g5_train_mode_batch_stats(state_6m):
    g = golden("g5_6m_train.npz")
    y = O.unet_forward(CFG_6M, state_6m, rand_input((1, 1, 32, 32, 32), 0), training=True)
    np.testing.assert_allclose(sub(y, 2), g["out_s2"], atol=2e-3, rtol=1e-3)


def test_g6_structured_inputs(state_6m):
    import importlib.util, os
    spec = importlib.util.spec_from_file_location(
        "make_golden", os.path.join(os.path.dirname(__file__), "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    g = golden("g6_6m_structured.npz")
    for name, x in mg.structured_inputs().items():
        y = O.unet_forward(CFG_6M, state_6m, x)
        np.testing.assert_allclose(sub(y, 2), g[name], atol=TOL, rtol=1e-4, err_msg=name)


def test_g7_feature_taps_and_encode_only(state_6m):
    """Tap branch (network.py:475-529): norm-slot taps hold post-activation values (in-place ReLU),
    Upsample-slot taps the concat, `encode_only` stops at layers[-1]."""
    g = golden("g7_6m_taps.npz")
    ids = [int(i) for i in g["tap_ids"]]
    x = rand_input((1, 1, 32, 32, 32), 3)
    y, taps = O.unet_forward(CFG_6M, state_6m, x, layers=ids)
    np.testing.assert_allclose(sub(y, 2), g["out_s2"], atol=TOL, rtol=1e-4)
    assert len(taps) == len(ids)
    for i, t in zip(ids, taps):
        got = sub(t, 2) if t.shape[-1] > 4 else t.numpy()
        np.testing.assert_allclose(got, g[f"tap{i}"], atol=TOL, rtol=1e-4, err_msg=f"tap {i}")
    assert taps[0].min() >= 0 and torch.equal(taps[0], taps[1])          # slot 1 (BatchNorm) aliases slot 2 (ReLU)
    only = O.unet_forward(CFG_6M, state_6m, x, layers=[int(i) for i in g["enc_ids"]], encode_only=True)
    assert isinstance(only, list) and len(only) == int(g["enc_count"]) == 2
    for k, t in enumerate(only):
        np.testing.assert_allclose(sub(t, 2), g[f"enc{k}"], atol=TOL, rtol=1e-4, err_msg=f"encode_only tap {k}")


def test_c_restatement_matches_golden_g1(state_6m):
    g = golden("g1_6m_32.npz")
    y = unet_forward_c(CFG_6M, state_6m, rand_input((1, 1, 32, 32, 32), 0).numpy())
    np.testing.assert_allclose(y, g["out"], atol=TOL, rtol=1e-4)


def test_c_restatement_instance_variant_small():
    cfg = dict(CFG_94M, num_downs=2, ngf=16, output_nc=8)
    sd = O.random_state(cfg, seed=3)
    x = rand_input((1, 1, 16, 8, 24), 5)
    want = O.unet_forward(cfg, sd, x).numpy()
    got = unet_forward_c(cfg, sd, x.numpy())
    np.testing.assert_allclose(got, want, atol=5e-4, rtol=1e-3)


def test_c_restatement_rejects_bad_shape(state_6m):
    with pytest.raises(ValueError):
        unet_forward_c(CFG_6M, state_6m, np.zeros((1, 1, 16, 32, 32), np.float32))


def test_engine_rounding_mode_stays_close(state_6m):
    """The bf16-emulating variant must sit within the loose gate of the fp32 oracle."""
    x = rand_input((1, 1, 32, 32, 32), 0)
    a = O.unet_forward(CFG_6M, state_6m, x)
    b = O.unet_forward(CFG_6M, state_6m, x, engine_rounding=True)
    rel = ((a - b).norm() / a.norm()).item()
    assert rel < 3e-2, rel
