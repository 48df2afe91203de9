"""The drop-in `Unet` / `load_from_hf` surface on CPU (boundary contract,
SURVEY.md section 8(b)); the engine is not involved here."""
import io
import os
import sys
import contextlib
import subprocess

import numpy as np
import pytest
import torch
import torch.nn as nn

from conftest import CFG_6M, CFG_94M, ROOT, golden, rand_input
from anatomix_b200 import Unet, ANATOMIX_VARIANTS
from anatomix_b200.hf import _load_handling_compile
from anatomix_b200.topology import make_plan

REF = "/root/reference"


def quiet(**kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return Unet(**kw)


def test_ctor_prints_and_attributes(capsys):
    m = Unet(**CFG_6M)
    out = capsys.readouterr().out
    assert "Encoder skip connect id [8, 15, 22, 29]" in out
    assert "Decoder skip connect id [37, 44, 51, 58]" in out
    assert len(m.model) == 66 and isinstance(m.model, nn.Sequential)
    assert m.encoder_idx == [8, 15, 22, 29] and m.decoder_idx == [37, 44, 51, 58]
    assert m.res_source[:3] == [0, 3, 6] and m.res_dest[:3] == [2, 5, 8]
    assert m.use_bias is False and m.use_skip_connection and not m.residual_connection
    assert m.training            # like the reference, nobody calls .eval() for you
    assert m.model[2] is m.model[5]   # one shared activation instance (network.py:301)


def test_state_dict_contract(state_6m):
    m = quiet(**CFG_6M)
    sd = m.state_dict()
    assert list(sd.keys()) == list(state_6m.keys())
    assert all(sd[k].shape == state_6m[k].shape for k in sd)
    m.load_state_dict(state_6m, strict=True)
    prefixed = {"_orig_mod." + k: v for k, v in state_6m.items()}
    _load_handling_compile(quiet(**CFG_6M), prefixed)     # load_from_hf.py:43-47


def test_dev_variant_layout():
    m = quiet(**CFG_94M)
    assert len(m.model) == 80 and m.use_bias
    assert m.encoder_idx == [8, 15, 22, 29, 36] and m.decoder_idx == [44, 51, 58, 65, 72]
    assert len(m.state_dict()) == 48
    assert isinstance(m.model[1], nn.InstanceNorm3d) and m.model[1].eps == 1e-2
    assert isinstance(m.model[9], nn.AvgPool3d) and m.model[44].mode == "trilinear"


def test_cpu_forward_matches_reference_golden(state_6m):
    m = quiet(**CFG_6M); m.load_state_dict(state_6m); m.eval()
    g = golden("g1_6m_32.npz")
    x = rand_input((1, 1, 32, 32, 32), 0)
    with torch.no_grad():
        y = m(x)
        y2, taps = m(x, layers=[int(i) for i in g["tap_ids"]])
        only = m(x, layers=[0, 8], encode_only=True)
    np.testing.assert_allclose(y.numpy(), g["out"], atol=2e-4, rtol=1e-4)
    assert torch.equal(y, y2) and len(taps) == len(g["tap_ids"])
    np.testing.assert_allclose(taps[1][:, :, ::2, ::2, ::2].numpy(), g["tap8"], atol=2e-4)
    assert len(only) == 2 and only[1].shape == (1, 16, 32, 32, 32)
    assert taps[4].shape[1] == 384      # tap on a decoder_idx returns the concatenated tensor


def test_train_mode_uses_batch_statistics(state_6m):
    m = quiet(**CFG_6M); m.load_state_dict(state_6m)       # stays in train mode
    g = golden("g5_6m_train.npz")
    with torch.no_grad():
        y = m(rand_input((1, 1, 32, 32, 32), 0))
    np.testing.assert_allclose(y[:, :, ::2, ::2, ::2].numpy(), g["out_s2"], atol=2e-3, rtol=1e-3)


def test_reference_failures_are_preserved(state_6m):
    m = quiet(**CFG_6M).eval()
    with torch.no_grad():
        with pytest.raises(RuntimeError):
            m(torch.zeros(1, 1, 16, 16, 16))       # reflect pad on a size-1 bottleneck
        with pytest.raises(RuntimeError):
            m(torch.zeros(1, 1, 72, 32, 32))       # pool/upsample mismatch at the concat


def test_other_dimensions_and_flags_still_work():
    m2 = quiet(dimension=2, input_nc=3, output_nc=5, num_downs=2, ngf=8, norm="none", activation="lrelu",
               residual_connection=True)
    with torch.no_grad():
        assert m2(torch.rand(2, 3, 16, 16)).shape == (2, 5, 16, 16)
    m1 = quiet(dimension=1, input_nc=1, output_nc=2, num_downs=1, ngf=4, doubleconv=False,
               use_skip_connection=False, final_act="tanh")
    with torch.no_grad():
        assert m1(torch.rand(1, 1, 8)).shape == (1, 2, 8)
    assert isinstance(m1.model[-1], nn.Tanh)


def test_eligibility_reasons_on_cpu(state_6m):
    m = quiet(**CFG_6M).eval()
    assert m.engine_ineligible_reason(torch.zeros(1, 1, 32, 32, 32)) == "input is not a CUDA tensor"


def test_topology_plan_matches_survey_appendix_a():
    p = make_plan(1, 16, 4, 16)
    convs = [(s.index, s.cin, s.cout, s.level) for s in p.convs]
    assert convs[0] == (0, 1, 16, 0) and convs[-1] == (65, 16, 16, 0)
    assert (38, 384, 128, 3) in convs and (59, 48, 16, 0) in convs and (34, 256, 256, 4) in convs
    p = make_plan(1, 32, 5, 32)
    assert [s.index for s in p.convs][-1] == 79 and (45, 1536, 512, 4) in [(s.index, s.cin, s.cout, s.level) for s in p.convs]


def test_variant_registry():
    assert ANATOMIX_VARIANTS["anatomix"]["unet_kwargs"] == CFG_6M
    assert ANATOMIX_VARIANTS["anatomix-dev"]["unet_kwargs"] == CFG_94M
    from anatomix_b200 import load_from_hf
    with pytest.raises(ValueError):
        load_from_hf("nope")


def test_import_path_shim():
    code = ("from anatomix.model.network import Unet, get_norm_layer, get_actvn_layer, ConvBlock;"
            "from anatomix.model.load_from_hf import load_from_hf, ANATOMIX_VARIANTS, _load_handling_compile, DEFAULT_REPO;"
            "import anatomix_b200.unet as u; assert Unet is u.Unet; print('ok')")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "shims"), ROOT]))
    r = subprocess.run([sys.executable, "-c", code], cwd="/tmp", capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this machine")
def test_repo_on_path_does_not_shadow_the_reference_package():
    """reference installed + this repository on PYTHONPATH.  `anatomix` must stay the
    reference's package (all of its subpackages importable as far as their own dependencies allow) and
    `patch_reference()` must patch the reference's class."""
    code = f"""
import importlib.util, os, sys
import anatomix, anatomix.model.network as net
assert os.path.realpath(anatomix.__file__).startswith({REF!r}), anatomix.__file__
for sub in ("anatomix.registration", "anatomix.segmentation", "anatomix.model.vit3d", "anatomix.model.load_from_hf"):
    assert importlib.util.find_spec(sub) is not None, sub        # resolvable: nothing shadows the reference tree
import anatomix_b200
cls = anatomix_b200.patch_reference()
assert cls is net.Unet and cls._anx_patched and cls.__module__ == "anatomix.model.network"
print("ok")
"""
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([ROOT, REF]))
    r = subprocess.run([sys.executable, "-c", code], cwd="/tmp", capture_output=True, text=True, env=env)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-2000:]


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this machine")
def test_same_modules_and_init_as_reference():
    """Module-by-module repr and seeded default init equal the reference's
    (runs in a subprocess so the reference's `anatomix` package can be imported)."""
    code = f"""
import sys, io, contextlib, torch
sys.path.insert(0, {REF!r})
from anatomix.model.network import Unet as R
sys.path.insert(0, {ROOT!r})
from anatomix_b200.unet import Unet as M
for kw in ({CFG_6M!r}, {CFG_94M!r}, dict(dimension=2, input_nc=2, output_nc=3, num_downs=2, ngf=8, norm='none', doubleconv=False)):
    with contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(7); r = R(**kw)
        torch.manual_seed(7); m = M(**kw)
    assert str(r.model) == str(m.model)
    assert (r.encoder_idx, r.decoder_idx, r.res_source, r.res_dest) == (m.encoder_idx, m.decoder_idx, m.res_source, m.res_dest)
    a, b = r.state_dict(), m.state_dict()
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
print('same')
"""
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0 and "same" in r.stdout, r.stderr[-2000:]
