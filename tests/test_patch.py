"""Patch mode (anatomix_b200/patch.py): wrapping the forward of a reference-style class in place.  The stand-in
class below has the reference's attributes and loop (network.py:467-548) and none of this package's hooks, like
the class `patch_reference` meets when the reference package is installed."""
import contextlib
import io
import os

import pytest
import torch
import torch.nn as nn

from conftest import CFG_6M, CFG_94M, ROOT, rand_input, rel_l2
from anatomix_b200 import Unet
from anatomix_b200.patch import _cfg_from_reference_module, patch_reference


def make_reflike():
    class RefLike(nn.Module):
        def __init__(self, **kw):
            super().__init__()
            with contextlib.redirect_stdout(io.StringIO()):
                u = Unet(**kw)
            self.model = u.model
            self.encoder_idx, self.decoder_idx = u.encoder_idx, u.decoder_idx
            self.res_source, self.res_dest = u.res_source, u.res_dest
            self.residual_connection, self.use_skip_connection = u.residual_connection, u.use_skip_connection

        def forward(self, input, layers=[], encode_only=False, verbose=False):
            feat, taps, skips = input, [], []
            for idx, layer in enumerate(self.model):
                feat = layer(feat)
                if idx in self.encoder_idx:
                    skips.append(feat)
                if idx in self.decoder_idx:
                    feat = torch.cat((skips.pop(), feat), dim=1)
                if idx in layers:
                    taps.append(feat)
                    if idx == layers[-1] and encode_only:
                        return taps
            return (feat, taps) if len(layers) else feat
    return RefLike


@pytest.mark.parametrize("cfg", [CFG_6M, CFG_94M, dict(CFG_6M, num_downs=2, ngf=32, norm="none", activation="lrelu", pooling="Avg")])
def test_constructor_arguments_are_recovered_from_a_built_module(cfg):
    m = make_reflike()(**cfg)
    got = _cfg_from_reference_module(m)
    for k, v in cfg.items():
        assert got[k] == pytest.approx(v) if isinstance(v, float) else got[k] == v, (k, got[k], v)
    assert got["pad_type"] == "reflect" and got["doubleconv"] and got["final_act"] == "none"


def test_patched_forward_on_cpu_is_the_stock_loop(state_6m):
    cls = patch_reference(make_reflike())
    assert patch_reference(cls) is cls and cls._anx_patched            # idempotent
    m = cls(**CFG_6M)
    m.load_state_dict(state_6m)
    m.eval()
    x = rand_input((1, 1, 32, 32, 32), 4)
    with torch.no_grad():
        y = m(x)
        y2, taps = m(x, layers=[8, 65])
        only = m(x, layers=[8], encode_only=True)
    assert torch.equal(y, y2) and torch.equal(taps[1], y) and len(only) == 1


@pytest.mark.gpu
def test_patched_forward_on_gpu_uses_the_engine(state_6m):
    from oracle import unet_oracle as O
    cls = patch_reference(make_reflike())
    m = cls(**CFG_6M)
    m.load_state_dict(state_6m)
    m = m.cuda().eval()
    x = rand_input((1, 1, 32, 32, 32), 4)
    with torch.no_grad():
        y = m(x.cuda())
        from anatomix_b200.engine import binding_for
        assert len(binding_for(m, m.__dict__["_anx_cfg"]).engines) == 1     # the engine took the call
        y2, taps = m(x.cuda(), layers=[8, 65])                              # stored tensors: engine as well
        y3, taps3 = m(x.cuda(), layers=[3])                                 # pre-norm conv output: stock loop
    want, wt = O.unet_forward(CFG_6M, state_6m, x, layers=[8, 65])
    assert rel_l2(y.cpu(), want) <= 3e-2 and rel_l2(y2.cpu(), want) <= 3e-2 and rel_l2(y3.cpu(), want) <= 3e-2
    assert rel_l2(taps[0].cpu(), wt[0]) <= 3e-2 and taps[1] is y2 and taps3[0].shape == (1, 16, 32, 32, 32)


REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this machine")
def test_patch_mode_on_the_real_reference_class():
    """`patch_reference()` against the UNMODIFIED reference class (imported from /root/reference in a
    subprocess so its `anatomix` package is the one on the path): constructor kwargs recovered exactly for both
    released variants, the patched CPU forward (stock loop) bit-equal to the unpatched one incl. feature taps,
    and a patched model still deep-copies and pickles."""
    code = f"""
import contextlib, copy, io, pickle, sys, torch
sys.path.insert(0, {REF!r}); sys.path.insert(0, {ROOT!r})
from anatomix.model.network import Unet
from anatomix.model.load_from_hf import ANATOMIX_VARIANTS
import anatomix_b200
from anatomix_b200.patch import _cfg_from_reference_module
stock = Unet.forward
for name in ("anatomix", "anatomix-dev"):
    kw = ANATOMIX_VARIANTS[name]["unet_kwargs"]
    with contextlib.redirect_stdout(io.StringIO()):
        torch.manual_seed(1); m = Unet(**kw)
    got = _cfg_from_reference_module(m)
    for k, v in kw.items():
        assert got[k] == v or abs(got[k] - v) < 1e-12, (name, k, got[k], v)
    assert got["pad_type"] == "reflect" and got["doubleconv"] and got["final_act"] == "none"
with contextlib.redirect_stdout(io.StringIO()):
    torch.manual_seed(2); m = Unet(**ANATOMIX_VARIANTS["anatomix"]["unet_kwargs"]).eval()
x = torch.rand(1, 1, 32, 32, 32, generator=torch.Generator().manual_seed(3))
with torch.no_grad():
    want = m(x); want_t = m(x, layers=[8, 20])
cls = anatomix_b200.patch_reference()
assert cls is Unet and Unet.forward is not stock and anatomix_b200.patch_reference() is cls
with torch.no_grad():
    got = m(x); got_t = m(x, layers=[8, 20])
assert torch.equal(got, want) and all(torch.equal(a, b) for a, b in zip(got_t[1], want_t[1]))
assert m.__dict__["_anx_cfg"]["ngf"] == 16
m2 = copy.deepcopy(m); m3 = pickle.loads(pickle.dumps(m))
with torch.no_grad():
    assert torch.equal(m2(x), want) and torch.equal(m3(x), want)
print("ok")
"""
    import subprocess, sys
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, cwd="/tmp")
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr[-3000:]
