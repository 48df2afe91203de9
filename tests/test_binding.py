"""Host-side behaviour of the engine bindings that needs no GPU: modules stay copyable / picklable once a
binding exists, in-place edits through `.data` are noticed, module-side eligibility (hooks, BatchNorm
submodule modes, parameter dtype)."""
import contextlib
import copy
import ctypes
import io
import pickle

import pytest
import torch
import torch.nn as nn

from conftest import CFG_6M
from anatomix_b200 import Unet
from anatomix_b200 import engine as E
from anatomix_b200.heads import FusedHeadSequential, UnetOutBlock, _HEAD_ENGINES, scaled_features


def quiet(**kw):
    with contextlib.redirect_stdout(io.StringIO()):
        return Unet(**kw)


class FakeEngine:
    """What a real `Engine` holds: a loaded CDLL and a raw handle -- neither can be pickled or deep-copied."""
    def __init__(self):
        self.lib = ctypes.CDLL(None)
        self._h = ctypes.c_void_p(1234)


def small():
    return quiet(dimension=3, input_nc=1, output_nc=16, num_downs=1, ngf=16)


def test_module_with_a_live_binding_still_copies_pickles_and_saves(tmp_path):
    m = small()
    b = m._engine_binding()
    b.engines[("cuda:0", 0)] = FakeEngine()
    with pytest.raises(Exception):
        pickle.dumps(b.engines[("cuda:0", 0)])                     # the thing that must stay off the module
    m2 = copy.deepcopy(m)
    m3 = pickle.loads(pickle.dumps(m))
    torch.save(m, tmp_path / "m.pt")
    m4 = torch.load(tmp_path / "m.pt", weights_only=False)
    for other in (m2, m3, m4):
        assert other._engine_binding() is not b and not other._engine_binding().engines
        assert all(torch.equal(a, c) for a, c in zip(m.state_dict().values(), other.state_dict().values()))
    assert not any(isinstance(v, (E.ModuleBinding, FakeEngine)) for v in vars(m).values())


def test_head_wrappers_keep_engines_off_the_module():
    u = small()
    seq = FusedHeadSequential(u, UnetOutBlock(3, 16, 3))
    _HEAD_ENGINES.setdefault(seq, {})["cuda:0"] = (FakeEngine(), E.PackStamp())
    assert set(copy.deepcopy(seq).state_dict()) == set(seq.state_dict())
    pickle.loads(pickle.dumps(seq))
    sc = scaled_features(u, 0.1)
    _HEAD_ENGINES.setdefault(sc, {})["cuda:0"] = (FakeEngine(), E.PackStamp())
    pickle.loads(pickle.dumps(sc))


def test_bindings_do_not_keep_modules_alive():
    import gc, weakref
    m = small()
    m._engine_binding()
    r = weakref.ref(m)
    del m
    gc.collect()
    assert r() is None


def test_pack_stamp_sees_version_bumps_storage_moves_and_data_edits(monkeypatch):
    monkeypatch.setattr(E, "VERIFY_EVERY", 4)
    m = small()
    ts = list(m.parameters()) + list(m.buffers())
    st = E.PackStamp()
    assert st.needs_repack(ts)                       # first use
    assert not st.needs_repack(ts)
    with torch.no_grad():
        m.model[0].weight.mul_(2.0)                  # optimizer-style in-place update: version bump
    assert st.needs_repack(ts) and not st.needs_repack(ts)
    # `.data` edits bump nothing (reference pretraining_networks.py:695-713 initialises this way) ...
    v = m.model[0].weight._version
    nn.init.normal_(m.model[0].weight.data, 0.0, 0.02)
    assert m.model[0].weight._version == v
    seen = [st.needs_repack(ts) for _ in range(4)]
    assert seen.count(True) == 1                     # ... the periodic content check catches them
    m.model[0].weight.data.mul_(3.0)
    st.invalidate()                                  # or immediately, when the caller says so
    assert st.needs_repack(ts)
    m.model[0].weight.data = m.model[0].weight.data.clone()     # new storage
    assert st.needs_repack(ts)


def test_invalidate_reaches_every_engine_of_a_module():
    m = small()
    b = m._engine_binding()
    b.stamps[("cuda:0", 0)] = E.PackStamp()
    b.stamps[("cuda:0", 0)].dirty = False
    m.invalidate_engine()
    assert b.stamps[("cuda:0", 0)].dirty


def test_module_side_eligibility():
    cpu = torch.device("cpu")
    m = quiet(**CFG_6M)
    assert "train mode" in E.module_ineligible_reason(m, m._anx_cfg, cpu)
    m.eval()
    assert E.module_ineligible_reason(m, m._anx_cfg, cpu) is None
    m.model[1].train()                               # one BatchNorm submodule back in train mode
    assert "train mode" in E.module_ineligible_reason(m, m._anx_cfg, cpu)
    m.eval()
    h = m.model[3].register_forward_hook(lambda mod, i, o: None)
    assert "hooks" in E.module_ineligible_reason(m, m._anx_cfg, cpu)
    h.remove()
    h = m.model[0].register_forward_pre_hook(lambda mod, i: None)
    assert "hooks" in E.module_ineligible_reason(m, m._anx_cfg, cpu)
    h.remove()
    assert E.module_ineligible_reason(m, m._anx_cfg, cpu) is None
    m.half()
    assert "fp32" in E.module_ineligible_reason(m, m._anx_cfg, cpu)
    m.float()
    assert "fp32" in E.module_ineligible_reason(m, m._anx_cfg, torch.device("meta"))    # parameters elsewhere
