"""Host logic of the consumers behind the U-Net (anatomix_b200/heads.py) on CPU: module / state-dict
contract of the restated MONAI output block, head detection, stock-torch behaviour off the GPU."""
import contextlib
import io

import torch
import torch.nn as nn
import torch.nn.functional as F

from conftest import CFG_6M, rand_input
from anatomix_b200 import Unet
from anatomix_b200.heads import (FusedHeadSequential, UnetOutBlock, avg_pool3d_scaled, fuse_output_head,
                                 pointwise_conv_of, scaled_features)


def quiet():
    with contextlib.redirect_stdout(io.StringIO()):
        return Unet(**CFG_6M)


def test_out_block_mirrors_monai_layout():
    blk = UnetOutBlock(3, 16, 5, False)                # the reference's call, segmentation_utils.py:114
    assert list(blk.state_dict()) == ["conv.conv.weight", "conv.conv.bias"]
    assert blk.conv.conv.weight.shape == (5, 16, 1, 1, 1)
    x = rand_input((1, 16, 4, 4, 4), 0)
    assert torch.allclose(blk(x), F.conv3d(x, blk.conv.conv.weight, blk.conv.conv.bias))


def test_head_detection():
    assert pointwise_conv_of(UnetOutBlock(3, 16, 5)) is not None
    assert pointwise_conv_of(nn.Conv3d(16, 4, 1)) is not None
    assert pointwise_conv_of(nn.Conv3d(16, 4, 3, padding=1)) is None
    assert pointwise_conv_of(nn.Sequential(nn.Conv3d(16, 4, 1), nn.ReLU())) is None


def test_fused_sequential_is_a_plain_sequential_on_cpu(state_6m):
    m = quiet(); m.load_state_dict(state_6m); m.eval()
    head = UnetOutBlock(3, 16, 3)
    seq = fuse_output_head(m, head)
    assert isinstance(seq, FusedHeadSequential) and isinstance(seq, nn.Sequential)
    keys = list(seq.state_dict())
    assert keys[0] == "0.model.0.weight" and keys[-1] == "1.conv.conv.bias"     # as nn.Sequential(model, fin_layer)
    x = rand_input((1, 1, 32, 32, 32), 1)
    with torch.no_grad():
        assert seq.fused_ineligible_reason(x) == "input is not a CUDA tensor"
        assert torch.equal(seq(x), head(m(x)))
        assert torch.allclose(scaled_features(m, 0.1)(x), m(x) * 0.1)


def test_avg_pool_falls_back_to_torch_on_cpu():
    x = rand_input((1, 2, 9, 8, 7), 2)
    assert torch.allclose(avg_pool3d_scaled(x, 2, 0.5), F.avg_pool3d(x, 2, stride=2) * 0.5)


def test_tap_requests_on_cpu_walk_the_stock_modules(state_6m):
    m = quiet(); m.load_state_dict(state_6m); m.eval()
    x = rand_input((1, 1, 32, 32, 32), 3)
    with torch.no_grad():
        y, taps = m(x, layers=[1, 2, 37])
        only = m(x, layers=[8, 22, 15], encode_only=True)
    assert torch.equal(taps[0], taps[1]) and taps[2].shape == (1, 384, 4, 4, 4) and len(only) == 2
