"""Consumers right behind the U-Net forward.

* Segmentation: the reference finetunes ``nn.Sequential(Unet, UnetOutBlock(3, C, n_classes + 1, False))``
  (``anatomix/segmentation/segmentation_utils.py:114-115``) and runs it as the sliding-window predictor
  (``train_segmentation.py:194-199``).  `UnetOutBlock` is MONAI's block (a bias-carrying 1x1x1 conv; MONAI is
  not importable offline, so it is restated here with the same state-dict keys) and `FusedHeadSequential`
  is that two-module Sequential whose eligible forwards run on the engine with the head evaluated inside
  the last conv's epilogue (``anx_engine_set_head``): the 16-channel feature volume never reaches HBM.
* Registration: ``pred * downscale_feat_scalar`` then ``F.avg_pool3d(pred, grid_sp, stride=grid_sp)``
  (``anatomix/registration/run_convex_adam_with_network_feats.py:166-167, 198-205``):
  `scaled_features` folds the scalar into the same epilogue (a diagonal head), `avg_pool3d_scaled` is the
  streaming pooling kernel behind ``anx_avgpool3d_scale_f32``.

Everything that is not eligible for the engine (training, CPU tensors, autograd) runs the stock torch
modules, exactly like the reference.
"""
from __future__ import annotations

import ctypes as C
import weakref
from collections import OrderedDict
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib

HEAD_MAX = 32          # csrc/layout.cuh

# engines with a fused head, weakly keyed by the wrapper module (never stored on it: deepcopy / pickle /
# torch.save of the wrapper keep working after the engine has run): {module: {device: (Engine, PackStamp)}}
_HEAD_ENGINES: "weakref.WeakKeyDictionary[nn.Module, Dict[torch.device, tuple]]" = weakref.WeakKeyDictionary()


def _head_engine(owner: nn.Module, unet: nn.Module, device, tensors, head_weight_fn):
    """Engine of `unet` with the head returned by ``head_weight_fn() -> (weight, bias)`` fused in; re-packed
    when any of `tensors` changed (same staleness rules as `anatomix_b200.engine.ModuleBinding`)."""
    from .engine import Engine, PackStamp
    per_dev = _HEAD_ENGINES.setdefault(owner, {})
    eng, stamp = per_dev.get(device, (None, None))
    if eng is None:
        eng, stamp = Engine(unet._anx_cfg, device), PackStamp()
        per_dev[device] = (eng, stamp)
    if stamp.needs_repack(tensors):
        eng.load_state(unet.state_dict())
        w, b = head_weight_fn()
        eng.set_head(w, b)
    return eng


class UnetOutBlock(nn.Module):
    """``monai.networks.blocks.dynunet_block.UnetOutBlock``: one 1x1(x1) convolution with bias, no norm,
    no activation; parameters live at ``conv.conv.{weight,bias}`` as in MONAI."""

    def __init__(self, spatial_dims: int, in_channels: int, out_channels: int, dropout=None):
        super().__init__()
        if dropout not in (None, False, 0, 0.0):
            raise NotImplementedError("dropout in the output block is not restated (the reference passes False)")
        conv = getattr(nn, f"Conv{spatial_dims}d")(in_channels, out_channels, kernel_size=1, stride=1, bias=True)
        self.conv = nn.Sequential(OrderedDict(conv=conv))

    def forward(self, inp):
        return self.conv(inp)


def pointwise_conv_of(head: nn.Module) -> Optional[nn.Conv3d]:
    """The single 1x1x1 ``nn.Conv3d`` a head consists of, or None when it is anything else."""
    leaves = [m for m in head.modules() if not list(m.children()) and not isinstance(m, nn.Identity)]
    if len(leaves) != 1 or not isinstance(leaves[0], nn.Conv3d):
        return None
    c = leaves[0]
    if tuple(c.kernel_size) != (1, 1, 1) or tuple(c.stride) != (1, 1, 1) or c.groups != 1 \
            or c.padding not in ((0, 0, 0), "same", "valid"):
        return None
    return c


class FusedHeadSequential(nn.Sequential):
    """``nn.Sequential(unet, head)`` (same children, same state-dict keys ``0.*`` / ``1.*``) whose eligible
    forwards run on the engine with the head fused into the last conv."""

    def __init__(self, unet: nn.Module, head: nn.Module):
        super().__init__(unet, head)

    def fused_ineligible_reason(self, x) -> Optional[str]:
        unet, head = self[0], self[1]
        if not hasattr(unet, "engine_ineligible_reason"):
            return "the first module is not an anatomix_b200 Unet"
        why = unet.engine_ineligible_reason(x)
        if why is not None:
            return why
        conv = pointwise_conv_of(head)
        if conv is None:
            return "the head is not a single 1x1x1 convolution"
        if conv.out_channels > HEAD_MAX or unet._anx_cfg["output_nc"] > 16 or conv.in_channels != unet._anx_cfg["output_nc"]:
            return "head / feature widths outside the fused epilogue's range"
        if torch.is_grad_enabled() and any(p.requires_grad for p in head.parameters()):
            return "autograd is recording"
        return None

    def _engine(self, device):
        unet, conv = self[0], pointwise_conv_of(self[1])
        return _head_engine(self, unet, device, list(self.parameters()) + list(self.buffers()),
                            lambda: (conv.weight, conv.bias))

    def forward(self, x):
        import os
        if os.environ.get("ANATOMIX_B200_DISABLE") != "1" and self.fused_ineligible_reason(x) is None:
            return self._engine(x.device).forward(x)
        return super().forward(x)


def fuse_output_head(unet: nn.Module, head: nn.Module) -> FusedHeadSequential:
    """Drop-in for ``torch.nn.Sequential(model, fin_layer)`` of segmentation_utils.py:115."""
    return FusedHeadSequential(unet, head)


class _ScaledUnet(nn.Module):
    def __init__(self, unet: nn.Module, scale: float):
        super().__init__()
        self.unet, self.scale = unet, float(scale)

    def forward(self, x):
        import os
        unet = self.unet
        if os.environ.get("ANATOMIX_B200_DISABLE") == "1" or unet.engine_ineligible_reason(x) is not None \
                or unet._anx_cfg["output_nc"] > 16:
            return unet(x) * self.scale
        eng = _head_engine(self, unet, x.device, list(unet.parameters()) + list(unet.buffers()),
                           lambda: (torch.eye(unet._anx_cfg["output_nc"]) * self.scale, None))
        return eng.forward(x)


def scaled_features(unet: nn.Module, scale: float) -> nn.Module:
    """A predictor returning ``unet(x) * scale`` with the multiply folded into the last conv's epilogue
    (run_convex_adam_with_network_feats.py:166-167: ``downscale_feat_scalar``)."""
    return _ScaledUnet(unet, scale)


def normalize_features(x: torch.Tensor, mode: str = "unit", eps: Optional[float] = None, inplace: bool = False) -> torch.Tensor:
    """Voxelwise normalisation across channels of ``[N, C, D, H, W]`` features, as the reference prescribes for the dev
    models before registration / visualisation (README.md:13,49): ``mode="unit"`` is
    ``F.normalize(x, dim=1)`` (unit L2 norm), ``mode="zscore"`` is ``(x - x.mean(1, True)) / (x.std(1, keepdim=True) + eps)``.
    fp32 CUDA tensors go through the engine library's streaming kernel (anx_channel_normalize_f32), anything else
    through torch."""
    if mode not in ("unit", "zscore"):
        raise ValueError("mode is 'unit' or 'zscore'")
    if eps is None:
        eps = 1e-12 if mode == "unit" else 0.0
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 5 or (torch.is_grad_enabled() and x.requires_grad):
        if mode == "unit":
            return F.normalize(x, dim=1, eps=eps)
        return (x - x.mean(1, keepdim=True)) / (x.std(1, keepdim=True) + eps)
    lib = _lib.load()
    x = x.contiguous()
    n, c, d, h, w = x.shape
    out = x if inplace else torch.empty_like(x)
    with torch.cuda.device(x.device):
        st = lib.anx_channel_normalize_f32(x.data_ptr(), out.data_ptr(), n, c, d, h, w, 0 if mode == "unit" else 1,
                                           C.c_float(eps), torch.cuda.current_stream(x.device).cuda_stream)
    if st != _lib.ANX_OK:
        raise _lib.EngineError(st, "anx_channel_normalize_f32")
    return out


def features_behind(unet: nn.Module, x: torch.Tensor, front: torch.Tensor) -> torch.Tensor:
    """``torch.cat([front, unet(x)], dim=1)`` without copying the network features: the engine writes them straight
    behind the ``front`` channels of one buffer (reference instance_optimization.py:16-119:
    ``torch.concatenate([mind_fixed, pred_fixed], dim=1)``).  Falls back to the plain concat off the engine path."""
    import os
    if os.environ.get("ANATOMIX_B200_DISABLE") == "1" or not hasattr(unet, "engine_ineligible_reason") \
            or unet.engine_ineligible_reason(x) is not None or front.dtype != torch.float32 or front.device != x.device:
        return torch.cat([front, unet(x)], dim=1)
    eng = unet._engine_binding().engine_for(x.device)
    n, cf = front.shape[0], front.shape[1]
    dest = torch.empty((n, cf + eng.output_nc) + tuple(x.shape[2:]), dtype=torch.float32, device=x.device)
    dest[:, :cf].copy_(front)
    return eng.forward_into(x, dest, cf)


def avg_pool3d_scaled(x: torch.Tensor, k: int, scale: float = 1.0) -> torch.Tensor:
    """``scale * F.avg_pool3d(x, k, stride=k)`` for fp32 ``[N, C, D, H, W]``; CUDA tensors go through the
    engine library's streaming kernel, anything else through torch."""
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 5 or (torch.is_grad_enabled() and x.requires_grad) \
            or min(x.shape[2:]) < k:
        return F.avg_pool3d(x, k, stride=k) * scale
    lib = _lib.load()
    x = x.contiguous()
    n, c, d, h, w = x.shape
    out = torch.empty((n, c, d // k, h // k, w // k), dtype=torch.float32, device=x.device)
    with torch.cuda.device(x.device):
        st = lib.anx_avgpool3d_scale_f32(x.data_ptr(), out.data_ptr(), n * c, d, h, w, k, C.c_float(scale),
                                         torch.cuda.current_stream(x.device).cuda_stream)
    if st != _lib.ANX_OK:
        raise _lib.EngineError(st, "anx_avgpool3d_scale_f32")
    return out
