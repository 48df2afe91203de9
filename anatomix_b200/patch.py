"""Patch mode: give an already-imported *reference* ``anatomix`` package the
B200 engine without editing it.

``patch_reference()`` wraps ``anatomix.model.network.Unet.forward`` of the
reference (network.py:467) so eligible calls go to the engine and everything
else runs the reference's own loop.  The reference's ``load_from_hf``,
registration and segmentation code then use the engine with zero changes.
"""
from __future__ import annotations

import functools
import inspect
import os


def _cfg_from_reference_module(m):
    """Recover the constructor kwargs from a built reference Unet."""
    import torch.nn as nn
    convs = [l for l in m.model if isinstance(l, (nn.Conv1d, nn.Conv2d, nn.Conv3d))]
    first = convs[0]
    ndims = first.weight.dim() - 2
    norms = [l for l in m.model if "Norm" in l.__class__.__name__]
    pools = [l for l in m.model if "Pool" in l.__class__.__name__]
    ups = [l for l in m.model if isinstance(l, nn.Upsample)]
    acts = [l for l in m.model if isinstance(l, (nn.ReLU, nn.LeakyReLU, nn.ELU, nn.PReLU, nn.SELU, nn.Tanh))]
    if not norms:
        norm, eps = "none", 1e-5
    elif "Batch" in norms[0].__class__.__name__:
        norm, eps = "batch", norms[0].eps
    else:
        norm, eps = ("instance_affine" if norms[0].affine else "instance"), norms[0].eps
    act = "none"
    if acts:
        act = {"ReLU": "relu", "LeakyReLU": "lrelu", "ELU": "elu", "PReLU": "prelu", "SELU": "selu",
               "Tanh": "tanh"}[acts[0].__class__.__name__]
    last = list(m.model)[-1]
    final_act = "none" if last is convs[-1] else "other"
    n_enc = len(m.encoder_idx)
    per_level = (m.encoder_idx[1] - m.encoder_idx[0]) if n_enc > 1 else None
    slots_per_conv = 1 + (norm != "none") + (act != "none")
    doubleconv = True if per_level is None else per_level == 2 * slots_per_conv + 1
    return dict(dimension=ndims, input_nc=first.in_channels, output_nc=convs[-1].out_channels,
                num_downs=len(pools), ngf=first.out_channels, norm=norm, final_act=final_act,
                activation=act, pad_type=first.padding_mode, doubleconv=doubleconv,
                residual_connection=m.residual_connection,
                pooling="Max" if pools and "Max" in pools[0].__class__.__name__ else "Avg",
                interp=ups[0].mode if ups else "nearest",
                use_skip_connection=m.use_skip_connection, norm_eps=eps)


def patch_reference(unet_cls=None):
    """Wraps ``Unet.forward`` of the reference package in place; returns the class.
    Idempotent."""
    if unet_cls is None:
        from anatomix.model.network import Unet as unet_cls   # the reference's (or the shim's) class
    if getattr(unet_cls, "_anx_patched", False) or hasattr(unet_cls, "engine_ineligible_reason"):
        return unet_cls
    from .engine import binding_for, ineligible_reason
    stock_forward = unet_cls.forward

    @functools.wraps(stock_forward)
    def forward(self, input, layers=[], encode_only=False, verbose=False):
        if os.environ.get("ANATOMIX_B200_DISABLE") != "1":
            cfg = self.__dict__.get("_anx_cfg")
            if cfg is None:
                # a plain dict of constructor kwargs: the only thing patch mode leaves on the instance, so
                # deepcopy / pickle / torch.save of a patched model behave as before (the engine binding is
                # kept in anatomix_b200.engine's weakly keyed registry)
                cfg = _cfg_from_reference_module(self)
                self.__dict__["_anx_cfg"] = cfg
            tapping = len(layers) > 0
            if not (tapping and verbose) and ineligible_reason(self, cfg, input, layers) is None:
                binding = binding_for(self, cfg)
                if tapping:       # feature taps the engine stores (network.py:475-529)
                    return binding.forward_taps(input, layers, encode_only)
                return binding.forward(input)
        return stock_forward(self, input, layers, encode_only, verbose)

    unet_cls.forward = forward
    unet_cls._anx_patched = True
    return unet_cls
