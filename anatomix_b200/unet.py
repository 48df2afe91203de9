"""Drop-in mirror of ``anatomix.model.network`` backed by the B200 engine.

`Unet` here is a real ``nn.Module`` with the reference's constructor signature
(reference network.py:262-279), the same flat ``self.model`` Sequential (same
indices, same state-dict keys, same RNG consumption order at init), the same
bookkeeping attributes (``encoder_idx``, ``decoder_idx``, ``res_source``,
``res_dest``, ``use_bias`` ...) and the same ``forward(input, layers=[],
encode_only=False, verbose=False)`` (reference network.py:467-548).

What differs is *where the arithmetic runs*: when a call is eligible (SURVEY.md
section 8(b): 3-D, CUDA input, no autograd, no taps, released topology, BN in eval
mode) the whole forward is handed to the sm_100a engine through the C ABI in
``include/anatomix_b200.h``.  Everything else walks the Sequential with stock
torch modules, exactly like the reference; that keeps training / finetuning /
CPU use correct.  The engine itself has no CPU fallback: on a CUDA device with
the shared library missing, an eligible call raises.
"""
from __future__ import annotations

import os
from functools import partial

import torch
import torch.nn as nn

from .topology import make_plan

_NORMS = ("batch", "instance", "instance_affine", "none")


def get_norm_layer(ndims, norm="batch", eps=1e-5):
    """Factory ``Norm(num_features)`` for the requested normalisation, or None.

    Mirrors reference network.py:127-168 (BatchNormNd / InstanceNormNd /
    affine InstanceNormNd with ``eps`` bound, ``'none'`` -> None).
    """
    if norm not in _NORMS:
        raise ValueError(f"Currently unsupported normalization: {norm}")
    if norm == "none":
        return None
    if norm == "batch":
        return partial(getattr(nn, f"BatchNorm{ndims}d"), eps=eps)
    cls = getattr(nn, f"InstanceNorm{ndims}d")
    if norm == "instance_affine":
        return partial(cls, affine=True, eps=eps)
    return partial(cls, eps=eps)


_ACTS = {
    "relu": lambda: nn.ReLU(inplace=True),
    "elu": lambda: nn.ELU(),
    "prelu": lambda: nn.PReLU(),
    "selu": lambda: nn.SELU(inplace=True),
    "tanh": lambda: nn.Tanh(),
    "none": lambda: None,
}


def get_actvn_layer(activation="relu"):
    """One activation module (or None).  ``lrelu`` uses slope 0.3 as in
    reference network.py:171-204 (note: 0.3, not the 0.2 of ``ConvBlock``)."""
    if activation == "lrelu":
        return nn.LeakyReLU(0.3, inplace=True)
    assert activation in _ACTS, "Unsupported activation: {}".format(activation)
    return _ACTS[activation]()


class ConvBlock(nn.Module):
    """conv -> [norm] -> [activation]; kept for import-path compatibility
    (reference network.py:13-124; nothing in the reference instantiates it)."""

    def __init__(self, ndims, input_dim, output_dim, kernel_size, stride, bias,
                 padding=0, norm="none", activation="relu", pad_type="zeros"):
        super().__init__()
        assert ndims in [1, 2, 3], "ndims in 1--3. found: %d" % ndims
        self.use_bias = bias
        self.conv = getattr(nn, f"Conv{ndims}d")(
            input_dim, output_dim, kernel_size, stride, bias=bias,
            padding=padding, padding_mode=pad_type)
        if norm == "batch":
            self.norm = getattr(nn, f"BatchNorm{ndims}d")(output_dim)
        elif norm == "instance":
            self.norm = getattr(nn, f"InstanceNorm{ndims}d")(
                output_dim, track_running_stats=False)
        elif norm == "none":
            self.norm = None
        else:
            assert 0, "Unsupported normalization: {}".format(norm)
        if activation == "lrelu":
            self.activation = nn.LeakyReLU(0.2, inplace=True)   # 0.2 here, 0.3 in Unet
        else:
            assert activation in _ACTS, "Unsupported activation: {}".format(activation)
            self.activation = _ACTS[activation]()

    def forward(self, x):
        x = self.conv(x)
        if self.norm:
            x = self.norm(x)
        if self.activation:
            x = self.activation(x)
        return x


class Unet(nn.Module):
    """3-D (also 1-D / 2-D) U-Net feature extractor; see module docstring.

    Constructor arguments are those of reference network.py:262-279.
    """

    def __init__(self, dimension, input_nc, output_nc, num_downs, ngf=24,
                 norm="batch", final_act="none", activation="relu",
                 pad_type="reflect", doubleconv=True, residual_connection=False,
                 pooling="Max", interp="nearest", use_skip_connection=True,
                 norm_eps=1e-5):
        super().__init__()
        ndims = dimension
        assert ndims in [1, 2, 3], "ndims should be 1--3. found: %d" % ndims
        self.use_bias = norm == "instance"          # network.py:292
        self.residual_connection = residual_connection
        self.use_skip_connection = use_skip_connection

        Conv = getattr(nn, f"Conv{ndims}d")
        Pool = getattr(nn, f"{pooling}Pool{ndims}d")
        Norm = get_norm_layer(ndims, norm, eps=norm_eps)
        act = get_actvn_layer(activation)            # ONE shared instance (network.py:301)
        final = get_actvn_layer(final_act)

        plan = make_plan(input_nc, output_nc, num_downs, ngf,
                         has_norm=Norm is not None, has_act=act is not None,
                         has_final_act=final is not None, doubleconv=doubleconv,
                         use_skip_connection=use_skip_connection)
        mods = []
        for s in plan.slots:      # construction order == index order == RNG order
            if s.kind == "conv":
                mods.append(Conv(s.cin, s.cout, kernel_size=3, stride=1,
                                 bias=self.use_bias, padding="same",
                                 padding_mode=pad_type))
            elif s.kind == "norm":
                mods.append(Norm(s.cout))
            elif s.kind == "act":
                mods.append(act)
            elif s.kind == "pool":
                mods.append(Pool(2))
            elif s.kind == "up":
                mods.append(nn.Upsample(scale_factor=2, mode=interp))
            else:
                mods.append(final)
        self.encoder_idx = plan.encoder_idx
        self.decoder_idx = plan.decoder_idx
        self.res_source = plan.res_source
        self.res_dest = plan.res_dest
        print("Encoder skip connect id", self.encoder_idx)      # network.py:447-448
        print("Decoder skip connect id", self.decoder_idx)
        self.model = nn.Sequential(*mods)

        # engine-side description (plain python; not part of the state dict)
        self._anx_cfg = dict(
            dimension=ndims, input_nc=input_nc, output_nc=output_nc,
            num_downs=num_downs, ngf=ngf, norm=norm, final_act=final_act,
            activation=activation, pad_type=pad_type, doubleconv=doubleconv,
            residual_connection=residual_connection, pooling=pooling,
            interp=interp, use_skip_connection=use_skip_connection,
            norm_eps=norm_eps)

    # ------------------------------------------------------------------ engine
    def _engine_binding(self):
        """The engine binding of this module.  It lives in a weakly keyed registry next to the engine code,
        not on the module: ``copy.deepcopy``, ``pickle`` and ``torch.save(model)`` keep working after the
        engine has run (ctypes handles cannot be copied or pickled)."""
        from .engine import binding_for
        return binding_for(self, self._anx_cfg)

    def invalidate_engine(self):
        """Re-pack the engine's weights on the next forward.  Needed only after edits through ``.data``
        (which bypass the version counters the binding watches); such edits are otherwise picked up by the
        periodic content check (`anatomix_b200.engine.VERIFY_EVERY`)."""
        from .engine import invalidate
        invalidate(self)

    def engine_ineligible_reason(self, x, layers=()):
        """None when this call goes to the B200 engine, else why it does not
        (SURVEY.md section 8(b) eligibility list)."""
        from .engine import ineligible_reason
        return ineligible_reason(self, self._anx_cfg, x, layers)

    # ----------------------------------------------------------------- forward
    def forward(self, input, layers=[], encode_only=False, verbose=False):
        """Reference network.py:467-548.  Without ``layers`` returns the output
        tensor; with ``layers`` (module indices) returns ``(output, taps)``, or
        only ``taps`` when ``encode_only`` stops at ``layers[-1]``."""
        tapping = len(layers) > 0
        if os.environ.get("ANATOMIX_B200_DISABLE") != "1" and not (tapping and verbose) and \
                self.engine_ineligible_reason(input, layers) is None:
            if tapping:
                return self._engine_binding().forward_taps(input, layers, encode_only)
            return self._engine_binding().forward(input)

        feat, taps, skips, held = input, [], [], None
        for idx, layer in enumerate(self.model):
            feat = layer(feat)
            if tapping and verbose:
                print(idx, layer.__class__.__name__, feat.size())
            if self.residual_connection:
                if idx in self.res_source:
                    held = feat
                    if tapping and verbose:
                        print("Record skip connection input from %d" % idx)
                if idx in self.res_dest:
                    assert held.size() == feat.size()
                    feat = feat + 0.1 * held
                    if tapping and verbose:
                        print("Add skip connection input for %d" % idx)
            if self.use_skip_connection:
                if idx in self.encoder_idx:
                    skips.append(feat)
                if idx in self.decoder_idx:
                    feat = torch.cat((skips.pop(), feat), dim=1)   # encoder first
            if not tapping:
                continue
            if idx in layers:
                if verbose:
                    print("%d: adding the output of %s %d"
                          % (idx, layer.__class__.__name__, feat.size(1)), feat.size())
                taps.append(feat)
            elif verbose:
                print("%d: skipping %s" % (idx, layer.__class__.__name__), feat.size())
            if idx == layers[-1] and encode_only:
                if verbose:
                    print("encoder only return features")
                return taps
        return (feat, taps) if tapping else feat
