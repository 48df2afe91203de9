"""Variant registry and checkpoint loading, mirroring
``anatomix.model.load_from_hf`` (reference load_from_hf.py:8-79) so that
``load_from_hf("anatomix")`` keeps working and hands back the engine-backed
`Unet`.  Also offers `load_from_file` for offline checkpoints (the reference's
callers do the same by hand, convex_adam_utils.py:61-75).
"""
from __future__ import annotations

import torch

from .unet import Unet

DEFAULT_REPO = "neeldey/anatomix"

_UNET_6M = dict(dimension=3, input_nc=1, output_nc=16, num_downs=4, ngf=16)
_UNET_94M = dict(dimension=3, input_nc=1, output_nc=32, num_downs=5, ngf=32,
                 norm="instance", pooling="Avg", interp="trilinear",
                 norm_eps=1e-2)

# name -> constructor kwargs + feature width (reference load_from_hf.py:11-36).
# The ViT variant belongs to a different model family that this engine does not
# cover; it is listed so the registry keys match and resolves lazily to the
# reference's own PrimusV2 when that package is importable.
ANATOMIX_VARIANTS = {
    "anatomix": {"unet_kwargs": dict(_UNET_6M), "output_channels": 16},
    "anatomix-dev": {"unet_kwargs": dict(_UNET_94M), "output_channels": 32},
    "anatomix-dev-vit": {
        "vit_kwargs": dict(
            input_channels=1, num_classes=32, embed_dim=396, eva_depth=12,
            eva_numheads=6, patch_embed_size=(8, 8, 8),
            input_shape=(128, 128, 128), num_register_tokens=8,
            init_values=0.1, scale_attn_inner=True, qk_norm=True,
            out_norm="demean", out_norm_eps=1e-2, register_init_std=0.02,
            in_eps=1e-2),
        "output_channels": 32,
    },
}


def _load_handling_compile(model, state_dict):
    """strict load; keys saved from a ``torch.compile`` wrapper carry an
    ``_orig_mod.`` prefix which is dropped first (load_from_hf.py:39-49)."""
    keys = list(state_dict)
    if keys and keys[0].startswith("_orig_mod."):
        state_dict = {k.removeprefix("_orig_mod."): v for k, v in state_dict.items()}
    model.load_state_dict(state_dict, strict=True)
    return model


def _build(variant):
    if variant not in ANATOMIX_VARIANTS:
        raise ValueError(f"Unknown variant {variant!r}. Known: {sorted(ANATOMIX_VARIANTS)}")
    cfg = ANATOMIX_VARIANTS[variant]
    if "vit_kwargs" in cfg:
        try:
            from anatomix.model.vit3d import PrimusV2     # reference package, if present
        except Exception as e:                            # pragma: no cover
            raise NotImplementedError(
                "the ViT variant is outside the U-Net engine's scope and needs the "
                "reference package with dynamic-network-architectures") from e
        return PrimusV2(**cfg["vit_kwargs"])
    return Unet(**cfg["unet_kwargs"])


def load_from_file(variant, weights_path, map_location="cpu"):
    """Same as `load_from_hf` with the checkpoint already on disk."""
    model = _build(variant)
    sd = torch.load(weights_path, map_location=map_location)
    return _load_handling_compile(model, sd)


def load_from_hf(variant, repo_id=DEFAULT_REPO, revision=None, map_location="cpu"):
    """Fetch ``<variant>.pth`` from the Hub and return the loaded model (in
    train mode, like the reference: load_from_hf.py:52-79)."""
    if variant not in ANATOMIX_VARIANTS:
        raise ValueError(f"Unknown variant {variant!r}. Known: {sorted(ANATOMIX_VARIANTS)}")
    from huggingface_hub import hf_hub_download
    path = hf_hub_download(repo_id, f"{variant}.pth", revision=revision)
    return load_from_file(variant, path, map_location=map_location)
