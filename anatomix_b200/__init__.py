"""anatomix_b200: a Blackwell (sm_100a) engine for the forward path of the
anatomix 3-D U-Net feature extractor, behind the reference's own interface.

    from anatomix_b200 import Unet, load_from_hf     # drop-in for anatomix.model.*
    from anatomix_b200 import Engine                  # the C-ABI engine, directly
"""
from .hf import ANATOMIX_VARIANTS, load_from_file, load_from_hf
from .unet import ConvBlock, Unet, get_actvn_layer, get_norm_layer

__all__ = ["Unet", "ConvBlock", "get_norm_layer", "get_actvn_layer", "load_from_hf",
           "load_from_file", "ANATOMIX_VARIANTS", "Engine", "patch_reference"]


def __getattr__(name):
    if name == "Engine":
        from .engine import Engine
        return Engine
    if name == "patch_reference":
        from .patch import patch_reference
        return patch_reference
    raise AttributeError(name)
