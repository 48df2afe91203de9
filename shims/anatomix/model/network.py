"""``anatomix.model.network`` import path -> `anatomix_b200.unet`."""
from anatomix_b200.unet import ConvBlock, Unet, get_actvn_layer, get_norm_layer

__all__ = ["Unet", "ConvBlock", "get_norm_layer", "get_actvn_layer"]
