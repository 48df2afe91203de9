from .network import Unet

__all__ = ["Unet"]
