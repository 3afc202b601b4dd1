from .common import Model
from .edsr import EDSR
from .swinir import SwinIR

__all__ = ["Model", "SwinIR", "EDSR"]
