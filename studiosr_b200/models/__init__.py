from .common import Model
from .edsr import EDSR
from .rcan import RCAN
from .swinir import SwinIR

__all__ = ["Model", "SwinIR", "EDSR", "RCAN"]
