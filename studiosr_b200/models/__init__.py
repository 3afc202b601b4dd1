from .common import Model
from .edsr import EDSR
from .han import HAN
from .hat import HAT
from .rcan import RCAN
from .swinfir import SwinFIR
from .swinir import SwinIR

__all__ = ["Model", "SwinIR", "HAT", "EDSR", "RCAN", "HAN", "SwinFIR"]
