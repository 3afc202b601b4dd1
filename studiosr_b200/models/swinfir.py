"""Drop-in for studiosr.models.swinfir.SwinFIR (reference swinfir.py:83-128): SwinIR whose RSTB convs and conv_after_body are
SFB modules (spatial conv branch + Fourier branch + 1x1 fusion, swinfir.py:9-80).  Same constructor and state_dict; the SFBs run
in libssr_b200 as implicit-GEMM convs, 1x1 convs on the GEMM kernel and the 2-D real FFT pair of k_fft.cu.  Inference in the
fp32-class precisions (the reference trains and ships SwinFIR in fp32, swinfir.py:126); SURVEY.md 8 row f-3."""
from typing import Dict, List

import torch
import torch.nn as nn

from .. import _lib
from .swinir import SwinIR


class FourierUnit(nn.Module):  # swinfir.py:9-34 (parameter container)
    def __init__(self, embed_dim: int, fft_norm: str = "ortho") -> None:
        super().__init__()
        self.embed_dim = embed_dim
        self.conv_layer = nn.Conv2d(embed_dim * 2, embed_dim * 2, 1, 1, 0)
        self.fft_norm = fft_norm


class SpectralTransform(nn.Module):  # swinfir.py:37-50
    def __init__(self, embed_dim: int) -> None:
        super().__init__()
        self.conv_before_fft = nn.Sequential(nn.Conv2d(embed_dim, embed_dim // 2, 1, 1, 0), nn.LeakyReLU(0.2, inplace=True))
        self.fu = FourierUnit(embed_dim // 2)
        self.conv_after_fft = nn.Conv2d(embed_dim // 2, embed_dim, 1, 1, 0)


class SpatialB(nn.Module):  # swinfir.py:53-65
    def __init__(self, embed_dim: int, red: int = 1) -> None:
        super().__init__()
        if red != 1:
            raise NotImplementedError("SpatialB(red != 1) is not part of the native path (SwinFIR uses red = 1)")
        self.body = nn.Sequential(nn.Conv2d(embed_dim, embed_dim // red, 3, 1, 1), nn.LeakyReLU(0.2, inplace=True),
                                  nn.Conv2d(embed_dim // red, embed_dim, 3, 1, 1))


class SFB(nn.Module):  # swinfir.py:68-80
    def __init__(self, embed_dim: int, red: int = 1) -> None:
        super().__init__()
        self.S = SpatialB(embed_dim, red)
        self.F = SpectralTransform(embed_dim)
        self.fusion = nn.Conv2d(embed_dim * 2, embed_dim, 1, 1, 0)


class SwinFIR(SwinIR):
    ARCH = _lib.SSR_ARCH_SWINFIR
    TRAINABLE = False  # forward only

    def __init__(self, scale: int = 4, n_colors: int = 3, img_range: float = 1.0, embed_dim: int = 180,
                 depths: List[int] = [6, 6, 6, 6, 6, 6], num_heads: List[int] = [6, 6, 6, 6, 6, 6], window_size: int = 8,
                 mlp_ratio: float = 2.0, drop_rate: float = 0.0, attn_drop_rate: float = 0.0, drop_path_rate: float = 0.1,
                 upsampler: str = "pixelshuffle") -> None:
        super().__init__(scale=scale, n_colors=n_colors, img_range=img_range, embed_dim=embed_dim, depths=depths, num_heads=num_heads,
                         window_size=window_size, mlp_ratio=mlp_ratio, drop_rate=drop_rate, attn_drop_rate=attn_drop_rate,
                         drop_path_rate=drop_path_rate, upsampler=upsampler, resi_connection=SFB)
        self.conv_after_body = SFB(embed_dim)

    def _trainable(self) -> bool:
        return False

    def _resolve_precision(self, x: torch.Tensor) -> str:
        p = super()._resolve_precision(x)
        if p == "bf16":
            raise NotImplementedError("studiosr_b200: SwinFIR runs in the fp32-class precisions ('tf32x3' default, 'fp32', 'tf32'); "
                                      "the reference trains and ships it in fp32 (swinfir.py:126)")
        return p

    def get_training_config(self) -> Dict:
        return dict(batch_size=32, learning_rate=0.0002, beta1=0.9, beta2=0.99, weight_decay=0.0, max_iters=500000, gamma=0.5,
                    milestones=[250000, 400000, 450000, 475000], bfloat16=False)

    @classmethod
    def from_pretrained(cls, *args, **kwargs) -> "SwinFIR":
        raise NotImplementedError("the reference ships no SwinFIR weights (swinfir.py has no from_pretrained of its own)")
