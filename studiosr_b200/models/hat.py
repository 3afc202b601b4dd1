"""Drop-in for studiosr.models.HAT (reference hat.py:25-593): same constructor, same parameter / buffer
names and shapes (so reference checkpoints load unchanged), same public methods.  The modules are
parameter containers; HAT.forward runs in libssr_b200: window attention over 16x16 windows, the
overlapping cross-attention block (queries 16x16, keys / values 24x24 with zero padding), the channel
attention block (conv 3x3 -> GELU -> conv 3x3 -> squeeze / excite gate) and the MLPs as fused GEMM epilogues."""
import os
from typing import Dict, List

import torch
import torch.nn as nn

from .. import _lib
from .common import Mlp, Model, Upsampler
from .swinir import PatchEmbed, _relative_position_index


class ChannelAttention(nn.Module):  # hat.py:25-38
    def __init__(self, num_feat: int, squeeze_factor: int = 16) -> None:
        super().__init__()
        self.attention = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(num_feat, num_feat // squeeze_factor, 1, padding=0),
                                       nn.ReLU(inplace=True), nn.Conv2d(num_feat // squeeze_factor, num_feat, 1, padding=0),
                                       nn.Sigmoid())


class CAB(nn.Module):  # hat.py:41-52
    def __init__(self, num_feat: int, compress_ratio: int = 3, squeeze_factor: int = 30) -> None:
        super().__init__()
        self.cab = nn.Sequential(nn.Conv2d(num_feat, num_feat // compress_ratio, 3, 1, 1), nn.GELU(),
                                 nn.Conv2d(num_feat // compress_ratio, num_feat, 3, 1, 1), ChannelAttention(num_feat, squeeze_factor))


class WindowAttention(nn.Module):  # hat.py:55-110 (the index lives on the HAT module, not here)
    def __init__(self, dim: int, window_size: int, num_heads: int) -> None:
        super().__init__()
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * window_size - 1) ** 2, num_heads))
        self.qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)


class HAB(nn.Module):  # hat.py:113-195
    def __init__(self, dim: int, num_heads: int, window_size: int, shift_size: int, mlp_ratio: float, compress_ratio: int,
                 squeeze_factor: int, conv_scale: float) -> None:
        super().__init__()
        assert 0 <= shift_size < window_size, "shift_size must in 0-window_size"
        self.shift_size = shift_size
        self.conv_scale = conv_scale
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, window_size, num_heads)
        self.conv_block = CAB(dim, compress_ratio, squeeze_factor)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))


class OCAB(nn.Module):  # hat.py:198-293
    def __init__(self, dim: int, num_heads: int, window_size: int, mlp_ratio: float, overlap_ratio: float) -> None:
        super().__init__()
        self.overlap_win_size = int(window_size * overlap_ratio) + window_size
        self.norm1 = nn.LayerNorm(dim)
        self.qkv = nn.Linear(dim, dim * 3)
        self.relative_position_bias_table = nn.Parameter(torch.zeros((window_size + self.overlap_win_size - 1) ** 2, num_heads))
        self.proj = nn.Linear(dim, dim)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)


class AttenBlocks(nn.Module):  # hat.py:296-344
    def __init__(self, dim: int, depth: int, num_heads: int, window_size: int, mlp_ratio: float, compress_ratio: int,
                 squeeze_factor: int, conv_scale: float, overlap_ratio: float) -> None:
        super().__init__()
        self.blocks = nn.ModuleList(
            HAB(dim, num_heads, window_size, 0 if i % 2 == 0 else window_size // 2, mlp_ratio, compress_ratio, squeeze_factor,
                conv_scale) for i in range(depth))
        self.overlap_attn = OCAB(dim, num_heads, window_size, mlp_ratio, overlap_ratio)


class RHAG(nn.Module):  # hat.py:347-385
    def __init__(self, dim: int, depth: int, num_heads: int, window_size: int, mlp_ratio: float, compress_ratio: int,
                 squeeze_factor: int, conv_scale: float, overlap_ratio: float) -> None:
        super().__init__()
        self.residual_group = AttenBlocks(dim, depth, num_heads, window_size, mlp_ratio, compress_ratio, squeeze_factor,
                                          conv_scale, overlap_ratio)
        self.conv = nn.Conv2d(dim, dim, 3, 1, 1)


def _rpi_oca(ws: int, overlap_ratio: float) -> torch.Tensor:
    """int64 [ws*ws, wse*wse] buffer kept for state_dict compatibility (hat.py:490-513)."""
    wse = ws + int(overlap_ratio * ws)
    o, e = torch.arange(ws * ws), torch.arange(wse * wse)
    dy = (e // wse)[None, :] - (o // ws)[:, None] + ws - wse + 1  # the reference's shift: indices run over [-880, 640]
    dx = (e % wse)[None, :] - (o % ws)[:, None] + ws - wse + 1
    return (dy * (ws + wse - 1) + dx).long()


class HAT(Model):
    ARCH = _lib.SSR_ARCH_HAT

    def __init__(
        self,
        scale: int = 4,
        n_colors: int = 3,
        img_range: float = 1.0,
        embed_dim: int = 180,
        depths: List[int] = [6, 6, 6, 6, 6, 6],
        num_heads: List[int] = [6, 6, 6, 6, 6, 6],
        window_size: int = 16,
        mlp_ratio: float = 2.0,
        drop_rate: float = 0.0,
        attn_drop_rate: float = 0.0,
        drop_path_rate: float = 0.1,
        compress_ratio: int = 3,
        squeeze_factor: int = 30,
        conv_scale: float = 0.01,
        overlap_ratio: float = 0.5,
    ) -> None:
        super().__init__(scale, n_colors, img_range)
        if drop_rate != 0.0 or attn_drop_rate != 0.0:
            raise NotImplementedError("dropout > 0 is not part of the native path (reference default is 0.0)")
        assert len(depths) == len(num_heads) <= _lib.SSR_MAX_LAYERS
        self.embed_dim = embed_dim
        self.depths = list(depths)
        self.num_heads = list(num_heads)
        self.window_size = window_size
        self.mlp_ratio = mlp_ratio
        self.drop_rate = drop_rate
        self.attn_drop_rate = attn_drop_rate
        self.drop_path_rate = drop_path_rate  # stochastic depth is treated as identity by the native forward
        self.compress_ratio = compress_ratio
        self.squeeze_factor = squeeze_factor
        self.conv_scale = conv_scale
        self.overlap_ratio = overlap_ratio
        self.shift_size = window_size // 2
        self.register_buffer("relative_position_index_SA", _relative_position_index(window_size))
        self.register_buffer("relative_position_index_OCA", _rpi_oca(window_size, overlap_ratio))

        self.conv_first = nn.Conv2d(n_colors, embed_dim, 3, 1, 1)
        self.patch_embed = PatchEmbed(embed_dim)
        self.layers = nn.ModuleList(
            RHAG(embed_dim, depths[i], num_heads[i], window_size, mlp_ratio, compress_ratio, squeeze_factor, conv_scale,
                 overlap_ratio) for i in range(len(depths)))
        self.norm = nn.LayerNorm(embed_dim)
        self.conv_after_body = nn.Conv2d(embed_dim, embed_dim, 3, 1, 1)
        num_feat = 64
        self.conv_before_upsample = nn.Sequential(nn.Conv2d(embed_dim, num_feat, 3, 1, 1), nn.LeakyReLU(inplace=True))
        self.upsample = Upsampler(scale, num_feat)
        self.conv_last = nn.Conv2d(num_feat, n_colors, 3, 1, 1)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m: nn.Module) -> None:
        """Same init rule as hat.py:466-473."""
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def _native_config(self, precision: int) -> "_lib.ModelConfig":
        c = _lib.ModelConfig()
        c.arch, c.precision, c.scale, c.n_colors, c.img_range = self.ARCH, precision, self.scale, self.n_colors, self.img_range
        c.embed_dim, c.n_layers, c.window_size, c.mlp_ratio = self.embed_dim, len(self.depths), self.window_size, self.mlp_ratio
        for i, (d, h) in enumerate(zip(self.depths, self.num_heads)):
            c.depths[i], c.num_heads[i] = d, h
        c.upsampler = 0
        c.compress_ratio, c.squeeze_factor, c.conv_scale, c.overlap_ratio = (self.compress_ratio, self.squeeze_factor,
                                                                             self.conv_scale, self.overlap_ratio)
        return c

    def _pad_mode(self) -> int:
        return _lib.PAD_TRAIN  # hat.py:544: check_image_size (reflect pad to a multiple of the window) in BOTH modes

    def get_model_config(self) -> Dict:
        config = super().get_model_config()
        config.update(dict(
            embed_dim=self.embed_dim, depths=self.depths, num_heads=self.num_heads, window_size=self.window_size,
            mlp_ratio=self.mlp_ratio, drop_rate=self.drop_rate, attn_drop_rate=self.attn_drop_rate,
            drop_path_rate=self.drop_path_rate, compress_ratio=self.compress_ratio, squeeze_factor=self.squeeze_factor,
            conv_scale=self.conv_scale, overlap_ratio=self.overlap_ratio))
        return config

    @classmethod
    def from_pretrained(cls, scale: int = 4) -> "HAT":
        """Same file naming as hat.py:576-593; weights must already be under ./pretrained (no network here)."""
        model = cls(scale=scale)
        path = os.path.join("pretrained", f"HAT_SRx{scale}.pth")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not found (downloads are outside the native path; place the file there)")
        model.load_state_dict(torch.load(path, map_location="cpu")["params_ema"])
        return model
