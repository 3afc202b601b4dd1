"""Drop-in for studiosr.models.SwinIR (reference swinir.py:258-445): same constructor, same
parameter / buffer names and shapes (so reference checkpoints load unchanged), same public
methods -- but the modules below are parameter containers only; the arithmetic of
SwinIR.forward runs in libssr_b200 (fused window attention, tcgen05 GEMMs with LN / GELU /
residual epilogues, implicit-GEMM convs with PixelShuffle folded into the store)."""
import os
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from .. import _lib
from .common import Mlp, Model, Upsampler


def _relative_position_index(ws: int) -> torch.Tensor:
    """int64 [ws*ws, ws*ws] buffer kept for state_dict compatibility (swinir.py:57-67)."""
    t = torch.arange(ws * ws)
    y, x = t // ws, t % ws
    return ((y[:, None] - y[None, :] + ws - 1) * (2 * ws - 1) + (x[:, None] - x[None, :] + ws - 1)).long()


class PatchEmbed(nn.Module):
    def __init__(self, embed_dim: int) -> None:
        super().__init__()
        self.norm = nn.LayerNorm(embed_dim)


class WindowAttention(nn.Module):
    def __init__(self, dim: int, window_size: int, num_heads: int) -> None:
        super().__init__()
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * window_size - 1) ** 2, num_heads))
        self.register_buffer("relative_position_index", _relative_position_index(window_size))
        self.qkv = nn.Linear(dim, dim * 3)
        self.proj = nn.Linear(dim, dim)
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)


class SwinTransformerBlock(nn.Module):
    def __init__(self, dim: int, num_heads: int, window_size: int, shift_size: int, mlp_ratio: float) -> None:
        super().__init__()
        assert 0 <= shift_size < window_size, "shift_size must in 0-window_size"
        self.shift_size = shift_size
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, window_size, num_heads)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))


class BasicLayer(nn.Module):
    def __init__(self, dim: int, depth: int, num_heads: int, window_size: int, mlp_ratio: float) -> None:
        super().__init__()
        self.blocks = nn.ModuleList(
            SwinTransformerBlock(dim, num_heads, window_size, 0 if i % 2 == 0 else window_size // 2, mlp_ratio)
            for i in range(depth))


class RSTB(nn.Module):
    def __init__(self, dim: int, depth: int, num_heads: int, window_size: int, mlp_ratio: float, resi_connection=None) -> None:
        super().__init__()
        self.residual_group = BasicLayer(dim, depth, num_heads, window_size, mlp_ratio)
        self.conv = resi_connection(dim) if resi_connection else nn.Conv2d(dim, dim, 3, 1, 1)  # swinir.py:241


class SwinIR(Model):
    ARCH = _lib.SSR_ARCH_SWINIR
    TRAINABLE = True

    def __init__(
        self,
        scale: int = 4,
        n_colors: int = 3,
        img_range: float = 1.0,
        embed_dim: int = 180,
        depths: List[int] = [6, 6, 6, 6, 6, 6],
        num_heads: List[int] = [6, 6, 6, 6, 6, 6],
        window_size: int = 8,
        mlp_ratio: float = 2.0,
        drop_rate: float = 0.0,
        attn_drop_rate: float = 0.0,
        drop_path_rate: float = 0.1,
        upsampler: str = "pixelshuffle",
        resi_connection: Optional[nn.Module] = None,
    ) -> None:
        super().__init__(scale, n_colors, img_range)
        if resi_connection is not None and (getattr(resi_connection, "__name__", "") != "SFB" or self.ARCH != _lib.SSR_ARCH_SWINFIR):
            raise NotImplementedError("resi_connection: only SwinFIR (models/swinfir.py: SFB in every RSTB and as conv_after_body, "
                                      "swinfir.py:83-114) has a native path")
        if drop_rate != 0.0 or attn_drop_rate != 0.0:
            raise NotImplementedError("dropout > 0 is not part of the native path (reference default is 0.0)")
        assert upsampler in ("pixelshuffle", "pixelshuffledirect")
        assert len(depths) == len(num_heads) <= _lib.SSR_MAX_LAYERS
        self.embed_dim = embed_dim
        self.depths = list(depths)
        self.num_heads = list(num_heads)
        self.window_size = window_size
        self.mlp_ratio = mlp_ratio
        self.drop_rate = drop_rate
        self.attn_drop_rate = attn_drop_rate
        self.drop_path_rate = drop_path_rate  # stochastic depth: training only (swinir.py:296,137,171-172), see _draw_drop_path
        self.upsampler = upsampler

        self.conv_first = nn.Conv2d(n_colors, embed_dim, 3, 1, 1)
        self.patch_embed = PatchEmbed(embed_dim)
        self.layers = nn.ModuleList(
            RSTB(embed_dim, depths[i], num_heads[i], window_size, mlp_ratio, resi_connection) for i in range(len(depths)))
        self.norm = nn.LayerNorm(embed_dim)
        self.conv_after_body = nn.Conv2d(embed_dim, embed_dim, 3, 1, 1)
        if upsampler == "pixelshuffle":
            num_feat = 64
            self.conv_before_upsample = nn.Sequential(nn.Conv2d(embed_dim, num_feat, 3, 1, 1), nn.LeakyReLU(inplace=True))
            self.upsample = Upsampler(scale, num_feat)
            self.conv_last = nn.Conv2d(num_feat, n_colors, 3, 1, 1)
        else:
            self.upsample = Upsampler(scale, embed_dim, n_colors)
        self.apply(self._init_weights)

    @staticmethod
    def _init_weights(m: nn.Module) -> None:
        """Same init rule as swinir.py:333-340."""
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def _trainable(self) -> bool:
        return self.upsampler == "pixelshuffle" and self.window_size == 8  # what train.cu's bind_swinir accepts

    def _draw_drop_path(self, batch: int, device):
        """Per-sample stochastic-depth factors of one training step, [2 * n_blocks, B]: row 2k / 2k+1 scales the attention /
        MLP branch of block k.  Mirrors the reference: rates follow `torch.linspace(0, drop_path_rate, sum(depths))`
        (swinir.py:296), a block with rate 0 has nn.Identity (swinir.py:137) and draws nothing, every other block calls timm's
        DropPath twice per forward (swinir.py:171-172), each call drawing `x.new_empty((B,1,1,1)).bernoulli_(keep) / keep`
        from torch's global generator of the input's device -- the draws below consume that generator identically."""
        if self.drop_path_rate <= 0.0:
            return None
        n = sum(self.depths)
        rates = [v.item() for v in torch.linspace(0, self.drop_path_rate, n)]
        rows = []
        for p in rates:
            for _ in range(2):
                if p > 0.0:
                    keep = 1.0 - p
                    rows.append(torch.empty((batch, 1, 1, 1), dtype=torch.float32, device=device).bernoulli_(keep).div_(keep).view(batch))
                else:
                    rows.append(torch.ones(batch, dtype=torch.float32, device=device))
        self._last_drop_scale = torch.stack(rows).contiguous()
        return self._last_drop_scale

    def _native_config(self, precision: int) -> "_lib.ModelConfig":
        c = _lib.ModelConfig()
        c.arch, c.precision, c.scale, c.n_colors, c.img_range = self.ARCH, precision, self.scale, self.n_colors, self.img_range
        c.embed_dim, c.n_layers, c.window_size, c.mlp_ratio = self.embed_dim, len(self.depths), self.window_size, self.mlp_ratio
        for i, (d, h) in enumerate(zip(self.depths, self.num_heads)):
            c.depths[i], c.num_heads[i] = d, h
        c.upsampler = 0 if self.upsampler == "pixelshuffle" else 1
        return c

    def get_model_config(self) -> Dict:
        config = super().get_model_config()
        config.update(dict(
            embed_dim=self.embed_dim, depths=self.depths, num_heads=self.num_heads, window_size=self.window_size,
            mlp_ratio=self.mlp_ratio, drop_rate=self.drop_rate, attn_drop_rate=self.attn_drop_rate,
            drop_path_rate=self.drop_path_rate, upsampler=self.upsampler))
        return config

    def get_training_config(self) -> Dict:
        return dict(batch_size=32, learning_rate=0.0002, beta1=0.9, beta2=0.99, weight_decay=0.0, max_iters=500000,
                    gamma=0.5, milestones=[250000, 400000, 450000, 475000])

    @classmethod
    def from_pretrained(cls, scale: int = 4, light: bool = False, dataset: str = "DF2K", pretrained: bool = True) -> "SwinIR":
        """Same file naming as swinir.py:404-445; weights must already be under ./pretrained (no network here)."""
        assert scale in [2, 3, 4, 8]
        assert dataset in ["DIV2K", "DF2K"]
        config = {"scale": scale}
        img_size = 64 if dataset == "DF2K" else 48
        task, label = "001_classicalSR", "M"
        if light:
            config.update(depths=[6, 6, 6, 6], embed_dim=60, num_heads=[6, 6, 6, 6], upsampler="pixelshuffledirect")
            task, dataset, img_size, label = "002_lightweightSR", "DIV2K", 64, "S"
        model = cls(**config)
        if pretrained:
            path = os.path.join("pretrained", f"{task}_{dataset}_s{img_size}w8_SwinIR-{label}_x{scale}.pth")
            if not os.path.exists(path):
                raise FileNotFoundError(f"{path} not found (downloads are outside the native path; place the file there)")
            ckpt = torch.load(path, map_location="cpu")
            model.load_state_dict(ckpt["params"] if "params" in ckpt else ckpt, strict=False)
        return model
