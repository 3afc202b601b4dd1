"""Drop-in for studiosr.models.EDSR (reference edsr.py:12-111): identical constructor and
state_dict; forward = implicit-GEMM conv chain in libssr_b200 (MeanShift folded into the first /
last kernels, ReLU / res_scale / residual adds in the GEMM epilogues, PixelShuffle in the store)."""
import os
from typing import Dict

import torch.nn as nn

from .. import _lib
from .common import MeanShift, Model, ResBlock, Upsampler, conv2d


class EDSR(Model):
    ARCH = _lib.SSR_ARCH_EDSR
    TRAINABLE = True

    def __init__(self, scale: int = 4, n_colors: int = 3, img_range: float = 1.0, n_feats: int = 256,
                 n_resblocks: int = 32, res_scale: float = 0.1) -> None:
        super().__init__(scale, n_colors, img_range)
        self.n_feats = n_feats
        self.n_resblocks = n_resblocks
        self.res_scale = res_scale
        self.sub_mean = MeanShift(img_range)
        self.add_mean = MeanShift(img_range, sign=1)
        self.head = nn.Sequential(conv2d(n_colors, n_feats, 3))
        self.body = nn.Sequential(*[ResBlock(n_feats, 3, res_scale) for _ in range(n_resblocks)], conv2d(n_feats, n_feats, 3))
        self.tail = nn.Sequential(Upsampler(scale, n_feats), conv2d(n_feats, n_colors, 3))

    def _native_config(self, precision: int) -> "_lib.ModelConfig":
        c = _lib.ModelConfig()
        c.arch, c.precision, c.scale, c.n_colors, c.img_range = self.ARCH, precision, self.scale, self.n_colors, self.img_range
        c.n_feats, c.n_resblocks, c.res_scale = self.n_feats, self.n_resblocks, self.res_scale
        return c

    def _pad_mode(self) -> int:
        return _lib.PAD_EVAL  # EDSR has no padding logic (edsr.py:39-48)

    def get_model_config(self) -> Dict:
        config = super().get_model_config()
        config.update(dict(n_feats=self.n_feats, n_resblocks=self.n_resblocks, res_scale=self.res_scale))
        return config

    def get_training_config(self) -> Dict:
        return dict(batch_size=16, learning_rate=0.0001, beta1=0.9, beta2=0.99, weight_decay=0.0, max_iters=1000000,
                    gamma=0.5, milestones=[200000, 400000, 600000, 800000])

    @classmethod
    def from_pretrained(cls, scale: int = 4, dataset: str = "DIV2K") -> "EDSR":
        """Signature, file names and the img_range=255 of the DIV2K weights as edsr.py:77-111; the file must already be
        under ./pretrained."""
        assert scale in [2, 3, 4]
        assert dataset in ["DIV2K", "DF2K"]
        if dataset == "DIV2K":
            model, file_name = cls(scale=scale, img_range=255.0), f"r32f256x{scale}.pth"
        else:
            model, file_name = cls(scale=scale), f"EDSRx{scale}.pth"
        model.load_state_dict(cls._load_pretrained_file(os.path.join("pretrained", file_name)), strict=False)
        return model
