"""Drop-in for studiosr.models.RCAN (reference rcan.py:11-122): identical constructor and state_dict;
forward = implicit-GEMM conv chain in libssr_b200 with the channel-attention gate (global average pool ->
64 -> 4 -> 64 MLP -> sigmoid, common.py:156-170) and both residual adds fused into two small kernels per RCAB."""
import os
from typing import Dict

import torch.nn as nn

from .. import _lib
from .common import ChannelAttention, MeanShift, Model, Upsampler, conv2d


class RCAB(nn.Module):  # rcan.py:11-24 (parameter container)
    def __init__(self, n_feat: int, kernel_size: int, reduction: int) -> None:
        super().__init__()
        self.body = nn.Sequential(conv2d(n_feat, n_feat, kernel_size), nn.ReLU(True), conv2d(n_feat, n_feat, kernel_size),
                                  ChannelAttention(n_feat, reduction))


class ResidualGroup(nn.Module):  # rcan.py:27-36 (parameter container)
    def __init__(self, n_feat: int, kernel_size: int, reduction: int, n_resblocks: int) -> None:
        super().__init__()
        self.body = nn.Sequential(*[RCAB(n_feat, kernel_size, reduction) for _ in range(n_resblocks)],
                                  conv2d(n_feat, n_feat, kernel_size))


class RCAN(Model):
    ARCH = _lib.SSR_ARCH_RCAN
    TRAINABLE = True  # forward + backward (train.cu: residual groups of RCABs incl. the channel-attention gate)

    def __init__(self, scale: int = 4, n_colors: int = 3, img_range: float = 1.0, n_feats: int = 64, n_resblocks: int = 20,
                 n_resgroups: int = 10, reduction: int = 16) -> None:
        super().__init__(scale, n_colors, img_range)
        self.n_feats = n_feats
        self.n_resblocks = n_resblocks
        self.n_resgroups = n_resgroups
        self.reduction = reduction
        self.sub_mean = MeanShift(img_range)
        self.add_mean = MeanShift(img_range, sign=1)
        self.head = nn.Sequential(conv2d(n_colors, n_feats, 3))
        self.body = nn.Sequential(*[ResidualGroup(n_feats, 3, reduction, n_resblocks) for _ in range(n_resgroups)],
                                  conv2d(n_feats, n_feats, 3))
        self.tail = nn.Sequential(Upsampler(scale, n_feats), conv2d(n_feats, n_colors, 3))

    def _native_config(self, precision: int) -> "_lib.ModelConfig":
        c = _lib.ModelConfig()
        c.arch, c.precision, c.scale, c.n_colors, c.img_range = self.ARCH, precision, self.scale, self.n_colors, self.img_range
        c.n_feats, c.n_resblocks, c.n_resgroups, c.reduction = self.n_feats, self.n_resblocks, self.n_resgroups, self.reduction
        return c

    def _pad_mode(self) -> int:
        return _lib.PAD_EVAL  # RCAN has no padding logic (rcan.py:68-77)

    def get_model_config(self) -> Dict:
        config = super().get_model_config()
        config.update(dict(n_feats=self.n_feats, n_resblocks=self.n_resblocks, n_resgroups=self.n_resgroups,
                           reduction=self.reduction))
        return config

    def get_training_config(self) -> Dict:
        return dict(batch_size=16, learning_rate=0.0001, beta1=0.9, beta2=0.99, weight_decay=0.0, max_iters=1000000,
                    gamma=0.5, milestones=[200000, 400000, 600000, 800000])

    @classmethod
    def from_pretrained(cls, scale: int = 4) -> "RCAN":
        """File layout and img_range=255 as rcan.py:107-119; the archive must already be extracted under ./pretrained."""
        path = os.path.join("pretrained", "models_ECCV2018RCAN", f"RCAN_BIX{scale}.pt")
        model = cls(scale=scale, img_range=255.0)
        model.load_state_dict(cls._load_pretrained_file(path), False)
        return model
