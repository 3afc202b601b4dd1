"""Drop-in for studiosr.models.han.HAN (reference han.py:55-161): identical constructor and state_dict; forward = the RCAN
trunk in libssr_b200 (implicit-GEMM convs, fused channel-attention gates) whose eleven body outputs feed the layer-attention
module (LAM, han.py:12-33: 11 x 11 gram matrix over C*H*W, softmax, re-mix -- three small kernels) and the channel-spatial
attention module (CSAM, han.py:36-52: one 3x3x3 Conv3d over the (C, H, W) volume + sigmoid gate), then `last_conv`
(704 -> 64) and `last` (128 -> 64) on the same implicit-GEMM kernel.  Trainable under the Trainer's bf16 autocast like RCAN."""
import os
from typing import Dict

import torch
import torch.nn as nn

from .. import _lib
from .common import MeanShift, Model, Upsampler, conv2d
from .rcan import ResidualGroup


class LAM_Module(nn.Module):  # han.py:12-33 (parameter container)
    def __init__(self, in_dim: int) -> None:
        super().__init__()
        self.chanel_in = in_dim
        self.gamma = nn.Parameter(torch.zeros(1))


class CSAM_Module(nn.Module):  # han.py:36-52 (parameter container)
    def __init__(self, in_dim: int) -> None:
        super().__init__()
        self.chanel_in = in_dim
        self.conv = nn.Conv3d(1, 1, 3, 1, 1)
        self.gamma = nn.Parameter(torch.zeros(1))


class HAN(Model):
    ARCH = _lib.SSR_ARCH_HAN
    TRAINABLE = True  # forward + backward (train.cu: the RCAN executor + the layer / channel-spatial attention adjoints)

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
        self.csa = CSAM_Module(n_feats)
        self.la = LAM_Module(n_feats)
        self.last_conv = nn.Conv2d(n_feats * 11, n_feats, 3, 1, 1)  # han.py:87: eleven stacked maps, i.e. n_resgroups == 10
        self.last = nn.Conv2d(n_feats * 2, n_feats, 3, 1, 1)

    def _native_config(self, precision: int) -> "_lib.ModelConfig":
        c = _lib.ModelConfig()
        c.arch, c.precision, c.scale, c.n_colors, c.img_range = self.ARCH, precision, self.scale, self.n_colors, self.img_range
        c.n_feats, c.n_resblocks, c.n_resgroups, c.reduction = self.n_feats, self.n_resblocks, self.n_resgroups, self.reduction
        return c

    def _pad_mode(self) -> int:
        return _lib.PAD_EVAL  # HAN has no padding logic (han.py:90-113)

    def get_model_config(self) -> Dict:
        config = super().get_model_config()
        config.update(dict(n_feats=self.n_feats, n_resblocks=self.n_resblocks, n_resgroups=self.n_resgroups,
                           reduction=self.reduction))
        return config

    def get_training_config(self) -> Dict:
        return dict(batch_size=16, learning_rate=0.0001, beta1=0.9, beta2=0.99, weight_decay=0.0, max_iters=1000000,
                    gamma=0.5, milestones=[200000, 400000, 600000, 800000])

    @classmethod
    def from_pretrained(cls, scale: int = 4) -> "HAN":
        """File name and img_range=255 as han.py:144-161; the file must already be under ./pretrained (no download here)."""
        path = os.path.join("pretrained", f"HAN_BIX{scale}.pt")
        model = cls(scale=scale, img_range=255.0)
        model.load_state_dict(cls._load_pretrained_file(path), strict=False)
        return model
