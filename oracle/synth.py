"""TEST INFRASTRUCTURE ONLY -- deterministic synthetic weights and inputs.

Golden fixtures cannot carry 47 MB of SwinIR weights, so every parity case is
defined by (config, seed): this module regenerates bit-identical fp32 weights
from a torch CPU generator on any box with the same torch build (the GPU box
runs the same image).  oracle/make_golden.py loads these dicts into the real
reference modules with ``load_state_dict(strict=True)``, which pins the key /
shape list below against the reference's own ``state_dict``.

Values are deliberately *not* the reference's init (trunc-normal 0.02, zero
bias, LN 1/0): larger linear weights, non-trivial LN affine, non-zero biases
and a visible relative-position-bias table make every term of the path matter
in the parity error.
"""
from collections import OrderedDict
from typing import Dict, List

import torch


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _randn(g, *shape, std=1.0):
    return torch.randn(*shape, generator=g, dtype=torch.float32) * std


def _conv(g, sd, name, cout, cin, k=3, gain=1.0):
    fan_in = cin * k * k
    sd[name + ".weight"] = _randn(g, cout, cin, k, k, std=gain / fan_in**0.5)
    sd[name + ".bias"] = _randn(g, cout, std=0.05)


def _linear(g, sd, name, cout, cin, gain=1.0):
    sd[name + ".weight"] = _randn(g, cout, cin, std=gain / cin**0.5)
    sd[name + ".bias"] = _randn(g, cout, std=0.05)


def _ln(g, sd, name, c):
    sd[name + ".weight"] = 1.0 + _randn(g, c, std=0.1)
    sd[name + ".bias"] = _randn(g, c, std=0.1)


def relative_position_index(ws: int) -> torch.Tensor:
    """Buffer contents of swinir.py:57-67 restated: index = (dy+ws-1)*(2ws-1) + (dx+ws-1)."""
    idx = torch.arange(ws * ws)
    yy, xx = idx // ws, idx % ws
    dy = yy[:, None] - yy[None, :] + ws - 1
    dx = xx[:, None] - xx[None, :] + ws - 1
    return (dy * (2 * ws - 1) + dx).to(torch.int64)


SWINIR_DEFAULT = dict(
    scale=4, n_colors=3, img_range=1.0, embed_dim=180, depths=[6] * 6, num_heads=[6] * 6,
    window_size=8, mlp_ratio=2.0, upsampler="pixelshuffle",
)
SWINIR_TINY = dict(
    scale=4, n_colors=3, img_range=1.0, embed_dim=60, depths=[2, 2], num_heads=[6, 6],
    window_size=8, mlp_ratio=2.0, upsampler="pixelshuffle",
)


def swinir_config(**over) -> Dict:
    cfg = dict(SWINIR_DEFAULT)
    cfg.update(over)
    return cfg


def upsampler_convs(scale: int, n_feats: int, num_out_ch=None) -> List:
    """(index-in-Sequential, cout, r) of each conv in common.py:124-137."""
    out = []
    if num_out_ch is not None:
        out.append((0, scale * scale * num_out_ch, scale))
    elif scale & (scale - 1) == 0:
        i = 0
        s = scale
        while s > 1:
            out.append((i, 4 * n_feats, 2))
            i += 2
            s //= 2
    else:
        out.append((0, scale * scale * n_feats, scale))
    return out


def swinir_weights(cfg: Dict, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """state_dict of reference SwinIR(**cfg) (swinir.py:258-340), synthetic values."""
    g = _gen(seed)
    sd = OrderedDict()
    C = cfg["embed_dim"]
    ws = cfg["window_size"]
    hid = int(C * cfg["mlp_ratio"])
    nc = cfg["n_colors"]
    _conv(g, sd, "conv_first", C, nc)
    _ln(g, sd, "patch_embed.norm", C)
    for li, depth in enumerate(cfg["depths"]):
        nh = cfg["num_heads"][li]
        for bi in range(depth):
            p = f"layers.{li}.residual_group.blocks.{bi}"
            _ln(g, sd, p + ".norm1", C)
            sd[p + ".attn.relative_position_bias_table"] = _randn(g, (2 * ws - 1) ** 2, nh, std=0.5)
            sd[p + ".attn.relative_position_index"] = relative_position_index(ws)
            _linear(g, sd, p + ".attn.qkv", 3 * C, C, gain=1.0)
            _linear(g, sd, p + ".attn.proj", C, C, gain=0.5)
            _ln(g, sd, p + ".norm2", C)
            _linear(g, sd, p + ".mlp.fc1", hid, C, gain=1.0)
            _linear(g, sd, p + ".mlp.fc2", C, hid, gain=0.5)
        _conv(g, sd, f"layers.{li}.conv", C, C, gain=0.5)
    _ln(g, sd, "norm", C)
    _conv(g, sd, "conv_after_body", C, C, gain=0.5)
    if cfg["upsampler"] == "pixelshuffle":
        _conv(g, sd, "conv_before_upsample.0", 64, C)
        for i, cout, _ in upsampler_convs(cfg["scale"], 64):
            _conv(g, sd, f"upsample.{i}", cout, 64)
        _conv(g, sd, "conv_last", nc, 64)
    else:  # pixelshuffledirect
        for i, cout, _ in upsampler_convs(cfg["scale"], C, nc):
            _conv(g, sd, f"upsample.{i}", cout, C)
    return sd


def _conv1(g, sd, name, cout, cin, gain=1.0):
    _conv(g, sd, name, cout, cin, k=1, gain=gain)


def swinfir_weights(cfg: Dict, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """state_dict of reference SwinFIR(**cfg) (swinfir.py:83-114): SwinIR whose RSTB convs and conv_after_body are SFB modules
    (S = SpatialB, F = SpectralTransform with the FourierUnit, fusion), synthetic values."""
    g = _gen(seed + 500)
    sd = swinir_weights(cfg, seed)
    C = cfg["embed_dim"]
    out = OrderedDict()

    def sfb(pre):
        _conv(g, out, pre + ".S.body.0", C, C, gain=0.7)
        _conv(g, out, pre + ".S.body.2", C, C, gain=0.5)
        _conv1(g, out, pre + ".F.conv_before_fft.0", C // 2, C)
        _conv1(g, out, pre + ".F.fu.conv_layer", C, C)
        _conv1(g, out, pre + ".F.conv_after_fft", C, C // 2, gain=0.7)
        _conv1(g, out, pre + ".fusion", C, 2 * C, gain=0.5)

    for k, v in sd.items():  # keep the module order of the reference's state_dict
        if k.endswith(".conv.weight") and k.startswith("layers.") and k.count(".") == 3:
            sfb(k[: -len(".weight")])
        elif k == "conv_after_body.weight":
            sfb("conv_after_body")
        elif (k.endswith(".conv.bias") and k.startswith("layers.") and k.count(".") == 3) or k == "conv_after_body.bias":
            continue
        else:
            out[k] = v
    return out


EDSR_DEFAULT = dict(scale=4, n_colors=3, img_range=1.0, n_feats=256, n_resblocks=32, res_scale=0.1)
EDSR_TINY = dict(scale=4, n_colors=3, img_range=1.0, n_feats=64, n_resblocks=3, res_scale=0.1)
RGB_MEAN = (0.4488, 0.4371, 0.4040)


def edsr_weights(cfg: Dict, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """state_dict of reference EDSR(**cfg) (edsr.py:12-37), synthetic values."""
    g = _gen(seed)
    sd = OrderedDict()
    F = cfg["n_feats"]
    nc = cfg["n_colors"]
    mean = torch.tensor(RGB_MEAN) * cfg["img_range"]
    sd["sub_mean.weight"] = torch.eye(3).view(3, 3, 1, 1)
    sd["sub_mean.bias"] = -mean
    sd["add_mean.weight"] = torch.eye(3).view(3, 3, 1, 1)
    sd["add_mean.bias"] = mean.clone()
    _conv(g, sd, "head.0", F, nc)
    for i in range(cfg["n_resblocks"]):
        _conv(g, sd, f"body.{i}.body.0", F, F)
        _conv(g, sd, f"body.{i}.body.2", F, F)
    _conv(g, sd, f"body.{cfg['n_resblocks']}", F, F, gain=0.5)
    for i, cout, _ in upsampler_convs(cfg["scale"], F):
        _conv(g, sd, f"tail.0.{i}", cout, F)
    _conv(g, sd, "tail.1", nc, F)
    return sd


HAT_DEFAULT = dict(scale=4, n_colors=3, img_range=1.0, embed_dim=180, depths=[6] * 6, num_heads=[6] * 6, window_size=16,
                   mlp_ratio=2.0, compress_ratio=3, squeeze_factor=30, conv_scale=0.01, overlap_ratio=0.5)
HAT_TINY = dict(scale=4, n_colors=3, img_range=1.0, embed_dim=60, depths=[2, 2], num_heads=[6, 6], window_size=16,
                mlp_ratio=2.0, compress_ratio=3, squeeze_factor=30, conv_scale=0.01, overlap_ratio=0.5)


def hat_rpi_sa(ws: int) -> torch.Tensor:
    """hat.py:475-488 restated: index = (dy + ws - 1) * (2 ws - 1) + (dx + ws - 1)."""
    return relative_position_index(ws)


def hat_rpi_oca(ws: int, overlap_ratio: float) -> torch.Tensor:
    """hat.py:490-513 restated, quirk included: the reference shifts the (key - query) offsets by
    ws - wse + 1 (= -7 for 16 / 24), NOT to zero, so the index runs over [-880, 640] and the bias table is
    read with Python's negative indexing (table[i] == table[i + 1521] for i < 0)."""
    wse = ws + int(overlap_ratio * ws)
    o = torch.arange(ws * ws)
    e = torch.arange(wse * wse)
    dy = (e // wse)[None, :] - (o // ws)[:, None] + ws - wse + 1
    dx = (e % wse)[None, :] - (o % ws)[:, None] + ws - wse + 1
    return (dy * (ws + wse - 1) + dx).to(torch.int64)


def hat_weights(cfg: Dict, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """state_dict of reference HAT(**cfg) (hat.py:388-470), synthetic values; integer buffers restated."""
    g = _gen(seed)
    sd = OrderedDict()
    C, ws, nc = cfg["embed_dim"], cfg["window_size"], cfg["n_colors"]
    hid = int(C * cfg["mlp_ratio"])
    wse = ws + int(cfg["overlap_ratio"] * ws)
    sd["relative_position_index_SA"] = hat_rpi_sa(ws)
    sd["relative_position_index_OCA"] = hat_rpi_oca(ws, cfg["overlap_ratio"])
    _conv(g, sd, "conv_first", C, nc)
    _ln(g, sd, "patch_embed.norm", C)
    for li, (depth, nh) in enumerate(zip(cfg["depths"], cfg["num_heads"])):
        for bi in range(depth):
            p = f"layers.{li}.residual_group.blocks.{bi}"
            _ln(g, sd, p + ".norm1", C)
            sd[p + ".attn.relative_position_bias_table"] = _randn(g, (2 * ws - 1) ** 2, nh, std=0.5)
            _linear(g, sd, p + ".attn.qkv", 3 * C, C)
            _linear(g, sd, p + ".attn.proj", C, C, gain=0.5)
            _conv(g, sd, p + ".conv_block.cab.0", C // cfg["compress_ratio"], C)
            _conv(g, sd, p + ".conv_block.cab.2", C, C // cfg["compress_ratio"], gain=8.0)  # x conv_scale 0.01: keep it visible
            _conv(g, sd, p + ".conv_block.cab.3.attention.1", C // cfg["squeeze_factor"], C, k=1)
            _conv(g, sd, p + ".conv_block.cab.3.attention.3", C, C // cfg["squeeze_factor"], k=1, gain=2.0)
            _ln(g, sd, p + ".norm2", C)
            _linear(g, sd, p + ".mlp.fc1", hid, C)
            _linear(g, sd, p + ".mlp.fc2", C, hid, gain=0.5)
        p = f"layers.{li}.residual_group.overlap_attn"
        sd[p + ".relative_position_bias_table"] = _randn(g, (ws + wse - 1) ** 2, nh, std=0.5)
        _ln(g, sd, p + ".norm1", C)
        _linear(g, sd, p + ".qkv", 3 * C, C)
        _linear(g, sd, p + ".proj", C, C, gain=0.5)
        _ln(g, sd, p + ".norm2", C)
        _linear(g, sd, p + ".mlp.fc1", hid, C)
        _linear(g, sd, p + ".mlp.fc2", C, hid, gain=0.5)
        _conv(g, sd, f"layers.{li}.conv", C, C, gain=0.5)
    _ln(g, sd, "norm", C)
    _conv(g, sd, "conv_after_body", C, C, gain=0.5)
    _conv(g, sd, "conv_before_upsample.0", 64, C)
    for i, cout, _ in upsampler_convs(cfg["scale"], 64):
        _conv(g, sd, f"upsample.{i}", cout, 64)
    _conv(g, sd, "conv_last", nc, 64)
    return sd


RCAN_DEFAULT = dict(scale=4, n_colors=3, img_range=1.0, n_feats=64, n_resblocks=20, n_resgroups=10, reduction=16)
RCAN_TINY = dict(scale=4, n_colors=3, img_range=1.0, n_feats=64, n_resblocks=2, n_resgroups=2, reduction=16)


def rcan_weights(cfg: Dict, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """state_dict of reference RCAN(**cfg) (rcan.py:39-66, RCAB :11-24, ResidualGroup :27-36,
    ChannelAttention common.py:156-170), synthetic values."""
    g = _gen(seed)
    sd = OrderedDict()
    F = cfg["n_feats"]
    nc = cfg["n_colors"]
    R = F // cfg["reduction"]
    mean = torch.tensor(RGB_MEAN) * cfg["img_range"]
    sd["sub_mean.weight"] = torch.eye(3).view(3, 3, 1, 1)
    sd["sub_mean.bias"] = -mean
    sd["add_mean.weight"] = torch.eye(3).view(3, 3, 1, 1)
    sd["add_mean.bias"] = mean.clone()
    _conv(g, sd, "head.0", F, nc)
    for gi in range(cfg["n_resgroups"]):
        for bi in range(cfg["n_resblocks"]):
            p = f"body.{gi}.body.{bi}.body"
            _conv(g, sd, p + ".0", F, F)
            _conv(g, sd, p + ".2", F, F, gain=0.5)
            _conv(g, sd, p + ".3.conv_du.0", R, F, k=1)
            _conv(g, sd, p + ".3.conv_du.2", F, R, k=1, gain=2.0)
        _conv(g, sd, f"body.{gi}.body.{cfg['n_resblocks']}", F, F, gain=0.5)
    _conv(g, sd, f"body.{cfg['n_resgroups']}", F, F, gain=0.5)
    for i, cout, _ in upsampler_convs(cfg["scale"], F):
        _conv(g, sd, f"tail.0.{i}", cout, F)
    _conv(g, sd, "tail.1", nc, F)
    return sd


HAN_DEFAULT = dict(scale=4, n_colors=3, img_range=1.0, n_feats=64, n_resblocks=20, n_resgroups=10, reduction=16)
# (han.py:87 hard-codes `n_feats * 11` input channels for last_conv, i.e. n_resgroups must be 10; the tiny model keeps that)
HAN_TINY = dict(scale=4, n_colors=3, img_range=1.0, n_feats=64, n_resblocks=1, n_resgroups=10, reduction=16)


def han_weights(cfg: Dict, seed: int = 0) -> "OrderedDict[str, torch.Tensor]":
    """state_dict of reference HAN(**cfg) (han.py:55-88: the RCAN trunk plus csa = CSAM_Module :36-52, la = LAM_Module :12-33,
    last_conv, last), synthetic values.  The two gammas are zero-initialised upstream; here they are O(1) so that both attention
    modules take part in the output."""
    g = _gen(seed)
    sd = rcan_weights(cfg, seed + 1000)
    tail = OrderedDict((k, sd.pop(k)) for k in list(sd.keys()) if k.startswith("tail."))
    F = cfg["n_feats"]
    sd["csa.gamma"] = torch.tensor([0.7])
    sd["csa.conv.weight"] = torch.randn(1, 1, 3, 3, 3, generator=g) * 0.3
    sd["csa.conv.bias"] = torch.randn(1, generator=g) * 0.1
    sd["la.gamma"] = torch.tensor([0.5])
    _conv(g, sd, "last_conv", F, 11 * F, gain=0.5)
    _conv(g, sd, "last", F, 2 * F, gain=0.7)
    sd.update(tail)
    return sd


def image_batch(shape, seed: int = 1234) -> torch.Tensor:
    """Synthetic LR input in [0,1] (torch.rand, seed 1234)."""
    return torch.rand(*shape, generator=_gen(seed), dtype=torch.float32)


def smooth_image_u8(h: int, w: int, seed: int = 7):
    """Band-limited uint8 HWC test image (numpy) for Model.inference / tiler cases."""
    import numpy as np

    rng = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = np.zeros((h, w, 3))
    for c in range(3):
        for _ in range(6):
            fy, fx = rng.uniform(0.01, 0.35, 2)
            ph = rng.uniform(0, 6.28)
            img[..., c] += rng.uniform(0.2, 1.0) * np.sin(fy * yy + fx * xx + ph)
    img = (img - img.min()) / (img.max() - img.min())
    img = img * 255.0 + rng.normal(0, 3.0, img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)
