"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference hot path.

Nothing under studiosr_b200/ may import this module; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do,
and only as the checker.

Parity status: PINNED.  The reference holds no golden vectors for this path
(tests/models/* assert shapes only), so the oracle is pinned by
fixtures produced by *executing the unmodified reference* in the build
container (oracle/make_golden.py -> tests/golden/*.npz); tests/test_oracle.py
checks every function below against them.

The restatement is functional (weights are a plain dict keyed like the
reference ``state_dict``), works in fp32 or fp64, and derives the shifted-window
mask and relative-position bias from token coordinates rather than from
materialised index / mask tensors.  Each function cites the reference lines it
follows (paths relative to /root/reference/studiosr/models/).
"""
import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

RGB_MEAN = (0.4488, 0.4371, 0.4040)  # common.py:223


# --------------------------------------------------------------------------- padding
def pad_for_eval(x: torch.Tensor, ws: int) -> torch.Tensor:
    """swinir.py:249-255 -- ALWAYS pads by (L//ws+1)*ws-L in [1,ws] rows/cols with an
    edge-inclusive mirror (row L+i == row L-1-i)."""
    _, _, h, w = x.shape
    H = (h // ws + 1) * ws
    W = (w // ws + 1) * ws
    ridx = torch.tensor([i if i < h else 2 * h - 1 - i for i in range(H)])
    cidx = torch.tensor([j if j < w else 2 * w - 1 - j for j in range(W)])
    return x[:, :, ridx][:, :, :, cidx]


def pad_for_train(x: torch.Tensor, ws: int) -> torch.Tensor:
    """common.py:277-282 -- reflect pad (edge-exclusive mirror) up to the next multiple."""
    _, _, h, w = x.shape
    H = (h + ws - 1) // ws * ws
    W = (w + ws - 1) // ws * ws
    ridx = torch.tensor([i if i < h else 2 * (h - 1) - i for i in range(H)])
    cidx = torch.tensor([j if j < w else 2 * (w - 1) - j for j in range(W)])
    return x[:, :, ridx][:, :, :, cidx]


# --------------------------------------------------------------------------- attention
def _region(p: torch.Tensor, L: int, ws: int, shift: int) -> torch.Tensor:
    """common.py:253-262 -- slice id of a coordinate in the *shifted* frame."""
    r = torch.zeros_like(p)
    r = torch.where(p >= L - ws, torch.ones_like(p), r)
    r = torch.where(p >= L - shift, torch.full_like(p, 2), r)
    return r


def shift_mask(H: int, W: int, ws: int, shift: int, dtype) -> torch.Tensor:
    """common.py:250-274 -- [nW, N, N] additive mask, -100 where region ids differ.
    With shift == 0 the reference's slices degenerate (slice(-0,None) overwrites
    everything with one id) and the mask is all zero."""
    nW = (H // ws) * (W // ws)
    N = ws * ws
    if shift == 0:
        return torch.zeros(nW, N, N, dtype=dtype)
    t = torch.arange(N)
    iy, ix = t // ws, t % ws
    out = torch.empty(nW, N, N, dtype=dtype)
    for wy in range(H // ws):
        for wx in range(W // ws):
            rid = 3 * _region(wy * ws + iy, H, ws, shift) + _region(wx * ws + ix, W, ws, shift)
            diff = rid[None, :] != rid[:, None]
            out[wy * (W // ws) + wx] = torch.where(diff, -100.0, 0.0).to(dtype)
    return out


def rel_pos_bias(table: torch.Tensor, ws: int) -> torch.Tensor:
    """swinir.py:57-67,86-91 -- bias[h, i, j] = table[(yi-yj+ws-1)*(2ws-1) + (xi-xj+ws-1), h]."""
    t = torch.arange(ws * ws)
    iy, ix = t // ws, t % ws
    idx = (iy[:, None] - iy[None, :] + ws - 1) * (2 * ws - 1) + (ix[:, None] - ix[None, :] + ws - 1)
    return table[idx.reshape(-1)].reshape(ws * ws, ws * ws, -1).permute(2, 0, 1)


def window_attention(P: Dict, pre: str, xw: torch.Tensor, heads: int, ws: int, mask) -> torch.Tensor:
    """swinir.py:78-105 -- xw [nWB, N, C] -> [nWB, N, C]."""
    nWB, N, C = xw.shape
    d = C // heads
    qkv = xw @ P[pre + ".qkv.weight"].t() + P[pre + ".qkv.bias"]
    qkv = qkv.reshape(nWB, N, 3, heads, d)
    q = qkv[:, :, 0].transpose(1, 2) * (d**-0.5)  # [nWB, heads, N, d]
    k = qkv[:, :, 1].transpose(1, 2)
    v = qkv[:, :, 2].transpose(1, 2)
    s = q @ k.transpose(-1, -2) + rel_pos_bias(P[pre + ".relative_position_bias_table"], ws)[None]
    if mask is not None:
        nW = mask.shape[0]
        s = (s.reshape(nWB // nW, nW, heads, N, N) + mask[None, :, None]).reshape(nWB, heads, N, N)
    p = torch.softmax(s, dim=-1)
    o = (p @ v).transpose(1, 2).reshape(nWB, N, C)
    return o @ P[pre + ".proj.weight"].t() + P[pre + ".proj.bias"]


def to_windows(x: torch.Tensor, ws: int) -> torch.Tensor:
    """common.py:236-240 -- [B,H,W,C] -> [B*nW, ws*ws, C]."""
    B, H, W, C = x.shape
    x = x.reshape(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(-1, ws * ws, C)


def from_windows(xw: torch.Tensor, ws: int, B: int, H: int, W: int) -> torch.Tensor:
    """common.py:243-247."""
    C = xw.shape[-1]
    x = xw.reshape(B, H // ws, W // ws, ws, ws, C).permute(0, 1, 3, 2, 4, 5)
    return x.reshape(B, H, W, C)


def layer_norm(x, w, b, eps=1e-5):
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu(x):
    """nn.GELU() default = exact erf form (common.py:186)."""
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def swin_block(P: Dict, pre: str, x: torch.Tensor, heads: int, ws: int, shift: int, drop=None) -> torch.Tensor:
    """swinir.py:146-174 -- x [B,H,W,C].  drop = (s_attn [B], s_mlp [B]): the per-sample factors timm's DropPath multiplies
    the two residual branches with in training (0 or 1/keep_prob; swinir.py:171-172), None = identity."""
    B, H, W, C = x.shape
    y = layer_norm(x, P[pre + ".norm1.weight"], P[pre + ".norm1.bias"])
    if shift > 0:
        y = torch.roll(y, (-shift, -shift), (1, 2))
    mask = shift_mask(H, W, ws, shift, x.dtype)  # the reference adds a mask on every block (swinir.py:161)
    a = window_attention(P, pre + ".attn", to_windows(y, ws), heads, ws, mask)
    y = from_windows(a, ws, B, H, W)
    if shift > 0:
        y = torch.roll(y, (shift, shift), (1, 2))
    if drop is not None:
        y = y * drop[0].to(y.dtype).view(-1, 1, 1, 1)
    x = x + y
    y = layer_norm(x, P[pre + ".norm2.weight"], P[pre + ".norm2.bias"])
    y = gelu(y @ P[pre + ".mlp.fc1.weight"].t() + P[pre + ".mlp.fc1.bias"])
    y = y @ P[pre + ".mlp.fc2.weight"].t() + P[pre + ".mlp.fc2.bias"]
    if drop is not None:
        y = y * drop[1].to(y.dtype).view(-1, 1, 1, 1)
    return x + y


def conv3x3(P: Dict, name: str, x: torch.Tensor) -> torch.Tensor:
    """nn.Conv2d(cin, cout, 3, 1, 1) on NCHW."""
    return F.conv2d(x, P[name + ".weight"], P[name + ".bias"], padding=1)


def pixel_shuffle(x: torch.Tensor, r: int) -> torch.Tensor:
    """nn.PixelShuffle: out[b, c, y*r+i, x*r+j] = in[b, c*r*r + i*r + j, y, x]."""
    B, Crr, H, W = x.shape
    C = Crr // (r * r)
    return x.reshape(B, C, r, r, H, W).permute(0, 1, 4, 2, 5, 3).reshape(B, C, H * r, W * r)


def _upsampler(P: Dict, pre: str, x: torch.Tensor, scale: int, n_feats: int, num_out_ch=None):
    """common.py:124-137."""
    from .synth import upsampler_convs

    for i, _, r in upsampler_convs(scale, n_feats, num_out_ch):
        x = pixel_shuffle(conv3x3(P, f"{pre}.{i}", x), r)
    return x


# --------------------------------------------------------------------------- SwinIR
def conv1x1(P: Dict, name: str, x: torch.Tensor) -> torch.Tensor:
    """nn.Conv2d(cin, cout, 1, 1, 0) on NCHW."""
    return F.conv2d(x, P[name + ".weight"], P[name + ".bias"])


def sfb(P: Dict, pre: str, x: torch.Tensor) -> torch.Tensor:
    """SwinFIR's SFB (swinfir.py:68-80) on NCHW: fusion(cat[SpatialB(x) :53-65, SpectralTransform(x) :37-50]) with the
    FourierUnit :9-34 (rfftn over (H, W), norm "ortho"; 1x1 conv + LeakyReLU(0.2) on [real | imag]; irfftn back to H x W)."""
    sp = conv3x3(P, pre + ".S.body.2", F.leaky_relu(conv3x3(P, pre + ".S.body.0", x), 0.2)) + x
    y = F.leaky_relu(conv1x1(P, pre + ".F.conv_before_fft.0", x), 0.2)
    c = y.shape[1]
    f = torch.fft.rfftn(y, dim=(-2, -1), norm="ortho")
    f = F.leaky_relu(conv1x1(P, pre + ".F.fu.conv_layer", torch.cat((f.real, f.imag), dim=1)), 0.2)
    fo = torch.fft.irfftn(torch.complex(f[:, :c], f[:, c:]), s=y.shape[-2:], dim=(-2, -1), norm="ortho")
    sf = conv1x1(P, pre + ".F.conv_after_fft", fo + y)
    return conv1x1(P, pre + ".fusion", torch.cat([sp, sf], dim=1))


def swinir_forward(P: Dict, x: torch.Tensor, cfg: Dict, training: bool = False, drop_scale=None) -> torch.Tensor:
    """swinir.py:353-372 (+ forward_features :342-351, RSTB :245-246).  drop_scale [2 * n_blocks, B] = the stochastic-depth
    factors of this step (row 2k / 2k+1: attention / MLP branch of block k), None = drop_path off.  cfg["sfb"] = True: SwinFIR
    (swinfir.py:83-114): every RSTB conv and conv_after_body is an SFB.

    x: [B, n_colors, H, W] float -> [B, n_colors, s*H, s*W]."""
    dt = x.dtype
    P = {k: (v.to(dt) if v.is_floating_point() else v) for k, v in P.items()}
    ws, s, C = cfg["window_size"], cfg["scale"], cfg["embed_dim"]
    body_conv = sfb if cfg.get("sfb") else conv3x3
    h0, w0 = x.shape[2:]
    x = pad_for_train(x, ws) if training else pad_for_eval(x, ws)
    mean = torch.tensor(RGB_MEAN, dtype=dt).view(1, 3, 1, 1)
    x = x / cfg["img_range"] - mean  # common.py:228-230
    x0 = conv3x3(P, "conv_first", x)
    t = x0.permute(0, 2, 3, 1)
    t = layer_norm(t, P["patch_embed.norm.weight"], P["patch_embed.norm.bias"])
    k = 0
    for li, depth in enumerate(cfg["depths"]):
        g = t
        for bi in range(depth):
            shift = 0 if bi % 2 == 0 else ws // 2
            drop = None if drop_scale is None else (drop_scale[2 * k], drop_scale[2 * k + 1])
            t = swin_block(P, f"layers.{li}.residual_group.blocks.{bi}", t, cfg["num_heads"][li], ws, shift, drop)
            k += 1
        t = body_conv(P, f"layers.{li}.conv", t.permute(0, 3, 1, 2)).permute(0, 2, 3, 1) + g
    t = layer_norm(t, P["norm.weight"], P["norm.bias"])
    y = body_conv(P, "conv_after_body", t.permute(0, 3, 1, 2)) + x0
    if cfg["upsampler"] == "pixelshuffle":
        y = F.leaky_relu(conv3x3(P, "conv_before_upsample.0", y), 0.01)
        y = conv3x3(P, "conv_last", _upsampler(P, "upsample", y, s, 64))
    else:
        y = _upsampler(P, "upsample", y, s, C, cfg["n_colors"])
    y = (y + mean) * cfg["img_range"]  # common.py:232-233
    return y[:, :, : h0 * s, : w0 * s]


# --------------------------------------------------------------------------- EDSR
def edsr_forward(P: Dict, x: torch.Tensor, cfg: Dict) -> torch.Tensor:
    """edsr.py:39-48 with ResBlock common.py:140-153 and MeanShift common.py:108-121."""
    dt = x.dtype
    P = {k: v.to(dt) for k, v in P.items()}
    x = F.conv2d(x, P["sub_mean.weight"], P["sub_mean.bias"])
    x = conv3x3(P, "head.0", x)
    r = x
    n = cfg["n_resblocks"]
    for i in range(n):
        b = conv3x3(P, f"body.{i}.body.2", torch.relu(conv3x3(P, f"body.{i}.body.0", r)))
        r = b * cfg["res_scale"] + r
    r = conv3x3(P, f"body.{n}", r) + x
    y = conv3x3(P, "tail.1", _upsampler(P, "tail.0", r, cfg["scale"], cfg["n_feats"]))
    return F.conv2d(y, P["add_mean.weight"], P["add_mean.bias"])


# --------------------------------------------------------------------------- HAT
def _hat_window_attention(P, pre, xw, heads, ws, mask):
    """hat.py:85-110 -- same arithmetic as SwinIR's WindowAttention with an explicit 16x16 rpi."""
    B_, N, C = xw.shape
    d = C // heads
    qkv = (xw @ P[pre + ".qkv.weight"].t() + P[pre + ".qkv.bias"]).reshape(B_, N, 3, heads, d).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * d**-0.5, qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1) + rel_pos_bias(P[pre + ".relative_position_bias_table"], ws)[None]
    if mask is not None:
        nW = mask.shape[0]
        attn = (attn.view(B_ // nW, nW, heads, N, N) + mask[None, :, None]).view(-1, heads, N, N)
    x = (torch.softmax(attn, -1) @ v).transpose(1, 2).reshape(B_, N, C)
    return x @ P[pre + ".proj.weight"].t() + P[pre + ".proj.bias"]


def _hat_cab(P, pre, x_nhwc):
    """hat.py:41-52 (+ ChannelAttention :25-38) on NHWC input; returns NHWC."""
    x = x_nhwc.permute(0, 3, 1, 2)
    t = gelu(conv3x3(P, pre + ".cab.0", x))
    t = conv3x3(P, pre + ".cab.2", t)
    y = t.mean(dim=(2, 3), keepdim=True)
    y = torch.relu(F.conv2d(y, P[pre + ".cab.3.attention.1.weight"], P[pre + ".cab.3.attention.1.bias"]))
    y = torch.sigmoid(F.conv2d(y, P[pre + ".cab.3.attention.3.weight"], P[pre + ".cab.3.attention.3.bias"]))
    return (t * y).permute(0, 2, 3, 1)


def hat_hab(P, pre, x, heads, ws, shift, conv_scale):
    """hat.py:154-195 on [B,H,W,C]."""
    B, H, W, C = x.shape
    y = layer_norm(x, P[pre + ".norm1.weight"], P[pre + ".norm1.bias"])
    conv_x = _hat_cab(P, pre + ".conv_block", y)
    if shift > 0:
        y = torch.roll(y, (-shift, -shift), (1, 2))
    mask = shift_mask(H, W, ws, shift, x.dtype) if shift > 0 else None  # hat.py:171: no mask on unshifted blocks
    a = _hat_window_attention(P, pre + ".attn", to_windows(y, ws), heads, ws, mask)
    y = from_windows(a, ws, B, H, W)
    if shift > 0:
        y = torch.roll(y, (shift, shift), (1, 2))
    x = x + y + conv_x * conv_scale
    y = layer_norm(x, P[pre + ".norm2.weight"], P[pre + ".norm2.bias"])
    y = gelu(y @ P[pre + ".mlp.fc1.weight"].t() + P[pre + ".mlp.fc1.bias"])
    return x + y @ P[pre + ".mlp.fc2.weight"].t() + P[pre + ".mlp.fc2.bias"]


def hat_ocab(P, pre, x, heads, ws, overlap_ratio):
    """hat.py:240-293 on [B,H,W,C]: queries in ws x ws windows, keys / values in the (1 + overlap_ratio) ws windows
    around them (nn.Unfold with zero padding: out-of-image keys are ZERO vectors that still take part in the softmax
    with their bias term)."""
    B, H, W, C = x.shape
    d = C // heads
    wse = ws + int(ws * overlap_ratio)
    pad = (wse - ws) // 2
    y = layer_norm(x, P[pre + ".norm1.weight"], P[pre + ".norm1.bias"])
    qkv = y @ P[pre + ".qkv.weight"].t() + P[pre + ".qkv.bias"]  # [B,H,W,3C]
    q = to_windows(qkv[..., :C], ws)  # [nW*B, ws*ws, C]
    kv = F.pad(qkv[..., C:].permute(0, 3, 1, 2), (pad, pad, pad, pad))  # [B,2C,H+2p,W+2p], zeros
    nwy, nwx = H // ws, W // ws
    kw = kv.unfold(2, wse, ws).unfold(3, wse, ws)  # [B,2C,nwy,nwx,wse,wse]
    kw = kw.permute(0, 2, 3, 4, 5, 1).reshape(B * nwy * nwx, wse * wse, 2 * C)
    k, v = kw[..., :C], kw[..., C:]
    B_ = q.shape[0]
    qh = q.reshape(B_, ws * ws, heads, d).permute(0, 2, 1, 3) * d**-0.5
    kh = k.reshape(B_, wse * wse, heads, d).permute(0, 2, 1, 3)
    vh = v.reshape(B_, wse * wse, heads, d).permute(0, 2, 1, 3)
    table = P[pre + ".relative_position_bias_table"]
    o = torch.arange(ws * ws)
    e = torch.arange(wse * wse)
    # hat.py:508-512: offsets are shifted by ws - wse + 1 (negative!), the table is read with negative-index wrap-around
    dy = (e // wse)[None, :] - (o // ws)[:, None] + ws - wse + 1
    dx = (e % wse)[None, :] - (o % ws)[:, None] + ws - wse + 1
    idx = (dy * (ws + wse - 1) + dx).reshape(-1)
    bias = table[idx % table.shape[0]].reshape(ws * ws, wse * wse, heads).permute(2, 0, 1)
    attn = torch.softmax(qh @ kh.transpose(-2, -1) + bias[None], -1)
    a = (attn @ vh).transpose(1, 2).reshape(B_, ws * ws, C)
    y = from_windows(a, ws, B, H, W)
    x = y @ P[pre + ".proj.weight"].t() + P[pre + ".proj.bias"] + x
    y = layer_norm(x, P[pre + ".norm2.weight"], P[pre + ".norm2.bias"])
    y = gelu(y @ P[pre + ".mlp.fc1.weight"].t() + P[pre + ".mlp.fc1.bias"])
    return x + y @ P[pre + ".mlp.fc2.weight"].t() + P[pre + ".mlp.fc2.bias"]


def hat_forward(P: Dict, x: torch.Tensor, cfg: Dict) -> torch.Tensor:
    """hat.py:542-554 (forward), :519-540 (forward_features), RHAG :379-385, AttenBlocks :340-344.
    Reflect padding to a multiple of the window in BOTH modes (common.py:277-282)."""
    dt = x.dtype
    P = {k: (v.to(dt) if v.is_floating_point() else v) for k, v in P.items()}
    h0, w0 = x.shape[2:]
    ws, C, s = cfg["window_size"], cfg["embed_dim"], cfg["scale"]
    x = pad_for_train(x, ws)
    mean = torch.tensor([0.4488, 0.4371, 0.4040], dtype=dt).view(1, 3, 1, 1)
    x = x / cfg["img_range"] - mean
    x = conv3x3(P, "conv_first", x)
    t = layer_norm(x.permute(0, 2, 3, 1), P["patch_embed.norm.weight"], P["patch_embed.norm.bias"])
    for li, (depth, nh) in enumerate(zip(cfg["depths"], cfg["num_heads"])):
        g = t
        for bi in range(depth):
            t = hat_hab(P, f"layers.{li}.residual_group.blocks.{bi}", t, nh, ws, 0 if bi % 2 == 0 else ws // 2, cfg["conv_scale"])
        t = hat_ocab(P, f"layers.{li}.residual_group.overlap_attn", t, nh, ws, cfg["overlap_ratio"])
        t = conv3x3(P, f"layers.{li}.conv", t.permute(0, 3, 1, 2)).permute(0, 2, 3, 1) + g
    t = layer_norm(t, P["norm.weight"], P["norm.bias"])
    y = conv3x3(P, "conv_after_body", t.permute(0, 3, 1, 2)) + x
    y = F.leaky_relu(conv3x3(P, "conv_before_upsample.0", y), 0.01)
    y = conv3x3(P, "conv_last", _upsampler(P, "upsample", y, s, 64))
    y = (y + mean) * cfg["img_range"]
    return y[:, :, : h0 * s, : w0 * s]


# --------------------------------------------------------------------------- RCAN
def channel_attention(P: Dict, pre: str, x: torch.Tensor) -> torch.Tensor:
    """common.py:156-170 -- x * sigmoid(W2 relu(W1 avgpool(x) + b1) + b2), 1x1 convs on the pooled [B,C,1,1]."""
    y = x.mean(dim=(2, 3), keepdim=True)
    y = torch.relu(F.conv2d(y, P[pre + ".conv_du.0.weight"], P[pre + ".conv_du.0.bias"]))
    y = torch.sigmoid(F.conv2d(y, P[pre + ".conv_du.2.weight"], P[pre + ".conv_du.2.bias"]))
    return x * y


def rcan_forward(P: Dict, x: torch.Tensor, cfg: Dict) -> torch.Tensor:
    """rcan.py:68-77 with RCAB :11-24, ResidualGroup :27-36 and MeanShift common.py:108-121."""
    dt = x.dtype
    P = {k: v.to(dt) for k, v in P.items()}
    x = F.conv2d(x, P["sub_mean.weight"], P["sub_mean.bias"])
    x = conv3x3(P, "head.0", x)
    g = x
    nb, ng = cfg["n_resblocks"], cfg["n_resgroups"]
    for gi in range(ng):
        r = g
        for bi in range(nb):
            p = f"body.{gi}.body.{bi}.body"
            t = conv3x3(P, p + ".2", torch.relu(conv3x3(P, p + ".0", r)))
            r = channel_attention(P, p + ".3", t) + r
        g = conv3x3(P, f"body.{gi}.body.{nb}", r) + g
    r = conv3x3(P, f"body.{ng}", g) + x
    y = conv3x3(P, "tail.1", _upsampler(P, "tail.0", r, cfg["scale"], cfg["n_feats"]))
    return F.conv2d(y, P["add_mean.weight"], P["add_mean.bias"])


# --------------------------------------------------------------------------- HAN
def han_lam(x5: torch.Tensor, gamma: torch.Tensor) -> torch.Tensor:
    """LAM_Module (han.py:12-33): attention between the N stacked feature maps.  x5 [B, N, C, H, W] -> [B, N*C, H, W].
    energy = Q Q^T over the flattened C*H*W axis, softmax of (row max - energy), out = gamma * (attention @ Q) + x."""
    B, N, C, H, W = x5.shape
    q = x5.reshape(B, N, -1)
    energy = torch.bmm(q, q.transpose(1, 2))
    energy_new = energy.max(dim=-1, keepdim=True)[0].expand_as(energy) - energy
    att = torch.softmax(energy_new, dim=-1)
    out = torch.bmm(att, q).reshape(B, N, C, H, W)
    return (gamma * out + x5).reshape(B, N * C, H, W)


def han_csam(P: Dict, x: torch.Tensor) -> torch.Tensor:
    """CSAM_Module (han.py:36-52): one 3x3x3 Conv3d over the (C, H, W) volume, sigmoid, x * (gamma * s) + x."""
    s = torch.sigmoid(F.conv3d(x.unsqueeze(1), P["csa.conv.weight"], P["csa.conv.bias"], padding=1))
    return x * (P["csa.gamma"] * s).reshape(x.shape) + x


def han_forward(P: Dict, x: torch.Tensor, cfg: Dict) -> torch.Tensor:
    """han.py:90-113: the RCAN trunk whose ng + 1 body outputs are stacked (newest first, :94-99), layer attention + last_conv
    on the stack, channel-spatial attention on the trunk output, `last` on their concatenation, long skip, tail."""
    dt = x.dtype
    P = {k: v.to(dt) for k, v in P.items()}
    x = F.conv2d(x, P["sub_mean.weight"], P["sub_mean.bias"])
    x = conv3x3(P, "head.0", x)
    nb, ng = cfg["n_resblocks"], cfg["n_resgroups"]
    res, stack = x, []
    for gi in range(ng):
        r = res
        for bi in range(nb):
            p = f"body.{gi}.body.{bi}.body"
            t = conv3x3(P, p + ".2", torch.relu(conv3x3(P, p + ".0", r)))
            r = channel_attention(P, p + ".3", t) + r
        res = conv3x3(P, f"body.{gi}.body.{nb}", r) + res
        stack.insert(0, res)
    res = conv3x3(P, f"body.{ng}", res)
    stack.insert(0, res)
    out2 = conv3x3(P, "last_conv", han_lam(torch.stack(stack, dim=1), P["la.gamma"]))
    out1 = han_csam(P, res)
    res = conv3x3(P, "last", torch.cat([out1, out2], dim=1)) + x
    y = conv3x3(P, "tail.1", _upsampler(P, "tail.0", res, cfg["scale"], cfg["n_feats"]))
    return F.conv2d(y, P["add_mean.weight"], P["add_mean.bias"])


# --------------------------------------------------------------------------- Model.inference + tiler
def quantize_u8(y: torch.Tensor, img_range: float) -> torch.Tensor:
    """common.py:44-45 -- [3,H,W] float -> [H,W,3] uint8 (round half to even, clip)."""
    scale = 255.0 if img_range == 1.0 else 1.0
    return (y.permute(1, 2, 0) * scale).round().clip(0, 255).to(torch.uint8)


def inference_u8(forward, image_u8, img_range: float = 1.0):
    """common.py:36-48 -- numpy uint8 [H,W,3] -> numpy uint8 [sH,sW,3]; `forward` is an
    eval-mode fp32 model forward."""
    import numpy as np

    scale = 255.0 if img_range == 1.0 else 1.0
    x = torch.from_numpy(image_u8.astype(np.float32) / scale).permute(2, 0, 1).unsqueeze(0)
    return quantize_u8(forward(x)[0], img_range).numpy()


def tile_starts(L: int, tile: int, stride: int):
    """Tile origins along one axis: 0, stride, ... with the last tile clamped to L - tile."""
    if L <= tile:
        return [0]
    s = list(range(0, L - tile, stride)) + [L - tile]
    return s


def ramp_weight(tile: int, overlap: int, first: bool, last: bool, scale: int) -> torch.Tensor:
    """1-D blend weight over scale*tile output samples: linear ramp across the `overlap`
    LR pixels shared with a neighbour, flat 1 elsewhere (no ramp at the frame border)."""
    n, o = tile * scale, overlap * scale
    w = torch.ones(n, dtype=torch.float64)
    ramp = (torch.arange(o, dtype=torch.float64) + 0.5) / o
    if not first:
        w[:o] = ramp
    if not last:
        w[n - o:] = ramp.flip(0)
    return w


def tiled_upscale(forward, x: torch.Tensor, scale: int, tile: int = 64, overlap: int = 16) -> torch.Tensor:
    """NEW capability (the reference has no tiler): the oracle is this tiler
    calling the reference-equivalent eval forward per tile.  x [1,3,H,W] float.

    Tiles are tile x tile LR crops at stride tile-overlap (last one clamped to the
    border); outputs are blended with separable linear ramps and normalised by the
    accumulated weight.  When the last tile is clamped its overlap with the previous
    tile is larger than `overlap`; the ramp still spans `overlap` pixels, so the
    normalisation by the weight sum is what keeps the blend exact."""
    _, C, H, W = x.shape
    ys, xs = tile_starts(H, tile, tile - overlap), tile_starts(W, tile, tile - overlap)
    th, tw = min(tile, H), min(tile, W)
    acc = torch.zeros(C, H * scale, W * scale, dtype=torch.float64)
    wsum = torch.zeros(H * scale, W * scale, dtype=torch.float64)
    for iy, y0 in enumerate(ys):
        wy = ramp_weight(th, min(overlap, th), iy == 0, iy == len(ys) - 1, scale)
        for ix, x0 in enumerate(xs):
            wx = ramp_weight(tw, min(overlap, tw), ix == 0, ix == len(xs) - 1, scale)
            out = forward(x[:, :, y0 : y0 + th, x0 : x0 + tw])[0].to(torch.float64)
            w2 = wy[:, None] * wx[None, :]
            acc[:, y0 * scale : (y0 + th) * scale, x0 * scale : (x0 + tw) * scale] += out * w2
            wsum[y0 * scale : (y0 + th) * scale, x0 * scale : (x0 + tw) * scale] += w2
    return (acc / wsum).to(x.dtype).unsqueeze(0)


def blend_tiles(tiles: torch.Tensor, H: int, W: int, scale: int, tile: int = 64, overlap: int = 16, row_begin: int = 0,
                row_end: int = -1) -> torch.Tensor:
    """The blend of `tiled_upscale` from an explicit row-major list of tile outputs [n_tiles, C, th*s, tw*s], restricted to
    the output rows [row_begin, row_end): the checker of the sharded protocol (studiosr_b200/sharding.py), where a rank
    blends only its band of the frame.  Returns [C, row_end - row_begin, W*s] (float64)."""
    ys, xs = tile_starts(H, tile, tile - overlap), tile_starts(W, tile, tile - overlap)
    th, tw = min(tile, H), min(tile, W)
    row_end = H * scale if row_end < 0 else row_end
    C = tiles.shape[1]
    acc = torch.zeros(C, H * scale, W * scale, dtype=torch.float64)
    wsum = torch.zeros(H * scale, W * scale, dtype=torch.float64)
    for iy, y0 in enumerate(ys):
        if y0 * scale >= row_end or (y0 + th) * scale <= row_begin:
            continue
        wy = ramp_weight(th, min(overlap, th), iy == 0, iy == len(ys) - 1, scale)
        for ix, x0 in enumerate(xs):
            wx = ramp_weight(tw, min(overlap, tw), ix == 0, ix == len(xs) - 1, scale)
            w2 = wy[:, None] * wx[None, :]
            acc[:, y0 * scale : (y0 + th) * scale, x0 * scale : (x0 + tw) * scale] += tiles[iy * len(xs) + ix].to(torch.float64) * w2
            wsum[y0 * scale : (y0 + th) * scale, x0 * scale : (x0 + tw) * scale] += w2
    return acc[:, row_begin:row_end] / wsum[row_begin:row_end]


def to_y(image):
    """utils/metrics.py:11-17: BT.601 luma of an HWC image (uint8 -> float32 / 255 first), float64 result."""
    import numpy as np

    if not (image.ndim == 3 and image.shape[-1] == 3):
        return image
    if image.dtype == np.uint8:
        image = image.astype(np.float32) / 255.0
    return np.dot(image, [65.481, 128.553, 24.966]) + 16.0


def compute_psnr(im1, im2, y_only: bool = False, crop_border: int = 0) -> float:
    """utils/metrics.py:20-49 (crop_img_to_equal + compute_psnr) on numpy HWC images."""
    import numpy as np

    h, w = min(im1.shape[0], im2.shape[0]), min(im1.shape[1], im2.shape[1])
    im1, im2 = im1[:h, :w], im2[:h, :w]
    if crop_border:
        im1 = im1[crop_border:-crop_border, crop_border:-crop_border]
        im2 = im2[crop_border:-crop_border, crop_border:-crop_border]
    if y_only:
        im1, im2 = to_y(im1), to_y(im2)
    elif im1.dtype != np.uint8:
        im1, im2 = im1 * 255.0, im2 * 255.0
    error = np.mean((im1.astype(np.float32) - im2.astype(np.float32)) ** 2)
    return float("inf") if error == 0 else float(20 * np.log10(255.0 / np.sqrt(error)))


def augment_pair(lq, gt, size: int, scale: int, xs: int, ys: int, flags: int):
    """data/transforms.py:8-68 with the random draws made explicit: crop at (xs, ys), then fliplr (flags & 1), flipud (flags & 2),
    np.rot90 (flags & 4) in the order of dataset.py:50-58, then array2tensor (CHW float32 / 255)."""
    import numpy as np

    a = lq[ys:ys + size, xs:xs + size]
    b = gt[ys * scale:(ys + size) * scale, xs * scale:(xs + size) * scale]
    if flags & 1:
        a, b = np.fliplr(a), np.fliplr(b)
    if flags & 2:
        a, b = np.flipud(a), np.flipud(b)
    if flags & 4:
        a, b = np.rot90(a), np.rot90(b)
    f = lambda t: np.ascontiguousarray(t.transpose(2, 0, 1)).astype(np.float32) / 255
    return f(a), f(b)


def psnr(a: torch.Tensor, b: torch.Tensor, peak: float = 255.0) -> float:
    """utils/metrics.py:36-49 on already-cropped arrays (no Y conversion)."""
    mse = ((a.double() - b.double()) ** 2).mean().item()
    return float("inf") if mse == 0 else 10.0 * math.log10(peak * peak / mse)
