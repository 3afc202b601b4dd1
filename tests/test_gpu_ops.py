"""-m gpu: op-level parity of the CUDA kernels (through the C ABI) against the oracle restatement
on the same seeded inputs.  Tolerances: fp32 CUDA-core path ~1e-5 (summation order only);
tf32 / bf16 tensor-core paths are judged relative to the output scale."""
import math

import numpy as np

import pytest
import torch
import torch.nn.functional as F

from oracle import sr_oracle as O
from tests import gpu_util as G

pytestmark = pytest.mark.gpu
PRECS = ["fp32", "tf32", "bf16"]
# max-abs error allowed relative to the reference output's RMS
REL_TOL = {"fp32": 2e-5, "tf32": 2e-3, "bf16": 2.5e-2}


def _gen(seed):
    return torch.Generator().manual_seed(seed)


def _close(y, ref, prec, what):
    scale = ref.pow(2).mean().sqrt().item() + 1e-6
    err = (y.cpu() - ref).abs().max().item()
    assert math.isfinite(err) and err <= REL_TOL[prec] * scale * 4, f"{what} [{prec}]: max err {err:.3e} vs rms {scale:.3e}"


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("M,K,N,act,use_res,use_ln", [
    (5184, 180, 540, 0, False, False),   # qkv at cfg1 size
    (5184, 180, 180, 0, True, True),     # proj + residual + LN epilogue
    (1000, 180, 360, 3, False, False),   # fc1 + GELU, ragged M
    (777, 360, 180, 0, True, True),      # fc2 + residual + LN, ragged M
    (130, 60, 120, 3, False, False),     # tiny config
    (64, 120, 60, 0, True, True),
    (1, 180, 180, 0, False, False),      # single row
])
def test_linear(prec, M, K, N, act, use_res, use_ln):
    g = _gen(M + K + N)
    x = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K**0.5
    b = torch.randn(N, generator=g) * 0.1
    res = torch.randn(M, N, generator=g) if use_res else None
    lw = 1 + 0.1 * torch.randn(N, generator=g) if use_ln else None
    lb = 0.1 * torch.randn(N, generator=g) if use_ln else None
    v = x @ W.t() + b
    if act == 3:
        v = O.gelu(v)
    if use_res:
        v = v + res
    c = lambda t: None if t is None else t.cuda()
    y, y_ln = G.op_linear(prec, c(x), c(W), c(b), c(res), act, c(lw), c(lb))
    _close(y, v, prec, "linear")
    if use_ln:
        _close(y_ln, O.layer_norm(v, lw, lb), prec, "linear+LN")


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("B,Cin,Cout,H,W,act,ps_r,use_res", [
    (1, 180, 180, 72, 72, 0, 0, True),    # RSTB conv at cfg1 size + residual
    (2, 180, 64, 24, 40, 2, 0, False),    # conv_before_upsample + LeakyReLU
    (2, 64, 256, 16, 24, 0, 2, False),    # upsample conv + PixelShuffle(2)
    (1, 64, 576, 9, 13, 0, 3, False),     # x3 upsampler, ragged size
    (3, 60, 60, 8, 8, 1, 0, True),        # tiny config, ReLU
    (1, 256, 256, 12, 20, 1, 0, False),   # EDSR body conv
])
def test_conv3x3(prec, B, Cin, Cout, H, W, act, ps_r, use_res):
    g = _gen(B + Cin + Cout + H)
    x = torch.randn(B, Cin, H, W, generator=g)
    Wt = torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    res = torch.randn(B, Cout, H, W, generator=g) if use_res else None
    alpha = 0.1 if use_res else 1.0
    v = F.conv2d(x, Wt, b, padding=1)
    if act == 1:
        v = torch.relu(v)
    elif act == 2:
        v = F.leaky_relu(v, 0.01)
    v = v * alpha
    if use_res:
        v = v + res
    if ps_r > 1:
        v = O.pixel_shuffle(v, ps_r)
    y = G.op_conv3x3(prec, x.cuda(), Wt.cuda(), b.cuda(), None if res is None else res.cuda(), act, alpha, ps_r)
    assert y.shape == v.shape
    _close(y, v, prec, "conv3x3")


@pytest.mark.parametrize("prec", PRECS)
@pytest.mark.parametrize("B,H,W,C,heads,shift", [
    (1, 72, 72, 180, 6, 0), (1, 72, 72, 180, 6, 4), (2, 16, 24, 180, 6, 4), (3, 8, 8, 64, 4, 4), (2, 24, 16, 60, 6, 4),
])
def test_window_attention(prec, B, H, W, C, heads, shift):
    ws, d = 8, C // heads
    g = _gen(B + H + W + C + shift)
    qkv = torch.randn(B, H, W, 3 * C, generator=g)
    table = torch.randn((2 * ws - 1) ** 2, heads, generator=g) * 0.5
    # oracle: roll -> windows -> attention core -> reverse -> roll (no linear layers)
    q = torch.roll(qkv, (-shift, -shift), (1, 2)) if shift else qkv
    qw = O.to_windows(q, ws).reshape(-1, ws * ws, 3, heads, d)
    Q = qw[:, :, 0].transpose(1, 2) * d**-0.5
    K = qw[:, :, 1].transpose(1, 2)
    V = qw[:, :, 2].transpose(1, 2)
    s = Q @ K.transpose(-1, -2) + O.rel_pos_bias(table, ws)[None]
    mask = O.shift_mask(H, W, ws, shift, torch.float32)
    nW = mask.shape[0]
    s = (s.reshape(B, nW, heads, 64, 64) + mask[None, :, None]).reshape(-1, heads, 64, 64)
    o = (torch.softmax(s, -1) @ V).transpose(1, 2).reshape(-1, 64, C)
    o = O.from_windows(o, ws, B, H, W)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    y = G.op_window_attention(prec, qkv.cuda(), table.cuda(), heads, ws, shift)
    _close(y, o, prec, "window attention")


def _attention_oracle(qkv, table, B, H, W, C, heads, shift):
    """roll -> windows -> attention core -> reverse -> roll (swinir.py:83-102, 154-168), fp32."""
    ws, d = 8, C // heads
    q = torch.roll(qkv, (-shift, -shift), (1, 2)) if shift else qkv
    qw = O.to_windows(q, ws).reshape(-1, ws * ws, 3, heads, d)
    Q = qw[:, :, 0].transpose(1, 2) * d**-0.5
    K = qw[:, :, 1].transpose(1, 2)
    V = qw[:, :, 2].transpose(1, 2)
    s = Q @ K.transpose(-1, -2) + O.rel_pos_bias(table, ws)[None]
    mask = O.shift_mask(H, W, ws, shift, torch.float32)
    nW = mask.shape[0]
    s = (s.reshape(B, nW, heads, 64, 64) + mask[None, :, None]).reshape(-1, heads, 64, 64)
    o = (torch.softmax(s, -1) @ V).transpose(1, 2).reshape(-1, 64, C)
    o = O.from_windows(o, ws, B, H, W)
    return torch.roll(o, (shift, shift), (1, 2)) if shift else o


@pytest.mark.parametrize("B,H,W,shift", [
    (1, 72, 72, 0),   # cfg1: 81 windows -> the last work item holds a single window
    (1, 72, 72, 4),   # shifted: seam windows on the bottom row / right column / corner
    (2, 16, 24, 4),
    (3, 8, 8, 4),     # one window per image: every window touches both seams
    (5, 24, 16, 0),
])
def test_swin_attn_fused(B, H, W, shift):
    """k_swin_attn.cu: qkv projection + (shifted-)window attention in one tcgen05 kernel, bf16."""
    C, heads = 180, 6
    g = _gen(B + H + W + shift)
    xn = torch.randn(B, H, W, C, generator=g)
    Wq = torch.randn(3 * C, C, generator=g) / C**0.5
    bq = torch.randn(3 * C, generator=g) * 0.2
    table = torch.randn(225, heads, generator=g) * 0.5
    ref = _attention_oracle(xn @ Wq.t() + bq, table, B, H, W, C, heads, shift)
    y = G.op_swin_attn(xn.cuda(), Wq.cuda(), bq.cuda(), table.cuda(), heads, shift)
    _close(y, ref, "bf16", "fused window attention")


@pytest.mark.parametrize("M,with_ln", [(128, True), (300, True), (5184, False), (148 * 128 + 77, True)])
def test_swin_tail_fused(M, with_ln):
    """k_swin_tail.cu: proj + residual + LN2 + fc1 + GELU + fc2 + residual (+ next LN) in one tcgen05 kernel, bf16."""
    C, heads, hid = 180, 6, 360
    g = _gen(M)
    r = lambda *s: torch.randn(*s, generator=g)
    o, res = r(M, C), r(M, C)
    Wp, W1, W2 = r(C, C) / C**0.5, r(hid, C) / C**0.5, r(C, hid) / hid**0.5
    bp, b1, b2 = r(C) * 0.1, r(hid) * 0.1, r(C) * 0.1
    g2, be2, g3, be3 = 1 + 0.1 * r(C), 0.1 * r(C), 1 + 0.1 * r(C), 0.1 * r(C)
    t1 = o @ Wp.t() + bp + res
    h = O.gelu(O.layer_norm(t1, g2, be2) @ W1.t() + b1)
    y_ref = t1 + h @ W2.t() + b2
    yl_ref = O.layer_norm(y_ref, g3, be3) if with_ln else y_ref
    c = lambda t: t.cuda()
    y, yl = G.op_swin_mlp(c(o), c(res), c(Wp), c(bp), c(g2), c(be2), c(W1), c(b1), c(W2), c(b2),
                          c(g3) if with_ln else None, c(be3) if with_ln else None, heads, hid)
    _close(y, y_ref, "bf16", "fused swin tail")
    _close(yl, yl_ref, "bf16", "fused swin tail (second output)")


def test_device_psnr_and_augmentation_match_reference():
    """studiosr_b200.data (k_data.cu) against the reference's own outputs (tests/golden/data_ops.npz, made by
    oracle/make_golden_data.py) and the oracle restatement: compute_psnr incl. y_only / crop_border / unequal sizes, and the
    batched crop + flip + rot90 + array2tensor kernel with the reference's seeded draws."""
    import random

    from tests.conftest import load_golden
    from studiosr_b200.data import PairedAugment, compute_psnr

    g = load_golden("data_ops")
    gt, sr, big = g["psnr_gt"], g["psnr_sr"], g["psnr_big"]
    vals = []
    for a, b in ((sr, gt), (big, gt)):
        a_d, b_d = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
        for y_only in (False, True):
            for cb in (0, 4):
                vals.append(compute_psnr(a_d, b_d, y_only=y_only, crop_border=cb))
    assert np.allclose(vals, g["psnr_values"], rtol=0, atol=1e-4), (vals, g["psnr_values"])
    assert compute_psnr(gt, gt) == float("inf")
    lq, hr = g["aug_lq"], g["aug_gt"]
    lq_d, hr_d = torch.from_numpy(lq).cuda(), torch.from_numpy(hr).cuda()
    seeds = [int(s) for s in g["aug_seeds"]]
    aug = PairedAugment(size=12, scale=4)
    params = [PairedAugment(size=12, scale=4, rng=random.Random(s)).draw(lq.shape[0], lq.shape[1]) for s in seeds]
    assert len({p[2] for p in params}) > 3  # the seeds cover several flip / rotation combinations
    x, y = aug([lq_d] * len(seeds), [hr_d] * len(seeds), params=params)
    assert torch.equal(x.cpu(), torch.from_numpy(g["aug_x"])) and torch.equal(y.cpu(), torch.from_numpy(g["aug_y"]))
    # every flag combination against the oracle, on a non-square source
    combos = [(3, 5, f) for f in range(8)]
    x, y = aug([lq_d] * 8, [hr_d] * 8, params=combos)
    for i, (xs, ys, f) in enumerate(combos):
        xo, yo = O.augment_pair(lq, hr, 12, 4, xs, ys, f)
        assert np.array_equal(x[i].cpu().numpy(), xo) and np.array_equal(y[i].cpu().numpy(), yo), f
