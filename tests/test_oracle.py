"""CPU: pin the oracle restatement (oracle/sr_oracle.py) against fixtures produced by
executing the unmodified reference (oracle/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import sr_oracle as O
from oracle import synth
from tests.conftest import load_golden

SWINIR_CASES = [
    "swinir_tiny_x4_eval_2x20x28", "swinir_tiny_x4_eval_1x16x16", "swinir_tiny_x4_train_1x12x12",
    "swinir_tiny_x4_train_2x16x24", "swinir_tiny_x2_eval_1x12x12", "swinir_tiny_x3_eval_1x8x8",
    "swinir_tiny_x8_eval_1x8x8", "swinir_light_x4_eval_1x12x20", "swinir_full_x4_eval_cfg1",
]
EDSR_CASES = ["edsr_tiny_x4_2x12x20", "edsr_tiny_x2_1x9x11", "edsr_tiny_x3_1x8x8", "edsr_full_x4_1x24x24"]
HAT_CASES = ["hat_tiny_x4_eval_2x20x40", "hat_tiny_x4_train_1x32x32", "hat_tiny_x2_eval_1x16x48", "hat_tiny_x3_eval_1x17x17",
             "hat_full_x4_eval_1x64x64"]
RCAN_CASES = ["rcan_tiny_x4_2x12x20", "rcan_tiny_x2_1x9x11", "rcan_tiny_x3_1x8x8", "rcan_full_x4_1x24x24"]
HAN_CASES = ["han_tiny_x4_2x12x20", "han_tiny_x2_1x9x11", "han_tiny_x3_1x8x8", "han_full_x4_1x16x16"]


@pytest.mark.parametrize("name", SWINIR_CASES)
def test_swinir_oracle_matches_reference_golden(name, golden_meta):
    c = golden_meta[name]
    P = synth.swinir_weights(c["cfg"], c["wseed"])
    x = synth.image_batch(c["shape"], c["xseed"])
    y = O.swinir_forward(P, x, c["cfg"], training=c["training"])
    ref = torch.from_numpy(load_golden(name)["y"])
    assert list(y.shape) == c["out_shape"]
    # same fp32 math, different op order: reference fp32-vs-fp64 noise is 3.6e-7
    assert (y - ref).abs().max().item() < 2e-5


@pytest.mark.parametrize("name", HAT_CASES)
def test_hat_oracle_matches_reference_golden(name, golden_meta):
    """oracle/sr_oracle.py:hat_forward (HAB + CAB + OCAB, hat.py) vs the reference's own HAT forward
    (fixtures: oracle/make_golden_hat.py; cfg3-class full model included)."""
    c = golden_meta[name]
    P = synth.hat_weights(c["cfg"], c["wseed"])
    x = synth.image_batch(c["shape"], c["xseed"])
    y = O.hat_forward(P, x, c["cfg"])
    ref = torch.from_numpy(load_golden(name)["y"])
    assert list(y.shape) == c["out_shape"]
    assert (y - ref).abs().max().item() <= 2e-5


def test_hat_index_buffers_match_reference():
    """relative_position_index_SA / _OCA as the reference computes them (the OCA one runs over [-880, 640])."""
    g = load_golden("hat_ops")
    assert (synth.hat_rpi_sa(16).numpy() == g["rpi_sa"]).all()
    assert (synth.hat_rpi_oca(16, 0.5).numpy() == g["rpi_oca"]).all()
    from studiosr_b200.models.hat import _rpi_oca

    assert (_rpi_oca(16, 0.5).numpy() == g["rpi_oca"]).all()


@pytest.mark.parametrize("name", RCAN_CASES)
def test_rcan_oracle_matches_reference_golden(name, golden_meta):
    """oracle/sr_oracle.py:rcan_forward vs the reference's own RCAN forward (fixtures: oracle/make_golden_rcan.py)."""
    c = golden_meta[name]
    P = synth.rcan_weights(c["cfg"], c["wseed"])
    x = synth.image_batch(c["shape"], c["xseed"])
    y = O.rcan_forward(P, x, c["cfg"])
    ref = torch.from_numpy(load_golden(name)["y"])
    assert list(y.shape) == c["out_shape"]
    assert (y - ref).abs().max().item() <= 2e-5


@pytest.mark.parametrize("name", HAN_CASES)
def test_han_oracle_matches_reference_golden(name, golden_meta):
    """oracle/sr_oracle.py:han_forward (LAM han.py:12-33, CSAM :36-52) vs the reference's own HAN forward (fixtures: oracle/make_golden_han.py)."""
    c = golden_meta[name]
    P = synth.han_weights(c["cfg"], c["wseed"])
    x = synth.image_batch(c["shape"], c["xseed"])
    y = O.han_forward(P, x, c["cfg"])
    ref = torch.from_numpy(load_golden(name)["y"])
    assert list(y.shape) == c["out_shape"]
    assert (y - ref).abs().max().item() <= 2e-5


SWINFIR_CASES = ["swinfir_tiny_x4_eval_1x12x20", "swinfir_tiny_x2_eval_2x16x16", "swinfir_tiny_x4_train_1x16x24", "swinfir_c180_x4_eval_1x8x8"]


@pytest.mark.parametrize("name", SWINFIR_CASES)
def test_swinfir_oracle_matches_reference_golden(name, golden_meta):
    """oracle/sr_oracle.py:swinir_forward with cfg["sfb"] (SFB swinfir.py:68-80, FourierUnit :9-34) vs the reference's own
    SwinFIR forward (fixtures: oracle/make_golden_swinfir.py)."""
    c = golden_meta[name]
    P = synth.swinfir_weights(c["cfg"], c["wseed"])
    x = synth.image_batch(c["shape"], c["xseed"])
    y = O.swinir_forward(P, x, c["cfg"], training=c["training"])
    ref = torch.from_numpy(load_golden(name)["y"])
    assert list(y.shape) == c["out_shape"]
    assert (y - ref).abs().max().item() <= 2e-5


@pytest.mark.parametrize("name", EDSR_CASES)
def test_edsr_oracle_matches_reference_golden(name, golden_meta):
    c = golden_meta[name]
    P = synth.edsr_weights(c["cfg"], c["wseed"])
    x = synth.image_batch(c["shape"], c["xseed"])
    y = O.edsr_forward(P, x, c["cfg"])
    ref = torch.from_numpy(load_golden(name)["y"])
    assert (y - ref).abs().max().item() < 2e-5


def test_ops_match_reference(golden_meta):
    g = load_golden("swinir_ops")
    c = golden_meta["swinir_ops"]
    assert np.array_equal(O.shift_mask(24, 32, 8, 4, torch.float32).numpy(), g["mask_24x32_ws8_s4"])
    assert np.array_equal(O.shift_mask(16, 16, 8, 0, torch.float32).numpy(), g["mask_16x16_ws8_s0"])
    assert np.array_equal(synth.relative_position_index(8).numpy(), g["rpi_ws8"])
    xpad = synth.image_batch((1, 3, 13, 16), c["pad_seed"])
    assert np.array_equal(O.pad_for_eval(xpad, 8).numpy(), g["pad_eval_13x16"])
    assert np.array_equal(O.pad_for_train(xpad, 8).numpy(), g["pad_train_13x16"])
    P = synth.swinir_weights(c["cfg"], c["wseed"])
    xb = torch.randn(2, 16, 24, 60, generator=torch.Generator().manual_seed(c["xb_seed"]))
    yb = O.swin_block(P, c["block"], xb, 6, 8, 4)
    assert np.abs(yb.numpy() - g["block_shift4_out"]).max() < 1e-5
    xw = torch.randn(6, 64, 60, generator=torch.Generator().manual_seed(c["xw_seed"]))
    a = O.window_attention(P, c["block"] + ".attn", xw, 6, 8, O.shift_mask(16, 24, 8, 4, torch.float32))
    assert np.abs(a.numpy() - g["winattn_masked_out"]).max() < 1e-5
    a = O.window_attention(P, c["block"] + ".attn", xw, 6, 8, None)
    assert np.abs(a.numpy() - g["winattn_nomask_out"]).max() < 1e-5


def test_inference_u8_matches_reference(golden_meta):
    g = load_golden("swinir_tiny_x4_inference_u8")
    cfg = golden_meta["swinir_ops"]["cfg"]
    P = synth.swinir_weights(cfg, 11)
    out = O.inference_u8(lambda x: O.swinir_forward(P, x, cfg), g["img"])
    # uint8 rounding can flip on 1e-6 float noise: allow a handful of +-1 LSB
    d = np.abs(out.astype(np.int32) - g["out"].astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 1e-3


def test_fp64_oracle_close_to_fp32_golden(golden_meta):
    name = "swinir_tiny_x4_eval_2x20x28"
    c = golden_meta[name]
    P = synth.swinir_weights(c["cfg"], c["wseed"])
    x = synth.image_batch(c["shape"], c["xseed"]).double()
    y = O.swinir_forward(P, x, c["cfg"])
    ref = torch.from_numpy(load_golden(name)["y"]).double()
    assert (y - ref).abs().max().item() < 2e-5


def test_tiler_identity_and_tile_starts():
    assert O.tile_starts(960, 64, 48) == list(range(0, 896, 48)) + [896]
    assert len(O.tile_starts(960, 64, 48)) == 20 and len(O.tile_starts(540, 64, 48)) == 11
    assert O.tile_starts(50, 64, 48) == [0]
    # a forward that is a pure x4 nearest upsample must be reproduced exactly by the blend
    x = synth.image_batch((1, 3, 100, 130), 5)
    up = lambda t: t.repeat_interleave(4, 2).repeat_interleave(4, 3)
    y = O.tiled_upscale(up, x, 4, tile=64, overlap=16)
    assert (y - up(x)).abs().max().item() < 1e-6


GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


# ---- training step: the oracle's autograd against gradients produced by the reference's own loss.backward() ----
def _train_meta():
    with open(os.path.join(GOLD, "meta_train.json")) as f:
        return json.load(f)["cases"]


@pytest.mark.parametrize("name", sorted(_train_meta().keys()))
def test_oracle_backward_matches_reference_gradients(name):
    import torch.nn.functional as F

    c = _train_meta()[name]
    cfg, shape = c["cfg"], tuple(c["shape"])
    gold = np.load(os.path.join(GOLD, name + ".npz"))
    x = synth.image_batch(shape, c["xseed"])
    tgt = synth.image_batch((shape[0], 3, shape[2] * cfg["scale"], shape[3] * cfg["scale"]), c["xseed"] + 1)
    if c["arch"] == "edsr":
        P = synth.edsr_weights(cfg, c["wseed"])
        Q = {k: v.clone().requires_grad_(v.is_floating_point() and "mean" not in k) for k, v in P.items()}
        loss = F.l1_loss(O.edsr_forward(Q, x, cfg), tgt)
    elif c["arch"] in ("rcan", "han"):
        P = (synth.rcan_weights if c["arch"] == "rcan" else synth.han_weights)(cfg, c["wseed"])
        Q = {k: v.clone().requires_grad_(v.is_floating_point() and "mean" not in k) for k, v in P.items()}
        loss = F.l1_loss((O.rcan_forward if c["arch"] == "rcan" else O.han_forward)(Q, x, cfg), tgt)
    else:
        P = synth.swinir_weights(cfg, c["wseed"])
        Q = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in P.items()}
        drop = None
        if c["arch"] == "swinir_dp":  # the drop-in module draws the masks from torch's generator exactly as the reference does
            from studiosr_b200.models import SwinIR

            kw = {k: cfg[k] for k in ("scale", "n_colors", "img_range", "embed_dim", "depths", "num_heads", "window_size",
                                      "mlp_ratio", "upsampler")}
            mod = SwinIR(drop_path_rate=c["drop_path_rate"], **kw)  # (construction consumes the generator: seed afterwards)
            torch.manual_seed(c["drop_seed"])
            drop = mod._draw_drop_path(shape[0], torch.device("cpu"))
            assert (drop == 0).any() and (drop > 1).any()
        loss = F.l1_loss(O.swinir_forward(Q, x, cfg, training=True, drop_scale=drop), tgt)
    loss.backward()
    assert abs(loss.item() - float(gold["loss"][0])) < 1e-5
    checked = 0
    for k, v in Q.items():
        if k + "::norm" not in gold.files:
            continue
        g = v.grad.flatten()
        ref_norm = float(gold[k + "::norm"][0])
        assert abs(g.norm().item() - ref_norm) <= 2e-4 * ref_norm + 1e-9, k
        ref = torch.from_numpy(gold[k + "::sample"])
        assert (g[::c["stride"]] - ref).norm().item() <= 2e-4 * ref.norm().item() + 1e-9, k
        checked += 1
    assert checked == c["n_params"]


def test_oracle_psnr_and_augmentation_match_reference_goldens():
    """oracle.compute_psnr / augment_pair against values and patches produced by the unmodified reference
    (oracle/make_golden_data.py): utils/metrics.py:11-49, data/transforms.py:8-68 with dataset.py:50-58's transform order."""
    import random

    import numpy as np

    g = load_golden("data_ops")
    gt, sr, big = g["psnr_gt"], g["psnr_sr"], g["psnr_big"]
    vals = []
    for a, b in ((sr, gt), (big, gt)):
        for y_only in (False, True):
            for cb in (0, 4):
                vals.append(O.compute_psnr(a, b, y_only=y_only, crop_border=cb))
    assert np.allclose(vals, g["psnr_values"], rtol=0, atol=1e-5)
    assert O.compute_psnr(gt, gt) == float("inf") and np.isinf(g["psnr_identical"][0])
    # the host-side draw sequence of studiosr_b200.data.PairedAugment is the reference's: seeded runs cut the same patches
    from studiosr_b200.data import PairedAugment

    lq, hr = g["aug_lq"], g["aug_gt"]
    for i, seed in enumerate(g["aug_seeds"]):
        rng = random.Random(int(seed))
        xs, ys, flags = PairedAugment(size=12, scale=4, rng=rng, device="cpu").draw(lq.shape[0], lq.shape[1])
        x, y = O.augment_pair(lq, hr, 12, 4, xs, ys, flags)
        assert np.array_equal(x, g["aug_x"][i]) and np.array_equal(y, g["aug_y"][i]), seed


def _dft_pair_as_k_fft(y: np.ndarray):
    """numpy restatement of the four passes of studiosr_b200/csrc/k_fft.cu (direct DFTs with a twiddle table indexed by
    k * x mod L, norm "ortho", c2r ignoring the imaginary parts of the DC / Nyquist bins): y [H, W] -> (spectrum, round trip)."""
    H, W = y.shape
    Wf = W // 2 + 1
    sc = 1.0 / np.sqrt(H * W)
    twW = np.exp(-2j * np.pi * np.arange(W) / W)
    twH = np.exp(-2j * np.pi * np.arange(H) / H)
    a = np.zeros((H, Wf), dtype=np.complex128)
    for k in range(Wf):  # pass 1: real -> complex along W
        a[:, k] = (y * twW[(k * np.arange(W)) % W][None, :]).sum(axis=1) * sc
    spec = np.zeros_like(a)
    for k in range(H):  # pass 2: complex along H
        spec[k] = (a * twH[(k * np.arange(H)) % H][:, None]).sum(axis=0)
    b = np.zeros_like(spec)
    for h in range(H):  # pass 3: inverse along H
        b[h] = (spec * np.conj(twH[(h * np.arange(H)) % H])[:, None]).sum(axis=0)
    out = np.zeros((H, W))
    for x in range(W):  # pass 4: complex -> real along W
        acc = b[:, 0].real.copy()
        for k in range(1, Wf):
            t = np.conj(twW[(k * x) % W])
            if 2 * k == W:
                acc += b[:, k].real * t.real
            else:
                acc += 2.0 * (b[:, k].real * t.real - b[:, k].imag * t.imag)
        out[:, x] = acc * sc
    return spec, out


@pytest.mark.parametrize("H,W", [(8, 8), (16, 24), (9, 12), (12, 7), (5, 5)])
def test_fft_algorithm_of_k_fft_matches_torch_fft(H, W):
    """The algorithm of the SwinFIR FFT kernels (even and odd lengths) against torch.fft.rfftn / irfftn with norm="ortho" -- the
    calls swinfir.py:20,31 makes.  (The kernels themselves are checked on the GPU through the SwinFIR goldens, whose padded
    sizes are always multiples of 8; this pins the odd-length and Nyquist handling of the formulae.)"""
    rng = np.random.default_rng(H * 100 + W)
    y = rng.standard_normal((H, W))
    spec, back = _dft_pair_as_k_fft(y)
    ref = torch.fft.rfftn(torch.from_numpy(y), dim=(-2, -1), norm="ortho").numpy()
    assert np.abs(spec - ref).max() < 1e-12
    assert np.abs(back - y).max() < 1e-12
    # a spectrum that is NOT Hermitian-consistent (what the 1x1 conv + LeakyReLU of the FourierUnit produces): same answer as irfftn
    z = rng.standard_normal((H, W // 2 + 1)) + 1j * rng.standard_normal((H, W // 2 + 1))
    Hh, Wf = z.shape
    sc = 1.0 / np.sqrt(H * W)
    twW = np.exp(-2j * np.pi * np.arange(W) / W)
    twH = np.exp(-2j * np.pi * np.arange(H) / H)
    b = np.stack([(z * np.conj(twH[(h * np.arange(H)) % H])[:, None]).sum(axis=0) for h in range(H)])
    out = np.zeros((H, W))
    for x in range(W):
        acc = b[:, 0].real.copy()
        for k in range(1, Wf):
            t = np.conj(twW[(k * x) % W])
            acc += b[:, k].real * t.real if 2 * k == W else 2.0 * (b[:, k].real * t.real - b[:, k].imag * t.imag)
        out[:, x] = acc * sc
    ref2 = torch.fft.irfftn(torch.from_numpy(z), s=(H, W), dim=(-2, -1), norm="ortho").numpy()
    assert np.abs(out - ref2).max() < 1e-11


def test_han_lam_backward_formulae_match_autograd():
    """The closed-form adjoint of HAN's layer attention used by k_simt.cu (han_gram2 / han_lam_bwd_small / han_lam_bwd_apply):
    with D = <dOut_n, X_m>, A = softmax(max - E), E = X X^T:  dgamma = sum A.D;  dE = -A.(gamma D - rowsum(A.gamma D))  (the row-max
    term cancels);  dX = dOut + gamma A^T dOut + (dE + dE^T) X  -- against torch autograd over oracle.han_lam."""
    torch.manual_seed(3)
    B, N, C, H, W = 2, 11, 4, 3, 5
    x = (torch.randn(B, N, C, H, W, dtype=torch.float64) * 0.3).requires_grad_(True)
    gamma = torch.tensor([0.7], dtype=torch.float64, requires_grad=True)
    out = O.han_lam(x, gamma)
    g = torch.randn_like(out)
    out.backward(g)
    X = x.detach().reshape(B, N, -1)
    G = g.reshape(B, N, -1)
    E = X @ X.transpose(1, 2)
    A = torch.softmax(E.max(dim=-1, keepdim=True)[0] - E, dim=-1)
    D = G @ X.transpose(1, 2)
    dgamma = (A * D).sum()
    dA = gamma.detach() * D
    dE = -A * (dA - (A * dA).sum(dim=-1, keepdim=True))
    dX = G + gamma.detach() * (A.transpose(1, 2) @ G) + (dE + dE.transpose(1, 2)) @ X
    assert torch.allclose(dX.reshape(x.shape), x.grad, atol=1e-10)
    assert abs(dgamma.item() - gamma.grad.item()) < 1e-10


def test_han_csam_backward_formulae_match_autograd():
    """The adjoint of HAN's channel-spatial attention used by k_simt.cu (han_csam_bwd1 / bwd2): dpre = g x gamma s (1 - s);
    dx = g (1 + gamma s) + conv3d^T(dpre); dgamma = sum g x s; db = sum dpre; dW[tap] = sum dpre x[. + off(tap)]."""
    torch.manual_seed(5)
    B, C, H, W = 2, 6, 4, 5
    x = torch.randn(B, C, H, W, dtype=torch.float64, requires_grad=True)
    P = {"csa.conv.weight": (torch.randn(1, 1, 3, 3, 3, dtype=torch.float64) * 0.3).requires_grad_(True),
         "csa.conv.bias": torch.tensor([0.1], dtype=torch.float64, requires_grad=True),
         "csa.gamma": torch.tensor([0.7], dtype=torch.float64, requires_grad=True)}
    out = O.han_csam(P, x)
    g = torch.randn_like(out)
    out.backward(g)
    xd, Wt, gm = x.detach(), P["csa.conv.weight"].detach(), P["csa.gamma"].detach()
    pre = torch.nn.functional.conv3d(xd.unsqueeze(1), Wt, P["csa.conv.bias"].detach(), padding=1).squeeze(1)
    sg = torch.sigmoid(pre)
    dpre = g * xd * gm * sg * (1 - sg)
    dx = g * (1 + gm * sg) + torch.nn.functional.conv_transpose3d(dpre.unsqueeze(1), Wt, padding=1).squeeze(1)
    assert torch.allclose(dx, x.grad, atol=1e-10)
    assert abs((g * xd * sg).sum().item() - P["csa.gamma"].grad.item()) < 1e-9
    assert abs(dpre.sum().item() - P["csa.conv.bias"].grad.item()) < 1e-9
    xp = torch.nn.functional.pad(xd, (1, 1, 1, 1, 1, 1))
    dW = torch.stack([(dpre * xp[:, dc:dc + C, dy:dy + H, dx_:dx_ + W]).sum() for dc in range(3) for dy in range(3) for dx_ in range(3)])
    assert torch.allclose(dW, P["csa.conv.weight"].grad.flatten(), atol=1e-9)


def test_channel_attention_backward_formulae_match_autograd():
    """The adjoint of out = res + t * sigmoid(W2 relu(W1 mean(t) + b1) + b2) used by k_simt.cu (ca_bwd_reduce / gate / apply):
    dz2 = gate (1 - gate) sum_hw(G t);  dz1 = relu'(.) W2^T dz2;  dt = G gate + W1^T dz1 / HW;  dW2 = dz2 h^T, dW1 = dz1 p^T."""
    torch.manual_seed(7)
    B, C, R, H, W = 2, 8, 2, 3, 4
    t = torch.randn(B, C, H, W, dtype=torch.float64, requires_grad=True)
    P = {"ca.conv_du.0.weight": torch.randn(R, C, 1, 1, dtype=torch.float64, requires_grad=True),
         "ca.conv_du.0.bias": torch.randn(R, dtype=torch.float64, requires_grad=True),
         "ca.conv_du.2.weight": torch.randn(C, R, 1, 1, dtype=torch.float64, requires_grad=True),
         "ca.conv_du.2.bias": torch.randn(C, dtype=torch.float64, requires_grad=True)}
    out = O.channel_attention(P, "ca", t)
    G = torch.randn_like(out)
    out.backward(G)
    W1, b1 = P["ca.conv_du.0.weight"].detach().reshape(R, C), P["ca.conv_du.0.bias"].detach()
    W2, b2 = P["ca.conv_du.2.weight"].detach().reshape(C, R), P["ca.conv_du.2.bias"].detach()
    td = t.detach()
    p = td.mean(dim=(2, 3))
    z1 = p @ W1.t() + b1
    h = torch.relu(z1)
    gate = torch.sigmoid(h @ W2.t() + b2)
    ds = (G * td).sum(dim=(2, 3))
    dz2 = ds * gate * (1 - gate)
    dz1 = (dz2 @ W2) * (z1 > 0)
    dt = G * gate[:, :, None, None] + (dz1 @ W1)[:, :, None, None] / (H * W)
    assert torch.allclose(dt, t.grad, atol=1e-10)
    assert torch.allclose(dz2.t() @ h, P["ca.conv_du.2.weight"].grad.reshape(C, R), atol=1e-10)
    assert torch.allclose(dz1.t() @ p, P["ca.conv_du.0.weight"].grad.reshape(R, C), atol=1e-10)
    assert torch.allclose(dz2.sum(0), P["ca.conv_du.2.bias"].grad, atol=1e-10) and torch.allclose(dz1.sum(0), P["ca.conv_du.0.bias"].grad, atol=1e-10)


@pytest.mark.parametrize("h,w", [(16, 24), (20, 27), (9, 8)])
def test_input_gradient_formula_with_reflect_pad_matches_autograd(h, w):
    """dL/dx as conv_first_dgrad_kernel computes it (k_train.cu): the first conv's data gradient on the padded grid,
    dxp[ci][y][x] = sum_{ky,kx,n} G[y - ky + 1][x - kx + 1][n] W[n][ci][ky][kx], folded back through the training-mode reflect pad
    (padded (y, x) reads source (y < h ? y : 2 (h - 1) - y, same for x), common.py:277-282) and scaled by 1 / img_range."""
    torch.manual_seed(h * 31 + w)
    C, ws, img_range = 5, 8, 2.0
    x = torch.rand(1, 3, h, w, dtype=torch.float64, requires_grad=True)
    Wc = torch.randn(C, 3, 3, 3, dtype=torch.float64)
    mean = torch.tensor(synth.RGB_MEAN, dtype=torch.float64).view(1, 3, 1, 1)
    xp = O.pad_for_train(x, ws) / img_range - mean
    y = F.conv2d(xp, Wc, padding=1)
    G = torch.randn_like(y)
    y.backward(G)
    Hp, Wp = y.shape[2:]
    dx = torch.zeros(3, h, w, dtype=torch.float64)
    Gp = F.pad(G[0], (1, 1, 1, 1))
    for yy in range(Hp):
        sy = yy if yy < h else 2 * (h - 1) - yy
        for xx in range(Wp):
            sx = xx if xx < w else 2 * (w - 1) - xx
            acc = torch.zeros(3, dtype=torch.float64)
            for ky in range(3):
                for kx in range(3):
                    acc += Gp[:, yy - ky + 1 + 1, xx - kx + 1 + 1] @ Wc[:, :, ky, kx]
            dx[:, sy, sx] += acc / img_range
    assert torch.allclose(dx, x.grad[0], atol=1e-10)


def test_compact_relative_position_bias_addressing():
    """The addressing the fused attention kernel uses for the relative-position bias (k_swin_attn.cu + pack_attn_fused_host):
    window tokens in the order r = (tx / 4) * 32 + ty * 4 + tx % 4; B[i][j] = R[7 - yi + yj][7 - xi + xj] with R the 15 x 15 table
    reversed in both axes, stored in two copies shifted by (7 - xi) % 2 elements at a row pitch of 16, read as 4-element runs
    (one per key quad) -- against the reference gather table[relative_position_index] (swinir.py:57-67, 92-95)."""
    ws, heads = 8, 6
    g = torch.Generator().manual_seed(1)
    table = torch.randn((2 * ws - 1) ** 2, heads, generator=g)
    ref = O.rel_pos_bias(table, ws)  # [heads, 64, 64] in row-major window-token order t = ty * 8 + tx
    copies = np.zeros((heads, 2, 15, 16), dtype=np.float32)
    for hd in range(heads):
        for sft in range(2):
            for a in range(15):
                for b in range(15):
                    copies[hd, sft, a, b + sft] = table[(14 - a) * 15 + (14 - b), hd]

    def tok(r):  # kernel token index -> (y, x) inside the window
        return (r & 31) >> 2, (r >> 5) * 4 + (r & 3)

    for hd in range(heads):
        for ri in range(64):
            yi, xi = tok(ri)
            s_ = (7 - xi) & 1
            row0, col0 = 7 - yi, 7 - xi + s_
            assert col0 % 2 == 0  # every run starts 4-byte aligned (bf16)
            got = np.zeros(64, dtype=np.float32)
            for c in range(8):  # keys 8c .. 8c+7: key rows (2c) % 8 and + 1, key columns 4 (c / 4) .. + 3
                for u in range(2):
                    run = copies[hd, s_, row0 + ((2 * c) & 7) + u, col0 + 4 * (c >> 2): col0 + 4 * (c >> 2) + 4]
                    got[8 * c + 4 * u: 8 * c + 4 * u + 4] = run
            for rj in range(64):
                yj, xj = tok(rj)
                assert got[rj] == ref[hd, yi * 8 + xi, yj * 8 + xj].item(), (hd, ri, rj)
