"""-m gpu: whole-model parity of the drop-in modules (-> C ABI -> CUDA kernels) against the
golden fixtures produced by the reference and against the oracle restatement.

Tolerances (BASELINE.json north_star): fp32 output max-abs <= 1e-3 (the CUDA-core fp32 path is
held to 1e-4, the tf32 tensor-core path to 1e-3).  bf16 is bounded two ways, both of which can fail:
  (1) against the UNMODIFIED REFERENCE's own error under stock bf16 autocast on the same weights and
      inputs (tests/golden/bf16_autocast_ref.json, made by oracle/make_golden_bf16_ref.py): rms <= 1.0x,
      max-abs <= 1.5x -- the tensor-core path must be at least as accurate as what the reference's
      users get from `torch.autocast(bf16)` (measured 0.45 - 0.8x);
  (2) the PSNR delta of the uint8-quantised output versus the reference's fp32 output against a
      REALISTIC synthetic ground truth, gt = clip(ref + N(0, 0.025)) (PSNR(ref, gt) ~ 32-33 dB; a
      uniform-random gt sits at 8 dB where no error moves the PSNR).  Measured on these random-init
      models: 0.005 - 0.07 dB (0.045 dB on the cfg1 golden; RCAN-full 0.26 dB; the reference's own
      autocast lands at ~2.7x ours), i.e. the north-star 0.01 dB is NOT met with random-init weights,
      where the body's output is O(1) instead of a small residual (it needs rms <= 1.2e-3).  The test
      holds the delta to the value implied by bound (1): 10 log10(1 + rms_ref^2 / sigma^2)."""
import numpy as np
import pytest
import torch

from oracle import sr_oracle as O
from oracle import synth
from tests.conftest import load_golden

pytestmark = pytest.mark.gpu

SWINIR_CASES = [
    "swinir_tiny_x4_eval_2x20x28", "swinir_tiny_x4_eval_1x16x16", "swinir_tiny_x4_train_1x12x12",
    "swinir_tiny_x4_train_2x16x24", "swinir_tiny_x2_eval_1x12x12", "swinir_tiny_x3_eval_1x8x8",
    "swinir_tiny_x8_eval_1x8x8", "swinir_full_x4_eval_cfg1", "swinir_light_x4_eval_1x12x20",
]
EDSR_CASES = ["edsr_tiny_x4_2x12x20", "edsr_tiny_x2_1x9x11", "edsr_tiny_x3_1x8x8", "edsr_full_x4_1x24x24"]
# "fp32" (CUDA-core FMA) carries the north-star fp32 claim (<= 1e-3; held to 1e-4 here).  "tf32" is an
# opt-in tensor-core mode whose single-pass 10-bit-mantissa operands land at ~1e-3 (measured 9.5e-4 ..
# 1.1e-3 on these cases), i.e. AT the fp32 tolerance, so it is held to 2e-3 and never used for that claim.
# "tf32x3" (three tf32 MMAs per k-step on head / tail splits of both operands) is the tensor-core mode that carries the fp32 claim
# too: held to 2e-4 like a true fp32 path (measured ~1e-5).
ABS_TOL = {"fp32": 1e-4, "tf32": 2e-3, "tf32x3": 2e-4}


def _swinir(cfg, wseed):
    from studiosr_b200.models import SwinIR

    kw = {k: cfg[k] for k in ("scale", "n_colors", "img_range", "embed_dim", "depths", "num_heads", "window_size",
                              "mlp_ratio", "upsampler")}
    m = SwinIR(drop_path_rate=0.0, **kw)
    m.load_state_dict(synth.swinir_weights(cfg, wseed), strict=True)
    return m.cuda()


def _edsr(cfg, wseed):
    from studiosr_b200.models import EDSR

    m = EDSR(**cfg)
    m.load_state_dict(synth.edsr_weights(cfg, wseed), strict=True)
    return m.cuda().eval()


GT_SIGMA = 0.025  # noise of the synthetic ground truth around the reference output: PSNR(ref, gt) ~ 32 dB


def _psnr_delta(y, ref, seed=99, draws=8):
    """|PSNR(q(y), gt) - PSNR(q(ref), gt)| with the reference's uint8 quantisation (common.py:44-45) and its PSNR
    (utils/metrics.py:36-49) against gt = clip(ref + N(0, GT_SIGMA)), averaged over `draws` ground truths (on the small
    golden outputs the error / noise cross term of a single draw is as large as the effect being measured)."""
    q = lambda t: (t * 255.0).round().clip(0, 255)
    d = 0.0
    for i in range(draws):
        gt = (ref + GT_SIGMA * torch.randn(ref.shape, generator=torch.Generator().manual_seed(seed + i))).clip(0, 1)
        d += O.psnr(q(y), q(gt)) - O.psnr(q(ref), q(gt))
    return abs(d / draws)


_BF16_REF = None


def _check_bf16(name, y, ref):
    """The two bf16 bounds of the module docstring."""
    global _BF16_REF
    if _BF16_REF is None:
        import json
        import os

        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bf16_autocast_ref.json")) as f:
            _BF16_REF = json.load(f)["cases"]
    r = _BF16_REF[name]
    e = (y - ref).double()
    max_abs, rms = e.abs().max().item(), e.pow(2).mean().sqrt().item()
    assert rms <= 1.0 * r["rms"], f"{name} [bf16] rms {rms:.3e} > the reference's own bf16-autocast rms {r['rms']:.3e}"
    assert max_abs <= 1.5 * r["max_abs"], f"{name} [bf16] max-abs {max_abs:.3e} vs reference autocast {r['max_abs']:.3e}"
    import math

    bound = 10.0 * math.log10(1.0 + (r["rms"] / GT_SIGMA) ** 2) + 0.005  # + uint8 quantisation noise of a few mdB
    d = _psnr_delta(y, ref)
    assert d <= bound, f"{name} [bf16] PSNR delta {d:.4f} dB > {bound:.4f} dB (reference-autocast level)"


@pytest.mark.parametrize("prec", ["fp32", "tf32", "tf32x3", "bf16"])
@pytest.mark.parametrize("name", SWINIR_CASES)
def test_swinir_matches_reference_golden(name, prec, golden_meta):
    c = golden_meta[name]
    m = _swinir(c["cfg"], c["wseed"])
    m.train(c["training"])
    m.precision = prec
    x = synth.image_batch(c["shape"], c["xseed"]).cuda()
    with torch.no_grad():
        y = m(x).float().cpu()
    ref = torch.from_numpy(load_golden(name)["y"])
    assert list(y.shape) == c["out_shape"]
    err = (y - ref).abs().max().item()
    if prec in ABS_TOL:
        assert err <= ABS_TOL[prec], f"{name} [{prec}] max-abs {err:.3e}"
    else:
        _check_bf16(name, y, ref)


@pytest.mark.parametrize("prec", ["fp32", "tf32", "tf32x3", "bf16"])
@pytest.mark.parametrize("name", EDSR_CASES)
def test_edsr_matches_reference_golden(name, prec, golden_meta):
    c = golden_meta[name]
    m = _edsr(c["cfg"], c["wseed"])
    m.precision = prec
    x = synth.image_batch(c["shape"], c["xseed"]).cuda()
    with torch.no_grad():
        y = m(x).float().cpu()
    ref = torch.from_numpy(load_golden(name)["y"])
    err = (y - ref).abs().max().item()
    if prec in ABS_TOL:
        assert err <= ABS_TOL[prec], f"{name} [{prec}] max-abs {err:.3e}"
    else:
        _check_bf16(name, y, ref)


HAT_CASES = ["hat_tiny_x4_eval_2x20x40", "hat_tiny_x4_train_1x32x32", "hat_tiny_x2_eval_1x16x48", "hat_tiny_x3_eval_1x17x17",
             "hat_full_x4_eval_1x64x64"]


@pytest.mark.parametrize("prec", ["fp32", "tf32", "tf32x3", "bf16"])
@pytest.mark.parametrize("name", HAT_CASES)
def test_hat_matches_reference_golden(name, prec, golden_meta):
    """HAT forward (hat.py:542-554: HAB with channel attention, 16x16 (S)W-MSA, overlapping cross-attention) against the
    reference's own outputs; the last case is the full BASELINE.json config-3 model on one 64x64 tile."""
    from studiosr_b200.models import HAT

    c = golden_meta[name]
    m = HAT(drop_path_rate=0.0, **c["cfg"])
    m.load_state_dict(synth.hat_weights(c["cfg"], c["wseed"]), strict=True)
    m = m.cuda()
    m.train(c["training"])
    m.precision = prec
    x = synth.image_batch(c["shape"], c["xseed"]).cuda()
    with torch.no_grad():
        y = m(x).float().cpu()
    ref = torch.from_numpy(load_golden(name)["y"])
    assert list(y.shape) == c["out_shape"]
    err = (y - ref).abs().max().item()
    if prec in ABS_TOL:
        assert err <= ABS_TOL[prec], f"{name} [{prec}] max-abs {err:.3e}"
    else:
        _check_bf16(name, y, ref)


RCAN_CASES = ["rcan_tiny_x4_2x12x20", "rcan_tiny_x2_1x9x11", "rcan_tiny_x3_1x8x8", "rcan_full_x4_1x24x24"]


@pytest.mark.parametrize("prec", ["fp32", "tf32", "tf32x3", "bf16"])
@pytest.mark.parametrize("name", RCAN_CASES)
def test_rcan_matches_reference_golden(name, prec, golden_meta):
    """RCAN forward (rcan.py:68-77; 10 x 20 RCABs with channel attention) against the reference's own outputs."""
    from studiosr_b200.models import RCAN

    c = golden_meta[name]
    m = RCAN(**c["cfg"])
    m.load_state_dict(synth.rcan_weights(c["cfg"], c["wseed"]), strict=True)
    m = m.cuda().eval()
    m.precision = prec
    x = synth.image_batch(c["shape"], c["xseed"]).cuda()
    with torch.no_grad():
        y = m(x).float().cpu()
    ref = torch.from_numpy(load_golden(name)["y"])
    assert list(y.shape) == c["out_shape"]
    err = (y - ref).abs().max().item()
    if prec in ABS_TOL:
        tol = ABS_TOL[prec] * (4 if "full" in name else 1)  # 400 convs deep: summation-order noise accumulates
        assert err <= tol, f"{name} [{prec}] max-abs {err:.3e}"
    else:
        _check_bf16(name, y, ref)


HAN_CASES = ["han_tiny_x4_2x12x20", "han_tiny_x2_1x9x11", "han_tiny_x3_1x8x8", "han_full_x4_1x16x16"]


@pytest.mark.parametrize("prec", ["fp32", "tf32", "tf32x3", "bf16"])
@pytest.mark.parametrize("name", HAN_CASES)
def test_han_matches_reference_golden(name, prec, golden_meta):
    """HAN forward (han.py:90-113: the RCAN trunk, layer attention over its 11 outputs, channel-spatial attention, last_conv /
    last) against the reference's own outputs."""
    from studiosr_b200.models import HAN

    c = golden_meta[name]
    m = HAN(**c["cfg"])
    m.load_state_dict(synth.han_weights(c["cfg"], c["wseed"]), strict=True)
    m = m.cuda().eval()
    m.precision = prec
    x = synth.image_batch(c["shape"], c["xseed"]).cuda()
    with torch.no_grad():
        y = m(x).float().cpu()
    ref = torch.from_numpy(load_golden(name)["y"])
    assert list(y.shape) == c["out_shape"]
    err = (y - ref).abs().max().item()
    if prec in ABS_TOL:
        tol = ABS_TOL[prec] * (4 if "full" in name else 1)  # 400 convs deep: summation-order noise accumulates
        assert err <= tol, f"{name} [{prec}] max-abs {err:.3e}"
    else:  # no stored reference-autocast error for HAN: bf16 is held to RCAN's level on the same trunk
        rms = (y - ref).pow(2).mean().sqrt().item()
        assert err <= (1.2e-1 if "full" in name else 4e-2) and rms <= (2.5e-2 if "full" in name else 8e-3), (name, err, rms)


def test_han_inference_u8_and_train_mode_forward():
    """The reference-facing entry points on HAN: Model.inference (uint8 in / out) equals the quantised fp32 forward, and the
    train-mode forward + backward run."""
    from studiosr_b200.models import HAN

    cfg = synth.HAN_TINY
    m = HAN(**cfg)
    m.load_state_dict(synth.han_weights(cfg, 41), strict=True)
    m = m.cuda().eval()
    img = synth.smooth_image_u8(20, 28, seed=5)
    out = m.inference(img)
    x = torch.from_numpy(img.astype(np.float32) / 255.0).permute(2, 0, 1).unsqueeze(0)
    ref = O.quantize_u8(O.han_forward(synth.han_weights(cfg, 41), x, cfg)[0], 1.0).numpy()
    d = np.abs(out.astype(np.int32) - ref.astype(np.int32))
    assert out.shape == (80, 112, 3) and d.max() <= 1 and (d > 0).mean() < 2e-3
    m.train()
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = m(x.cuda())
    assert y.shape == (1, 3, 80, 112)
    y.sum().backward()  # the native backward (tests/test_gpu_train.py::test_han_backward checks the values)
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in m.parameters() if p.requires_grad)


SWINFIR_CASES = ["swinfir_tiny_x4_eval_1x12x20", "swinfir_tiny_x2_eval_2x16x16", "swinfir_tiny_x4_train_1x16x24", "swinfir_c180_x4_eval_1x8x8"]
_SWIN_KW = ("scale", "n_colors", "img_range", "embed_dim", "depths", "num_heads", "window_size", "mlp_ratio", "upsampler")


@pytest.mark.parametrize("prec", ["fp32", "tf32", "tf32x3"])
@pytest.mark.parametrize("name", SWINFIR_CASES)
def test_swinfir_matches_reference_golden(name, prec, golden_meta):
    """SwinFIR forward (swinfir.py:83-114: SFB = spatial convs + FourierUnit + fusion in every RSTB and as conv_after_body; the
    2-D real FFT pair is k_fft.cu) against the reference's own outputs."""
    from studiosr_b200.models import SwinFIR

    c = golden_meta[name]
    m = SwinFIR(drop_path_rate=0.0, **{k: c["cfg"][k] for k in _SWIN_KW})
    m.load_state_dict(synth.swinfir_weights(c["cfg"], c["wseed"]), strict=True)
    m = m.cuda().train(c["training"])
    m.precision = prec
    x = synth.image_batch(c["shape"], c["xseed"]).cuda()
    with torch.no_grad():
        y = m(x).float().cpu()
    ref = torch.from_numpy(load_golden(name)["y"])
    assert list(y.shape) == c["out_shape"]
    err = (y - ref).abs().max().item()
    assert err <= ABS_TOL[prec], f"{name} [{prec}] max-abs {err:.3e}"


def test_swinfir_reference_shape_test_and_bf16_refusal():
    """tests/models/test_swinfir.py of the reference on the drop-in (default full-size model, every scale), and the bf16 mode
    refuses loudly (the reference trains and ships SwinFIR in fp32, swinfir.py:126)."""
    from studiosr_b200.models import SwinFIR

    for scale in (2, 3, 4, 8):
        model = SwinFIR(scale=scale, n_colors=3).cuda()
        for hw in (8, 12):
            y = model(torch.randn(1, 3, hw, hw).cuda())
            assert y.shape == (1, 3, scale * hw, scale * hw) and torch.isfinite(y).all()
    with pytest.raises(NotImplementedError, match="fp32-class"):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            model(torch.randn(1, 3, 8, 8).cuda())


def test_f3_families_through_inference_and_tiler():
    """HAN and SwinFIR behind the caller-facing entry points: Model.inference (uint8 in / out) and the tiler equal the oracle's
    forward / tiler on the same weights (<= 1 LSB on a small share of pixels)."""
    from studiosr_b200.models import HAN, SwinFIR

    img = synth.smooth_image_u8(40, 52, seed=9)
    x = torch.from_numpy(img.astype(np.float32) / 255.0).permute(2, 0, 1).unsqueeze(0)
    cfg_h = dict(synth.HAN_TINY, scale=2)
    Ph = synth.han_weights(cfg_h, 42)
    han = HAN(**cfg_h)
    han.load_state_dict(Ph, strict=True)
    han = han.cuda().eval()
    cfg_s = dict(synth.swinir_config(**synth.SWINIR_TINY), scale=2, sfb=True)
    Ps = synth.swinfir_weights(cfg_s, 22)
    fir = SwinFIR(drop_path_rate=0.0, **{k: cfg_s[k] for k in _SWIN_KW})
    fir.load_state_dict(Ps, strict=True)
    fir = fir.cuda().eval()
    for name, m, fwd in (("han", han, lambda t: O.han_forward(Ph, t, cfg_h)), ("swinfir", fir, lambda t: O.swinir_forward(Ps, t, cfg_s))):
        m.precision = "fp32"
        ref = O.quantize_u8(fwd(x)[0], 1.0).numpy()
        d = np.abs(m.inference(img).astype(np.int32) - ref.astype(np.int32))
        assert d.max() <= 1 and (d > 0).mean() < 2e-3, (name, d.max(), (d > 0).mean())
        ref_t = O.quantize_u8(O.tiled_upscale(fwd, x, 2, tile=32, overlap=8)[0], 1.0).numpy()
        d = np.abs(m.inference_tiled(img, tile=32, overlap=8, precision="fp32").astype(np.int32) - ref_t.astype(np.int32))
        assert d.max() <= 1 and (d > 0).mean() < 2e-3, (name, "tiled", d.max(), (d > 0).mean())


def test_inference_u8_matches_reference(golden_meta):
    g = load_golden("swinir_tiny_x4_inference_u8")
    m = _swinir(golden_meta["swinir_ops"]["cfg"], 11)
    out = m.inference(g["img"])
    assert out.dtype == np.uint8 and out.shape == g["out"].shape
    d = np.abs(out.astype(np.int32) - g["out"].astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 2e-3
    ens = load_golden("swinir_tiny_x4_ensemble_u8")
    out = m.inference_with_self_ensemble(ens["img"])
    d = np.abs(out.astype(np.int32) - ens["out"].astype(np.int32))
    assert d.max() <= 1 and (d > 0).mean() < 2e-3


def test_tiled_inference_matches_oracle_tiler():
    """cfg5 semantics at a size the oracle finishes in seconds: 100x130 frame, 64x64 tiles, overlap 16."""
    cfg = synth.swinir_config(**synth.SWINIR_TINY)
    P = synth.swinir_weights(cfg, 11)
    m = _swinir(cfg, 11).eval()
    img = synth.smooth_image_u8(100, 130, seed=3)
    x = torch.from_numpy(img.astype(np.float32) / 255.0).permute(2, 0, 1).unsqueeze(0)
    ref = O.tiled_upscale(lambda t: O.swinir_forward(P, t, cfg), x, 4, tile=64, overlap=16)
    ref_u8 = O.quantize_u8(ref[0], 1.0).numpy()
    out = m.inference_tiled(img, tile=64, overlap=16, precision="fp32")
    d = np.abs(out.astype(np.int32) - ref_u8.astype(np.int32))
    assert out.shape == (400, 520, 3) and d.max() <= 1 and (d > 0).mean() < 2e-3
    out_bf = m.inference_tiled(img, tile=64, overlap=16, precision="bf16")
    assert abs(O.psnr(torch.from_numpy(out_bf.astype(np.float32)), torch.from_numpy(ref_u8.astype(np.float32)))) > 35.0
    # frame smaller than a tile -> single tile == plain inference
    small = synth.smooth_image_u8(20, 28, seed=7)
    assert np.array_equal(m.inference_tiled(small), m.inference(small))


@pytest.mark.parametrize("H,W,tile,overlap", [(40, 150, 64, 16), (97, 64, 48, 8), (64, 64, 64, 16), (130, 70, 32, 12)])
def test_tiled_inference_ragged_frames_and_chunking(H, W, tile, overlap):
    """Frames with one side below the tile size, tile / overlap pairs other than 64 / 16, clamped last tiles: the device tiler
    equals the oracle tiler, and processing the tile list in chunks (a bounded workspace) changes nothing, bit for bit."""
    cfg = synth.swinir_config(**synth.SWINIR_TINY)
    P = synth.swinir_weights(cfg, 11)
    m = _swinir(cfg, 11).eval()
    img = synth.smooth_image_u8(H, W, seed=H + W)
    x = torch.from_numpy(img.astype(np.float32) / 255.0).permute(2, 0, 1).unsqueeze(0)
    ref_u8 = O.quantize_u8(O.tiled_upscale(lambda t: O.swinir_forward(P, t, cfg), x, 4, tile=tile, overlap=overlap)[0], 1.0).numpy()
    out = m.inference_tiled(img, tile=tile, overlap=overlap, precision="fp32")
    d = np.abs(out.astype(np.int32) - ref_u8.astype(np.int32))
    assert out.shape == (4 * H, 4 * W, 3) and d.max() <= 1 and (d > 0).mean() < 2e-3, (d.max(), (d > 0).mean())
    frame = torch.from_numpy(img).cuda()
    for prec in ("fp32", "bf16"):
        nat = m._native(frame.device, prec)
        whole = nat.upscale_tiled_u8(frame, 4, tile, overlap)
        # the host entry point (two passes with the D2H of the finished rows under the last tile row's compute) == the device one
        assert np.array_equal(m.inference_tiled(img, tile=tile, overlap=overlap, precision=prec), whole.cpu().numpy()), prec
        for chunk in (1, 2, 3):
            assert torch.equal(nat.upscale_tiled_u8(frame, 4, tile, overlap, chunk), whole), (prec, chunk)


def test_reference_shape_tests_pass_unchanged():
    """The reference's own API gate (tests/models/test_swinir.py:8-26, test_edsr.py) on the drop-in:
    default full-size config, train mode, 8x8 and 12x12 inputs, every scale."""
    from studiosr_b200.models import EDSR, SwinIR

    for scale in (2, 3, 4, 8):
        model = SwinIR(scale=scale, n_colors=3).cuda()
        for hw in (8, 12):
            y = model(torch.randn(1, 3, hw, hw).cuda())
            assert y.shape == (1, 3, scale * hw, scale * hw)
    for scale in (2, 3, 4):
        model = EDSR(scale=scale, n_colors=3).cuda()
        y = model(torch.randn(1, 3, 8, 8).cuda())
        assert y.shape == (1, 3, scale * 8, scale * 8)


def test_reference_shape_tests_hat_rcan_pass_unchanged():
    """tests/models/test_hat.py:8-21, test_rcan.py:8-25 and test_han.py:8-25 of the reference on the drop-in: default full-size configs, train
    mode with grad enabled (as the reference's tests call them), every scale."""
    from studiosr_b200.models import HAT, RCAN

    for scale in (2, 3, 4, 8):
        model = HAT(scale=scale, n_colors=3).cuda()
        y = model(torch.randn(1, 3, 12, 12).cuda())
        assert y.shape == (1, 3, scale * 12, scale * 12) and torch.isfinite(y).all()
        del model
    for scale in (2, 3, 4, 8):
        model = RCAN(scale=scale, n_colors=3).cuda()
        for hw in (8, 12):
            y = model(torch.randn(1, 3, hw, hw).cuda())
            assert y.shape == (1, 3, scale * hw, scale * hw) and torch.isfinite(y).all()
        del model
    from studiosr_b200.models import HAN  # tests/models/test_han.py:8-25

    for scale in (2, 3, 4, 8):
        model = HAN(scale=scale, n_colors=3).cuda()
        for hw in (8, 12):
            y = model(torch.randn(1, 3, hw, hw).cuda())
            assert y.shape == (1, 3, scale * hw, scale * hw) and torch.isfinite(y).all()
        del model


def test_from_pretrained_never_returns_random_weights(tmp_path, monkeypatch):
    """Model.from_pretrained (common.py:79-81 and the per-model overrides): reference signatures, and a missing weight file is an
    error, not a randomly initialised model."""
    from studiosr_b200.models import EDSR, HAT, RCAN, SwinIR
    from studiosr_b200.models.common import Model

    monkeypatch.chdir(tmp_path)
    for call in (lambda: EDSR.from_pretrained(scale=4, dataset="DIV2K"), lambda: EDSR.from_pretrained(2, "DF2K"),
                 lambda: RCAN.from_pretrained(scale=3), lambda: SwinIR.from_pretrained(scale=4), lambda: HAT.from_pretrained(scale=4)):
        with pytest.raises(FileNotFoundError):
            call()
    with pytest.raises(NotImplementedError):
        Model.from_pretrained()
    # a weight file in place is loaded, with the reference's file layout and img_range (rcan.py:116-118)
    import os

    os.makedirs(os.path.join("pretrained", "models_ECCV2018RCAN"))
    src = RCAN(scale=2, img_range=255.0)
    torch.save(src.state_dict(), os.path.join("pretrained", "models_ECCV2018RCAN", "RCAN_BIX2.pt"))
    m = RCAN.from_pretrained(scale=2)
    assert m.img_range == 255.0
    for k, v in src.state_dict().items():
        assert torch.equal(m.state_dict()[k], v), k


def test_param_cache_sees_data_writes_and_model_copies():
    """writes through `.data` do not bump `_version`; the packed-weight cache must still notice them, and models
    must survive copy.deepcopy / pickling after their first forward (EMA copies)."""
    import copy
    import pickle

    m = _edsr(synth.EDSR_TINY, 5)
    x = synth.image_batch((1, 3, 12, 12), 9).cuda()
    with torch.no_grad():
        y0 = m(x).clone()
        for p in m.parameters():
            if p.requires_grad:
                p.data.mul_(0.5)  # invisible to p._version
        y1 = m(x).clone()
        assert (y1 - y0).abs().max().item() > 1e-4, "stale packed weights after a .data write"
        ema = copy.deepcopy(m)
        assert torch.equal(ema(x), y1)
        m2 = pickle.loads(pickle.dumps(m))
        assert torch.equal(m2.cuda()(x), y1)
        m.trust_param_versions = True  # opt-out of the checksum: .data writes then need an explicit invalidate_native()
        for p in m.parameters():
            if p.requires_grad:
                p.data.mul_(2.0)
        m.invalidate_native()
        assert torch.allclose(m(x), y0, atol=1e-5)


def test_tiled_full_model_matches_oracle_tiler():
    """The path bench.py times (BASELINE.json config 5) on the FULL 180/6 model: 112x160 frame -> 2x3 tiles of 64x64 (overlap 16)
    through ssr_model_upscale_tiled_u8(_host) against the oracle tiler calling the reference-equivalent forward per tile.
    fp32: <= 1 LSB; bf16: uint8 PSNR against the fp32 oracle output (rms error below ~1 LSB) and the realistic-GT PSNR delta."""
    cfg = synth.swinir_config()
    P = synth.swinir_weights(cfg, 0)
    m = _swinir(cfg, 0).eval()
    img = synth.smooth_image_u8(112, 160, seed=5)
    x = torch.from_numpy(img.astype(np.float32) / 255.0).permute(2, 0, 1).unsqueeze(0)
    with torch.inference_mode():
        ref = O.tiled_upscale(lambda t: O.swinir_forward(P, t, cfg), x, 4, tile=64, overlap=16)
    ref_u8 = O.quantize_u8(ref[0], 1.0).numpy()
    out = m.inference_tiled(img, tile=64, overlap=16, precision="fp32")
    d = np.abs(out.astype(np.int32) - ref_u8.astype(np.int32))
    assert out.shape == (448, 640, 3) and d.max() <= 1 and (d > 0).mean() < 2e-3
    out_bf = m.inference_tiled(img, tile=64, overlap=16, precision="bf16")
    f = lambda a: torch.from_numpy(a.astype(np.float32))
    assert O.psnr(f(out_bf), f(ref_u8)) > 46.0  # rms error < 1.3 LSB
    gt = (ref[0] + GT_SIGMA * torch.randn(ref[0].shape, generator=torch.Generator().manual_seed(99))).clip(0, 1)
    gt_u8 = O.quantize_u8(gt, 1.0).numpy()
    delta = abs(O.psnr(f(out_bf), f(gt_u8)) - O.psnr(f(ref_u8), f(gt_u8)))
    assert delta < 0.12, delta  # the reference's own bf16 autocast sits at ~0.12 dB on this model (bf16_autocast_ref.json)


def test_sharded_tiles_and_band_blend_are_bit_identical_to_whole_frame():
    """ssr_model_tiles_u8 + ssr_model_blend_tiles_u8 (the per-rank halves of the multi-GPU protocol, studiosr_b200/sharding.py):
    the tile list computed in three ranges and the frame blended in three row bands equal the one-call result bit for bit, and
    so does the ShardedTiledUpscaler at world size 1."""
    from studiosr_b200.sharding import NativeTileBackend, ShardedTiledUpscaler, slot_partition

    cfg = synth.swinir_config(**synth.SWINIR_TINY)
    m = _swinir(cfg, 11).eval()
    img = synth.smooth_image_u8(100, 130, seed=3)
    frame = torch.from_numpy(img).cuda()
    for prec in ("fp32", "bf16"):
        nat = m._native(frame.device, prec)
        whole = nat.upscale_tiled_u8(frame, 4, 64, 16)
        be = NativeTileBackend(nat, 100, 130, 4, 64, 16)
        assert be.n_tiles == 6
        tiles = torch.empty((be.n_tiles, be.tile_elems), dtype=torch.float32, device="cuda")
        for b, e in slot_partition(be.n_tiles, 3)[1]:
            be.compute(frame, tiles[b:], b, e)
        out = torch.zeros((400, 520, 3), dtype=torch.uint8, device="cuda")
        for r0, r1 in ((0, 133), (133, 134), (134, 400)):
            be.blend(tiles, out, r0, r1)
        assert torch.equal(out, whole), prec
        up = ShardedTiledUpscaler(be)  # torch.distributed not initialised: world size 1
        assert torch.equal(up.upscale(frame), whole)
        h_out = torch.empty((400, 520, 3), dtype=torch.uint8).pin_memory()
        up.upscale_host(torch.from_numpy(img).pin_memory(), h_out)
        assert torch.equal(h_out, whole.cpu())
        # the same pass captured once and replayed as a CUDA graph (what bench.py's multi-GPU leg runs): a second, different
        # frame through the replay must equal its own eager result
        upg = ShardedTiledUpscaler(NativeTileBackend(nat, 100, 130, 4, 64, 16), graph=True)
        assert torch.equal(upg.upscale(frame), whole)
        img2 = synth.smooth_image_u8(100, 130, seed=4)
        frame2 = torch.from_numpy(img2).cuda()
        whole2 = nat.upscale_tiled_u8(frame2, 4, 64, 16)
        assert not torch.equal(whole2, whole)
        assert torch.equal(upg.upscale(frame2), whole2) and upg._graph is not None
        upg.upscale_host(torch.from_numpy(img).pin_memory(), h_out)
        assert torch.equal(h_out, whole.cpu())


def test_hat_full_batch_of_two():
    """The full BASELINE.json config-3 model at B = 2 (the golden holds B = 1): sample 0 is the golden input and must reproduce
    the golden output; sample 1 must equal its own B = 1 run (samples are independent)."""
    from studiosr_b200.models import HAT

    meta = __import__("json").load(open(__import__("os").path.join(__import__("os").path.dirname(__file__), "golden", "meta.json")))["cases"]
    c = meta["hat_full_x4_eval_1x64x64"]
    m = HAT(drop_path_rate=0.0, **c["cfg"])
    m.load_state_dict(synth.hat_weights(c["cfg"], c["wseed"]), strict=True)
    m = m.cuda().eval()
    x0 = synth.image_batch(c["shape"], c["xseed"])
    x1 = synth.image_batch(c["shape"], c["xseed"] + 1)
    ref = torch.from_numpy(load_golden("hat_full_x4_eval_1x64x64")["y"])
    for prec in ("bf16", "tf32"):
        m.precision = prec
        with torch.no_grad():
            y = m(torch.cat([x0, x1]).cuda()).float().cpu()
            y1 = m(x1.cuda()).float().cpu()
        assert (y[1:] - y1).abs().max().item() <= 1e-6, prec
        if prec == "bf16":
            _check_bf16("hat_full_x4_eval_1x64x64", y[:1], ref)
        else:
            assert (y[:1] - ref).abs().max().item() <= 4 * ABS_TOL["tf32"]
