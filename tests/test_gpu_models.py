"""-m gpu: whole-model parity of the drop-in modules (-> C ABI -> CUDA kernels) against the
golden fixtures produced by the reference and against the oracle restatement.

Tolerances (BASELINE.json north_star): fp32 output max-abs <= 1e-3 (the CUDA-core fp32 path is
held to 1e-4, the tf32 tensor-core path to 1e-3); bf16: PSNR delta <= 0.01 dB on the uint8-
quantised output versus the reference's fp32 output against a synthetic ground truth."""
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
ABS_TOL = {"fp32": 1e-4, "tf32": 2e-3}


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


def _psnr_delta(y, ref, seed=99):
    """|PSNR(q(y), gt) - PSNR(q(ref), gt)| with the reference's uint8 quantisation (common.py:44-45)."""
    gt = torch.rand(ref.shape, generator=torch.Generator().manual_seed(seed))
    q = lambda t: (t * 255.0).round().clip(0, 255)
    return abs(O.psnr(q(y), q(gt)) - O.psnr(q(ref), q(gt)))


@pytest.mark.parametrize("prec", ["fp32", "tf32", "bf16"])
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
        assert err <= 6e-2, f"{name} [bf16] max-abs {err:.3e}"
        assert _psnr_delta(y, ref) <= 0.01, f"{name} [bf16] PSNR delta {_psnr_delta(y, ref):.4f} dB"


@pytest.mark.parametrize("prec", ["fp32", "tf32", "bf16"])
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
        assert err <= 6e-2 and _psnr_delta(y, ref) <= 0.01, f"{name} [bf16] max-abs {err:.3e}"


HAT_CASES = ["hat_tiny_x4_eval_2x20x40", "hat_tiny_x4_train_1x32x32", "hat_tiny_x2_eval_1x16x48", "hat_tiny_x3_eval_1x17x17",
             "hat_full_x4_eval_1x64x64"]


@pytest.mark.parametrize("prec", ["fp32", "tf32", "bf16"])
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
        assert err <= 6e-2 and _psnr_delta(y, ref) <= 0.01, f"{name} [bf16] max-abs {err:.3e}"


RCAN_CASES = ["rcan_tiny_x4_2x12x20", "rcan_tiny_x2_1x9x11", "rcan_tiny_x3_1x8x8", "rcan_full_x4_1x24x24"]


@pytest.mark.parametrize("prec", ["fp32", "tf32", "bf16"])
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
        assert err <= 1e-1 and _psnr_delta(y, ref) <= 0.01, f"{name} [bf16] max-abs {err:.3e}"


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
    assert np.array_equal(m.inference_tiled(small, precision="fp32"), m.inference(small))


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
