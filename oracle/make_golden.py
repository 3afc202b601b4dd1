"""TEST INFRASTRUCTURE ONLY -- generate tests/golden/*.npz by EXECUTING THE UNMODIFIED
REFERENCE (imported from /root/reference through oracle/ref_shim.py).

Run in the build container (the only place /root/reference exists):

    python -m oracle.make_golden

Every fixture is defined by (model config, weight seed, input seed); weights are
re-synthesised from the seed by oracle/synth.py at test time, so only inputs that are
not seed-derived and the reference's OUTPUTS are stored.  Loading the synthetic
state_dict uses strict=True, which pins synth.py's key/shape list to the reference.
"""
import os

import numpy as np
import torch

from . import synth
from .ref_shim import import_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _ref_swinir(models, cfg, seed):
    m = models.SwinIR(drop_path_rate=0.0, **cfg)
    m.load_state_dict(synth.swinir_weights(cfg, seed), strict=True)
    return m


def main() -> None:
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref = import_reference()
    from studiosr.models import common as rc
    from studiosr.models import swinir as rs

    models = ref.models
    os.makedirs(OUT, exist_ok=True)
    meta = {}

    # ---- whole-model SwinIR cases --------------------------------------------------
    cases = {
        # name: (cfg, weight seed, input shape, input seed, training)
        "swinir_tiny_x4_eval_2x20x28": (synth.swinir_config(**synth.SWINIR_TINY), 11, (2, 3, 20, 28), 101, False),
        "swinir_tiny_x4_eval_1x16x16": (synth.swinir_config(**synth.SWINIR_TINY), 11, (1, 3, 16, 16), 102, False),
        "swinir_tiny_x4_train_1x12x12": (synth.swinir_config(**synth.SWINIR_TINY), 11, (1, 3, 12, 12), 103, True),
        "swinir_tiny_x4_train_2x16x24": (synth.swinir_config(**synth.SWINIR_TINY), 11, (2, 3, 16, 24), 104, True),
        "swinir_tiny_x2_eval_1x12x12": (synth.swinir_config(**dict(synth.SWINIR_TINY, scale=2)), 12, (1, 3, 12, 12), 105, False),
        "swinir_tiny_x3_eval_1x8x8": (synth.swinir_config(**dict(synth.SWINIR_TINY, scale=3)), 13, (1, 3, 8, 8), 106, False),
        "swinir_tiny_x8_eval_1x8x8": (synth.swinir_config(**dict(synth.SWINIR_TINY, scale=8)), 14, (1, 3, 8, 8), 107, False),
        "swinir_light_x4_eval_1x12x20": (
            synth.swinir_config(**dict(synth.SWINIR_TINY, upsampler="pixelshuffledirect")), 15, (1, 3, 12, 20), 108, False),
        "swinir_full_x4_eval_cfg1": (synth.swinir_config(), 0, (1, 3, 64, 64), 1234, False),
    }
    for name, (cfg, wseed, shape, xseed, training) in cases.items():
        m = _ref_swinir(models, cfg, wseed)
        m.train(training)
        x = synth.image_batch(shape, xseed)
        with torch.no_grad():
            y = m(x)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), y=y.numpy())
        meta[name] = dict(cfg=cfg, wseed=wseed, shape=list(shape), xseed=xseed, training=training,
                          out_shape=list(y.shape), absmax=float(y.abs().max()))
        print(name, tuple(y.shape), float(y.abs().max()))

    # ---- Model.inference (uint8 in/out) on the tiny model ----------------------------
    cfg = synth.swinir_config(**synth.SWINIR_TINY)
    m = _ref_swinir(models, cfg, 11)
    img = synth.smooth_image_u8(20, 28, seed=7)
    out = m.inference(img)
    np.savez_compressed(os.path.join(OUT, "swinir_tiny_x4_inference_u8.npz"), img=img, out=out)
    ens = m.inference_with_self_ensemble(synth.smooth_image_u8(16, 16, seed=8))
    np.savez_compressed(os.path.join(OUT, "swinir_tiny_x4_ensemble_u8.npz"),
                        img=synth.smooth_image_u8(16, 16, seed=8), out=ens)

    # ---- op-level cases from the reference's own functions / modules ------------------
    ops = {}
    ops["mask_24x32_ws8_s4"] = rc.calculate_mask((24, 32), 8, 4).numpy()
    ops["mask_16x16_ws8_s0"] = rc.calculate_mask((16, 16), 8, 0).numpy()
    xpad = synth.image_batch((1, 3, 13, 16), 201)
    ops["pad_eval_13x16"] = rs.check_image_size_for_eval(xpad, 8).numpy()
    ops["pad_train_13x16"] = rc.check_image_size(xpad, 8).numpy()
    # one shifted block of the tiny model (C=60, 6 heads, d=10) on [2,16,24,60]
    blk_pre = "layers.1.residual_group.blocks.1"
    blk = m.layers[1].residual_group.blocks[1]
    assert blk.shift_size == 4
    xb = torch.randn(2, 16, 24, 60, generator=torch.Generator().manual_seed(202))
    with torch.no_grad():
        ops["block_shift4_out"] = blk.eval()(xb).numpy()
        xw = torch.randn(6, 64, 60, generator=torch.Generator().manual_seed(203))
        msk = rc.calculate_mask((16, 24), 8, 4)
        ops["winattn_masked_out"] = blk.attn(xw, mask=msk).numpy()
        ops["winattn_nomask_out"] = blk.attn(xw, mask=None).numpy()
    ops["rpi_ws8"] = blk.attn.relative_position_index.numpy()
    np.savez_compressed(os.path.join(OUT, "swinir_ops.npz"), **ops)
    meta["swinir_ops"] = dict(cfg=cfg, wseed=11, block=blk_pre, xb_seed=202, xw_seed=203, pad_seed=201)

    # ---- EDSR ------------------------------------------------------------------------
    ecases = {
        "edsr_tiny_x4_2x12x20": (dict(synth.EDSR_TINY), 21, (2, 3, 12, 20), 301),
        "edsr_tiny_x2_1x9x11": (dict(synth.EDSR_TINY, scale=2), 22, (1, 3, 9, 11), 302),
        "edsr_tiny_x3_1x8x8": (dict(synth.EDSR_TINY, scale=3), 23, (1, 3, 8, 8), 303),
        "edsr_full_x4_1x24x24": (dict(synth.EDSR_DEFAULT), 24, (1, 3, 24, 24), 304),
    }
    for name, (cfg, wseed, shape, xseed) in ecases.items():
        m = models.EDSR(**cfg)
        m.load_state_dict(synth.edsr_weights(cfg, wseed), strict=True)
        m.eval()
        x = synth.image_batch(shape, xseed)
        with torch.no_grad():
            y = m(x)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), y=y.numpy())
        meta[name] = dict(cfg=cfg, wseed=wseed, shape=list(shape), xseed=xseed, out_shape=list(y.shape),
                          absmax=float(y.abs().max()))
        print(name, tuple(y.shape), float(y.abs().max()))

    import json

    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(dict(torch=torch.__version__, cases=meta), f, indent=1, sort_keys=True)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
