"""Golden fixtures for SwinFIR: executes the UNMODIFIED reference (`/root/reference`, imported
through oracle/ref_shim.py) on seed-defined synthetic weights / inputs (oracle/synth.py) and commits the outputs under
tests/golden/.  Test infrastructure only; run in the build container (the reference does not exist on the GPU box):
    python -m oracle.make_golden_swinfir"""
import json
import os

import numpy as np
import torch

from oracle import synth
from oracle.make_golden import OUT, import_reference

KEYS = ("scale", "n_colors", "img_range", "embed_dim", "depths", "num_heads", "window_size", "mlp_ratio", "upsampler")


def main() -> None:
    torch.manual_seed(0)
    torch.set_num_threads(8)
    import_reference()
    from studiosr.models.swinfir import SwinFIR

    with open(os.path.join(OUT, "meta.json")) as f:
        doc = json.load(f)
    tiny = synth.swinir_config(**synth.SWINIR_TINY)
    cases = {
        # eval forward: the input is flip-padded to the next multiple of 8 (12x20 -> 16x24: FFT lengths 16 / 24, not all powers of 2)
        "swinfir_tiny_x4_eval_1x12x20": (tiny, 21, (1, 3, 12, 20), 601, False),
        "swinfir_tiny_x2_eval_2x16x16": (dict(tiny, scale=2), 22, (2, 3, 16, 16), 602, False),
        "swinfir_tiny_x4_train_1x16x24": (tiny, 23, (1, 3, 16, 24), 603, True),
        "swinfir_c180_x4_eval_1x8x8": (synth.swinir_config(embed_dim=180, depths=[2, 2], num_heads=[6, 6]), 24, (1, 3, 8, 8), 604, False),
    }
    for name, (cfg, wseed, shape, xseed, train) in cases.items():
        m = SwinFIR(drop_path_rate=0.0, **{k: cfg[k] for k in KEYS})
        m.load_state_dict(synth.swinfir_weights(cfg, wseed), strict=True)
        m.train(train)
        x = synth.image_batch(shape, xseed)
        with torch.no_grad():
            y = m(x)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), y=y.numpy())
        doc["cases"][name] = dict(cfg=dict(cfg, sfb=True), wseed=wseed, shape=list(shape), xseed=xseed, out_shape=list(y.shape),
                                  absmax=float(y.abs().max()), training=train)
        print(name, tuple(y.shape), float(y.abs().max()))
    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
