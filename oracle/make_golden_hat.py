"""Golden fixtures for HAT: executes the UNMODIFIED reference
(`/root/reference`, imported through oracle/ref_shim.py) on seed-defined synthetic weights / inputs
(oracle/synth.py) and commits the outputs under tests/golden/.  Test infrastructure only; run in the
build container (the reference does not exist on the GPU box):  python -m oracle.make_golden_hat"""
import json
import os

import numpy as np
import torch

from oracle import synth
from oracle.make_golden import OUT, import_reference


def main() -> None:
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref = import_reference()
    models = ref.models
    with open(os.path.join(OUT, "meta.json")) as f:
        doc = json.load(f)
    cases = {
        # name: (cfg, weight seed, input shape, input seed, training)
        "hat_tiny_x4_eval_2x20x40": (dict(synth.HAT_TINY), 41, (2, 3, 20, 40), 501, False),
        "hat_tiny_x4_train_1x32x32": (dict(synth.HAT_TINY), 41, (1, 3, 32, 32), 502, True),
        "hat_tiny_x2_eval_1x16x48": (dict(synth.HAT_TINY, scale=2), 42, (1, 3, 16, 48), 503, False),
        "hat_tiny_x3_eval_1x17x17": (dict(synth.HAT_TINY, scale=3), 43, (1, 3, 17, 17), 504, False),
        "hat_full_x4_eval_1x64x64": (dict(synth.HAT_DEFAULT), 44, (1, 3, 64, 64), 505, False),
    }
    for name, (cfg, wseed, shape, xseed, training) in cases.items():
        m = models.HAT(drop_path_rate=0.0, **cfg)
        m.load_state_dict(synth.hat_weights(cfg, wseed), strict=True)
        m.train(training)
        x = synth.image_batch(shape, xseed)
        with torch.no_grad():
            y = m(x)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), y=y.numpy())
        doc["cases"][name] = dict(cfg=cfg, wseed=wseed, shape=list(shape), xseed=xseed, training=training,
                                  out_shape=list(y.shape), absmax=float(y.abs().max()))
        print(name, tuple(y.shape), float(y.abs().max()))
    # the index buffers a freshly constructed reference model computes for itself (hat.py:475-513) -- the OCA one
    # carries the negative-index quirk the restatement has to reproduce
    fresh = models.HAT(embed_dim=60, depths=[2], num_heads=[6], window_size=16)
    np.savez_compressed(os.path.join(OUT, "hat_ops.npz"), rpi_sa=fresh.relative_position_index_SA.numpy(),
                        rpi_oca=fresh.relative_position_index_OCA.numpy())
    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
