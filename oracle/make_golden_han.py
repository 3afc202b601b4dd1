"""Golden fixtures for HAN: executes the UNMODIFIED reference (`/root/reference`, imported
through oracle/ref_shim.py) on seed-defined synthetic weights / inputs (oracle/synth.py) and commits the outputs under
tests/golden/.  Test infrastructure only; run in the build container (the reference does not exist on the GPU box):
    python -m oracle.make_golden_han"""
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
    from studiosr.models.han import HAN  # (not re-exported by every version of studiosr.models)

    with open(os.path.join(OUT, "meta.json")) as f:
        doc = json.load(f)
    cases = {
        "han_tiny_x4_2x12x20": (dict(synth.HAN_TINY), 41, (2, 3, 12, 20), 501),
        "han_tiny_x2_1x9x11": (dict(synth.HAN_TINY, scale=2), 42, (1, 3, 9, 11), 502),
        "han_tiny_x3_1x8x8": (dict(synth.HAN_TINY, scale=3, reduction=8, n_resblocks=2), 43, (1, 3, 8, 8), 503),
        "han_full_x4_1x16x16": (dict(synth.HAN_DEFAULT), 44, (1, 3, 16, 16), 504),
    }
    for name, (cfg, wseed, shape, xseed) in cases.items():
        m = HAN(**cfg)
        m.load_state_dict(synth.han_weights(cfg, wseed), strict=True)
        m.eval()
        x = synth.image_batch(shape, xseed)
        with torch.no_grad():
            y = m(x)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), y=y.numpy())
        doc["cases"][name] = dict(cfg=cfg, wseed=wseed, shape=list(shape), xseed=xseed, out_shape=list(y.shape),
                                  absmax=float(y.abs().max()))
        print(name, tuple(y.shape), float(y.abs().max()))
    with open(os.path.join(OUT, "meta.json"), "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
