"""TEST INFRASTRUCTURE ONLY -- golden GRADIENTS of the training step, produced by EXECUTING THE UNMODIFIED REFERENCE
(imported from /root/reference through oracle/ref_shim.py): model.train(), fp32, `loss = nn.L1Loss()(model(x), y)`,
`loss.backward()` exactly as trainer.py:45,101-104 does, with drop_path_rate = 0 so the step is deterministic.

    python -m oracle.make_golden_train [case ...]      (no names: every case; names: only those, the rest of meta_train.json is kept)

Stored per case: the loss and, for every parameter, the gradient's L2 norm plus a strided sample of its entries
(every `stride`-th element of the flattened gradient) -- enough to pin the oracle's autograd (tests/test_oracle.py) and,
through it, the CUDA backward (tests/test_gpu_train.py) without committing megabytes of gradients."""
import json
import os
import sys

import numpy as np
import torch

from . import synth
from .ref_shim import import_reference

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
STRIDE = 7

CASES = {
    # name: (arch, cfg, weight seed, input shape, input seed)
    "train_edsr_tiny_x4_2x24x20": ("edsr", dict(synth.EDSR_TINY), 5, (2, 3, 24, 20), 77),
    "train_edsr_tiny_x3_1x12x16": ("edsr", dict(synth.EDSR_TINY, scale=3, n_resblocks=1), 5, (1, 3, 12, 16), 77),
    "train_rcan_tiny_x4_2x12x20": ("rcan", dict(synth.RCAN_TINY), 9, (2, 3, 12, 20), 55),
    "train_han_tiny_x4_2x12x16": ("han", dict(synth.HAN_TINY), 13, (2, 3, 12, 16), 57),
    "train_swinir_tiny_x4_2x16x24": ("swinir", synth.swinir_config(**synth.SWINIR_TINY), 11, (2, 3, 16, 24), 101),
    "train_swinir_tiny_x4_pad_1x20x28": ("swinir", synth.swinir_config(**synth.SWINIR_TINY), 11, (1, 3, 20, 28), 101),
    "train_swinir_c180_x4_1x16x16": ("swinir", synth.swinir_config(embed_dim=180, depths=[2, 2], num_heads=[6, 6]), 11,
                                     (1, 3, 16, 16), 101),
    # stochastic depth ON (reference default is drop_path_rate=0.1; 0.5 here so that several branches really drop):
    # torch.manual_seed(DROP_SEED) right before the forward fixes the masks the reference's DropPath modules draw
    "train_swinir_tiny_x4_droppath_4x16x16": ("swinir_dp", synth.swinir_config(**synth.SWINIR_TINY), 11, (4, 3, 16, 16), 101),
}
DROP_SEED, DROP_RATE = 77, 0.5


def case_inputs(cfg, shape, xseed):
    x = synth.image_batch(shape, xseed)
    tgt = synth.image_batch((shape[0], 3, shape[2] * cfg["scale"], shape[3] * cfg["scale"]), xseed + 1)
    return x, tgt


def main() -> None:
    torch.manual_seed(0)
    torch.set_num_threads(8)
    ref = import_reference()
    models = ref.models
    meta = {}
    only = set(sys.argv[1:])
    if only:
        with open(os.path.join(OUT, "meta_train.json")) as f:
            meta = json.load(f)["cases"]
    for name, (arch, cfg, wseed, shape, xseed) in CASES.items():
        if only and name not in only:
            continue
        if arch == "edsr":
            m = models.EDSR(**cfg)
            m.load_state_dict(synth.edsr_weights(cfg, wseed), strict=True)
        elif arch == "rcan":
            m = models.RCAN(**cfg)
            m.load_state_dict(synth.rcan_weights(cfg, wseed), strict=True)
        elif arch == "han":
            from studiosr.models.han import HAN

            m = HAN(**cfg)
            m.load_state_dict(synth.han_weights(cfg, wseed), strict=True)
        else:
            m = models.SwinIR(drop_path_rate=DROP_RATE if arch == "swinir_dp" else 0.0, **cfg)
            m.load_state_dict(synth.swinir_weights(cfg, wseed), strict=True)
        m.train()
        x, tgt = case_inputs(cfg, shape, xseed)
        if arch == "swinir_dp":
            torch.manual_seed(DROP_SEED)
        loss = torch.nn.L1Loss()(m(x), tgt)
        loss.backward()
        arrays = {}
        for k, p in m.named_parameters():
            if p.grad is None:
                continue
            g = p.grad.detach().flatten()
            arrays[k + "::norm"] = np.asarray([g.norm().item()], dtype=np.float64)
            arrays[k + "::sample"] = g[::STRIDE].numpy().astype(np.float32)
        np.savez_compressed(os.path.join(OUT, name + ".npz"), loss=np.asarray([loss.item()]), **arrays)
        meta[name] = dict(arch=arch, cfg=cfg, wseed=wseed, shape=list(shape), xseed=xseed, stride=STRIDE,
                          n_params=len(arrays) // 2)
        if arch == "swinir_dp":
            meta[name].update(drop_seed=DROP_SEED, drop_path_rate=DROP_RATE)
        print(name, "loss", loss.item(), "params", len(arrays) // 2)
    with open(os.path.join(OUT, "meta_train.json"), "w") as f:
        json.dump(dict(cases=meta), f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
