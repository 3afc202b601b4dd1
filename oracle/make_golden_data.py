"""TEST INFRASTRUCTURE ONLY -- golden vectors for the evaluation metric and the training transform, produced by EXECUTING THE
UNMODIFIED REFERENCE (studiosr.utils.metrics.compute_psnr, studiosr.data.transforms) through oracle/ref_shim.py:

    python -m oracle.make_golden_data          (build container only)

Writes tests/golden/data_ops.npz: seeded uint8 image pairs, the reference's PSNR values for (y_only, crop_border) combinations,
and the patches the reference's Compose([RandomCrop, RandomHorizontalFlip, RandomVerticalFlip, RandomRotation90]) + array2tensor
cuts with `random.seed(s)` for a list of seeds."""
import os
import random

import numpy as np

from . import synth
from .ref_shim import import_reference

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
SEEDS = [0, 1, 2, 3, 7, 11, 12, 13]


def main() -> None:
    import_reference()
    from studiosr.data import transforms as T
    from studiosr.utils import metrics

    out = {}
    # ---- PSNR: an "SR" image = the ground truth + noise, and one pair of unequal sizes (crop_img_to_equal)
    gt = synth.smooth_image_u8(96, 120, seed=31)
    rs = np.random.RandomState(5)
    sr = np.clip(gt.astype(np.int32) + rs.randint(-6, 7, gt.shape), 0, 255).astype(np.uint8)
    big = np.concatenate([sr, sr[-3:]], axis=0)[:, : 118]  # 99 x 118 against 96 x 120
    out["psnr_gt"], out["psnr_sr"], out["psnr_big"] = gt, sr, big
    vals = []
    for a, b in ((sr, gt), (big, gt)):
        for y_only in (False, True):
            for cb in (0, 4):
                vals.append(float(metrics.compute_psnr(a, b, y_only=y_only, crop_border=cb)))
    out["psnr_values"] = np.array(vals, dtype=np.float64)
    out["psnr_identical"] = np.array([float(metrics.compute_psnr(gt, gt))])
    # ---- training transform (dataset.py:50-58) + array2tensor, one pair per seed
    size, scale = 12, 4
    lq = synth.smooth_image_u8(20, 28, seed=41)
    hr = synth.smooth_image_u8(80, 112, seed=42)
    tf = T.Compose([T.RandomCrop(size, scale), T.RandomHorizontalFlip(), T.RandomVerticalFlip(), T.RandomRotation90()])
    xs, ys = [], []
    for s in SEEDS:
        random.seed(s)
        a, b = tf(lq, hr)
        xs.append(T.array2tensor(np.ascontiguousarray(a)).numpy())
        ys.append(T.array2tensor(np.ascontiguousarray(b)).numpy())
    out["aug_lq"], out["aug_gt"], out["aug_seeds"] = lq, hr, np.array(SEEDS)
    out["aug_x"], out["aug_y"] = np.stack(xs), np.stack(ys)
    np.savez_compressed(os.path.join(GOLD, "data_ops.npz"), **out)
    print("psnr", vals, "aug", out["aug_x"].shape, out["aug_y"].shape)


if __name__ == "__main__":
    main()
