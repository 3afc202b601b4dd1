"""Device-side counterparts of the reference's evaluation metric and training transform:

  compute_psnr      studiosr/utils/metrics.py:36-49 (with to_y :11-17 and crop_img_to_equal :20-33) on uint8 HWC images that are
                    already on the GPU -- an evaluation loop (evaluator.py:53-79) then moves one float per image over PCIe instead
                    of the 4x upscaled image;
  PairedAugment     the training transform of dataset.py:50-58 (RandomCrop -> RandomHorizontalFlip -> RandomVerticalFlip ->
                    RandomRotation90, transforms.py:8-61) + array2tensor (:64-68) for a whole batch of (LR, HR) pairs in one launch.
                    The random draws stay on the host with python's `random` in the reference's order (randint for x, randint
                    for y, then one random() per flip / rotation), so a seeded run reproduces the reference's patches.

Both need CUDA tensors (there is no CPU fallback; numpy inputs are uploaded)."""
import math
import random
from typing import List, Sequence, Tuple

import numpy as np
import torch

from . import _lib


def _dev_u8(img, device) -> torch.Tensor:
    t = torch.from_numpy(np.ascontiguousarray(img)) if isinstance(img, np.ndarray) else img
    if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
        raise ValueError("expected a uint8 HWC image with 3 channels")
    return t.to(device).contiguous()


def compute_psnr(im1, im2, y_only: bool = False, crop_border: int = 0, device="cuda") -> float:
    """metrics.py:36-49 for uint8 HWC images (torch CUDA tensors, or numpy arrays which are uploaded)."""
    lib = _lib.load()
    a, b = _dev_u8(im1, device), _dev_u8(im2, device)
    if not a.is_cuda:
        raise RuntimeError("studiosr_b200.data.compute_psnr runs on CUDA tensors only (no CPU fallback)")
    mse = torch.empty(1, dtype=torch.float64, device=a.device)
    ws = torch.empty(lib.ssr_psnr_workspace_bytes(), dtype=torch.uint8, device=a.device)
    with torch.cuda.device(a.device):
        _lib.check(lib.ssr_psnr_mse_u8(a.data_ptr(), a.shape[0], a.shape[1], b.data_ptr(), b.shape[0], b.shape[1], int(crop_border),
                                       int(bool(y_only)), mse.data_ptr(), ws.data_ptr(), ws.numel(),
                                       torch.cuda.current_stream(a.device).cuda_stream))
    e = float(np.float32(mse.item()))  # the reference's error is a float32 mean
    return float("inf") if e == 0 else 20.0 * math.log10(255.0 / math.sqrt(e))


class PairedAugment:
    """Compose([RandomCrop(size, scale), RandomHorizontalFlip(), RandomVerticalFlip(), RandomRotation90()]) + ToTensor of
    dataset.py:50-60 for a batch: __call__(lq_images, gt_images) -> (x [B,3,size,size], y [B,3,size*scale,size*scale]) fp32 / 255.
    Images are uint8 HWC torch CUDA tensors (numpy arrays are uploaded); `rng` is a `random.Random` (default: the module-level
    generator the reference's transforms use)."""

    def __init__(self, size: int = 48, scale: int = 4, p: float = 0.5, rng=None, device="cuda"):
        self.size, self.scale, self.p, self.rng, self.device = size, scale, p, rng or random, torch.device(device)

    def draw(self, h: int, w: int) -> Tuple[int, int, int]:
        """(xs, ys, flags) with the reference's sequence of draws (transforms.py:14-15,31,42,53)."""
        xs = self.rng.randint(0, w - self.size)
        ys = self.rng.randint(0, h - self.size)
        flags = 0
        for bit in (1, 2, 4):
            if self.rng.random() < self.p:
                flags |= bit
        return xs, ys, flags

    def __call__(self, lq_images: Sequence, gt_images: Sequence, params: List[Tuple[int, int, int]] = None):
        lib = _lib.load()
        lqs = [_dev_u8(im, self.device) for im in lq_images]
        gts = [_dev_u8(im, self.device) for im in gt_images]
        n, s, k = len(lqs), self.size, self.scale
        if params is None:
            params = [self.draw(im.shape[0], im.shape[1]) for im in lqs]
        table = np.zeros((n, 5), dtype=np.int64)  # struct ssr_aug_pair (include/ssr_b200.h), 40 bytes
        for i, (lq, gt, (xs, ys, flags)) in enumerate(zip(lqs, gts, params)):
            if gt.shape[0] < lq.shape[0] * k or gt.shape[1] < lq.shape[1] * k:
                raise ValueError("the HR image must be at least `scale` times the LR image")
            table[i] = (lq.data_ptr(), gt.data_ptr(), lq.shape[1] | (gt.shape[1] << 32), xs | (ys << 32), flags)
        dev_table = torch.from_numpy(table).to(self.device)
        x = torch.empty((n, 3, s, s), dtype=torch.float32, device=self.device)
        y = torch.empty((n, 3, s * k, s * k), dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            _lib.check(lib.ssr_augment_pairs_u8(dev_table.data_ptr(), n, s, k, x.data_ptr(), y.data_ptr(),
                                                torch.cuda.current_stream(self.device).cuda_stream))
        self._keep = (lqs, gts, dev_table)  # the launch is asynchronous: keep its inputs alive until the next call
        return x, y
