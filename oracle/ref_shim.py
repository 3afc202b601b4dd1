"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (studiosr_b200/).

Makes the *unmodified* reference package at /root/reference importable in this
container, where three of its third-party dependencies are absent (no network):

  * timm.layers.DropPath / trunc_normal_  (call sites swinir.py:7,75,137,335; hat.py:9)
  * gdown                                  (edsr.py:4, hat.py:5, utils/helpers.py:8)
  * skimage.metrics                        (utils/metrics.py:4)

The shim only provides the missing names; none of the reference's own arithmetic
is replaced.  DropPath follows timm's published semantics (per-sample Bernoulli
keep-mask scaled by 1/keep_prob, identity in eval / p == 0).

/root/reference exists only in the build container, NOT on the GPU box: this
module is used solely by oracle/make_golden.py (fixture generation) and by the
optional `reference_available()` cross-checks in the CPU test-suite.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("STUDIOSR_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "studiosr", "models"))


def _install_shims() -> None:
    import torch.nn as nn

    if "timm" not in sys.modules:
        try:
            import timm  # noqa: F401
        except Exception:
            timm = types.ModuleType("timm")
            layers = types.ModuleType("timm.layers")

            class DropPath(nn.Module):
                def __init__(self, drop_prob: float = 0.0, scale_by_keep: bool = True):
                    super().__init__()
                    self.drop_prob = drop_prob
                    self.scale_by_keep = scale_by_keep

                def forward(self, x):
                    if self.drop_prob == 0.0 or not self.training:
                        return x
                    keep = 1.0 - self.drop_prob
                    mask = x.new_empty((x.shape[0],) + (1,) * (x.ndim - 1)).bernoulli_(keep)
                    if keep > 0.0 and self.scale_by_keep:
                        mask.div_(keep)
                    return x * mask

            def trunc_normal_(tensor, mean=0.0, std=1.0, a=-2.0, b=2.0):
                return nn.init.trunc_normal_(tensor, mean=mean, std=std, a=a, b=b)

            layers.DropPath = DropPath
            layers.trunc_normal_ = trunc_normal_
            timm.layers = layers
            sys.modules["timm"] = timm
            sys.modules["timm.layers"] = layers
    if "gdown" not in sys.modules:
        try:
            import gdown  # noqa: F401
        except Exception:
            sys.modules["gdown"] = types.ModuleType("gdown")
    if "skimage" not in sys.modules:
        try:
            import skimage.metrics  # noqa: F401
        except Exception:
            sk = types.ModuleType("skimage")
            skm = types.ModuleType("skimage.metrics")

            def _missing(*a, **k):
                raise RuntimeError("skimage is not installed in this image")

            skm.structural_similarity = _missing
            skm.peak_signal_noise_ratio = _missing
            sk.metrics = skm
            sys.modules["skimage"] = sk
            sys.modules["skimage.metrics"] = skm


def import_reference():
    """Return the reference's `studiosr` package (imported from REFERENCE_ROOT)."""
    if not reference_available():
        raise RuntimeError(f"reference not found under {REFERENCE_ROOT}")
    _install_shims()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import studiosr  # noqa: E402

    return studiosr
