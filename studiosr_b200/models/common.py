"""Host-side mirror of studiosr/models/common.py: the `Model` base class keeps the reference's
public surface (inference, inference_with_self_ensemble, get_model_config, get_training_config,
from_pretrained, export; common.py:29-98) while every forward is executed by libssr_b200."""
import os
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from .. import _lib
from ..native import NativeModel

# precision of the fp32 (non-autocast) forward: "tf32x3" = tcgen05 kind::tf32 with head / tail operand splits (three MMAs per
# k-step, fp32-level accuracy: 5e-5 max-abs on the cfg1 golden), "fp32" = CUDA-core FMA, "tf32" = single-pass tcgen05 kind::tf32
DEFAULT_FP32_MODE = os.environ.get("STUDIOSR_B200_FP32_MODE", "tf32x3")
TRUST_PARAM_VERSIONS = os.environ.get("STUDIOSR_B200_TRUST_PARAM_VERSIONS", "0") == "1"

GRAPH_MAX_LR_PIXELS = int(os.environ.get("STUDIOSR_B200_GRAPH_MAX_LR_PIXELS", str(4 * 96 * 96)))

# bumped whenever any nn.Module registers a parameter / buffer / submodule: validates the per-model tensor-list caches in O(1)
_STRUCT_VERSION = [0]


def _bump_struct_version(*_args):
    _STRUCT_VERSION[0] += 1
    return None


try:
    from torch.nn.modules.module import (register_module_buffer_registration_hook, register_module_module_registration_hook,
                                         register_module_parameter_registration_hook)

    register_module_parameter_registration_hook(_bump_struct_version)
    register_module_buffer_registration_hook(_bump_struct_version)
    register_module_module_registration_hook(_bump_struct_version)
    _HAVE_REG_HOOKS = True
except ImportError:  # no hooks: every call re-walks the module tree
    _HAVE_REG_HOOKS = False


def diverge_images(image: torch.Tensor) -> List[torch.Tensor]:
    """The 8 rot90 / fliplr variants of an HWC image, in the order of common.py:10-16."""
    out = []
    for k in range(4):
        r = torch.rot90(image, k=k, dims=[0, 1])
        out += [r, torch.fliplr(r)]
    return out


def converge_images(images: List[torch.Tensor]) -> torch.Tensor:
    """Inverse of diverge_images followed by the mean (common.py:19-26)."""
    back = []
    for i, im in enumerate(images):
        if i & 1:
            im = torch.fliplr(im)
        back.append(torch.rot90(im, k=i // 2, dims=[1, 0]))
    return torch.mean(torch.stack(back), dim=0)


class _NativeTrainStep(torch.autograd.Function):
    """`out = model(x)` / `loss.backward()` of the Trainer step (trainer.py:97-105) through libssr_b200:
    forward keeps every GEMM operand in a workspace, backward produces all parameter gradients (fp32, PyTorch
    layouts) with the tensor-core dgrad / wgrad kernels, and dL/dx when the input asks for it."""

    @staticmethod
    def forward(ctx, x, model, nat, drop_scale, *tensors):
        y, ws = nat.train_forward(x, [t.detach() for t in tensors], model.scale, drop_scale)
        ctx.nat, ctx.ws, ctx.shape, ctx.drop_scale, ctx.model = nat, ws, tuple(x.shape), drop_scale, model
        ctx.xdtype = x.dtype
        ctx.meta = [(t.shape, t.numel(), t.requires_grad) for t in tensors]
        return y

    @staticmethod
    def backward(ctx, dy):
        # one flat buffer: grads are views into it, every slice on a 16-byte boundary -- the layout engine.FusedAdam keeps its
        # parameters in, so the optimizer update is one launch and a data-parallel exchange one all-reduce
        total = sum((n + 3) // 4 * 4 for _, n, rg in ctx.meta if rg)
        flat = torch.zeros(total, dtype=torch.float32, device=dy.device)
        grads, off = [], 0
        for shape, n, rg in ctx.meta:
            if rg:
                grads.append(flat[off:off + n].view(shape))
                off += (n + 3) // 4 * 4
            else:
                grads.append(None)
        ctx.nat.train_backward(dy, grads, ctx.shape, ctx.ws, ctx.drop_scale)
        dx = ctx.nat.train_input_grad(ctx.shape, dy.device).to(ctx.xdtype) if ctx.needs_input_grad[0] else None  # (reads the workspace)
        ctx.ws = None
        sync = getattr(ctx.model, "_grad_sync", None)  # engine.DistributedDataParallel: the gradient mean over ranks
        if sync is not None:
            sync(flat)
        return (dx, None, None, None, *grads)


class _NativeForwardOnly(torch.autograd.Function):
    """Differentiable-looking forward for the modes whose backward kernels are not built (fp32 / tf32 contractions, HAT):
    the forward runs natively (so the reference's own shape tests, which call the model in train mode with grad
    enabled, pass unchanged), a backward through it fails loudly instead of returning something else."""

    @staticmethod
    def forward(ctx, x, model, precision, pad_mode, *params):
        ctx.why = f"{type(model).__name__} in '{precision}' mode"
        return model._native(x.device, precision).forward(x, model.scale, pad_mode)

    @staticmethod
    def backward(ctx, *grads):
        raise NotImplementedError(
            f"studiosr_b200: no backward kernels for {ctx.why}; the training path covers EDSR, RCAN, HAN and SwinIR under the Trainer's "
            "bf16 autocast (trainer.py:69,80) -- run under torch.autocast('cuda', dtype=torch.bfloat16) or set "
            "model.precision = 'bf16'")


class Model(nn.Module):
    ARCH = -1
    TRAINABLE = False  # True for the model families whose backward is built (EDSR, RCAN, SwinIR)

    def __init__(self, scale: int = 4, n_colors: int = 3, img_range: float = 1.0) -> None:
        super().__init__()
        self.scale: int = scale
        self.n_colors: int = n_colors
        self.img_range: float = img_range
        self.precision: Optional[str] = None  # None = auto: bf16 under bf16 autocast, else DEFAULT_FP32_MODE
        self._natives: Dict = {}
        # True skips the per-forward checksum (one device sync) of the packed-weight cache key: safe when parameters are only
        # changed through autograd-visible ops / load_state_dict, or when invalidate_native() is called after `.data` writes
        self.trust_param_versions: bool = False
        # CUDA-graph replay of the inference launch sequence: None = automatic (small inputs, where ~90 launches of 10-20 us
        # kernels are launch-bound: up to GRAPH_MAX_LR_PIXELS LR pixels per call), True / False = always / never
        self.cuda_graphs: Optional[bool] = None

    # ---- native plumbing ------------------------------------------------------------------
    def _native_config(self, precision: int) -> "_lib.ModelConfig":
        raise NotImplementedError

    def _float_tensors(self):
        """The floating-point state_dict tensors, cached: walking ~500 modules costs 1.4 ms per forward.  The cache is dropped
        whenever ANY module registers a parameter, buffer or submodule (global registration hooks bump _STRUCT_VERSION)."""
        c = self.__dict__.get("_ft_cache") if _HAVE_REG_HOOKS else None
        if c is None or c[0] != _STRUCT_VERSION[0]:
            c = (_STRUCT_VERSION[0], [t for t in self.state_dict(keep_vars=True).values() if t.is_floating_point()])
            self.__dict__["_ft_cache"] = c
        return c[1]

    def _param_version(self):
        """Key of the packed-weight cache.  `_version` alone misses writes through `.data` (p.data.mul_(), p.data.copy_(),
        p.data = ...), so the storage pointers and a checksum of the parameter values (ssr_tensors_checksum: two launches over a
        cached device pointer table + one 16-byte read-back, ~50 us) are part of the key; `trust_param_versions` skips the
        checksum, `invalidate_native()` forces a re-pack explicitly."""
        ts = self._float_tensors()
        key = tuple((t._version, t.data_ptr()) for t in ts)
        if self.trust_param_versions or TRUST_PARAM_VERSIONS:
            return key
        cuda = [t for t in ts if t.is_cuda and t.dtype == torch.float32 and t.is_contiguous()]
        if cuda:
            dev = cuda[0].device
            ptrs = tuple(t.data_ptr() for t in cuda)
            tab = self.__dict__.get("_ck_table")
            if tab is None or tab[0] != ptrs:
                n = len(cuda)
                tab = (ptrs, torch.tensor(ptrs, dtype=torch.int64, device=dev), torch.tensor([t.numel() for t in cuda], dtype=torch.int64, device=dev),
                       torch.empty(2 * n, dtype=torch.float64, device=dev), torch.empty(2, dtype=torch.float64, device=dev))
                self.__dict__["_ck_table"] = tab
            lib = _lib.load()
            with torch.cuda.device(dev):
                _lib.check(lib.ssr_tensors_checksum(tab[1].data_ptr(), tab[2].data_ptr(), len(cuda), tab[3].data_ptr(), tab[4].data_ptr(),
                                                    torch.cuda.current_stream(dev).cuda_stream))
            key += tuple(tab[4].tolist())
        return key

    def invalidate_native(self) -> None:
        """Drop the packed native copies of the weights (they are rebuilt by the next forward)."""
        self._natives = {}

    # ctypes handles cannot be copied or pickled: EMA copies (copy.deepcopy), torch.save(model) and DDP's module pickling
    # carry the parameters only and rebuild the native side lazily
    def __getstate__(self):
        state = self.__dict__.copy()
        state["_natives"] = {}
        state.pop("_ft_cache", None)
        state.pop("_ck_table", None)
        return state

    def __deepcopy__(self, memo):
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k in ("_ft_cache", "_ck_table"):
                continue
            new.__dict__[k] = {} if k == "_natives" else copy.deepcopy(v, memo)
        return new

    def _native(self, device, precision: str) -> NativeModel:
        device = torch.device(device)
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device(), precision)
        nat = self._natives.get(key)
        if nat is None:
            nat = NativeModel(self._native_config(_lib.PRECISIONS[precision]), device)
            self._natives[key] = nat
        ver = self._param_version()
        if nat.version != ver:
            nat.load_state(self.state_dict(keep_vars=True), ver)
        return nat

    def _resolve_precision(self, x: torch.Tensor) -> str:
        if self.precision is not None:
            return self.precision
        if torch.is_autocast_enabled() and torch.get_autocast_dtype("cuda") == torch.bfloat16:
            return "bf16"
        return DEFAULT_FP32_MODE

    def _pad_mode(self) -> int:
        return _lib.PAD_TRAIN if self.training else _lib.PAD_EVAL

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not x.is_cuda:
            raise RuntimeError(
                f"{type(self).__name__}: studiosr_b200 executes on a CUDA sm_100 (B200) device only; "
                "there is no CPU / PyTorch fallback path")
        assert x.dim() == 4 and x.shape[1] == self.n_colors, "expected [B, n_colors, H, W]"
        precision = self._resolve_precision(x)
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return self._train_forward(x, precision)
        return self._native(x.device, precision).forward(x, self.scale, self._pad_mode(), graph=self._use_graph(x.shape[0] * x.shape[2] * x.shape[3]))

    def _use_graph(self, lr_pixels: int) -> bool:
        if self.cuda_graphs is not None:
            return bool(self.cuda_graphs)
        return lr_pixels <= GRAPH_MAX_LR_PIXELS and not torch.cuda.is_current_stream_capturing()

    def _trainable(self) -> bool:
        return self.TRAINABLE

    def _train_forward(self, x: torch.Tensor, precision: str) -> torch.Tensor:
        """Differentiable forward (the Trainer's `model(x)`, trainer.py:101-102)."""
        # the backward kernels mirror the TRAINING branch of the forward (reflect padding, stochastic depth); a differentiable
        # call in eval mode keeps the eval forward and fails loudly on backward
        if precision != "bf16" or not self._trainable() or not self.training:
            params = [p for p in self.parameters() if p.requires_grad]
            return _NativeForwardOnly.apply(x, self, precision, self._pad_mode(), *params)
        named = {k: v for k, v in self.state_dict(keep_vars=True).items() if v.is_floating_point()}
        dev = torch.device(x.device)
        key = (dev.type, dev.index if dev.index is not None else torch.cuda.current_device(), precision, "train")
        nat = self._natives.get(key)
        if nat is None:
            nat = NativeModel(self._native_config(_lib.PRECISIONS[precision]), dev)
            nat.load_state(named, None)  # establishes the packed layouts; the values are re-packed on device every step
            nat.train_bind(named)
            self._natives[key] = nat
        names = list(named.keys())
        drop_scale = self._draw_drop_path(x.shape[0], dev) if self.training else None
        return _NativeTrainStep.apply(x, self, nat, drop_scale, *[named[k] for k in names])

    def _draw_drop_path(self, batch: int, device) -> Optional[torch.Tensor]:
        """Stochastic-depth factors of this step, or None (models without DropPath)."""
        return None

    def _device(self) -> torch.device:
        return next(self.parameters()).device

    # ---- reference API (common.py:36-98) -----------------------------------------------------
    @torch.inference_mode()
    def inference(self, image: np.ndarray) -> np.ndarray:
        """uint8 HWC -> uint8 HWC (common.py:36-48).  /255, forward, *255, round, clip and the uint8
        cast are fused into the first / last kernels; only the uint8 image crosses PCIe."""
        self.eval()
        dev = self._device()
        img = torch.from_numpy(np.ascontiguousarray(image)).to(dev)
        precision = self.precision or DEFAULT_FP32_MODE
        out = self._native(dev, precision).upscale_u8(img.unsqueeze(0), self.scale, graph=self._use_graph(img.shape[0] * img.shape[1]))[0]
        return out.cpu().numpy()

    @torch.inference_mode()
    def inference_with_self_ensemble(self, image: np.ndarray) -> np.ndarray:
        """8-way rot/flip self-ensemble (common.py:50-67)."""
        self.eval()
        dev = self._device()
        scale = 255.0 if self.img_range == 1.0 else 1.0
        img = torch.from_numpy(image.astype(np.float32) / scale).to(dev)
        # the 8 rot / flip variants come in at most two shapes (HxW and WxH): one batched native forward per shape instead of
        # eight launches sequences; samples of a batch are independent, so the result is unchanged
        variants = [im.permute(2, 0, 1).contiguous() for im in diverge_images(img)]
        outs = [None] * len(variants)
        by_shape = {}
        for i, v in enumerate(variants):
            by_shape.setdefault(tuple(v.shape), []).append(i)
        for idx in by_shape.values():
            y = self.forward(torch.stack([variants[i] for i in idx]))
            for k, i in enumerate(idx):
                outs[i] = y[k].permute(1, 2, 0)
        out = converge_images(outs) * scale
        return out.round().clip(0, 255).to(torch.uint8).cpu().numpy()

    @torch.inference_mode()
    def inference_tiled(self, image: np.ndarray, tile: int = 64, overlap: int = 16, precision: Optional[str] = None,
                        out: Optional[np.ndarray] = None) -> np.ndarray:
        """Full-frame tiled inference (BASELINE.json config 5; the reference has no tiler): host uint8
        HWC -> host uint8 HWC through ONE C-ABI call (H2D, batched tiles, blend, D2H)."""
        self.eval()
        dev = self._device()
        H, W, _ = image.shape
        if out is None:
            out = np.empty((H * self.scale, W * self.scale, 3), dtype=np.uint8)
        nat = self._native(dev, precision or self.precision or DEFAULT_FP32_MODE)
        return nat.upscale_tiled_u8_host(np.ascontiguousarray(image), out, self.scale, tile, overlap)

    def sharded_tiler(self, H: int, W: int, tile: int = 64, overlap: int = 16, precision: Optional[str] = None, group=None,
                      chunk: int = 0):
        """One frame's tile list strong-scaled over the ranks of a torch.distributed group (BASELINE.json config 5 at
        N > 1): returns a studiosr_b200.sharding.ShardedTiledUpscaler bound to this model's packed weights on
        this rank's device.  `upscale_host(frame, out)` on it is the multi-GPU form of `inference_tiled`."""
        import torch.distributed as dist

        from ..sharding import NativeTileBackend, ShardedTiledUpscaler

        self.eval()
        nat = self._native(self._device(), precision or self.precision or DEFAULT_FP32_MODE)
        return ShardedTiledUpscaler(NativeTileBackend(nat, H, W, self.scale, tile, overlap, chunk), dist, group)

    def get_model_config(self) -> Dict:
        return dict(scale=self.scale, n_colors=self.n_colors, img_range=self.img_range)

    def get_training_config(self) -> Dict:
        return dict()

    @classmethod
    def from_pretrained(cls, scale: int = 4) -> "Model":
        """Every model family overrides this with the reference's signature and file naming; the base class must not hand
        back a randomly initialised model as if it were pretrained (the CLI would write garbage images)."""
        raise NotImplementedError(f"{cls.__name__}.from_pretrained: no pretrained weights are defined for this class")

    @staticmethod
    def _load_pretrained_file(path: str):
        """torch.load of a weight file that must already exist (downloads are outside the native path, no network here)."""
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} not found (downloads are outside the native path; place the file there)")
        return torch.load(path, map_location="cpu")

    def export(self, path: Optional[str] = None, input_shape: List[int] = [1, 3, 256, 256], format: str = "onnx") -> str:
        raise NotImplementedError("ONNX export of the native path is out of scope (SURVEY.md §2, common.py:84-98)")


BaseModule = Model


def conv2d(in_channels: int, out_channels: int, kernel_size: int) -> nn.Module:
    return nn.Conv2d(in_channels, out_channels, kernel_size, padding=kernel_size // 2)


RGB_MEAN = (0.4488, 0.4371, 0.4040)


class MeanShift(nn.Conv2d):
    """Frozen 1x1 conv holding +-img_range*mean (common.py:108-121); executed as a fused affine."""

    def __init__(self, img_range: float, rgb_mean=RGB_MEAN, rgb_std=(1.0, 1.0, 1.0), sign: int = -1) -> None:
        super().__init__(3, 3, kernel_size=1)
        std = torch.tensor(rgb_std, dtype=torch.float32)
        self.weight.data = torch.eye(3).view(3, 3, 1, 1) / std.view(3, 1, 1, 1)
        self.bias.data = sign * img_range * torch.tensor(rgb_mean, dtype=torch.float32) / std
        for p in self.parameters():
            p.requires_grad = False


def upsampler_stages(scale: int, n_feats: int, num_out_ch: Optional[int] = None):
    """[(cout, r)] of the conv + PixelShuffle stages of common.py:124-137."""
    if num_out_ch is not None:
        return [(scale * scale * num_out_ch, scale)]
    if scale & (scale - 1) == 0:
        out, s = [], scale
        while s > 1:
            out.append((4 * n_feats, 2))
            s //= 2
        return out
    return [(scale * scale * n_feats, scale)]


class Upsampler(nn.Sequential):
    """Parameter container with the reference's Sequential indices (conv at 0, 2, ...)."""

    def __init__(self, scale: int, n_feats: int, num_out_ch: Optional[int] = None) -> None:
        mods = []
        cin = n_feats
        for cout, r in upsampler_stages(scale, n_feats, num_out_ch):
            mods += [conv2d(cin, cout, 3), nn.PixelShuffle(r)]
        super().__init__(*mods)


class Mlp(nn.Module):
    def __init__(self, in_features: int, hidden_features: int) -> None:
        super().__init__()
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.fc2 = nn.Linear(hidden_features, in_features)


class ResBlock(nn.Module):
    def __init__(self, n_feats: int, kernel_size: int, res_scale: float = 1.0) -> None:
        super().__init__()
        self.body = nn.Sequential(conv2d(n_feats, n_feats, kernel_size), nn.ReLU(True), conv2d(n_feats, n_feats, kernel_size))
        self.res_scale = res_scale


class ChannelAttention(nn.Module):
    """Parameter container with the reference's key names (common.py:156-170): avg-pool -> 1x1 conv -> ReLU ->
    1x1 conv -> sigmoid gate.  The arithmetic runs in libssr_b200 (pool + gate kernels)."""

    def __init__(self, channel: int, reduction: int = 16) -> None:
        super().__init__()
        self.avg_pool = nn.AdaptiveAvgPool2d(1)
        self.conv_du = nn.Sequential(nn.Conv2d(channel, channel // reduction, 1), nn.ReLU(True),
                                     nn.Conv2d(channel // reduction, channel, 1), nn.Sigmoid())

