"""Drop-ins for the remaining pieces of the reference's Trainer step (studiosr/engine/trainer.py:89-109,133-145):

    with ctx: out = model(x); loss = criterion(out, y)        criterion  -> L1Loss            (nn.L1Loss, trainer.py:45)
    loss.backward()                                           DDP        -> DistributedDataParallel (trainer.py:89-91)
    optimizer.step(); optimizer.zero_grad(set_to_none=True)   optimizer  -> FusedAdam          (torch.optim.Adam, :133-139)
    scheduler.step()                                          unchanged  (MultiStepLR only rewrites param_groups[i]["lr"])

The native backward (models/common.py:_NativeTrainStep) returns every parameter gradient as a view of ONE flat fp32 buffer;
everything here works on that buffer: the data-parallel exchange is a single NCCL all-reduce (no reducer buckets, no gradient
copies, `broadcast_buffers` has nothing to do -- the models have no trainable buffers), and Adam is one kernel over flat
parameter / moment buffers (ssr_adam_step, 28 B per parameter).  A reference maintainer re-points three names in
trainer.py; the Trainer loop itself does not change.  No CPU fallback: these classes need CUDA tensors."""
from typing import Iterable, List, Optional

import torch
import torch.nn as nn

from . import _lib


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


# ----------------------------------------------------------------------------------------------------------------------
class _L1Fused(torch.autograd.Function):
    """mean |out - y| with the backward seed sign(out - y) / N produced by the same pass."""

    @staticmethod
    def forward(ctx, out, y):
        lib = _lib.load()
        o = out.detach().float().contiguous()
        t = y.detach().float().contiguous()
        loss = torch.empty((), dtype=torch.float32, device=o.device)
        seed = torch.empty_like(o) if out.requires_grad else None
        ws = torch.empty(lib.ssr_l1_loss_workspace_bytes(), dtype=torch.uint8, device=o.device)
        with torch.cuda.device(o.device):
            _lib.check(lib.ssr_l1_loss(o.data_ptr(), t.data_ptr(), o.numel(), loss.data_ptr(), None if seed is None else seed.data_ptr(),
                                       ws.data_ptr(), ws.numel(), _stream(o.device)))
        ctx.seed, ctx.dtype = seed, out.dtype
        return loss

    @staticmethod
    def backward(ctx, g):
        seed, ctx.seed = ctx.seed, None
        if seed is None:
            return None, None
        return seed.mul_(g).to(ctx.dtype), None  # g is the 0-dim root gradient (1 for loss.backward()): one in-place scale


class L1Loss(nn.Module):
    """nn.L1Loss() (the Trainer's default criterion, trainer.py:45) with the loss and its backward seed in one kernel."""

    def forward(self, out: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
        if not out.is_cuda:
            raise RuntimeError("studiosr_b200.engine.L1Loss runs on CUDA tensors only (no CPU fallback)")
        return _L1Fused.apply(out, y)


# ----------------------------------------------------------------------------------------------------------------------
class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam(params, lr, betas, eps, weight_decay) (trainer.py:133-139) as one kernel per step.

    Construction moves the parameters into one flat fp32 buffer (each `p.data` becomes a view of it, values unchanged) and
    allocates flat exp_avg / exp_avg_sq; `state_dict()` keeps torch.optim.Adam's layout (per-parameter step / exp_avg /
    exp_avg_sq), so Trainer.save / Trainer.load checkpoints are interchangeable with the stock optimizer.  When the
    gradients of a step are the views of the flat buffer the native backward filled (the normal case) the update is ONE
    launch; otherwise each parameter's slice is updated by its own launch of the same kernel."""

    def __init__(self, params: Iterable, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay)
        super().__init__(params, defaults)
        self._lib = _lib.load()
        ps = [p for g in self.param_groups for p in g["params"]]
        if not ps or not all(p.is_cuda and p.dtype == torch.float32 for p in ps):
            raise RuntimeError("studiosr_b200.engine.FusedAdam needs fp32 CUDA parameters (no CPU fallback)")
        self._device = ps[0].device
        # every slice starts on a 16-byte boundary so that the float4 kernel may also run per parameter
        self._offsets, total = {}, 0
        for p in ps:
            self._offsets[p] = total
            total += (p.numel() + 3) // 4 * 4
        self._total = total
        self._flat_p = torch.zeros(total, dtype=torch.float32, device=self._device)
        self._flat_m = torch.zeros_like(self._flat_p)
        self._flat_v = torch.zeros_like(self._flat_p)
        for p in ps:
            o, n = self._offsets[p], p.numel()
            self._flat_p[o:o + n].copy_(p.data.reshape(-1))
            p.data = self._flat_p[o:o + n].view(p.shape)
        self._step = 0
        self._step_t = torch.tensor(0.0)  # ONE host tensor shared by every parameter's state["step"] (torch.optim.Adam's layout)
        self._ps = ps
        self._flat_ok = False  # the full pointer check of the flat gradient layout has passed once

    def _state_for(self, p):
        st = self.state[p]
        if not st:
            o, n = self._offsets[p], p.numel()
            st["step"] = self._step_t
            st["exp_avg"] = self._flat_m[o:o + n].view(p.shape)
            st["exp_avg_sq"] = self._flat_v[o:o + n].view(p.shape)
        return st

    def _launch(self, off: int, n: int, g_ptr: int, group, grad_scale: float) -> None:
        b1, b2 = group["betas"]
        with torch.cuda.device(self._device):
            _lib.check(self._lib.ssr_adam_step(self._flat_p.data_ptr() + 4 * off, g_ptr, self._flat_m.data_ptr() + 4 * off,
                                               self._flat_v.data_ptr() + 4 * off, n, float(group["lr"]), b1, b2, group["eps"],
                                               group["weight_decay"], self._step, grad_scale, _stream(self._device)))

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        self._step += 1
        self._step_t.fill_(float(self._step))
        ps = self._ps
        first, last = ps[0], ps[-1]
        # ONE launch when there is one hyper-parameter group and the gradients sit in one buffer at the parameters' own offsets --
        # the layout _NativeTrainStep produces.  The full pointer check runs once; afterwards the two ends are re-checked per step.
        one = len(self.param_groups) == 1 and first.grad is not None and last.grad is not None
        if one:
            base = first.grad.data_ptr() - 4 * self._offsets[first]
            one = base % 16 == 0 and last.grad.data_ptr() == base + 4 * self._offsets[last]
            if one and not self._flat_ok:
                one = all(p.grad is not None and p.grad.dtype == torch.float32 and p.grad.is_contiguous() and
                          p.grad.data_ptr() == base + 4 * self._offsets[p] for p in ps)
                self._flat_ok = one
        if one:
            self._launch(0, self._total, base, self.param_groups[0], grad_scale)
            return loss
        self._flat_ok = False
        for group in self.param_groups:
            for p in group["params"]:
                if p.grad is None:
                    continue
                grad = p.grad.float().contiguous()
                if grad.data_ptr() % 16:
                    grad = grad.clone()
                self._launch(self._offsets[p], p.numel(), grad.data_ptr(), group, grad_scale)
        return loss

    def grad_layout(self):
        """(total elements, {parameter: element offset}) of the flat buffers: a backward that writes gradients at these offsets of
        one buffer gets the single-launch update."""
        return self._total, dict(self._offsets)

    def state_dict(self):
        for g in self.param_groups:
            for p in g["params"]:
                self._state_for(p)
        return super().state_dict()

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        steps = []
        for g in self.param_groups:
            for p in g["params"]:
                st = self.state.get(p)
                if not st:
                    continue
                o, n = self._offsets[p], p.numel()
                self._flat_m[o:o + n].copy_(st["exp_avg"].reshape(-1))
                self._flat_v[o:o + n].copy_(st["exp_avg_sq"].reshape(-1))
                st["exp_avg"] = self._flat_m[o:o + n].view(p.shape)
                st["exp_avg_sq"] = self._flat_v[o:o + n].view(p.shape)
                steps.append(int(float(st["step"])))
                st["step"] = self._step_t
        if steps:
            self._step = max(steps)
            self._step_t.fill_(float(self._step))


# ----------------------------------------------------------------------------------------------------------------------
class DistributedDataParallel(nn.Module):
    """torch.nn.parallel.DistributedDataParallel(model, device_ids=[d], output_device=d) as the Trainer uses it
    (trainer.py:89-91) for the native training path: parameters are broadcast from rank 0 once, and the gradient mean over
    ranks is ONE all-reduce of the flat gradient buffer, enqueued on the compute stream right behind the native backward
    (47.6 MB for SwinIR: ~0.2 ms over NVLink, against the 5.5 ms that DDP's bucket copies + serialized all-reduce cost
    behind a single autograd node).  `no_sync()` and `.module` behave as in DDP."""

    def __init__(self, module: nn.Module, device_ids=None, output_device=None, process_group=None, broadcast_buffers: bool = False,
                 **_unused):
        super().__init__()
        import torch.distributed as dist

        if not dist.is_initialized():
            raise RuntimeError("DistributedDataParallel needs torch.distributed to be initialised")
        self.module, self._dist, self._group, self._sync = module, dist, process_group, True
        self._world = dist.get_world_size(process_group)
        src = dist.get_global_rank(process_group, 0) if process_group is not None else 0
        with torch.no_grad():
            ts = [t for t in module.state_dict(keep_vars=True).values() if t.is_floating_point()]
            if ts:  # one flat broadcast instead of one per tensor
                flat = torch.cat([t.detach().reshape(-1).float() for t in ts])
                dist.broadcast(flat, src=src, group=process_group)
                off = 0
                for t in ts:
                    t.detach().copy_(flat[off:off + t.numel()].view(t.shape))
                    off += t.numel()
        module._grad_sync = self._all_reduce_flat

    def _all_reduce_flat(self, flat: torch.Tensor) -> None:
        if self._sync and self._world > 1:
            self._dist.all_reduce(flat, op=self._dist.ReduceOp.AVG, group=self._group)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def no_sync(self):
        ddp = self

        class _NoSync:
            def __enter__(self):
                self.prev, ddp._sync = ddp._sync, False

            def __exit__(self, *exc):
                ddp._sync = self.prev

        return _NoSync()

    def state_dict(self, *args, **kwargs):
        return self.module.state_dict(*args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        return self.module.load_state_dict(*args, **kwargs)


def build_optimizer(trainer) -> "tuple":
    """Trainer.build_optimizer (trainer.py:133-145) with FusedAdam in place of torch.optim.Adam; MultiStepLR is the stock one."""
    opt = FusedAdam(trainer.model.parameters(), lr=trainer.learning_rate, betas=trainer.betas, weight_decay=trainer.weight_decay)
    sch = torch.optim.lr_scheduler.MultiStepLR(opt, milestones=trainer.milestones, gamma=trainer.gamma)
    return opt, sch
