#!/usr/bin/env python
"""Secondary benchmark lines (not the driver's headline): BASELINE.json configs 2, 3 and 4.

  python scripts/bench_extra.py --workload cfg1   # SwinIR-x4 single 64x64 image latency (bf16)
  python scripts/bench_extra.py --workload cfg2   # EDSR-x4 fwd+bwd, 16 x 3x48x48, bf16 autocast, L1 loss (+ Adam step)
  python scripts/bench_extra.py --workload cfg3   # HAT-x4 bf16 inference, batch 32 of 64x64 LR tiles
  python scripts/bench_extra.py --workload cfg4   # SwinIR-x4 Trainer step, 32 x 3x64x64 per GPU (torchrun: DDP)

Prints one JSON line: ms per step (CUDA events, after warm-up), algorithmic TFLOP/s (SURVEY 8d FLOP counts) as a
fraction of the measured bf16 peak, output Mpix/s, and the per-kernel-class profile from ssr_profile_begin/end."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4"])
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-optimizer", action="store_true")
    args = ap.parse_args()
    import torch
    import torch.nn.functional as F

    from bench import peaks
    from studiosr_b200 import _lib
    from studiosr_b200.models import EDSR, HAT, SwinIR

    torch.manual_seed(0)
    if args.workload == "cfg2":
        model, B, H, W = EDSR(scale=4), 16, 48, 48
        step_flops = 16 * 694.66e9  # fwd+bwd per 48x48 patch
        name = "EDSR-x4 forward+backward (+Adam), batch 16 of 48x48 LR patches, bf16 autocast, L1 loss"
    elif args.workload == "cfg1":
        model, B, H, W = SwinIR(scale=4), 1, 64, 64
        step_flops = 135.56e9  # one 64x64 image, padded to 72x72 by the eval forward
        name = "SwinIR-x4 inference latency, one 3x64x64 LR image (BASELINE config 1, the reference's CPU-runnable case), bf16"
    elif args.workload == "cfg3":
        model, B, H, W = HAT(scale=4), 32, 64, 64
        step_flops = 32 * 207.76e9  # forward per 64x64 tile
        name = "HAT-x4 bf16 inference, batch 32 of 64x64 LR tiles (overlapping cross-attention + channel attention)"
    else:
        model, B, H, W = SwinIR(scale=4), 32, 64, 64
        step_flops = 32 * 321.30e9
        name = "SwinIR-x4 Trainer step, batch 32 of 64x64 LR patches per GPU, bf16 autocast, L1 loss"
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    infer = args.workload in ("cfg1", "cfg3")
    model = model.cuda().eval() if infer else model.cuda().train()
    if infer:
        model.precision = "bf16"
    dist = None
    if world > 1 and not infer:  # data-parallel replicas, gradient all-reduce by DDP over NCCL (trainer.py:89-91)
        import torch.distributed as dist
        from torch.nn.parallel import DistributedDataParallel as DDP

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        model = DDP(model, device_ids=[local], output_device=local)
    opt = torch.optim.Adam(model.parameters(), lr=1e-4, betas=(0.9, 0.99))
    g = torch.Generator().manual_seed(1234 + rank)
    x = torch.rand(B, 3, H, W, generator=g).cuda()
    y = torch.rand(B, 3, 4 * H, 4 * W, generator=g).cuda()
    lib = _lib.load()

    def step():
        if infer:
            with torch.inference_mode():
                return model(x).sum()
        with torch.autocast("cuda", dtype=torch.bfloat16):
            loss = F.l1_loss(model(x), y)
        loss.backward()
        if not args.no_optimizer:
            opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if dist is not None:
        dist.barrier()
        torch.cuda.synchronize()
    l0 = lib.ssr_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    if dist is not None:  # max over ranks of the device-side time
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    launches = (lib.ssr_launch_count() - l0) // args.steps
    if rank != 0:  # the profiled step below still all-reduces: every rank takes part, rank 0 reports
        step()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        return
    lib.ssr_profile_begin()
    step()
    buf = _lib.ctypes.create_string_buffer(1 << 16)
    _lib.check(lib.ssr_profile_end(buf, len(buf)))
    prof = json.loads(buf.value.decode())
    kernels = {k: {"launches": v["launches"], "ms": round(v["ms"], 4), "tflops": round(v["flops"] / v["ms"] / 1e9, 1) if v["ms"] > 0 else 0.0,
                   "gbs": round(v["bytes"] / v["ms"] / 1e6, 1) if v["ms"] > 0 else 0.0} for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
    pk = peaks()
    step_flops *= world
    tf = step_flops / (ms / 1e3) / 1e12
    print(json.dumps({
        "metric": f"{args.workload}_{'inference' if infer else 'train'}_step", "n_gpus": world, "ms_per_step": ms, "achieved_tflops": tf,
        "frac_of_sustained_peak": tf / pk["tf_sust"] / world, "scaling": "weak",
        "output_mpix_per_s": world * B * 16 * H * W / 1e6 / (ms / 1e3), "alg_tflop_per_step": step_flops / 1e12, "loss": float(loss.detach()),
        "launches_per_step": int(launches), "optimizer_in_step": (not args.no_optimizer) and not infer, "config": {"workload": name},
        "profiled_kernel_ms": sum(v["ms"] for v in prof.values()), "kernels": kernels}))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
