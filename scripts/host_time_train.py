#!/usr/bin/env python
"""Developer aid: where the HOST time of a training step goes (Python prologue, native enqueue of forward / backward,
optimizer) next to the device time of the step -- if their sum approaches the step time the step is launch-bound."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.nn.functional as F

from studiosr_b200 import native
from studiosr_b200.models import EDSR, SwinIR

wl = sys.argv[1] if len(sys.argv) > 1 else "cfg4"
model, B, H, W = (SwinIR(scale=4), 32, 64, 64) if wl == "cfg4" else (EDSR(scale=4), 16, 48, 48)
model = model.cuda().train()
opt = torch.optim.Adam(model.parameters(), lr=1e-4)
x = torch.rand(B, 3, H, W).cuda()
y = torch.rand(B, 3, 4 * H, 4 * W).cuda()
acc = {}


def timed(name, fn):
    def w(*a, **k):
        t = time.perf_counter()
        r = fn(*a, **k)
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t
        return r
    return w


native.NativeModel.train_forward = timed("enqueue_forward", native.NativeModel.train_forward)
native.NativeModel.train_backward = timed("enqueue_backward", native.NativeModel.train_backward)
type(model)._train_forward = timed("python_forward_total", type(model)._train_forward)
opt.step = timed("optimizer_step", opt.step)


def step():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        loss = F.l1_loss(model(x), y)
    t = time.perf_counter()
    loss.backward()
    acc["backward_total"] = acc.get("backward_total", 0.0) + time.perf_counter() - t
    opt.step()
    opt.zero_grad(set_to_none=True)


for _ in range(3):
    step()
torch.cuda.synchronize()
acc.clear()
n = 10
t0 = time.perf_counter()
for _ in range(n):
    step()
host = time.perf_counter() - t0
torch.cuda.synchronize()
wall = time.perf_counter() - t0
print(f"{wl}: wall {1e3 * wall / n:.2f} ms/step, host loop {1e3 * host / n:.2f} ms/step")
for k, v in sorted(acc.items()):
    print(f"   {k:24s} {1e3 * v / n:8.2f} ms/step")
