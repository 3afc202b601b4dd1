#!/usr/bin/env python
"""Developer aid: time one conv3x3 configuration of the implicit-GEMM kernel through the op-level C ABI (the packing /
layout-conversion launches of the op wrapper are excluded: only the `gemm_tc_conv3x3` class of the launch profile is shown).
  python scripts/conv_bench.py B Cin Cout H W ps_r [act] [res]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from studiosr_b200 import _lib
from tests import gpu_util as G

B, Cin, Cout, H, W, ps = [int(v) for v in sys.argv[1:7]]
act = int(sys.argv[7]) if len(sys.argv) > 7 else 0
use_res = len(sys.argv) > 8 and sys.argv[8] == "1"
x = torch.randn(B, Cin, H, W, device="cuda")
w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05
b = torch.randn(Cout, device="cuda") * 0.1
res = torch.randn(B, Cout, H, W, device="cuda") if use_res else None
lib = _lib.load()
for _ in range(2):
    G.op_conv3x3("bf16", x, w, b, res, act, 1.0, ps)
lib.ssr_profile_begin()
for _ in range(5):
    G.op_conv3x3("bf16", x, w, b, res, act, 1.0, ps)
buf = _lib.ctypes.create_string_buffer(1 << 16)
_lib.check(lib.ssr_profile_end(buf, len(buf)))
prof = json.loads(buf.value.decode())
for k, v in prof.items():
    if "gemm" in k:
        print(f"{k}: {v['ms'] / v['launches']:.4f} ms/launch, {v['flops'] / v['ms'] / 1e9:.0f} TFLOP/s, {v['bytes'] / v['ms'] / 1e6:.0f} GB/s "
              f"(HALO={os.environ.get('STUDIOSR_B200_HALO')}, NO_TMA_STORE={os.environ.get('STUDIOSR_B200_NO_TMA_STORE')})")
