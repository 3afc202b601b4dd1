"""Developer aid: per-item epilogue phase timing of gemm_tc for the RSTB conv (180 -> 180, residual + LN-free variant)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from studiosr_b200 import _lib
from tests import gpu_util as G

lib = _lib.load()
lib.ssr_debug_set_buffer.argtypes = [ctypes.c_void_p]
B, C, H, W = 64, 180, 72, 72
res = len(sys.argv) > 1
g = torch.Generator().manual_seed(0)
x = torch.randn(B, C, H, W, generator=g).cuda(); Wt = (torch.randn(C, C, 3, 3, generator=g) / (9 * C) ** 0.5).cuda(); b = torch.randn(C, generator=g).cuda()
r = torch.randn(B, C, H, W, generator=g).cuda() if res else None
dbg = torch.zeros(148 * 64 * 16, dtype=torch.int64, device="cuda")
for it in range(2):
    dbg.zero_()
    lib.ssr_debug_set_buffer(dbg.data_ptr())
    G.op_conv3x3("bf16", x, Wt, b, r)
    lib.ssr_debug_set_buffer(None)
d = dbg.view(148, 64, 16).cpu().double()
ok = d[:, :, 0] > 0
dd = d[ok]
f = lambda a, b: (dd[:, a] - dd[:, b]).mean().item()
print("items sampled", dd.shape[0], "| wait_full %.0f  chunks %.0f  ln %.0f | chunk1: tmem_ld %.0f math+res %.0f stores %.0f" % (f(1, 0), f(2, 1), f(3, 2), f(5, 4), f(6, 5), f(7, 6)))
c0 = d[0]
t0 = c0[0, 0]
print("CTA0 timeline (start, wait, chunks):", [(int(c0[i, 0] - t0), int(c0[i, 1] - c0[i, 0]), int(c0[i, 2] - c0[i, 1])) for i in range(0, 10) if c0[i, 0] > 0])
