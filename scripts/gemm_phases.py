"""Developer aid: per-item epilogue phase timing of gemm_tc (clock64 stamps) for one big linear launch."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from studiosr_b200 import _lib
from tests import gpu_util as G

lib = _lib.load()
lib.ssr_debug_set_buffer.argtypes = [ctypes.c_void_p]
M, K, N = 8910 * 128 // 4, 180, int(sys.argv[1]) if len(sys.argv) > 1 else 540
res = len(sys.argv) > 2
g = torch.Generator().manual_seed(0)
x = torch.randn(M, K, generator=g).cuda(); W = (torch.randn(N, K, generator=g) / K**0.5).cuda(); b = torch.randn(N, generator=g).cuda()
r = torch.randn(M, N, generator=g).cuda() if res else None
lw = torch.ones(N).cuda() if res else None
dbg = torch.zeros(148 * 64 * 16, dtype=torch.int64, device="cuda")
for it in range(2):
    dbg.zero_()
    lib.ssr_debug_set_buffer(dbg.data_ptr())
    G.op_linear("bf16", x, W, b, r, 0, lw, lw)
    lib.ssr_debug_set_buffer(None)
d = dbg.view(148, 64, 16).cpu().double()
ok = d[:, :, 0] > 0
dd = d[ok]
print("items sampled", dd.shape[0])
f = lambda a, b: (dd[:, a] - dd[:, b]).mean().item()
print("wait_full %.0f  chunks %.0f  ln %.0f | chunk1: tmem_ld %.0f math+res %.0f stores %.0f" % (f(1, 0), f(2, 1), f(3, 2), f(5, 4), f(6, 5), f(7, 6)))
c0 = d[0]
print("CTA0 item timeline (start rel, wait, chunks, ln):")
t0 = c0[0, 0]
for i in range(0, 12):
    if c0[i, 0] > 0:
        print(i, int(c0[i, 0] - t0), int(c0[i, 1] - c0[i, 0]), int(c0[i, 2] - c0[i, 1]), int(c0[i, 3] - c0[i, 2]))
