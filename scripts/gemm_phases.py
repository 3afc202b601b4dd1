"""Developer aid: per-CTA phase timing of gemm_tc (clock64 stamps) for one big linear launch."""
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
nct = ((M + 127) // 128) * ((N + 191) // 192)
dbg = torch.zeros(nct * 8, dtype=torch.int64, device="cuda")
for it in range(2):
    lib.ssr_debug_set_buffer(dbg.data_ptr())
    G.op_linear("bf16", x, W, b, r, 0, lw, lw)
    lib.ssr_debug_set_buffer(None)
d = dbg.view(-1, 8).cpu().double()
t0 = d[:, 0].min()
print("CTAs", nct, "span cycles", (d[:, 4].max() - t0).item())
print("setup->epi_wait_start %.0f  wait_tmem_full %.0f  epilogue %.0f  tail %.0f  total %.0f" % (
    (d[:, 1] - d[:, 0]).mean(), (d[:, 2] - d[:, 1]).mean(), (d[:, 3] - d[:, 2]).mean(), (d[:, 4] - d[:, 3]).mean(), (d[:, 4] - d[:, 0]).mean()))
sm0 = d[d[:, 5] == d[0, 5]]
sm0 = sm0[sm0[:, 0].argsort()]
print("timeline on one SM (start, total, epi):")
for row in sm0[:12]:
    print(int(row[0] - t0), int(row[4] - row[0]), int(row[3] - row[2]))
