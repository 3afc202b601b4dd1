"""Developer aid: per-tile phase timing (clock64 stamps) of the fused Swin tail kernel (two tiles in flight):
IO warp 4, GELU warp 12 and the MMA issuer of every CTA."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from studiosr_b200 import _lib
from tests import gpu_util as G

lib = _lib.load()
lib.ssr_debug_set_buffer.argtypes = [ctypes.c_void_p]
M, C, heads, hid = 148 * 128 * 20, 180, 6, 360
g = torch.Generator().manual_seed(0)
r = lambda *s: torch.randn(*s, generator=g).cuda()
o, res = r(M, C), r(M, C)
Wp, W1, W2 = r(C, C) / C**0.5, r(hid, C) / C**0.5, r(C, hid) / hid**0.5
bp, b1, b2, g2, be2, g3, be3 = r(C), r(hid), r(C), r(C), r(C), r(C), r(C)
yo = torch.empty(M, C, device="cuda"); ylo = torch.empty(M, C, device="cuda")
ws = torch.empty(M * 192 * 16 + (1 << 22), dtype=torch.uint8, device="cuda")
R = 148 * 32 * 16
dbg = torch.zeros(3 * R, dtype=torch.int64, device="cuda")
ptrs = [t.data_ptr() for t in (o, res, Wp, bp, g2, be2, W1, b1, W2, b2, g3, be3)]
for it in range(2):
    dbg.zero_()
    lib.ssr_debug_set_buffer(dbg.data_ptr())
    _lib.check(lib.ssr_op_swin_mlp(*ptrs, yo.data_ptr(), ylo.data_ptr(), M, C, heads, hid, ws.data_ptr(), ws.numel(), G.stream()))
    torch.cuda.synchronize()
    lib.ssr_debug_set_buffer(None)
io, ge, mm = (dbg[k * R:(k + 1) * R].view(148, 32, 16).cpu().double()[:, 3:18] for k in range(3))
f = lambda x, a, b: (x[:, :, a] - x[:, :, b]).mean().item()
print("tile period %.0f cycles" % ((io[:, 1:, 0] - io[:, :-1, 0]).mean().item()))
print("IO warps  E1: wait PFULL %.0f | acc + residual (waits) + Y store %.0f | stats + wait xn2 free %.0f | xn2 -> smem %.0f  (E1 total %.0f)"
      % (f(io, 1, 0), f(io, 2, 1), f(io, 3, 2), f(io, 4, 3), f(io, 4, 0)))
print("IO warps  E3: wait YFULL %.0f | Y ld + fp32 stage + 3 stores %.0f | LN stats %.0f | drain %.0f | bf16 out + residual refill %.0f  (E3 total %.0f)"
      % (f(io, 9, 8), f(io, 10, 9), f(io, 11, 10), f(io, 12, 11), f(io, 13, 12), f(io, 13, 8)))
print("GELU warps: " + " | ".join("c%d wait %.0f work %.0f" % (c, f(ge, 2 * c + 1, 2 * c), (ge[:, :, 2 * c + 2] - ge[:, :, 2 * c + 1]).mean().item()) for c in range(6)))
print("MMA issuer: wait XNREADY %.0f | fc1 c0,c1 %.0f | (YFREE wait +) proj(next) %.0f | " % (f(mm, 1, 0), f(mm, 2, 1), f(mm, 3, 2))
      + " ".join("c%d %.0f" % (c, f(mm, 4 + c, 3 + c)) for c in range(6)) + " | period %.0f" % ((mm[:, 1:, 0] - mm[:, :-1, 0]).mean().item()))
