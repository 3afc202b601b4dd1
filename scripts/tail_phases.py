"""Developer aid: per-tile phase timing (clock64 stamps of epilogue warp 0) of the fused Swin tail kernel."""
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
dbg = torch.zeros(148 * 32 * 16, dtype=torch.int64, device="cuda")
ptrs = [t.data_ptr() for t in (o, res, Wp, bp, g2, be2, W1, b1, W2, b2, g3, be3)]
for it in range(2):
    dbg.zero_()
    lib.ssr_debug_set_buffer(dbg.data_ptr())
    _lib.check(lib.ssr_op_swin_mlp(*ptrs, yo.data_ptr(), ylo.data_ptr(), M, C, heads, hid, ws.data_ptr(), ws.numel(), G.stream()))
    torch.cuda.synchronize()
    lib.ssr_debug_set_buffer(None)
d = dbg.view(148, 32, 16).cpu().double()
x = d[:, 2:20]  # steady-state tiles
f = lambda a, b: (x[:, :, a] - x[:, :, b]).mean().item()
print("per tile (cycles): wait PFULL %.0f | E1 main %.0f | E1 LN+xn2 %.0f | wait X0 %.0f gelu0 %.0f | wait X1 %.0f gelu1 %.0f | wait X2 %.0f gelu2 %.0f | wait Y %.0f | E3 %.0f"
      % (f(1, 0), f(2, 1), f(3, 2), f(4, 3), f(5, 4), f(6, 5), f(7, 6), f(8, 7), f(9, 8), f(10, 9), f(11, 10)))
print("tile period %.0f" % ((x[:, 1:, 0] - x[:, :-1, 0]).mean().item()))
