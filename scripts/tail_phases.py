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
dbg = torch.zeros(2 * 148 * 32 * 16, dtype=torch.int64, device="cuda")
ptrs = [t.data_ptr() for t in (o, res, Wp, bp, g2, be2, W1, b1, W2, b2, g3, be3)]
for it in range(2):
    dbg.zero_()
    lib.ssr_debug_set_buffer(dbg.data_ptr())
    _lib.check(lib.ssr_op_swin_mlp(*ptrs, yo.data_ptr(), ylo.data_ptr(), M, C, heads, hid, ws.data_ptr(), ws.numel(), G.stream()))
    torch.cuda.synchronize()
    lib.ssr_debug_set_buffer(None)
d = dbg[:148 * 32 * 16].view(148, 32, 16).cpu().double()
d2 = dbg[148 * 32 * 16:].view(148, 32, 16).cpu().double()[:, 2:20]
x = d[:, 2:20]  # steady-state tiles
f = lambda a, b: (x[:, :, a] - x[:, :, b]).mean().item()
print("per tile (cycles): wait PFULL %.0f | E1 main %.0f | E1 LN+xn2 %.0f | wait X0 %.0f gelu0 %.0f | wait X1 %.0f gelu1 %.0f | wait X2 %.0f gelu2 %.0f | wait Y %.0f | E3 %.0f"
      % (f(1, 0), f(2, 1), f(3, 2), f(4, 3), f(5, 4), f(6, 5), f(7, 6), f(8, 7), f(9, 8), f(10, 9), f(11, 10)))
print("E3 detail: tmem+fp32 stage %.0f | stats+barriers %.0f | pack+drain %.0f | bf16 out + res issue %.0f" % (f(13, 10), f(14, 13), f(15, 14), f(11, 15)))
print("tile period %.0f" % ((x[:, 1:, 0] - x[:, :-1, 0]).mean().item()))
print("E1 detail: tmem ld %.0f |" % (d2[:, :, 0] - x[:, :, 1]).mean().item(), " ".join("wait %.0f work %.0f |" % ((d2[:, :, 1 + 2 * k] - (d2[:, :, 2 * k] if k else d2[:, :, 0])).mean().item(), (d2[:, :, 2 + 2 * k] - d2[:, :, 1 + 2 * k]).mean().item()) for k in range(3)), "rest %.0f" % (x[:, :, 2] - d2[:, :, 6]).mean().item())
g2 = lambda a, b: (d2[:, :, a] - d2[:, :, b]).mean().item()
print("E3 fine: tmem ld+mask %.0f | 24 STS %.0f | fence.proxy.async %.0f | syncwarp %.0f | 3 TMA stores+commit %.0f | stats math %.0f | named barrier %.0f"
      % ((d2[:, :, 8] - x[:, :, 10]).mean().item(), g2(9, 8), g2(10, 9), g2(11, 10), g2(12, 11), g2(13, 12), g2(14, 13)))
print("GELU chunk 0: tmem ld %.0f of %.0f" % ((d2[:, :, 15] - x[:, :, 4]).mean().item(), f(5, 4)))
