"""Developer aid: per-head-pair phase timing (clock64 stamps of epilogue warp 0) of the fused attention kernel."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from studiosr_b200 import _lib
from tests import gpu_util as G

lib = _lib.load()
lib.ssr_debug_set_buffer.argtypes = [ctypes.c_void_p]
B, H, W, C, heads = 220, 72, 72, 180, 6
shift = int(sys.argv[1]) if len(sys.argv) > 1 else 4
g = torch.Generator().manual_seed(0)
xn = torch.randn(B, H, W, C, generator=g).cuda()
Wq = (torch.randn(3 * C, C, generator=g) / C**0.5).cuda(); bq = (torch.randn(3 * C, generator=g) * 0.2).cuda()
table = (torch.randn(225, heads, generator=g) * 0.5).cuda()
o = torch.empty(B, H, W, C, device="cuda")
ws = torch.empty(B * H * W * 192 * 8 + (1 << 22), dtype=torch.uint8, device="cuda")
dbg = torch.zeros(148 * 96 * 16, dtype=torch.int64, device="cuda")
for it in range(2):
    dbg.zero_()
    lib.ssr_debug_set_buffer(dbg.data_ptr())
    _lib.check(lib.ssr_op_swin_attn(xn.data_ptr(), Wq.data_ptr(), bq.data_ptr(), table.data_ptr(), o.data_ptr(), B, H, W, C, heads, shift,
                                    ws.data_ptr(), ws.numel(), G.stream()))
    torch.cuda.synchronize()
    lib.ssr_debug_set_buffer(None)
d = dbg.view(148, 96, 16).cpu().double()[:, 6:90]
f = lambda a, b: (d[:, :, a] - d[:, :, b]).mean().item()
print("per head pair (cycles): wait QKV %.0f | QKV epilogue %.0f | previous pair's output %.0f | wait S %.0f | softmax %.0f | pair period %.0f (x3 per 128-token item)"
      % (f(1, 0), f(2, 1), f(3, 2), f(4, 3), f(5, 4), (d[:, 1:, 0] - d[:, :-1, 0]).mean().item()))

for hp in range(3):
    dd = d[:, hp::3]
    ff = lambda a, b: (dd[:, :, a] - dd[:, :, b]).mean().item()
    print("  head pair %d of the item: wait QKV %.0f | QKV epilogue %.0f | output %.0f | wait S %.0f | softmax %.0f" % (hp, ff(1, 0), ff(2, 1), ff(3, 2), ff(4, 3), ff(5, 4)))

# one CTA's timeline relative to the epilogue's start of pair g (stamp 0): epilogue 0..5 | MMA issuer 8..13 | store warp 14..15 (store of pair g's output)
names = ["e:start", "e:QKVFULL", "e:OPREADY sent", "e:out staged", "e:SFULL", "e:PREADY sent", "e:OFULL(g-1) seen", "e:OSTFREE seen", "m:iter start", "m:OPREADY seen", "m:S issued",
         "m:proj(g+1) issued", "m:PREADY0 seen", "m:PREADY1 seen", "s:OSTAGED(g) seen", "s:store read done"]
for hp in range(3):
    dd = d[:, hp::3]
    rel = (dd - dd[:, :, 0:1]).mean(dim=(0, 1))
    print("  pair %d timeline: " % hp + " | ".join("%s %+.0f" % (names[k], rel[k].item()) for k in (1, 2, 6, 7, 3, 4, 5, 8, 9, 10, 11, 12, 13, 14, 15)))
