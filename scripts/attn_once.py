"""Developer aid: launch the fused attention kernel a few times on a frame-sized synthetic input (for ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from studiosr_b200 import _lib
from tests import gpu_util as G

lib = _lib.load()
B, H, W, C, heads = 48, 72, 72, 180, 6
shift = int(sys.argv[1]) if len(sys.argv) > 1 else 4
g = torch.Generator().manual_seed(0)
xn = torch.randn(B, H, W, C, generator=g).cuda()
Wq = (torch.randn(3 * C, C, generator=g) / C**0.5).cuda(); bq = (torch.randn(3 * C, generator=g) * 0.2).cuda()
table = (torch.randn(225, heads, generator=g) * 0.5).cuda()
o = torch.empty(B, H, W, C, device="cuda")
ws = torch.empty(B * H * W * 192 * 8 + (1 << 22), dtype=torch.uint8, device="cuda")
for it in range(3):
    _lib.check(lib.ssr_op_swin_attn(xn.data_ptr(), Wq.data_ptr(), bq.data_ptr(), table.data_ptr(), o.data_ptr(), B, H, W, C, heads, shift,
                                    ws.data_ptr(), ws.numel(), G.stream()))
torch.cuda.synchronize()
print("ok", float(o.abs().mean()))
