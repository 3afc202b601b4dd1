import ctypes, sys, os, json
sys.path.insert(0, "/root/repo")
import torch
from studiosr_b200 import _lib
from tests import gpu_util as G
lib = _lib.load()
C, heads, hid = 180, 6, 360
for M in (148 * 128 * 20, 220 * 72 * 72, 148 * 128 * 60):
    g = torch.Generator().manual_seed(0)
    r = lambda *s: torch.randn(*s, generator=g).cuda()
    o, res = r(M, C), r(M, C)
    Wp, W1, W2 = r(C, C) / C**0.5, r(hid, C) / C**0.5, r(C, hid) / hid**0.5
    bp, b1, b2, g2, be2, g3, be3 = r(C), r(hid), r(C), r(C), r(C), r(C), r(C)
    yo = torch.empty(M, C, device="cuda"); ylo = torch.empty(M, C, device="cuda")
    ws = torch.empty(M * 192 * 16 + (1 << 22), dtype=torch.uint8, device="cuda")
    ptrs = [t.data_ptr() for t in (o, res, Wp, bp, g2, be2, W1, b1, W2, b2, g3, be3)]
    for it in range(3):
        lib.ssr_profile_begin()
        _lib.check(lib.ssr_op_swin_mlp(*ptrs, yo.data_ptr(), ylo.data_ptr(), M, C, heads, hid, ws.data_ptr(), ws.numel(), G.stream()))
        buf = ctypes.create_string_buffer(1 << 16)
        _lib.check(lib.ssr_profile_end(buf, len(buf)))
    p = json.loads(buf.value.decode())["swin_tail"]
    tiles = (M + 127) // 128
    per = p["ms"] * 1e-3 / ((tiles + 147) // 148)
    print(f"M={M} tiles/CTA={(tiles+147)//148} swin_tail {p['ms']:.4f} ms -> {per*1e6:.2f} us per tile-wave")
