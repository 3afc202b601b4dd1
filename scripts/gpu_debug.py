"""Developer aid (GPU box): run each kernel check in its own subprocess (a trap poisons the CUDA
context) and print compact error statistics.  Usage: python scripts/gpu_debug.py [names...]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CHECKS = {}


def check(fn):
    CHECKS[fn.__name__] = fn
    return fn


def _stats(name, y, ref):
    import torch

    y = y.float().cpu()
    d = (y - ref).abs()
    rms = ref.pow(2).mean().sqrt().item()
    print(f"{name}: max_err={d.max().item():.3e} mean_err={d.mean().item():.3e} ref_rms={rms:.3e} "
          f"nan={int(torch.isnan(y).sum())} shape={tuple(y.shape)}")
    if d.max().item() > 0.05 * rms and y.dim() == 2:
        bad_rows = (d.max(1).values > 0.05 * rms).nonzero().flatten()
        bad_cols = (d.max(0).values > 0.05 * rms).nonzero().flatten()
        print(f"   bad rows: {bad_rows.numel()}/{y.shape[0]} first {bad_rows[:12].tolist()}  "
              f"bad cols: {bad_cols.numel()}/{y.shape[1]} first {bad_cols[:12].tolist()}")
        print("   y[0,:8]  ", [round(v, 4) for v in y[0, :8].tolist()])
        print("   ref[0,:8]", [round(v, 4) for v in ref[0, :8].tolist()])


def _linear(prec, M, K, N, act=0, res=False, ln=False):
    import torch

    from oracle import sr_oracle as O
    from tests import gpu_util as G

    g = torch.Generator().manual_seed(1)
    x = torch.randn(M, K, generator=g)
    W = torch.randn(N, K, generator=g) / K**0.5
    b = torch.randn(N, generator=g) * 0.1
    r = torch.randn(M, N, generator=g) if res else None
    lw = 1 + 0.1 * torch.randn(N, generator=g) if ln else None
    lb = 0.1 * torch.randn(N, generator=g) if ln else None
    v = x @ W.t() + b
    if act == 3:
        v = O.gelu(v)
    if res:
        v = v + r
    c = lambda t: None if t is None else t.cuda()
    y, yl = G.op_linear(prec, c(x), c(W), c(b), c(r), act, c(lw), c(lb))
    _stats(f"linear[{prec}] M{M} K{K} N{N} act{act} res{int(res)} ln{int(ln)}", y, v)
    if ln:
        _stats("   +LN", yl, O.layer_norm(v, lw, lb))


@check
def lin_fp32():
    _linear("fp32", 300, 180, 540)
    _linear("fp32", 300, 180, 180, res=True, ln=True)


@check
def lin_bf16_small():
    _linear("bf16", 128, 64, 64)


@check
def lin_bf16_k192():
    _linear("bf16", 256, 180, 180)


@check
def lin_bf16_full():
    _linear("bf16", 5184, 180, 540)
    _linear("bf16", 777, 360, 180, res=True, ln=True)
    _linear("bf16", 1000, 180, 360, act=3)


@check
def lin_tf32():
    _linear("tf32", 128, 32, 64)
    _linear("tf32", 777, 360, 180, res=True, ln=True)


def _conv(prec, B, Cin, Cout, H, W, act=0, ps=0, res=False):
    import torch
    import torch.nn.functional as F

    from oracle import sr_oracle as O
    from tests import gpu_util as G

    g = torch.Generator().manual_seed(2)
    x = torch.randn(B, Cin, H, W, generator=g)
    Wt = torch.randn(Cout, Cin, 3, 3, generator=g) / (9 * Cin) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    r = torch.randn(B, Cout, H, W, generator=g) if res else None
    v = F.conv2d(x, Wt, b, padding=1)
    if act == 1:
        v = torch.relu(v)
    if res:
        v = v + r
    if ps > 1:
        v = O.pixel_shuffle(v, ps)
    y = G.op_conv3x3(prec, x.cuda(), Wt.cuda(), b.cuda(), None if r is None else r.cuda(), act, 1.0, ps)
    _stats(f"conv[{prec}] B{B} {Cin}->{Cout} {H}x{W} act{act} ps{ps} res{int(res)}", y.reshape(-1, y.shape[-1]),
           v.reshape(-1, v.shape[-1]))


@check
def conv_fp32():
    _conv("fp32", 1, 180, 180, 24, 24, res=True)
    _conv("fp32", 2, 64, 256, 16, 24, ps=2)


@check
def conv_bf16():
    _conv("bf16", 1, 64, 64, 16, 16)
    _conv("bf16", 1, 180, 180, 72, 72, res=True)
    _conv("bf16", 2, 64, 256, 16, 24, ps=2)
    _conv("bf16", 1, 64, 576, 9, 13, ps=3)


@check
def conv_tf32():
    _conv("tf32", 1, 180, 180, 24, 24, res=True)


def _attn(prec, B, H, W, C, heads, shift):
    import torch

    from oracle import sr_oracle as O
    from tests import gpu_util as G

    ws, d = 8, C // heads
    g = torch.Generator().manual_seed(3)
    qkv = torch.randn(B, H, W, 3 * C, generator=g)
    table = torch.randn((2 * ws - 1) ** 2, heads, generator=g) * 0.5
    q = torch.roll(qkv, (-shift, -shift), (1, 2)) if shift else qkv
    qw = O.to_windows(q, ws).reshape(-1, 64, 3, heads, d)
    Q, K, V = (qw[:, :, i].transpose(1, 2) for i in range(3))
    s = (Q * d**-0.5) @ K.transpose(-1, -2) + O.rel_pos_bias(table, ws)[None]
    mask = O.shift_mask(H, W, ws, shift, torch.float32)
    s = (s.reshape(B, mask.shape[0], heads, 64, 64) + mask[None, :, None]).reshape(-1, heads, 64, 64)
    o = O.from_windows((torch.softmax(s, -1) @ V).transpose(1, 2).reshape(-1, 64, C), ws, B, H, W)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    y = G.op_window_attention(prec, qkv.cuda(), table.cuda(), heads, ws, shift)
    _stats(f"attn[{prec}] B{B} {H}x{W} C{C} h{heads} shift{shift}", y.reshape(-1, C), o.reshape(-1, C))


@check
def attn_fp32():
    _attn("fp32", 1, 16, 24, 180, 6, 0)
    _attn("fp32", 2, 16, 24, 180, 6, 4)


@check
def attn_bf16():
    _attn("bf16", 1, 16, 24, 180, 6, 0)
    _attn("bf16", 2, 16, 24, 180, 6, 4)
    _attn("bf16", 2, 24, 16, 60, 6, 4)


def _swin_mlp(M, with_ln):
    import ctypes

    import torch

    from oracle import sr_oracle as O
    from studiosr_b200 import _lib
    from tests import gpu_util as G

    lib = _lib.load()
    C, heads, hid = 180, 6, 360
    g = torch.Generator().manual_seed(5)
    o = torch.randn(M, C, generator=g)
    res = torch.randn(M, C, generator=g)
    Wp = torch.randn(C, C, generator=g) / C**0.5 * 0.5
    bp = torch.randn(C, generator=g) * 0.1
    g2 = 1 + 0.1 * torch.randn(C, generator=g); be2 = 0.1 * torch.randn(C, generator=g)
    W1 = torch.randn(hid, C, generator=g) / C**0.5
    b1 = torch.randn(hid, generator=g) * 0.1
    W2 = torch.randn(C, hid, generator=g) / hid**0.5 * 0.5
    b2 = torch.randn(C, generator=g) * 0.1
    g3 = 1 + 0.1 * torch.randn(C, generator=g); be3 = 0.1 * torch.randn(C, generator=g)
    t1 = o @ Wp.t() + bp + res
    h = O.gelu(O.layer_norm(t1, g2, be2) @ W1.t() + b1)
    y = t1 + h @ W2.t() + b2
    yl = O.layer_norm(y, g3, be3) if with_ln else y
    c = lambda t: t.cuda()
    yo = torch.empty(M, C, device="cuda"); ylo = torch.empty(M, C, device="cuda")
    ws = torch.empty(M * 192 * 16 + (1 << 22), dtype=torch.uint8, device="cuda")
    args = [c(o), c(res), c(Wp), c(bp), c(g2), c(be2), c(W1), c(b1), c(W2), c(b2)]
    args += [c(g3), c(be3)] if with_ln else [None, None]
    ptrs = [None if t is None else t.data_ptr() for t in args]
    _lib.check(lib.ssr_op_swin_mlp(*ptrs, yo.data_ptr(), ylo.data_ptr(), M, C, heads, hid, ws.data_ptr(), ws.numel(), G.stream()))
    torch.cuda.synchronize()
    _stats(f"swin_mlp M{M} ln{int(with_ln)}: y", yo, y)
    _stats("   y_ln / bf16 copy", ylo, yl)


@check
def mlp_fused_small():
    _swin_mlp(128, True)


@check
def mlp_fused():
    _swin_mlp(300, True)
    _swin_mlp(5184, False)
    _swin_mlp(148 * 128 * 3 + 77, True)


def _attn_fused(B, H, W, shift, C=180, heads=6):
    import torch

    from oracle import sr_oracle as O
    from studiosr_b200 import _lib
    from tests import gpu_util as G

    lib = _lib.load()
    ws, d = 8, C // heads
    g = torch.Generator().manual_seed(B + H + W + shift)
    xn = torch.randn(B, H, W, C, generator=g)
    Wq = torch.randn(3 * C, C, generator=g) / C**0.5
    bq = torch.randn(3 * C, generator=g) * 0.2
    table = torch.randn(225, heads, generator=g) * 0.5
    qkv = xn @ Wq.t() + bq
    q = torch.roll(qkv, (-shift, -shift), (1, 2)) if shift else qkv
    qw = O.to_windows(q, ws).reshape(-1, ws * ws, 3, heads, d)
    Q = qw[:, :, 0].transpose(1, 2) * d**-0.5
    K = qw[:, :, 1].transpose(1, 2)
    V = qw[:, :, 2].transpose(1, 2)
    s_ = Q @ K.transpose(-1, -2) + O.rel_pos_bias(table, ws)[None]
    mask = O.shift_mask(H, W, ws, shift, torch.float32)
    if mask is not None:
        nW = mask.shape[0]
        s_ = (s_.reshape(B, nW, heads, 64, 64) + mask[None, :, None]).reshape(-1, heads, 64, 64)
    o = (torch.softmax(s_, -1) @ V).transpose(1, 2).reshape(-1, 64, C)
    o = O.from_windows(o, ws, B, H, W)
    if shift:
        o = torch.roll(o, (shift, shift), (1, 2))
    y = torch.empty(B, H, W, C, device="cuda")
    wsb = torch.empty(B * H * W * 192 * 8 + (1 << 22), dtype=torch.uint8, device="cuda")
    args = [t.cuda() for t in (xn, Wq, bq, table)]
    _lib.check(lib.ssr_op_swin_attn(*[t.data_ptr() for t in args], y.data_ptr(), B, H, W, C, heads, shift, wsb.data_ptr(),
                                    wsb.numel(), G.stream()))
    torch.cuda.synchronize()
    _stats(f"swin_attn fused B{B} {H}x{W} shift{shift}", y.reshape(-1, C), o.reshape(-1, C))


@check
def attn_fused_small():
    _attn_fused(1, 8, 16, 0)


@check
def attn_fused():
    _attn_fused(1, 16, 24, 0)
    _attn_fused(2, 16, 24, 4)
    _attn_fused(3, 8, 8, 4)
    _attn_fused(1, 72, 72, 4)
    _attn_fused(5, 72, 72, 0)


@check
def attn_fused_odd():
    _attn_fused(1, 72, 72, 0)
    _attn_fused(1, 24, 24, 0)
    _attn_fused(1, 24, 24, 4)


@check
def model_cfg1():
    import json, os
    import numpy as np
    import torch

    from oracle import synth
    from studiosr_b200.models import SwinIR

    meta = json.load(open(os.path.join(ROOT, "tests", "golden", "meta.json")))["cases"]["swinir_full_x4_eval_cfg1"]
    cfg = meta["cfg"]
    kw = {k: cfg[k] for k in ("scale", "n_colors", "img_range", "embed_dim", "depths", "num_heads", "window_size", "mlp_ratio", "upsampler")}
    m = SwinIR(drop_path_rate=0.0, **kw)
    m.load_state_dict(synth.swinir_weights(cfg, meta["wseed"]), strict=True)
    m = m.cuda().eval()
    m.precision = "bf16"
    x = synth.image_batch(meta["shape"], meta["xseed"]).cuda()
    ref = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "swinir_full_x4_eval_cfg1.npz"))["y"])
    for rep in range(3):
        with torch.no_grad():
            y = m(x).float().cpu()
        print("rep", rep, "nan", int(torch.isnan(y).sum()), "max err", float((y - ref).abs().nan_to_num(0).max()))


@check
def model_tiny():
    import torch

    from oracle import sr_oracle as O
    from oracle import synth
    from studiosr_b200.models import SwinIR

    cfg = synth.swinir_config(**synth.SWINIR_TINY)
    P = synth.swinir_weights(cfg, 11)
    x = synth.image_batch((2, 3, 20, 28), 101)
    ref = O.swinir_forward(P, x, cfg)
    kw = {k: cfg[k] for k in ("scale", "embed_dim", "depths", "num_heads", "window_size", "mlp_ratio", "upsampler")}
    m = SwinIR(drop_path_rate=0.0, **kw)
    m.load_state_dict(P)
    m = m.cuda().eval()
    for prec in sys.argv[3:] or ["fp32", "tf32", "bf16"]:
        m.precision = prec
        with torch.no_grad():
            y = m(x.cuda())
        _stats(f"swinir tiny [{prec}]", y.reshape(-1, y.shape[-1]), ref.reshape(-1, ref.shape[-1]))


if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        CHECKS[sys.argv[2]]()
        sys.exit(0)
    names = sys.argv[1:] or list(CHECKS)
    for n in names:
        print(f"=== {n}", flush=True)
        try:
            r = subprocess.run([sys.executable, __file__, "--one", n], capture_output=True, text=True, timeout=240)
            out = (r.stdout + r.stderr).strip().splitlines()
            keep = [l for l in out if "Warning" not in l][-14:]
            print("\n".join(keep))
            print(f"--- exit {r.returncode}", flush=True)
        except subprocess.TimeoutExpired:
            print("--- TIMEOUT", flush=True)
