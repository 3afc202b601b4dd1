"""Helpers for the -m gpu tests: call the op-level C ABI with torch CUDA tensors."""
import torch

from studiosr_b200 import _lib


def _p(t):
    return None if t is None else t.data_ptr()


def _ws(nbytes):
    return torch.empty(int(nbytes), dtype=torch.uint8, device="cuda")


def stream():
    return torch.cuda.current_stream().cuda_stream


def op_linear(prec, x, W, b, res=None, act=0, ln_w=None, ln_b=None):
    lib = _lib.load()
    M, K = x.shape
    N = W.shape[0]
    y = torch.empty(M, N, device="cuda")
    y_ln = torch.empty(M, N, device="cuda") if ln_w is not None else None
    ws = _ws(lib.ssr_op_workspace_bytes(M * (K + 2 * N + 256) + (N + 64) * (K + 64) + 4096))
    _lib.check(lib.ssr_op_linear(_lib.PRECISIONS[prec], _p(x), _p(W), _p(b), _p(res), act, _p(ln_w), _p(ln_b), _p(y),
                                 _p(y_ln), M, K, N, _p(ws), ws.numel(), stream()))
    torch.cuda.synchronize()
    return y, y_ln


def op_conv3x3(prec, x, W, b, res=None, act=0, alpha=1.0, ps_r=0):
    lib = _lib.load()
    B, Cin, H, Wd = x.shape
    Cout = W.shape[0]
    r = ps_r if ps_r > 1 else 1
    y = torch.empty(B, Cout // (r * r), H * r, Wd * r, device="cuda")
    elems = B * H * Wd * (Cin + 64 + 3 * (Cout + 64) * 1) + (Cout + 64) * 9 * (Cin + 64)
    ws = _ws(lib.ssr_op_workspace_bytes(elems))
    _lib.check(lib.ssr_op_conv3x3(_lib.PRECISIONS[prec], _p(x), _p(W), _p(b), _p(res), _p(y), B, Cin, Cout, H, Wd, act,
                                  alpha, ps_r, _p(ws), ws.numel(), stream()))
    torch.cuda.synchronize()
    return y


def op_window_attention(prec, qkv, table, heads, ws_, shift):
    lib = _lib.load()
    B, H, W, C3 = qkv.shape
    C = C3 // 3
    o = torch.empty(B, H, W, C, device="cuda")
    ws = _ws(lib.ssr_op_workspace_bytes(B * H * W * 5 * (C + 64)))
    _lib.check(lib.ssr_op_window_attention(_lib.PRECISIONS[prec], _p(qkv), _p(table), _p(o), B, H, W, C, heads, ws_,
                                           shift, _p(ws), ws.numel(), stream()))
    torch.cuda.synchronize()
    return o


def op_swin_attn(xn, Wqkv, bqkv, table, heads, shift):
    lib = _lib.load()
    B, H, W, C = xn.shape
    o = torch.empty(B, H, W, C, device="cuda")
    ws = _ws(B * H * W * 192 * 8 + (1 << 22))
    _lib.check(lib.ssr_op_swin_attn(_p(xn), _p(Wqkv), _p(bqkv), _p(table), _p(o), B, H, W, C, heads, shift, _p(ws),
                                    ws.numel(), stream()))
    torch.cuda.synchronize()
    return o


def op_swin_mlp(o, res, Wp, bp, g2, be2, W1, b1, W2, b2, g3, be3, heads, hidden):
    lib = _lib.load()
    M, C = o.shape
    y = torch.empty(M, C, device="cuda")
    y2 = torch.empty(M, C, device="cuda")
    ws = _ws(M * 192 * 16 + (1 << 22))
    _lib.check(lib.ssr_op_swin_mlp(_p(o), _p(res), _p(Wp), _p(bp), _p(g2), _p(be2), _p(W1), _p(b1), _p(W2), _p(b2), _p(g3),
                                   _p(be3), _p(y), _p(y2), M, C, heads, hidden, _p(ws), ws.numel(), stream()))
    torch.cuda.synchronize()
    return y, y2


def op_conv_wgrad(dy, x, taps=9, alpha=1.0, want_bias=True):
    """dy [B,Cout,H,W], x [B,Cin,H,W] -> (dW [Cout,Cin,3,3] or [Cout,Cin], db [Cout])"""
    lib = _lib.load()
    B, Cout, H, Wd = dy.shape
    Cin = x.shape[1]
    dW = torch.empty((Cout, Cin, 3, 3) if taps == 9 else (Cout, Cin), device="cuda")
    db = torch.empty(Cout, device="cuda") if want_bias else None
    elems = B * H * Wd * (Cin + Cout + 128) + 2 * (Cout + 64) * taps * (Cin + 64) + 4096 + 2 * 592 * 2304
    ws = _ws(lib.ssr_op_workspace_bytes(elems))
    _lib.check(lib.ssr_op_conv3x3_wgrad(_p(dy), _p(x), _p(dW), _p(db), B, Cin, Cout, H, Wd, taps, alpha, _p(ws), ws.numel(),
                                        stream()))
    torch.cuda.synchronize()
    return dW, db
