"""Host-side glue between the drop-in nn.Modules and the C ABI: per-(device, precision) native
handles, weight (re)packing keyed on parameter versions, and grow-only workspaces.  PyTorch is
used only for device memory and streams here."""
from typing import Dict, Optional

import numpy as np
import torch

from . import _lib


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class NativeModel:
    """One packed copy of a module's weights inside libssr_b200 (for one device and precision)."""

    def __init__(self, cfg: "_lib.ModelConfig", device: torch.device):
        self.lib = _lib.load()
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("studiosr_b200 runs on CUDA sm_100 devices only (no CPU fallback)")
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.index = idx
        self.handle = _lib.c_void_p()
        _lib.check(self.lib.ssr_model_create(_lib.ctypes.byref(cfg), idx, _lib.ctypes.byref(self.handle)))
        self.version = None
        self._ws: Optional[torch.Tensor] = None
        self._graphs: Dict = {}  # (entry, B, C, H, W, pad_mode) -> (CUDAGraph, static input, static output, workspace)

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.ssr_model_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    # -- weights ------------------------------------------------------------------------------
    def load_state(self, state: Dict[str, torch.Tensor], version) -> None:
        with torch.cuda.device(self.index):
            for name, t in state.items():
                if not t.is_floating_point():
                    continue  # integer buffers (relative_position_index) are derived natively
                h = t.detach().to(device="cpu", dtype=torch.float32).contiguous()
                _lib.check(self.lib.ssr_model_set_param(self.handle, name.encode(), h.data_ptr(), h.numel()))
            _lib.check(self.lib.ssr_model_finalize(self.handle))
        self.version = version
        self._graphs.clear()  # finalize may have moved the packed weights: captured graphs hold their old addresses

    # -- workspace ----------------------------------------------------------------------------
    def workspace(self, nbytes: int) -> torch.Tensor:
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = None
            self._ws = torch.empty(int(nbytes), dtype=torch.uint8, device=self.device)
        return self._ws

    # -- entry points -------------------------------------------------------------------------
    def _graphed(self, key, x: torch.Tensor, out_shape, out_dtype, need_bytes: int, call):
        """Replay (capture on first use) the launch sequence of one entry point at one shape as a CUDA graph: at batch 1 a
        SwinIR forward is ~87 launches of 10-20 us kernels and the host launch path dominates (BASELINE.json config 1).
        `call(x_static, y_static, ws)` must enqueue everything on the current stream and touch no other memory."""
        g = self._graphs.get(key)
        if g is None:
            xs, ys = torch.empty_like(x), torch.empty(out_shape, dtype=out_dtype, device=x.device)
            ws = torch.empty(int(need_bytes), dtype=torch.uint8, device=x.device)  # owned by the graph: never re-allocated
            xs.copy_(x)
            side = torch.cuda.Stream(device=x.device)
            side.wait_stream(torch.cuda.current_stream(x.device))
            with torch.cuda.stream(side):
                call(xs, ys, ws)  # eager warm-up on the side stream: one-time cudaFuncSetAttribute calls happen outside the capture
            torch.cuda.current_stream(x.device).wait_stream(side)
            graph = torch.cuda.CUDAGraph()
            l0 = self.lib.ssr_launch_count()
            with torch.cuda.graph(graph):
                call(xs, ys, ws)
            g = (graph, xs, ys, ws, self.lib.ssr_launch_count() - l0)
            self._graphs[key] = g
        graph, xs, ys, _, kernels = g
        xs.copy_(x, non_blocking=True)
        graph.replay()
        self.lib.ssr_note_graph_replay(kernels)
        return ys.clone()

    def forward(self, x: torch.Tensor, scale: int, pad_mode: int, graph: bool = False) -> torch.Tensor:
        B, C, H, W = x.shape
        x = x.detach().to(torch.float32).contiguous()
        if graph:
            with torch.cuda.device(self.index):
                need = self.lib.ssr_model_workspace_bytes(self.handle, B, H, W, pad_mode)

                def call(xs, ys, ws):
                    _lib.check(self.lib.ssr_model_forward(self.handle, xs.data_ptr(), ys.data_ptr(), B, H, W, pad_mode, ws.data_ptr(),
                                                          ws.numel(), _stream_ptr(x.device)))

                return self._graphed(("fwd", B, C, H, W, pad_mode), x, (B, C, H * scale, W * scale), torch.float32, need, call)
        y = torch.empty((B, C, H * scale, W * scale), dtype=torch.float32, device=x.device)
        with torch.cuda.device(self.index):
            need = self.lib.ssr_model_workspace_bytes(self.handle, B, H, W, pad_mode)
            ws = self.workspace(need)
            _lib.check(self.lib.ssr_model_forward(self.handle, x.data_ptr(), y.data_ptr(), B, H, W, pad_mode,
                                                  ws.data_ptr(), ws.numel(), _stream_ptr(x.device)))
        return y

    # -- training step (ssr_model_train_*): the fp32 master parameters stay in PyTorch's device memory -----------
    def train_bind(self, named: Dict[str, torch.Tensor]) -> None:
        """Name the state_dict entries once; pointers are passed per call (parameters may be re-allocated)."""
        self._train_names = list(named.keys())
        n = len(self._train_names)
        names = (_lib.c_char_p * n)(*[k.encode() for k in self._train_names])
        numels = (_lib.c_int64 * n)(*[int(named[k].numel()) for k in self._train_names])
        with torch.cuda.device(self.index):
            _lib.check(self.lib.ssr_model_train_bind(self.handle, n, names, numels))

    def _ptr_array(self, tensors):
        arr = (_lib.c_void_p * len(tensors))()
        for i, t in enumerate(tensors):
            arr[i] = None if t is None else t.data_ptr()
        return arr

    def train_forward(self, x: torch.Tensor, params, scale: int, drop_scale: Optional[torch.Tensor] = None):
        """x fp32 [B,3,H,W]; params: tensors in bound order; drop_scale: fp32 [2*n_blocks, B] stochastic-depth factors or
        None.  Returns (y, workspace) -- the workspace holds the saved activations and must reach train_backward untouched."""
        B, C, H, W = x.shape
        x = x.detach().to(torch.float32).contiguous()
        y = torch.empty((B, C, H * scale, W * scale), dtype=torch.float32, device=x.device)
        with torch.cuda.device(self.index):
            need = self.lib.ssr_model_train_workspace_bytes(self.handle, B, H, W)
            ws = torch.empty(int(need), dtype=torch.uint8, device=x.device)
            _lib.check(self.lib.ssr_model_train_forward(self.handle, self._ptr_array(params),
                                                        None if drop_scale is None else drop_scale.data_ptr(), x.data_ptr(),
                                                        y.data_ptr(), B, H, W, ws.data_ptr(), ws.numel(), _stream_ptr(x.device)))
        return y, ws

    def train_backward(self, dy: torch.Tensor, grads, shape, ws: torch.Tensor, drop_scale: Optional[torch.Tensor] = None) -> None:
        B, _, H, W = shape
        dy = dy.detach().to(torch.float32).contiguous()
        with torch.cuda.device(self.index):
            _lib.check(self.lib.ssr_model_train_backward(self.handle, dy.data_ptr(),
                                                         None if drop_scale is None else drop_scale.data_ptr(), self._ptr_array(grads), B, H, W,
                                                         ws.data_ptr(), ws.numel(), _stream_ptr(dy.device)))

    def train_input_grad(self, shape, device) -> torch.Tensor:
        """dL/dx of the step whose train_backward ran last (its workspace must still be alive)."""
        B, _, H, W = shape
        dx = torch.empty((B, 3, H, W), dtype=torch.float32, device=device)
        with torch.cuda.device(self.index):
            _lib.check(self.lib.ssr_model_train_input_grad(self.handle, dx.data_ptr(), B, H, W, _stream_ptr(dx.device)))
        return dx

    def upscale_u8(self, img: torch.Tensor, scale: int, graph: bool = False) -> torch.Tensor:
        """img: uint8 [B,H,W,3] on the device -> uint8 [B,sH,sW,3]."""
        B, H, W, _ = img.shape
        if graph:
            img = img.contiguous()
            with torch.cuda.device(self.index):
                need = self.lib.ssr_model_workspace_bytes(self.handle, B, H, W, _lib.PAD_EVAL)

                def call(xs, ys, ws):
                    _lib.check(self.lib.ssr_model_upscale_u8(self.handle, xs.data_ptr(), ys.data_ptr(), B, H, W, ws.data_ptr(), ws.numel(),
                                                             _stream_ptr(img.device)))

                return self._graphed(("u8", B, 3, H, W, 0), img, (B, H * scale, W * scale, 3), torch.uint8, need, call)
        out = torch.empty((B, H * scale, W * scale, 3), dtype=torch.uint8, device=img.device)
        with torch.cuda.device(self.index):
            need = self.lib.ssr_model_workspace_bytes(self.handle, B, H, W, _lib.PAD_EVAL)
            ws = self.workspace(need)
            _lib.check(self.lib.ssr_model_upscale_u8(self.handle, img.data_ptr(), out.data_ptr(), B, H, W,
                                                     ws.data_ptr(), ws.numel(), _stream_ptr(img.device)))
        return out

    def upscale_tiled_u8(self, frame: torch.Tensor, scale: int, tile: int, overlap: int, chunk: int = 0) -> torch.Tensor:
        """frame: uint8 [H,W,3] on the device -> uint8 [sH,sW,3] (tiles batched, blended on device)."""
        H, W, _ = frame.shape
        out = torch.empty((H * scale, W * scale, 3), dtype=torch.uint8, device=frame.device)
        with torch.cuda.device(self.index):
            need = self.lib.ssr_model_tiled_workspace_bytes(self.handle, H, W, tile, overlap, chunk)
            ws = self.workspace(need)
            _lib.check(self.lib.ssr_model_upscale_tiled_u8(self.handle, frame.data_ptr(), out.data_ptr(), H, W, tile,
                                                           overlap, chunk, ws.data_ptr(), ws.numel(),
                                                           _stream_ptr(frame.device)))
        return out

    def upscale_tiled_u8_host(self, frame: np.ndarray, out: np.ndarray, scale: int, tile: int, overlap: int,
                              chunk: int = 0) -> np.ndarray:
        """HOST uint8 [H,W,3] -> HOST uint8 [sH,sW,3]; H2D + compute + D2H + sync inside the call."""
        H, W, _ = frame.shape
        assert frame.dtype == np.uint8 and frame.flags["C_CONTIGUOUS"]
        assert out.dtype == np.uint8 and out.flags["C_CONTIGUOUS"] and out.shape == (H * scale, W * scale, 3)
        with torch.cuda.device(self.index):
            need = self.lib.ssr_model_tiled_workspace_bytes(self.handle, H, W, tile, overlap, chunk)
            ws = self.workspace(need)
            _lib.check(self.lib.ssr_model_upscale_tiled_u8_host(
                self.handle, frame.ctypes.data, out.ctypes.data, H, W, tile, overlap, chunk, ws.data_ptr(), ws.numel(),
                _stream_ptr(self.device)))
        return out
