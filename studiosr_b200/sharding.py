"""One frame, several GPUs: BASELINE.json config 5 ("overlapping tiles sharded across 1/2/4/8 GPUs").

The reference has no multi-GPU inference (`Model.inference`, common.py:36-48, is single-device, batch 1); tiles are
independent units (window attention never crosses a tile), so a frame shards by its row-major tile list with exactly one
data exchange, the gather of the tile outputs:

  rank 0        frame (host uint8) --H2D--> device --broadcast (NCCL)--> every rank
  every rank    tiles [b_r, e_r)  --ssr_model_tiles_u8-->  its slot of the full fp32 tile list
  all ranks     all_gather of the slots (in place, NCCL over NVLink / NVSwitch)
  every rank    blends its band of output rows  --ssr_model_blend_tiles_u8-->  its slot of the uint8 frame
  all ranks     all_gather of the bands (in place)  ->  rank 0 --D2H--> host

Both partitions are "equal slots, last one short" so the gathered buffers ARE the row-major lists (no index map) and the
collectives are plain equal-count all-gathers.  The blend is a per-pixel gather over the tiles covering it, so the result is
bit-identical to the single-GPU call whatever the number of ranks.

`dist` is torch.distributed (plumbing only); the compute is behind a small backend so the protocol itself is covered on CPU
with gloo (tests/test_multiproc.py) and on GPUs with NCCL (tests/test_gpu_models.py)."""
from typing import List, Optional, Tuple

import torch


def slot_partition(n: int, world: int) -> Tuple[int, List[Tuple[int, int]]]:
    """Split range(n) into `world` contiguous slots of ceil(n / world) items (trailing slots short or empty)."""
    per = (n + world - 1) // world
    return per, [(min(n, r * per), min(n, (r + 1) * per)) for r in range(world)]


class NativeTileBackend:
    """The C-ABI pair ssr_model_tiles_u8 / ssr_model_blend_tiles_u8 of one NativeModel."""

    def __init__(self, nat, H: int, W: int, scale: int, tile: int, overlap: int, chunk: int = 0):
        from . import _lib

        self.nat, self.H, self.W, self.scale, self.tile, self.overlap, self.chunk = nat, H, W, scale, tile, overlap, chunk
        self.lib, self._lib = nat.lib, _lib
        self.device = nat.device
        self.n_tiles = self.lib.ssr_tiled_num_tiles(H, W, tile, overlap)
        self.tile_elems = self.lib.ssr_tiled_tile_elems(nat.handle, H, W, tile)
        self._ws = None

    def compute(self, frame: torch.Tensor, tiles: torch.Tensor, begin: int, end: int) -> None:
        """tiles: fp32 view whose element 0 is tile `begin`."""
        if end <= begin:
            return
        n = end - begin
        per_pass = n if self.chunk <= 0 else min(n, self.chunk)
        with torch.cuda.device(self.nat.index):
            need = self.lib.ssr_model_tiles_workspace_bytes(self.nat.handle, self.H, self.W, self.tile, per_pass)
            if self._ws is None or self._ws.numel() < need:  # private (a captured graph keeps pointing at it), grown never shrunk
                self._ws = torch.empty(int(need), dtype=torch.uint8, device=self.device)
            ws = self._ws
            self._lib.check(self.lib.ssr_model_tiles_u8(self.nat.handle, frame.data_ptr(), tiles.data_ptr(), self.H, self.W, self.tile,
                                                        self.overlap, begin, end, self.chunk, ws.data_ptr(), ws.numel(),
                                                        torch.cuda.current_stream(self.device).cuda_stream))

    def blend(self, tiles_all: torch.Tensor, out_frame: torch.Tensor, row_begin: int, row_end: int) -> None:
        if row_end <= row_begin:
            return
        with torch.cuda.device(self.nat.index):
            self._lib.check(self.lib.ssr_model_blend_tiles_u8(self.nat.handle, tiles_all.data_ptr(), out_frame.data_ptr(), self.H, self.W,
                                                              self.tile, self.overlap, row_begin, row_end,
                                                              torch.cuda.current_stream(self.device).cuda_stream))


class ShardedTiledUpscaler:
    """Strong-scales one frame over the ranks of `group` (None = the default group; world size 1 works without
    torch.distributed being initialised).  Buffers are allocated once and re-used across frames."""

    def __init__(self, backend, dist=None, group=None, graph: bool = False):
        """graph=True: the whole pass of one frame (this rank's ~87 kernel launches per tile batch, both NCCL all-gathers and the
        blend) is captured once as a CUDA graph and replayed.  With the frame sharded over 8 GPUs a launch is only ~80 us
        long and the host launch path / inter-kernel gaps become a visible share of the 9 ms frame."""
        self.be, self.dist, self.group = backend, dist, group
        self.use_graph, self._graph, self._graph_kernels = bool(graph), None, 0
        live = dist is not None and dist.is_initialized()
        self.world = dist.get_world_size(group) if live else 1
        self.rank = dist.get_rank(group) if live else 0
        be = backend
        self.out_rows, self.out_cols = be.H * be.scale, be.W * be.scale
        self.tiles_per, self.tile_slots = slot_partition(be.n_tiles, self.world)
        self.rows_per, self.row_slots = slot_partition(self.out_rows, self.world)
        dev = be.device
        self.tiles_all = torch.empty((self.world * self.tiles_per, be.tile_elems), dtype=torch.float32, device=dev)
        self.frame_all = torch.empty((self.world * self.rows_per, self.out_cols, 3), dtype=torch.uint8, device=dev)
        self.frame_in = torch.empty((be.H, be.W, 3), dtype=torch.uint8, device=dev)

    def _all_gather_slots(self, buf: torch.Tensor, per: int) -> None:
        if self.world > 1:  # in place: rank r's input is slot r of the output
            self.dist.all_gather_into_tensor(buf, buf[self.rank * per:(self.rank + 1) * per], group=self.group)

    def _pass(self, frame: torch.Tensor) -> None:
        b, e = self.tile_slots[self.rank]
        self.be.compute(frame, self.tiles_all[b:], b, e)
        self._all_gather_slots(self.tiles_all, self.tiles_per)
        r0, r1 = self.row_slots[self.rank]
        self.be.blend(self.tiles_all, self.frame_all, r0, r1)
        self._all_gather_slots(self.frame_all, self.rows_per)

    def upscale(self, frame: Optional[torch.Tensor] = None) -> torch.Tensor:
        """frame: uint8 [H, W, 3] already on this rank's device (every rank passes the same frame), or None to use what
        `broadcast_frame` left in self.frame_in.  Returns the whole uint8 [sH, sW, 3] frame (a view; on every rank)."""
        if not (self.use_graph and self.frame_all.is_cuda):
            self._pass(self.frame_in if frame is None else frame)
            return self.frame_all[:self.out_rows]
        if frame is not None and frame.data_ptr() != self.frame_in.data_ptr():
            self.frame_in.copy_(frame, non_blocking=True)  # the graph reads the frame from its own static buffer
        if self._graph is None:
            dev = self.frame_all.device
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side):  # eager pass first: communicator set-up and one-time attribute calls stay outside the capture
                self._pass(self.frame_in)
            cur.wait_stream(side)
            torch.cuda.synchronize(dev)
            lib = getattr(self.be, "lib", None)
            l0 = lib.ssr_launch_count() if lib is not None else 0
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self._pass(self.frame_in)
            except Exception as exc:  # e.g. a collective algorithm that cannot be captured: stay eager (same collective order,
                # so ranks that did capture and ranks that did not still match)
                import warnings

                warnings.warn(f"studiosr_b200.sharding: CUDA-graph capture of the sharded pass failed ({exc!r}); running eagerly")
                self.use_graph = False
                torch.cuda.synchronize(dev)
                self._pass(self.frame_in)
                return self.frame_all[:self.out_rows]
            self._graph_kernels = (lib.ssr_launch_count() - l0) if lib is not None else 0
            self._graph = g
        self._graph.replay()
        if self._graph_kernels:
            self.be.lib.ssr_note_graph_replay(self._graph_kernels)
        return self.frame_all[:self.out_rows]

    def broadcast_frame(self, frame_host: Optional[torch.Tensor]) -> None:
        """Rank 0 uploads its (pinned) host frame; every rank receives it."""
        if self.rank == 0:
            self.frame_in.copy_(frame_host, non_blocking=True)
        if self.world > 1:
            self.dist.broadcast(self.frame_in, src=self.dist.get_global_rank(self.group, 0) if self.group is not None else 0,
                                group=self.group)

    def upscale_host(self, frame_host: Optional[torch.Tensor], out_host: Optional[torch.Tensor]) -> None:
        """Host in (rank 0) -> host out (rank 0), synchronous on rank 0: H2D, broadcast, tiles, all-gather, blend, all-gather,
        D2H.  Other ranks pass None, None."""
        self.broadcast_frame(frame_host)
        out = self.upscale(None)
        if self.rank == 0:
            out_host.copy_(out, non_blocking=True)
            if out.is_cuda:
                torch.cuda.current_stream(out.device).synchronize()
