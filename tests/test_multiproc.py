"""CPU (-m "not gpu"): the N > 1 plumbing with world_size-2 gloo process groups -- timings reduce to the max over ranks, only
rank 0 speaks, and the sharded-frame protocol of studiosr_b200/sharding.py (tile slots -> all-gather -> row-band blend ->
all-gather) reproduces the single-process tiler with the oracle standing in for the CUDA compute."""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    import bench

    dist.init_process_group("gloo", rank=rank, world_size=world)
    ms = bench.max_over_ranks(10.0 + 5.0 * rank, dist, torch.device("cpu"))  # rank 1 is the slow one
    frame = bench.synthetic_frame(rank)
    sig = torch.tensor([float(frame[:8, :8].astype("float64").sum())], dtype=torch.float64)
    sigs = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(sigs, sig)
    dist.barrier()
    q.put((rank, ms, [s.item() for s in sigs], frame.shape))
    dist.destroy_process_group()


def test_max_over_ranks_and_distinct_shards_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ms, sigs, shape in out:
        assert ms == 15.0  # every rank reports the slowest rank's time
        assert shape == (540, 960, 3)
        assert sigs[0] != sigs[1]  # weak scaling: each rank upscales its own frame


def test_reference_arm_under_torchrun_prints_once():
    """`bench.py --impl reference` launched as the driver launches it for N = 2: rank 0 runs the CPU arm and prints ONE JSON
    line, rank 1 exits 0 without work."""
    env = dict(os.environ, OMP_NUM_THREADS="4")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(31500 + os.getpid() % 2000), os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
           "--steps", "1", "--warmup", "1"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    line = lines[0]
    assert line["impl"] == "reference" and line["n_gpus"] == 2 and line["cpu_baseline"]["kind"] == "port"
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["value"] > 0


# ---- one frame sharded over two ranks (studiosr_b200/sharding.py) with the oracle as the compute backend -------------------
SH_H, SH_W, SH_TILE, SH_OVERLAP = 40, 56, 16, 4


def _sharding_case():
    from oracle import sr_oracle as O
    from oracle import synth

    cfg = synth.swinir_config(**synth.SWINIR_TINY)
    P = synth.swinir_weights(cfg, 11)
    frame = torch.from_numpy(synth.smooth_image_u8(SH_H, SH_W, seed=7))
    fwd = lambda x: O.swinir_forward(P, x, cfg)
    return O, cfg, frame, fwd


class _OracleBackend:
    """Same interface as sharding.NativeTileBackend, computing on the CPU through the oracle."""

    def __init__(self):
        self.O, cfg, _, self.fwd = _sharding_case()
        self.H, self.W, self.scale, self.tile, self.overlap = SH_H, SH_W, cfg["scale"], SH_TILE, SH_OVERLAP
        self.ys, self.xs = self.O.tile_starts(SH_H, SH_TILE, SH_TILE - SH_OVERLAP), self.O.tile_starts(SH_W, SH_TILE, SH_TILE - SH_OVERLAP)
        self.n_tiles = len(self.ys) * len(self.xs)
        self.tile_elems = 3 * (SH_TILE * self.scale) ** 2
        self.device = torch.device("cpu")

    def compute(self, frame, tiles, begin, end):
        x = frame.permute(2, 0, 1).float().div(255.0).unsqueeze(0)
        with torch.inference_mode():
            for t in range(begin, end):
                y0, x0 = self.ys[t // len(self.xs)], self.xs[t % len(self.xs)]
                tiles[t - begin] = self.fwd(x[:, :, y0:y0 + SH_TILE, x0:x0 + SH_TILE])[0].reshape(-1)

    def blend(self, tiles_all, out_frame, row_begin, row_end):
        ts = SH_TILE * self.scale
        t = tiles_all[:self.n_tiles].view(self.n_tiles, 3, ts, ts)
        band = self.O.blend_tiles(t, SH_H, SH_W, self.scale, SH_TILE, SH_OVERLAP, row_begin, row_end)
        out_frame[row_begin:row_end] = (band * 255.0).round().clip(0, 255).to(torch.uint8).permute(1, 2, 0)


def _shard_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    from studiosr_b200.sharding import ShardedTiledUpscaler

    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    up = ShardedTiledUpscaler(_OracleBackend(), dist)
    frame = _sharding_case()[2]
    out_host = torch.empty((SH_H * 4, SH_W * 4, 3), dtype=torch.uint8) if rank == 0 else None
    up.upscale_host(frame if rank == 0 else None, out_host)  # only rank 0 holds the frame; the broadcast delivers it
    q.put((rank, up.tile_slots, up.row_slots, out_host.numpy() if rank == 0 else None, up.upscale(None).clone().numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_slot_partition():
    from studiosr_b200.sharding import slot_partition

    assert slot_partition(220, 8) == (28, [(0, 28), (28, 56), (56, 84), (84, 112), (112, 140), (140, 168), (168, 196), (196, 220)])
    assert slot_partition(2160, 8)[0] == 270
    assert slot_partition(3, 4) == (1, [(0, 1), (1, 2), (2, 3), (3, 3)])  # more ranks than items: trailing slots empty
    for n, w in ((220, 1), (220, 2), (220, 4), (7, 3)):
        per, slots = slot_partition(n, w)
        assert slots[0][0] == 0 and slots[-1][1] == n and all(a[1] == b[0] for a, b in zip(slots, slots[1:]))
        assert all(e - b <= per for b, e in slots)


def test_sharded_frame_two_ranks_gloo_matches_single_process():
    O, cfg, frame, fwd = _sharding_case()
    with torch.inference_mode():
        ref = O.tiled_upscale(fwd, frame.permute(2, 0, 1).float().div(255.0).unsqueeze(0), cfg["scale"], SH_TILE, SH_OVERLAP)[0]
    ref_u8 = (ref.double() * 255.0).round().clip(0, 255).to(torch.uint8).permute(1, 2, 0).numpy()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=300) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, tile_slots, row_slots, host0, dev0), (_, _, _, _, dev1) = out
    assert tile_slots[0][1] == tile_slots[1][0] and tile_slots[1][1] == 15 and row_slots == [(0, 80), (80, 160)]
    assert (dev0 == dev1).all()  # every rank ends with the whole frame
    assert (host0 == dev0).all()
    diff = abs(host0.astype(int) - ref_u8.astype(int))
    assert diff.max() <= 1 and (diff > 0).mean() < 1e-3  # fp32 tile list vs fp64 accumulation of the single-process tiler
