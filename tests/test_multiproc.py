"""CPU (-m "not gpu"): the N > 1 plumbing of bench.py with world_size-2 gloo process groups -- the frame shards differ per
rank, timings reduce to the max over ranks, and only rank 0 speaks.  (The data path has no collective: SURVEY 8e.)"""
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
