"""CPU: bench plumbing that does not need a GPU — the reference arm and world_size-2 sharding."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_algorithmic_bytes_match_survey_figures():
    sys.path.insert(0, ROOT)
    import bench
    assert bench.algorithmic_bytes("bd_fps(6, 1, 50000, 2048)", 1) == 12 * 50000 + 4 * 2048  # 608 KB
    assert bench.algorithmic_bytes("bd_ball_query(6, 1, 50000, 2048, 64)", 1) == 1_148_864  # 1.149 MB idx-only
    assert bench.algorithmic_bytes("bd_fps(3, 1, 2048, 1024)", 1) is None


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    import bench
    # each rank draws its own scenes (weak scaling: no data-path collective) ...
    pool = bench.make_pool(2, 100000 * (rank + 1)) if False else None
    from butd_detr_b200 import synth
    mine = synth.synth_scene(100000 * (rank + 1), 512, 8)["point_clouds"]
    t = torch.tensor([float(mine.sum()), 10.0 + rank], dtype=torch.float64)
    gathered = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    # ... and the only reduction is the max over ranks of the elapsed time
    tm = torch.tensor([10.0 + rank], dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    q.put((rank, [float(g[0]) for g in gathered], float(tm)))
    dist.destroy_process_group()


def test_world_size_2_sharding_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=120) for _ in procs]
    [p.join(timeout=60) for p in procs]
    for rank, sums, tmax in res:
        assert sums[0] != sums[1], "ranks must process different scenes"
        assert tmax == 11.0
