"""CPU (gloo, world_size 2): the flat gradient arena — views, in-place accumulation, ONE all-reduce."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from butd_detr_b200.train import GradArena
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.ReLU(), torch.nn.Linear(7, 3))
    arena = GradArena(net)
    calls = []
    real = dist.all_reduce
    dist.all_reduce = lambda *a, **k: (calls.append(1), real(*a, **k))[1]
    x = torch.full((4, 5), float(rank + 1))
    arena.zero()
    net(x).sum().backward()
    net(x).sum().backward()  # accumulates in place, still inside the arena
    assert arena.check_views()
    local = arena.flat.clone()
    arena.all_reduce()
    dist.all_reduce = real
    gathered = [torch.zeros_like(local) for _ in range(world)]
    dist.all_gather(gathered, local)
    ok = torch.allclose(arena.flat, sum(gathered) / world) and len(calls) == 1
    ok = ok and arena.flat.numel() == sum(p.numel() for p in net.parameters())
    ok = ok and all(p.grad.data_ptr() >= arena.flat.data_ptr() for p in net.parameters())
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_grad_arena_single_all_reduce_two_ranks():
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out[0] and out[1]
