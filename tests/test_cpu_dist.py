"""N>1 path on CPU: two gloo ranks shard scenes one per rank (no data-path collective) and agree on the
max-over-ranks step time exactly like bench.py's torchrun path does with NCCL."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import bench
    sc = bench.make_scene(seed=rank, n=3000)                       # scene-per-rank sharding (seed == rank)
    n = len(sc["coord"])
    ms = torch.tensor([10.0 + 5.0 * rank], dtype=torch.float64)    # pretend per-rank step time
    dist.barrier()
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)                      # bench.py: time = max over ranks
    pts = torch.tensor([float(n)], dtype=torch.float64)
    dist.all_reduce(pts)                                           # whole-job units
    h = torch.tensor([float(np.abs(sc["grid_coord"]).sum())], dtype=torch.float64)
    gathered = [torch.zeros_like(h) for _ in range(world)]
    dist.all_gather(gathered, h)
    if rank == 0:
        out.put((float(ms.item()), float(pts.item()), [float(g.item()) for g in gathered]))
    dist.destroy_process_group()


def test_scene_sharding_and_max_over_ranks():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    ms, pts, hashes = q.get(timeout=120)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ms == 15.0                      # max over ranks
    assert pts == 6000.0                   # value = units of ALL ranks / that time  (weak scaling)
    assert hashes[0] != hashes[1]          # different scenes on different ranks
