"""Frame sharding across ranks: frame f -> rank f mod G, no data-path collective;
only per-rank timing vectors are reduced.  world_size-2 gloo run on CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_rank_frames_partition(host):
    for n, g in ((240, 8), (100, 4), (7, 2), (1, 8)):
        parts = [host.rank_frames(n, r, g) for r in range(g)]
        flat = sorted(f for p in parts for f in p)
        assert flat == list(range(n))
        assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def _worker(rank, world, port, n_frames, out):
    import importlib
    host = importlib.import_module("hevc-deep-learning-pipeline_b200.host")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = host.rank_frames(n_frames, rank, world)
    ctus = torch.tensor([len(mine) * 510.0])
    ms = torch.tensor([10.0 + rank])                       # pretend device time of this rank
    dist.barrier()
    dist.all_reduce(ctus, op=dist.ReduceOp.SUM)            # whole-job units
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)              # max over ranks
    if rank == 0:
        out.put((float(ctus), float(ms)))
    dist.destroy_process_group()


def test_two_rank_gloo_aggregation():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, 9, q)) for r in range(2)]
    [p.start() for p in ps]
    total, ms = q.get(timeout=120)
    [p.join(60) for p in ps]
    assert total == 9 * 510.0 and ms == 11.0
