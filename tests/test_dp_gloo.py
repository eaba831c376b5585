"""world_size-2 gloo test (CPU) of the data-parallel host logic: scene sharding, the global-batch divisor and the
gradient sum that the NCCL all-reduce performs on the GPU path."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import batcher, npe


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.util import balls_batches
    _, batches = balls_batches(60, 4, seed=77)
    this, ctx, y, *_ = batcher.make_batch(*batches[0], offset=5)
    params = npe.init_params(1234)
    # uneven shard of ONE reference batch of 50: 27 + 23, divisor stays the global 50
    lo, hi = (0, 27) if rank == 0 else (27, 50)
    g_local, _ = npe.bp(params, (this[lo:hi], ctx[lo:hi], y[lo:hi]), divisor_batch=50)
    t = torch.from_numpy(g_local.copy())
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    if rank == 0:
        g_full, _ = npe.bp(params, (this, ctx, y))
        out["err"] = float(np.max(np.abs(t.numpy() - g_full)) / np.max(np.abs(g_full)))
    dist.destroy_process_group()


def test_sharded_gradient_sum_equals_full_batch_gradient():
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    assert out["err"] < 1e-5
