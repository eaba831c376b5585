"""Multi-GPU parity check, launched by torchrun (one process per GPU):
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/dp_gpu_check.py
One reference batch of 50 is split unevenly across the ranks; each rank back-propagates with the GLOBAL divisor, the
flat gradients are summed by ONE ncclAllReduce (libnpe_cuda's own communicator) and compared with the oracle's
full-batch gradient; then 20 data-parallel Adam steps are compared with the oracle's single-process steps."""
import os
import sys

import numpy as np
import torch.distributed as dist
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamics_b200 import ModelParams, NPEModel  # noqa: E402
from oracle import batcher, npe, optim  # noqa: E402
from tests.util import balls_batches  # noqa: E402


def attach(m, rank, world, mode):
    idt = torch.zeros(128, dtype=torch.uint8)
    if mode == "nccl":
        if rank == 0:
            idt = torch.tensor(list(m.unique_id()), dtype=torch.uint8)
        dist.broadcast(idt, 0)
        m.comm_init(rank, world, bytes(idt.tolist()))
    else:
        mine = torch.tensor(list(m.peer_export()), dtype=torch.uint8)
        allh = [torch.zeros(64, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(allh, mine)
        m.peer_attach(rank, world, bytes(torch.cat(allh).tolist()))


def run(mode, rank, local, world):
    m = NPEModel(ModelParams(device=local))
    attach(m, rank, world, mode)
    _, batches = balls_batches(60, 5, seed=91)
    params = npe.init_params(1234)
    cuts = np.linspace(0, 50, world + 1).astype(int)
    cuts[1:-1] += 1                      # uneven shards
    lo, hi = cuts[rank], cuts[rank + 1]
    x = params.copy(); st = optim.adam_init(x.size)
    m.set_params(params); m.optim_reset(0.0)
    for it in range(20):
        this, ctx, y, *_ = batcher.make_batch(*batches[it % len(batches)], offset=it % 20)
        g_ref, _ = npe.bp(x, (this, ctx, y))
        m._upload((this[lo:hi], ctx[lo:hi], y[lo:hi]))
        if it == 0 and mode == "nccl":
            m.fp(None, (this[lo:hi], ctx[lo:hi], y[lo:hi]))
            m.bp(divisor_batch=50)
            m.allreduce_grads()
            g = m.get_grads()
            err = float(np.max(np.abs(g - g_ref)) / np.max(np.abs(g_ref)))
            assert err < 1e-4, f"rank {rank}: all-reduced gradient error {err}"
            m._upload((this[lo:hi], ctx[lo:hi], y[lo:hi]))
        m.train_step("adam", 3e-4, divisor_batch=50, want_loss=(it % 3 == 0))
        if mode == "p2p" and it == 0:
            g = m.get_grads()            # the fused kernel leaves the all-reduced gradient in grad_params
            err = float(np.max(np.abs(g - g_ref)) / np.max(np.abs(g_ref)))
            assert err < 1e-4, f"rank {rank}: peer all-reduced gradient error {err}"
        optim.adam_step(x, g_ref, st, 3e-4)
    got = m.get_params()
    worst = float(np.max(np.abs(got - x)) / np.max(np.abs(x)))
    assert worst < 1e-4, f"rank {rank} [{mode}]: parameters after 20 DP steps differ by {worst}"
    allp = [torch.zeros(got.size) for _ in range(world)]
    dist.all_gather(allp, torch.from_numpy(got))
    assert all(torch.equal(allp[0], p) for p in allp), f"[{mode}] replicas diverged"
    # the free-running loop (npe_train_steps: no host wait between steps, kernels of consecutive steps overlap through
    # programmatic dependent launch) must give the same bits as stepping with a host wait after every step
    from oracle import scenes
    stt = scenes.raw_to_state(scenes.gen_balls(60, 4, 12, seed=200 + rank))
    rng = np.random.default_rng(300 + rank)
    nb, nsteps = 7 + rank % 2, 60
    sc = rng.integers(0, 60, nb * nsteps); ob = rng.integers(0, 4, nb * nsteps); t0 = rng.integers(0, 9, nb * nsteps)
    gB = int(sum(7 + r % 2 for r in range(world)))
    m.upload_scenes(stt); m.upload_foci(sc, ob, t0)
    m.set_params(params); m.optim_reset(0.0)
    m.train_steps(0, nb, nsteps, "adam", 3e-4, divisor_batch=gB)
    free = m.get_params()
    dist.barrier()
    m.set_params(params); m.optim_reset(0.0)
    for s_ in range(nsteps):
        m.select_foci(s_ * nb, nb)
        m.train_step("adam", 3e-4, divisor_batch=gB, want_loss=True)
    stepped = m.get_params()
    assert np.array_equal(free, stepped), f"rank {rank} [{mode}]: free-running loop differs from stepped loop"
    allp = [torch.zeros(free.size) for _ in range(world)]
    dist.all_gather(allp, torch.from_numpy(free))
    assert all(torch.equal(allp[0], p) for p in allp), f"[{mode}] replicas diverged in the free-running loop"
    if rank == 0:
        print(f"dp_gpu_check ok [{mode}]: world={world} param error after 20 steps {worst:.2e}, replicas bit-identical, "
              f"free-running == stepped")
    dist.barrier()
    m.close()


def main():
    rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    for mode in ("nccl", "p2p"):
        run(mode, rank, local, world)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
