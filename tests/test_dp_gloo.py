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


def _train_worker(rank, world, port, root, out):
    """dynamics_b200.train under a 2-process launch with a stand-in for the device model: what each rank samples, the
    divisor it passes and who writes the logs."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), LOCAL_RANK=str(rank),
                      WORLD_SIZE=str(world))
    from dynamics_b200 import train
    from tests.test_train_host import FakeModel

    class DPFake(FakeModel):
        divisors = []

        def attach_data_parallel(self, d, r, w, mode="p2p"):
            DPFake.attached = (r, w, mode)

        fp_calls = 0

        def fp(self, params_, batch, sim=False):
            DPFake.fp_calls += 1
            return super().fp(params_, batch, sim)

        def train_step(self, opt="adam", lr=3e-4, divisor_batch=0, want_loss=True):
            DPFake.divisors.append(divisor_batch)
            return super().train_step(opt, lr, divisor_batch, want_loss)

        def train_steps(self, first, batch, nsteps, opt="adam", lr=3e-4, divisor_batch=0, want_losses=False):
            DPFake.divisors.append(divisor_batch)
            return FakeModel.train_steps(self, first, batch, nsteps, opt, lr, 0, want_losses)

    train.NPEModel = DPFake
    argv = ["-mode", "exp", "-nbrhd", "-rs", "-opt", "adam", "-name", "dp", "-data_root", root, "-logs_root",
            os.path.join(root, "logs"), "-dataset_folders", "{'balls_n3_t60_ex100_m'}", "-test_dataset_folders",
            "{'balls_n3_t60_ex100_m'}", "-max_iter", "12", "-print_every", "5", "-val_every", "10", "-save_every", "10",
            "-seed", "1"]
    e = train.Experiment(train.parse_args(argv))
    e.history = []
    e.run_experiment()
    out[rank] = {"ids": [(h[0], h[1]) for h in e.history], "div": sorted(set(DPFake.divisors) - {0}),
                 "attached": DPFake.attached, "val": list(e.val_losses), "device": e.mp.device, "fp_calls": DPFake.fp_calls,
                 "val_batches": 2 * (min(5000, e.val_loader.num_batches) + e.val_loader.num_batches + e.test_loader.num_batches)}
    dist.barrier()
    dist.destroy_process_group()


def test_training_driver_under_two_ranks(tmp_path):
    from tests.test_loader import _write_dataset
    _write_dataset(tmp_path, "balls_n3_t60_ex100_m", 3, 100, seed=1)
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + ((os.getpid() + 17) % 500)
    mp.spawn(_train_worker, args=(2, port, str(tmp_path), out), nprocs=2, join=True)
    a, b = out[0], out[1]
    assert a["attached"] == (0, 2, "p2p") and b["attached"] == (1, 2, "p2p") and (a["device"], b["device"]) == (0, 1)
    assert a["div"] == [100] and b["div"] == [100]                 # global batch = 2 x 50 is the divisor on both ranks
    assert len(a["ids"]) == len(b["ids"]) == 12 and a["ids"] != b["ids"]   # each rank samples its own batches
    assert a["val"] == b["val"]                                     # one all-reduced loss sum: identical on every rank
    # validation is sharded (main.lua:398-437 on 2 ranks): each rank evaluated half of the batches of the 2 validations
    assert a["fp_calls"] + b["fp_calls"] == a["val_batches"] and abs(a["fp_calls"] - b["fp_calls"]) <= 6
    logs = os.listdir(tmp_path / "logs" / "dp")
    assert "train.log" in logs and sum(f.startswith("epoch") for f in logs) == 2   # written once (rank 0)
