"""The training driver (dynamics_b200.train, mirror of main.lua) under a one-process-per-GPU launch:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tests/dp_train_check.py
Every rank samples its own batches, the gradient is exchanged inside the optimiser kernel; at the end the replicas must
be bit-identical, must have moved away from the initial weights, and rank 0 alone has written logs and checkpoints."""
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dynamics_b200 import train  # noqa: E402
from dynamics_b200.model import init_params  # noqa: E402
from tests.test_loader import _write_dataset  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)
    root = [tempfile.mkdtemp() if rank == 0 else None]
    dist.broadcast_object_list(root, 0)
    root = root[0]
    if rank == 0:
        from pathlib import Path
        _write_dataset(Path(root), "balls_n3_t60_ex100_m", 3, 100, seed=1)
        _write_dataset(Path(root), "balls_n4_t60_ex100_m", 4, 100, seed=2)
    dist.barrier()
    for mode in ("p2p", "nccl"):
        argv = ["-mode", "exp", "-nbrhd", "-rs", "-opt", "adam", "-name", "dp_" + mode, "-data_root", root, "-logs_root",
                os.path.join(root, "logs"), "-dataset_folders", "{'balls_n3_t60_ex100_m','balls_n4_t60_ex100_m'}",
                "-test_dataset_folders", "{'balls_n4_t60_ex100_m'}", "-max_iter", "23", "-print_every", "5", "-val_every", "10",
                "-save_every", "10", "-lrdecay_every", "4", "-lrdecayafter", "8", "-lrdecay", "0.5", "-seed", "3", "-dp", mode]
        e = train.Experiment(train.parse_args(argv))
        e.history = []
        e.run_experiment()
        p = e.model.get_params()
        allp = [torch.zeros(p.size) for _ in range(world)]
        dist.all_gather(allp, torch.from_numpy(p))
        assert all(torch.equal(allp[0], q) for q in allp), f"[{mode}] replicas diverged"
        assert np.max(np.abs(p - init_params(3))) > 1e-4 and np.isfinite(p).all()
        ids = [None] * world
        dist.all_gather_object(ids, [(h[0], h[1]) for h in e.history])
        assert len(ids[0]) == 23 and (world == 1 or ids[0] != ids[1]), "ranks must sample their own batches"
        vals = [None] * world
        dist.all_gather_object(vals, list(e.val_losses))
        assert all(v == vals[0] for v in vals) and len(vals[0]) == 3, "identical replicas validate identically"
        if rank == 0:
            logs = os.listdir(os.path.join(root, "logs", "dp_" + mode))
            assert "train.log" in logs and sum(f.startswith("epoch") for f in logs) == 3
            print(f"dp_train_check ok [{mode}]: world={world}, 23 iterations, replicas bit-identical, val losses {vals[0]}")
        e.model.close()
        dist.barrier()
    # eval.lua -mode sim on the p2p run's newest checkpoint: batch files sharded over the ranks == all files on one rank
    from dynamics_b200 import evaluate
    argv = ["-mode", "sim", "-name", "dp_p2p", "-test_dataset_folders", "{'balls_n4_t60_ex100_m'}", "-data_root", root,
            "-logs_root", os.path.join(root, "logs"), "-ns", "1", "-steps", "6"]
    sharded = evaluate.main(argv)["balls_n4_t60_ex100_m"]
    dist.barrier()
    if rank == 0:
        os.environ["WORLD_SIZE"] = "1"                       # the same driver, unsharded, on rank 0's GPU
        single = evaluate.main(argv)["balls_n4_t60_ex100_m"]
        os.environ["WORLD_SIZE"] = str(world)
        for k in sharded:
            assert np.array_equal(sharded[k], single[k], equal_nan=True), k
        log = open(os.path.join(root, "logs", "dp_p2p", "balls_n4_t60_ex100_m_predictions", "gt_divergence.log")).read()
        assert len(log.splitlines()) == 7
        print(f"dp_train_check ok [rollout]: {sharded['loss'].shape[0]} batch files sharded over {world} ranks == one rank")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
