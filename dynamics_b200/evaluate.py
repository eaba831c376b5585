"""Host-side mirror of ``th eval.lua -mode sim`` (src/lua/eval.lua:33-117,237-454,468-533): load the newest checkpoint
of an experiment, roll every test batch forward `-steps` steps on the GPU, write gt_*/pred_* trajectories for the
renderer and the per-step gt_divergence.log.

    python -m dynamics_b200.evaluate -mode sim -name balls -test_dataset_folders "{'balls_n6_t60_ex50000_m'}" \
        -data_root ../../data -logs_root logs -ns 3 -steps 58

Per batch file the reference runs N x steps fp_batch calls (every object of the 50 examples as focus, eval.lua:296-376);
here that is one npe_rollout_slots call on the file's 50 example scenes, whose per-slot metric sums reproduce the
reference's per-batch averages (model.slot_metrics_to_batch_row).  The json names use the batch FILE id: the reference
names them after `current_sampled_id`, which load_batch_id_first_offset never sets (data_sampler.lua:149-151,198-199),
so all of its -ns files collapse onto batch0.json.
"""
import argparse
import os
import sys

import numpy as np

from . import dataprep, t7
from .loader import MAXWINSIZE, MAXWINSIZE_LONG, GeneralDataSampler
from .model import ModelParams, NPEModel, slot_metrics_to_batch_row
from .train import Logger, get_last_snapshot, parse_lua_list

LOG_COLUMNS = (("MSE Error", "loss"), ("Cosine Difference", "cos"), ("Magnitude Difference", "mag"),
               ("Velocity Error", "vel"), ("Angular Velocity Error", "angvel"))       # eval.lua:437-444


def build_parser():
    p = argparse.ArgumentParser(prog="dynamics_b200.evaluate", description="eval.lua options (same names and defaults)")
    a = p.add_argument
    a("-mode", default="sim", choices=["sim"])
    a("-debug", action="store_true")
    a("-logs_root", default="logs"); a("-data_root", default="../../data")
    a("-name", default=""); a("-test_dataset_folders", default="")
    a("-ns", type=int, default=3); a("-steps", type=int, default=58)
    a("-device", type=int, default=0)
    a("-precision", default="fp32", choices=["fp32", "tf32x3", "tf32"])
    return p


def save_ex_pred_json(scenes, traj, numsteps, jsonfile, dataset_name, subfolder, num_past=2, num_future=1):
    """eval.lua:468-497: past frames + prediction -> pred_<file>, past + ground truth -> gt_<file>."""
    flags = dataprep.dataset_flags(dataset_name)
    config = {"num_past": num_past, "num_future": num_future, "env": flags["env"], "numObj": flags.get("n"),
              "gravity": False, "friction": False, "pairwise": False}
    past = scenes[:, :, :num_past]
    dataprep.record_trajectories(np.concatenate([past, traj], 2), os.path.join(subfolder, "pred_" + jsonfile), config)
    dataprep.record_trajectories(scenes[:, :, :num_past + numsteps], os.path.join(subfolder, "gt_" + jsonfile), config)


def simulate_all(model, sampler, numsteps, savedir, ns=3, saveoutput=True, dist=None, rank=0, world=1):
    """eval.lua:237-454 for one dataset's sampler.  Returns the (total_batches, numsteps) tables by log column.
    Under a one-process-per-GPU launch the batch files are sharded by trajectory (file i goes to rank i mod world; no
    collective on the data path), the tables are summed over ranks at the end and rank 0 writes the log."""
    name = os.path.basename(os.path.normpath(sampler.dataset_folder))
    subfolder = os.path.join(savedir, name + "_predictions")
    os.makedirs(subfolder, exist_ok=True)
    assert numsteps <= sampler.maxwinsize - sampler.num_past, \
        f"Number of predictive steps should be less than {sampler.maxwinsize - sampler.num_past + 1}"
    tables = {k: np.zeros((sampler.total_batches, numsteps)) for _, k in LOG_COLUMNS}
    for i in range(sampler.total_batches):
        sampler.sample_sequential_batch()                          # every rank walks the same order
        if i % world != rank:
            continue
        file_id = int(sampler.batch_idxs[sampler.current_batch - 1])
        scenes = sampler.file_scenes(file_id)                      # (B, N, T, 22): object 0 = focus, then contexts
        model.upload_scenes(scenes)
        sampler.store.dirty = True                                 # the training store no longer owns the device scenes
        traj, slots = model.rollout_slots(0, numsteps, use_gt=True, want_traj=saveoutput and i < ns)
        row = slot_metrics_to_batch_row(slots, scenes.shape[0])
        for _, k in LOG_COLUMNS:
            tables[k][i] = row[k]
        if saveoutput and i < ns:
            save_ex_pred_json(scenes, traj, numsteps, f"batch{file_id}.json", name, subfolder, sampler.num_past,
                              sampler.num_future)
    if dist is not None:
        import torch
        for _, k in LOG_COLUMNS:                                   # each row was filled by exactly one rank
            t = torch.from_numpy(np.nan_to_num(tables[k], nan=0.0)); nanmask = torch.from_numpy(np.isnan(tables[k]).astype(np.float64))
            dist.all_reduce(t); dist.all_reduce(nanmask)
            tables[k] = np.where(nanmask.numpy() > 0, np.nan, t.numpy())
    logger = Logger(os.path.join(subfolder, "gt_divergence.log") if rank == 0 else None)
    means = {k: tables[k].mean(0) for _, k in LOG_COLUMNS}
    for tt in range(numsteps):
        row = {"Timesteps": tt + 1}
        row.update({col: float(means[k][tt]) for col, k in LOG_COLUMNS})
        logger.add(row)
    return tables


def predict_simulate_all(mp):
    """eval.lua:513-526."""
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dist = None
    if world > 1:
        import torch.distributed as dist
        if not dist.is_initialized():
            dist.init_process_group("gloo", rank=rank, world_size=world)
        mp.device = int(os.environ.get("LOCAL_RANK", rank))
    savedir = os.path.join(mp.logs_root, mp.name)
    snapshot = get_last_snapshot(savedir)
    if snapshot is None:
        raise FileNotFoundError(f"no epoch*.t7 snapshot under {savedir}")
    print("Last Snapshot: " + snapshot)
    saved = dict(t7.load(os.path.join(savedir, "args.t7"))["mp"]) if os.path.exists(os.path.join(savedir, "args.t7")) else {}
    mparams = ModelParams(batch_size=int(saved.get("batch_size", 50)), nbrhd=bool(saved.get("nbrhd", True)),
                          nbrhdsize=float(saved.get("nbrhdsize", 3.5)), device=mp.device)
    model = NPEModel.create(mparams, True, snapshot)
    model.set_precision(mp.precision)
    maxwin = MAXWINSIZE_LONG if "tower" in savedir else MAXWINSIZE                    # eval.lua:93-100
    args = {"data_root": mp.data_root, "dataset_folders": parse_lua_list(mp.test_dataset_folders), "maxwinsize": maxwin,
            "winsize": 3, "num_past": 2, "num_future": 1, "relative": True, "subdivide": False, "shuffle": True,
            "sim": True, "model": model, "batch_size": mparams.batch_size, "rng": np.random.default_rng(123)}
    test_loader = GeneralDataSampler.create("testset", args)
    out = {}
    for folder, sampler in zip(test_loader.dataset_folders, test_loader.datasamplers):
        print("Evaluating " + folder)
        out[folder] = simulate_all(model, sampler, mp.steps, savedir, mp.ns, True, dist, rank, world)
    model.close()
    return out


def main(argv=None):
    mp = build_parser().parse_args(argv)
    mp.name = mp.name.replace("{", "").replace("}", "").replace("'", "")
    return predict_simulate_all(mp)


if __name__ == "__main__":
    main(sys.argv[1:])
