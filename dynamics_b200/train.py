"""Host-side mirror of src/lua/main.lua (the caller of the hot path): same options, modes and training protocol, with
the per-iteration work (sample -> fp -> bp -> optimizer) enqueued on the GPU through libnpe_cuda.so.

    python -m dynamics_b200.train -mode exp -model npe -nbrhd -rs -opt adam -name balls \
        -dataset_folders "{'balls_n3_t60_ex50000_m','balls_n4_t60_ex50000_m'}" \
        -test_dataset_folders "{'balls_n6_t60_ex50000_m'}" -data_root ../../data -logs_root logs

  options                     main.lua:23-73          debug overrides        main.lua:78-103
  inittrain / initsavebatches main.lua:154-245        feval_train            main.lua:247-269
  train                       main.lua:271-396        test / validate        main.lua:398-437
  run_experiment(_load)       main.lua:456-540        modes                  main.lua:563-574

Data parallel (the reference is single-process): launched with one process per GPU
(`python -m torch.distributed.run --nproc-per-node N -m dynamics_b200.train ...`), every rank samples its own batches
from the same folders, back-propagates with the GLOBAL batch (N x batch_size) as divisor and the gradient is exchanged
inside the optimiser kernel, so the replicas stay bit-identical; rank 0 alone prints, logs and writes checkpoints.

With -rs (uniform batch sampling) and -L2 0 the samples of the next chunk of iterations are drawn up front (they do
not depend on the losses), their focus triples are uploaded once and the whole chunk runs as ONE npe_train_steps call;
a chunk ends wherever the reference prints, validates, saves or decays the learning rate, so every host-visible event
happens at the same iteration with the same values.  Priority sampling (-rs off) depends on the previous loss and
runs one iteration per call.
"""
import argparse
import glob
import math
import os
import sys

import numpy as np

from . import dataprep, t7
from .loader import MAXWINSIZE, GeneralDataSampler
from .model import ModelParams, NPEModel, init_params

CONFIG_ARGS = {"maxwinsize": MAXWINSIZE, "subdivide": True, "shuffle": True}     # config.lua:26-30
MAX_CHUNK = 4096


def parse_lua_list(s):
    """-dataset_folders "{'a','b'}"  (main.lua:147-148 evaluates the string as a Lua table constructor)."""
    if isinstance(s, (list, tuple)):
        return list(s)
    return [x for x in (p.strip().strip("'\"") for p in s.strip().strip("{}").split(",")) if x]


def build_parser():
    p = argparse.ArgumentParser(prog="dynamics_b200.train", description="main.lua options (same names and defaults)")
    a = p.add_argument
    a("-mode", default="exp", choices=["exp", "expload", "save"])
    a("-debug", action="store_true")
    a("-logs_root", default="logs"); a("-data_root", default="../../data")
    a("-model", default="npe"); a("-name", default=""); a("-seed", type=int, default=0)
    a("-dataset_folders", default=""); a("-test_dataset_folders", default="")
    a("-rnn_dim", type=int, default=50); a("-nbrhd", action="store_true"); a("-nbrhdsize", type=float, default=3.5)
    a("-layers", type=int, default=5); a("-num_past", type=int, default=2); a("-num_future", type=int, default=1)
    a("-opt", default="rmsprop", choices=["rmsprop", "adam"]); a("-batch_size", type=int, default=50)
    a("-max_iter", type=int, default=1200000); a("-L2", type=float, default=0.0); a("-lr", type=float, default=0.0003)
    a("-lrdecay", type=float, default=0.99); a("-val_window", type=int, default=10); a("-val_eps", type=float, default=1e-6)
    a("-vlambda", type=float, default=1.0); a("-lambda", dest="lambda_", type=float, default=1.0)
    a("-rs", action="store_true"); a("-sharpen", type=float, default=1.0)
    a("-print_every", type=int, default=500); a("-save_every", type=int, default=100000)
    a("-val_every", type=int, default=100000); a("-lrdecay_every", type=int, default=2500)
    a("-lrdecayafter", type=int, default=50000); a("-fast", action="store_true")
    a("-device", type=int, default=0)
    a("-dp", default="p2p", choices=["p2p", "nccl"], help="gradient exchange when launched with one process per GPU")
    return p


def parse_args(argv=None):
    mp = build_parser().parse_args(argv)
    if mp.debug:                                                   # main.lua:78-101 (layers / rnn_dim stay 5 / 50:
        mp.num_future = 1; mp.batch_size = 5; mp.max_iter = 60     #  the kernels are built for the paper's topology)
        mp.nbrhd = True; mp.lr = 3e-3; mp.lrdecay = 0.5; mp.lrdecayafter = 20; mp.lrdecay_every = 20
        mp.val_window = 5; mp.val_eps = 2e-5; mp.print_every = 1; mp.save_every = 20; mp.val_every = 20
        mp.rs = True; mp.fast = True
    mp.winsize = mp.num_past + mp.num_future
    mp.name = mp.name.replace("{", "").replace("}", "").replace("'", "")
    mp.savedir = os.path.join(mp.logs_root, mp.name)
    mp.dataset_folders = parse_lua_list(mp.dataset_folders)
    mp.test_dataset_folders = parse_lua_list(mp.test_dataset_folders)
    return mp


def chunk_len(done, periods, remaining, cap=MAX_CHUNK):
    """Iterations that can run back to back after `done` finished ones before the reference would print, validate,
    save or decay the learning rate: the distance to the next multiple of any period."""
    return max(1, min(min(p - done % p for p in periods), remaining, cap))


class Logger:
    """optim.Logger: a header of tab-separated names, then one tab-separated row per add().  path None: discard (the
    ranks other than 0 of a data-parallel run)."""

    def __init__(self, path):
        self.fh = None
        if path is not None:
            os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
            self.fh = open(path, "w")
        self.names = None

    def add(self, row):
        if self.fh is None:
            return
        if self.names is None:
            self.names = list(row)
            self.fh.write("\t".join(self.names) + "\t\n")
        self.fh.write("\t".join(f"{row[k]:.6g}" if isinstance(row[k], float) else str(row[k]) for k in self.names) + "\t\n")
        self.fh.flush()


def get_last_snapshot(savedir):
    """getLastSnapshot (main.lua:553-560): newest file whose name contains 'epoch'."""
    files = [f for f in glob.glob(os.path.join(savedir, "*")) if "epoch" in os.path.basename(f).lower()]
    return max(files, key=os.path.getmtime) if files else None


class Experiment:
    def __init__(self, mp):
        self.mp = mp
        self.train_losses, self.val_losses, self.test_losses = [], [], []
        self.model = None
        # one process per GPU under torchrun: RANK / LOCAL_RANK / WORLD_SIZE
        self.rank = int(os.environ.get("RANK", 0)); self.world = int(os.environ.get("WORLD_SIZE", 1))
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            if not dist.is_initialized():
                dist.init_process_group("gloo", rank=self.rank, world_size=self.world)
            self.dist = dist
            mp.device = int(os.environ.get("LOCAL_RANK", self.rank))
        self.rng = np.random.default_rng(mp.seed + 1000 * self.rank)      # training batches differ per rank
        self.rng_eval = np.random.default_rng(mp.seed)                    # validation order is the same on every rank
        self.global_batch = mp.batch_size * self.world
        self.lr = mp.lr
        self.say = print if self.rank == 0 else (lambda *a, **k: None)
        self.stepwise = False          # True: one device call per iteration even with -rs (what priority sampling does)
        self.history = None            # set to [] to record (dataset, sampled_id, lr, loss) of every iteration

    # ---- main.lua:216-245 ----
    def initsavebatches(self):
        mp = self.mp
        if self.rank != 0:                      # rank 0 writes the batch files, the others wait for them
            self.dist.barrier()
            return
        for folder in list(mp.dataset_folders) + list(mp.test_dataset_folders):
            root = os.path.join(mp.data_root, folder)
            if os.path.isdir(os.path.join(root, "batches")):
                self.say(f"Batches for {folder} already made")
                continue
            jsons = sorted(glob.glob(os.path.join(root, "jsons", "*.json")))
            self.say(f"Saving batches of size {mp.batch_size} from {root}/jsons into {root}/batches")
            dataprep.json_to_batch_files(jsons, root, batch_size=mp.batch_size, seed=mp.seed)
        if self.dist is not None:
            self.dist.barrier()

    # ---- main.lua:154-214 ----
    def inittrain(self, preload=False, model_path=None, iters=None):
        mp = self.mp
        self.say("Network parameters:"); self.say(vars(mp))
        mparams = ModelParams(batch_size=mp.batch_size, rnn_dim=mp.rnn_dim, layers=mp.layers, num_past=mp.num_past,
                              num_future=mp.num_future, nbrhd=mp.nbrhd, nbrhdsize=mp.nbrhdsize, vlambda=mp.vlambda,
                              lambda_=mp.lambda_, model=mp.model, device=mp.device)
        self.model = NPEModel.create(mparams, preload, model_path, None if preload else init_params(mp.seed))
        if self.world > 1:
            self.model.attach_data_parallel(self.dist, self.rank, self.world, mp.dp)
        args = {"data_root": mp.data_root, "dataset_folders": mp.dataset_folders, "maxwinsize": CONFIG_ARGS["maxwinsize"],
                "winsize": mp.winsize, "num_past": mp.num_past, "num_future": mp.num_future, "relative": True,
                "sim": False, "subdivide": CONFIG_ARGS["subdivide"], "shuffle": CONFIG_ARGS["shuffle"],
                "model": self.model, "batch_size": mp.batch_size, "rng": self.rng}
        eval_args = dict(args); eval_args["rng"] = self.rng_eval
        test_args = dict(eval_args); test_args["dataset_folders"] = mp.test_dataset_folders
        # without -rs the batch is drawn from the loss table (main.lua:249-253): kept on the device
        self.train_loader = GeneralDataSampler.create("trainset", dict(args, device_priority=not mp.rs))
        self.val_loader = GeneralDataSampler.create("valset", dict(eval_args))
        self.test_loader = GeneralDataSampler.create("testset", test_args)
        self.train_test_loader = GeneralDataSampler.create("trainset", dict(eval_args))
        lead = self.rank == 0
        if lead:
            os.makedirs(mp.savedir, exist_ok=True)
        self.train_logger = Logger(os.path.join(mp.savedir, f"train_{iters}.log" if iters else "train.log") if lead else None)
        self.experiment_logger = Logger(os.path.join(mp.savedir, "experiment.log") if lead else None)
        if lead:
            t7.save(os.path.join(mp.savedir, f"args{iters}.t7" if iters else "args.t7"),
                    {"mp": _plain(vars(mp)), "config_args": dict(CONFIG_ARGS)})
        self.model.optim_reset(0.0)
        self.say("Initialized Network")

    # ---- one reference iteration: feval_train + optimizer (main.lua:247-269,301-304) ----
    def _sample(self):
        if self.mp.rs:
            return self.train_loader.sample_random_batch()[0]
        return self.train_loader.sample_priority_batch(self.mp.sharpen)[0]

    def _step(self):
        m, mp = self.model, self.mp
        batch = self._sample()
        if mp.L2 > 0:
            if self.world > 1:
                raise NotImplementedError("-L2 > 0 runs the optimiser from a host-edited gradient: single GPU only")
            loss, _, _, _ = m.fp(None, batch)
            grad = m.bp(batch)
            params = m.get_params()
            loss = loss + mp.L2 * float(np.dot(params.astype(np.float64), params)) / 2
            m.set_grads(grad + mp.L2 * params)
            (m.adam_step if mp.opt == "adam" else m.rmsprop_step)(self.lr)
        else:
            batch.select(m)
            loss = float(m.train_step(mp.opt, self.lr, divisor_batch=self.global_batch)[0])
        self.train_loader.update_batch_weight(loss)
        if self.history is not None:
            self.history.append((self.train_loader.current_dataset, self.train_loader.current_sampled_id, self.lr, loss))
        return loss

    def _chunk(self, n):
        """n iterations of uniform sampling as one device call; returns the n losses."""
        m, tl = self.model, self.train_loader
        scene, t0, who = [], [], []
        for _ in range(n):
            b, ds = tl.sample_random_batch()
            scene.append(b.scene); t0.append(b.t0); who.append((ds, tl.current_sampled_id))
        b.store.commit()
        scene = np.concatenate(scene); t0 = np.concatenate(t0)
        m.upload_foci(scene, np.zeros_like(scene), t0)
        losses = m.train_steps(0, self.mp.batch_size, n, self.mp.opt, self.lr, divisor_batch=self.global_batch,
                               want_losses=True)[:, 0]
        for (ds, sid), loss in zip(who, losses):                  # update_batch_weight of every iteration, in order
            tl.datasamplers[ds - 1].priority_sampler.update_batch_weight(sid, float(loss))
            if self.history is not None:
                self.history.append((ds, sid, self.lr, float(loss)))
        return [float(x) for x in losses]

    # ---- main.lua:271-396 ----
    def train(self, start_iter=1, epoch_num=1):
        mp, m = self.mp, self.model
        self.say("Start iter:", start_iter); self.say("Start epoch num:", epoch_num)
        if start_iter == 1:
            self._validate_and_record()
            self._save(epoch_num, 0, self.val_losses[-1], 0)
        periods = [mp.print_every, mp.val_every, mp.lrdecay_every]
        chunked = mp.rs and mp.L2 == 0 and not self.stepwise
        t = start_iter
        while t <= mp.max_iter:
            done = t - start_iter                                  # iterations finished so far
            n = chunk_len(done, periods, mp.max_iter - t + 1) if chunked else 1
            losses = self._chunk(n) if n > 1 else [self._step()]
            for loss in losses:
                self.train_logger.add({"log MSE loss (train set)": math.log(loss) if loss > 0 else float("-inf")})
            t_last = t + n - 1
            k = t_last - start_iter + 1
            if k % mp.print_every == 0:
                hb = self.train_loader.get_hardest_batch()
                self.say("epoch %2d  iteration %2d  loss = %6.8f  gradnorm = %6.4e  batch = %d-%d    hardest batch: %d-%d    "
                      "with loss %6.8f lr = %6.4e" % (epoch_num, t_last, losses[-1], float(np.linalg.norm(m.get_grads())),
                                                     self.train_loader.current_dataset, self.train_loader.current_sampled_id,
                                                     hb[2], hb[1], hb[0], self.lr))
            if k % mp.val_every == 0:
                self._validate_and_record()
                assert mp.save_every % mp.val_every == 0 or mp.val_every % mp.save_every == 0
                if k % mp.save_every == 0:
                    self._save(epoch_num, t_last, self.val_losses[-1], t_last)
                if len(self.val_losses) >= mp.val_window and self._converged():
                    break
            if t_last >= mp.lrdecayafter and k % mp.lrdecay_every == 0:
                self.lr *= mp.lrdecay
                mp.lr = self.lr
                self.say(f"Learning rate is now {self.lr}")
            nb = self.train_loader.num_batches
            if t_last // nb > (t - 1) // nb:                        # some iteration of this chunk had t % nb == 0
                epoch_num = (t_last // nb) + 1
            t = t_last + 1
        return epoch_num

    def _converged(self):                                           # main.lua:350-377
        mp = self.mp
        w = np.asarray(self.val_losses[-mp.val_window:], np.float64)
        delta = float(np.mean(w[1:] - w[:-1]))
        self.say(f"Average change in val loss over {mp.val_window} validations: {delta}")
        self.say("Loss is decreasing" if delta < 0 and int(np.argmax(w)) < int(np.argmin(w)) else "Loss is increasing")
        spread = float(w.max() - w.min())
        self.say(f"Val loss difference in a window of {mp.val_window}: {spread}")
        if spread < mp.val_eps:
            self.say(f"That is less than {mp.val_eps}. Converged.")
            return True
        return False

    def _save(self, epoch_num, step, val_loss, iters):              # main.lua:288-296,333-346
        path = "%s/epoch%d_step%d_%.7f.t7" % (self.mp.savedir, epoch_num, step, val_loss)
        if self.rank != 0:
            return path
        self.say("saving checkpoint to " + path)
        t7.save_checkpoint(path, self.model.get_params(), mp=_plain(vars(self.mp)), iters=iters,
                           train_losses=self.train_losses, val_losses=self.val_losses, test_losses=self.test_losses)
        self.say("Saved model")
        return path

    # ---- main.lua:398-437 ----
    VAL_GROUP = 64     # reference batches evaluated by ONE device call (forward only, per-example losses back in one copy)

    def _eval_group(self, batches):
        """model:fp losses (npe.lua:233-246) of several batches from one forward over all their examples.  Per batch the
        loss is (sum vel + sum angvel) / (3B) with double accumulation, the arithmetic of the device's loss kernel."""
        m = self.model
        if not (hasattr(m, "forward_per_example_current") and all(hasattr(b, "select") for b in batches)
                and len({id(b.store) for b in batches}) == 1):
            return [float(m.fp(None, b)[0]) for b in batches]
        batches[0].store.commit()
        m.upload_foci(np.concatenate([b.scene for b in batches]), np.concatenate([b.obj for b in batches]),
                      np.concatenate([b.t0 for b in batches]))
        m.select_foci(0, sum(len(b) for b in batches))
        _, le = m.forward_per_example_current()
        out, a = [], 0
        for b in batches:
            e = le[a:a + len(b)].astype(np.float64); a += len(b)
            out.append(float(np.float32((2.0 * e[:, 1].sum() + e[:, 2].sum()) / (3.0 * len(b)))))
        return out

    def test(self, loader, num_batches=None):
        """test() of main.lua:398-414, sharded: every rank walks the same sequential batch order (so the loaders stay in
        step), evaluates batches i with i % world == rank in groups of VAL_GROUP, and the per-rank loss sums meet in ONE
        scalar all-reduce."""
        num_batches = num_batches or loader.num_batches
        if self.mp.fast:
            num_batches = min(5000, num_batches)
        self.say(f"Testing {num_batches} batches")
        total, mine = 0.0, []
        for i in range(num_batches):
            batch, _ = loader.sample_sequential_batch(False)
            if i % self.world == self.rank:
                mine.append(batch)
                if len(mine) == self.VAL_GROUP:
                    total += sum(self._eval_group(mine)); mine = []
        if mine:
            total += sum(self._eval_group(mine))
        if self.world > 1:
            import torch
            t = torch.tensor([total], dtype=torch.float64)
            self.dist.all_reduce(t)
            total = float(t[0])
        return total / num_batches

    def validate(self):
        tr = self.test(self.train_test_loader, min(5000, self.val_loader.num_batches))
        va = self.test(self.val_loader, self.val_loader.num_batches)
        te = self.test(self.test_loader, self.test_loader.num_batches)
        self.say(f"train loss\t{tr}\tval loss\t{va}\ttest_loss\t{te}")
        self.experiment_logger.add({"log MSE loss (train set)": math.log(tr), "log MSE loss (val set)": math.log(va),
                                    "log MSE loss (test set)": math.log(te)})
        return tr, va, te

    def _validate_and_record(self):
        tr, va, te = self.validate()
        self.train_losses.append(tr); self.val_losses.append(va); self.test_losses.append(te)

    # ---- main.lua:456-459,490-540 ----
    def run_experiment(self):
        self.inittrain(False)
        return self.train()

    def run_experiment_load(self):
        mp = self.mp
        snapshot = get_last_snapshot(mp.savedir)
        if snapshot is None:
            raise FileNotFoundError(f"no epoch*.t7 snapshot under {mp.savedir}")
        self.say(snapshot)
        ck = t7.load(snapshot)
        for k, v in dict(ck["mp"]).items():                         # "completely overwrite" (main.lua:497)
            setattr(mp, k, v)
        for k in ("batch_size", "rnn_dim", "layers", "num_past", "num_future", "winsize", "max_iter", "print_every",
                  "save_every", "val_every", "lrdecay_every", "lrdecayafter", "val_window", "seed", "device"):
            if hasattr(mp, k):
                setattr(mp, k, int(getattr(mp, k)))
        for k in ("dataset_folders", "test_dataset_folders"):
            v = getattr(mp, k)
            setattr(mp, k, [v[i] for i in sorted(v)] if isinstance(v, dict) else list(v))
        mp.mode = "expload"
        iters = int(ck["iters"]) + 1
        as_list = lambda v: [float(v[i]) for i in sorted(v)] if isinstance(v, dict) else [float(x) for x in v]
        self.train_losses, self.val_losses, self.test_losses = (as_list(ck[k]) for k in ("train_losses", "val_losses", "test_losses"))
        if (iters - 1) >= mp.lrdecayafter and (iters - 1) % mp.lrdecay_every == 0:
            mp.lr = mp.lr * mp.lrdecay
        self.lr = mp.lr
        self.say(f"Learning rate is now {self.lr}")
        self.inittrain(True, snapshot, iters)
        for tr, va, te in zip(self.train_losses, self.val_losses, self.test_losses):
            self.experiment_logger.add({"log MSE loss (train set)": math.log(tr), "log MSE loss (val set)": math.log(va),
                                        "log MSE loss (test set)": math.log(te)})
        epoch_num = iters // self.train_loader.num_batches + 1
        return self.train(iters, epoch_num)


def _plain(d):
    return {k: (list(v) if isinstance(v, (list, tuple)) else v) for k, v in d.items()
            if isinstance(v, (int, float, str, bool, list, tuple)) or v is None}


def main(argv=None):
    mp = parse_args(argv)
    exp = Experiment(mp)
    if mp.mode == "exp":
        exp.initsavebatches()
        print("Running experiment.")
        exp.run_experiment()
    elif mp.mode == "expload":
        exp.run_experiment_load()
    elif mp.mode == "save":
        exp.initsavebatches()
    return exp


if __name__ == "__main__":
    main(sys.argv[1:])
