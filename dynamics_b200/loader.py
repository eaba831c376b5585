"""Host-side mirror of the reference's loader boundary over a DEVICE-RESIDENT example store.

  PrioritySampler      <- src/lua/priority_sampler.lua:4-63
  DataSampler          <- src/lua/data_sampler.lua:23-227   (one dataset folder, one of trainset / valset / testset)
  GeneralDataSampler   <- src/lua/general_data_sampler.lua:23-139 (``D.create(name, args)`` of main.lua:180-183)

The reference re-reads a 60-frame ``.t7`` batch file from disk on every iteration to use 3 frames of it
(data_sampler.lua:156-157).  Here every batch file of every folder is read ONCE, stored in HBM as "example scenes"
(object 0 = the focus, objects 1.. = its contexts in file order) and a sampled batch is just 50 (scene, object 0, t0)
triples: the window cut, the relative target (data_process.lua:87-96), the 12-nearest trim (:51-70) and the
neighbourhood mask are computed by the pair-builder kernel when the batch is selected.  Method names, id arithmetic
(sub-batch id -> file id, offset; data_sampler.lua:198-203) and the priority table follow the reference; the random
streams are numpy Generators (Lua's math.random / torch.randperm streams are not reproducible outside Torch7).
"""
import os
import re

import numpy as np

from . import t7

MAXWINSIZE = 60          # config.lua:27
MAXWINSIZE_LONG = 120    # config.lua:28
MAX_CTX = 12             # data_sampler.lua:169


class PrioritySampler:
    """priority_sampler.lua: a table of the last loss seen per batch; 10 % uniform, else multinomial(weights^pow).
    With `model` (an NPEModel) and a table index the weights live on the DEVICE and the draw (scan of weight^pow + binary
    search, counter-based random numbers) runs there (npe_priority_*); the host keeps only which batches have been seen."""

    _next_table = {}

    def __init__(self, num_batches, rng=None, model=None):
        self.batch_weights = np.zeros(num_batches, np.float64)
        self.alpha = 0.1
        self.epc_num = 1
        self.table_is_full = False
        self.num_batches = num_batches
        self.rng = rng or np.random.default_rng(0)
        self.model, self.table = None, None
        if model is not None and hasattr(model, "priority_init"):
            t = PrioritySampler._next_table.get(id(model), 0)
            if t < 16:
                PrioritySampler._next_table[id(model)] = t + 1
                model.priority_init(t, num_batches, self.alpha)
                self.model, self.table = model, t

    def update_batch_weight(self, batch_id, weight):            # 1-based id, priority_sampler.lua:18-22
        self.batch_weights[batch_id - 1] = weight
        self.table_is_full = bool(self.batch_weights.min() > 0)
        if self.model is not None:
            self.model.priority_update(self.table, [batch_id], [weight])

    def sample(self, pow_=1):                                     # priority_sampler.lua:37-48
        if self.model is not None:
            return int(self.model.priority_sample(self.table, pow_ or 1, int(self.rng.integers(0, 2 ** 63)), 1)[0])
        if self.rng.integers(1, 101) / 100.0 < self.alpha:
            return int(self.rng.integers(1, self.num_batches + 1))
        sharpened = np.power(self.batch_weights, pow_ if pow_ else 1)
        if sharpened.min() <= 0:
            raise AssertionError("priority sampler: every batch needs a positive weight before sampling")
        p = sharpened / sharpened.sum()
        return int(self.rng.choice(self.num_batches, p=p)) + 1

    def get_hardest_batch(self):                                  # priority_sampler.lua:50-54 -> {max, argmax}
        if self.model is not None:
            return self.model.priority_hardest(self.table)
        i = int(np.argmax(self.batch_weights))
        return [float(self.batch_weights[i]), i + 1]

    def set_alpha(self, v):
        self.alpha = v
        if self.model is not None:                                 # re-create the device table with the new alpha, same weights
            self.model.priority_init(self.table, self.num_batches, v)
            seen = np.nonzero(self.batch_weights)[0]
            if seen.size:
                self.model.priority_update(self.table, seen + 1, self.batch_weights[seen])

    def set_epcnum(self, v):
        self.epc_num = v


class SceneStore:
    """All examples of every registered batch file as one padded scene tensor in the model's HBM.

    add() queues host arrays, commit() pads them to a common object count and uploads once (npe_scenes_upload);
    afterwards only 12-byte-per-example focus triples travel per sampled batch."""

    def __init__(self, model):
        self.model = model
        self.parts = []          # (states (S,N,T,22))
        self.bases = []
        self.n_scenes = 0
        self.dirty = False

    @classmethod
    def of(cls, model):
        st = getattr(model, "_scene_store", None)
        if st is None:
            st = model._scene_store = cls(model)
        return st

    def add(self, states):
        base = self.n_scenes
        self.parts.append(np.asarray(states, np.float32))
        self.bases.append(base)
        self.n_scenes += states.shape[0]
        self.dirty = True
        return base

    def commit(self):
        if not self.dirty:
            return
        Nmax = max(p.shape[1] for p in self.parts)
        T = max(p.shape[2] for p in self.parts)
        states = np.zeros((self.n_scenes, Nmax, T, 22), np.float32)
        n_obj = np.zeros(self.n_scenes, np.int32)
        for base, p in zip(self.bases, self.parts):
            states[base:base + p.shape[0], :p.shape[1], :p.shape[2]] = p
            n_obj[base:base + p.shape[0]] = p.shape[1]
        self.model.upload_scenes(states, n_obj)
        self.dirty = False

    def host_example(self, scene):
        for base, p in zip(reversed(self.bases), reversed(self.parts)):
            if scene >= base:
                return p[scene - base]
        raise IndexError(scene)


class DeviceBatch:
    """What a sampler returns in place of the reference's 7-tuple: the focus triples of one batch.  NPEModel.fp /
    fp_batch / bp accept it directly; host() materialises the reference tuple for inspection."""

    def __init__(self, store, scene, t0, num_past=2, num_future=1, relative=True, sim=False):
        self.store = store
        self.scene = np.ascontiguousarray(scene, np.int32)
        self.obj = np.zeros_like(self.scene)
        self.t0 = np.ascontiguousarray(t0, np.int32)
        self.num_past, self.num_future, self.relative, self.sim = num_past, num_future, relative, sim

    def __len__(self):
        return self.scene.size

    def select(self, model=None):
        m = model or self.store.model
        self.store.commit()
        m.upload_foci(self.scene, self.obj, self.t0)
        return m.select_foci(0, self.scene.size)

    def host(self):
        """{this, trimmed context, y, trimmed context_future, mask, original_batch, closest_indices}
        (data_sampler.lua:189-190); the trim indices come from the device pair table (1-based like torch)."""
        m = self.store.model
        self.select(m)
        idx, nbr, cnt = m.pair_table()
        ex = [self.store.host_example(int(s)) for s in self.scene]
        o = int(self.t0[0])
        p, f = self.num_past, self.num_future
        this = np.stack([e[0, o:o + p] for e in ex])
        ctx = np.stack([e[1:, o:o + p] for e in ex])
        end = None if self.sim else o + p + f
        y = np.stack([e[0, o + p:end] for e in ex]).copy()
        ctx_f = np.stack([e[1:, o + p:end] for e in ex])
        if self.relative and not self.sim:
            y[:, :, :6] -= this[:, -1:, :6]
        k = int(cnt[0])
        closest = idx[:, :k].astype(np.int64)
        rows = np.arange(len(ex))[:, None]
        mask = np.zeros(MAX_CTX, np.float32); mask[k - 1] = 1 if k > 0 else 0
        original = (this, ctx, y, ctx_f)
        return this, ctx[rows, closest], y, ctx_f[rows, closest], mask, original, closest + 1


def _count_batch_files(folder):
    if not os.path.isdir(folder):
        raise FileNotFoundError(f"no batch folder {folder} (run -mode save first, main.lua:216-245)")
    ids = sorted(int(m.group(1)) for m in (re.fullmatch(r"batch(\d+)", f) for f in os.listdir(folder)) if m)
    if ids != list(range(1, len(ids) + 1)):
        raise ValueError(f"{folder}: batch files are not numbered 1..{len(ids)}")
    return len(ids)


class DataSampler:
    """data_sampler.lua for one dataset folder.  args: dataset_folder, winsize, num_past, num_future, relative,
    shuffle, subdivide, sim, maxwinsize, model (NPEModel), batch_size, rng (numpy Generator)."""

    def __init__(self, dataset_name, args):
        self.dataset_folder = args["dataset_folder"]
        self.dataset_name = dataset_name
        self.maxwinsize = MAXWINSIZE_LONG if "tower" in self.dataset_folder else MAXWINSIZE   # data_sampler.lua:29-33
        self.winsize = args["winsize"]; self.num_past = args["num_past"]; self.num_future = args["num_future"]
        self.relative = args.get("relative", True); self.shuffle = args.get("shuffle", True)
        self.subdivide = args.get("subdivide", True); self.sim = args.get("sim", False)
        self.batch_size = args.get("batch_size", 50)
        self.rng = args.get("rng") or np.random.default_rng(0)
        assert self.num_past + self.num_future <= self.winsize
        assert self.winsize < args["maxwinsize"]
        if self.subdivide:
            assert self.shuffle
        self.savefolder = os.path.join(self.dataset_folder, "batches", self.dataset_name)
        self.num_batches = _count_batch_files(self.savefolder)
        self.num_subbatches_per_batch = self.maxwinsize // self.winsize
        self.num_subbatches = self.num_batches * self.num_subbatches_per_batch
        self.total_batches = self.num_subbatches if self.subdivide else self.num_batches
        self.batch_idxs = (self.rng.permutation(self.total_batches) + 1 if self.shuffle
                           else np.arange(1, self.total_batches + 1))
        # the loss table of priority sampling lives on the device when asked for (train loader of a -rs-off run)
        self.priority_sampler = PrioritySampler(self.total_batches, self.rng,
                                                args.get("model") if args.get("device_priority") else None)
        self.current_sampled_id = 0
        self.current_batch = 0
        self.current_subbatch = 0
        self.current_dataset = 1
        self.has_reported = False
        # ---- the one-time read of every batch file into the device store ----
        self.store = SceneStore.of(args["model"])
        self.file_base = []
        for i in range(1, self.num_batches + 1):
            focus, context = t7.load_batch_file(os.path.join(self.savefolder, f"batch{i}"))
            if focus.shape[0] != self.batch_size:
                raise ValueError(f"{self.savefolder}/batch{i}: {focus.shape[0]} examples, batch_size is {self.batch_size}")
            assert focus.shape[1] >= self.winsize and context.shape[2] >= self.winsize       # data_sampler.lua:86
            self.file_base.append(self.store.add(np.concatenate([focus[:, None], context], 1)))

    @classmethod
    def create(cls, dataset_name, args):
        return cls(dataset_name, args)

    # ---- sampling (data_sampler.lua:111-151) ----
    def sample_random_batch(self):
        self.current_batch = int(self.rng.integers(1, self.total_batches + 1))
        return self.load_batch_id(int(self.batch_idxs[self.current_batch - 1]))

    def sample_priority_batch(self, pow_=1):
        if self.priority_sampler.table_is_full:
            batch = self.load_batch_id(self.priority_sampler.sample(pow_))
        else:
            batch = self.sample_sequential_batch()
        if self.priority_sampler.table_is_full and not self.has_reported:
            print(self.dataset_folder + " has seen all batches")
            self.has_reported = True
        return batch

    def sample_sequential_batch(self):
        self.current_batch = (self.current_batch % self.total_batches) + 1
        return self.load_batch_id(int(self.batch_idxs[self.current_batch - 1]))

    def load_batch_id(self, id_):
        return self.load_subbatch_id(id_) if self.subdivide else self.load_batch_id_first_offset(id_)

    def load_batch_id_first_offset(self, id_):
        return self.load_subbatch_id_any_offset(id_, 1)

    def subbatch_location(self, id_):
        """sub-batch id -> (file id, offset), both 1-based (data_sampler.lua:199-201)."""
        batch_id = (id_ - 1) // self.num_subbatches_per_batch + 1
        offset = (id_ - 1) - self.num_subbatches_per_batch * (batch_id - 1) + 1
        return batch_id, offset

    def load_subbatch_id(self, id_):
        self.current_sampled_id = id_
        return self.load_subbatch_id_any_offset(*self.subbatch_location(id_))

    def load_subbatch_id_any_offset(self, id_, offset):
        assert (offset - 1) + self.num_past + self.num_future <= self.maxwinsize                  # data_sampler.lua:87
        base = self.file_base[id_ - 1]
        scene = np.arange(base, base + self.batch_size, dtype=np.int32)
        return DeviceBatch(self.store, scene, np.full(self.batch_size, offset - 1, np.int32), self.num_past,
                           self.num_future, self.relative, self.sim)

    def file_scenes(self, id_):
        """(batch_size, N, T, 22) host copy of batch file `id_` (1-based): the rollout's input (eval.lua:263-283)."""
        base = self.file_base[id_ - 1]
        return np.stack([self.store.host_example(s) for s in range(base, base + self.batch_size)])

    def get_hardest_batch(self):
        return self.priority_sampler.get_hardest_batch()

    def update_batch_weight(self, weight):
        self.priority_sampler.update_batch_weight(self.current_sampled_id, weight)


class GeneralDataSampler:
    """general_data_sampler.lua: one DataSampler per dataset folder.  args as DataSampler plus data_root,
    dataset_folders (list)."""

    def __init__(self, dataset_name, args):
        self.args = dict(args)
        self.dataset_name = dataset_name
        self.current_batch = 0
        self.data_root = args["data_root"]
        self.dataset_folders = list(args["dataset_folders"])
        self.maxwinsize = args["maxwinsize"]; self.winsize = args["winsize"]
        self.num_past = args["num_past"]; self.num_future = args["num_future"]
        self.relative = args.get("relative", True); self.sim = args.get("sim", False)
        self.rng = self.args.setdefault("rng", np.random.default_rng(0))
        assert self.num_past + self.num_future <= self.winsize
        assert self.winsize < self.maxwinsize
        self._make_samplers()

    @classmethod
    def create(cls, dataset_name, args):
        return cls(dataset_name, args)

    def _make_samplers(self):
        self.datasamplers = []
        for folder in self.dataset_folders:
            a = dict(self.args); a["dataset_folder"] = os.path.join(self.data_root, folder)
            self.datasamplers.append(DataSampler(self.dataset_name, a))
        self.num_batches = sum(d.total_batches for d in self.datasamplers)
        self.total_batches = self.num_batches
        print(f"{self.dataset_name}: num_batches: {self.num_batches}")
        self.current_sampled_id = None
        self.current_dataset = 1
        self.has_seen_all_batches = False
        self.has_reported = False

    def reset(self):
        self._make_samplers()
        self.current_batch = 0

    def _note_seen(self):
        if all(d.has_reported for d in self.datasamplers) and not self.has_reported:
            self.has_seen_all_batches = True
            self.has_reported = True
            print("Seen all batches")

    def sample_priority_batch(self, pow_=1):
        self.current_dataset = int(self.rng.integers(1, len(self.datasamplers) + 1))
        d = self.datasamplers[self.current_dataset - 1]
        batch = d.sample_priority_batch(pow_)
        self.current_sampled_id = d.current_sampled_id
        self._note_seen()
        return batch, self.current_dataset

    def sample_random_batch(self):
        self.current_dataset = int(self.rng.integers(1, len(self.datasamplers) + 1))
        d = self.datasamplers[self.current_dataset - 1]
        batch = d.sample_random_batch()
        self.current_sampled_id = d.current_sampled_id
        self._note_seen()
        return batch, self.current_dataset

    def get_hardest_batch(self):
        h = self.datasamplers[self.current_dataset - 1].get_hardest_batch()
        return [h[0], h[1], self.current_dataset]

    def update_batch_weight(self, weight):
        d = self.datasamplers[self.current_dataset - 1]
        assert self.current_sampled_id == d.current_sampled_id
        d.update_batch_weight(weight)

    def sample_sequential_batch(self, modulo=False):
        n = len(self.datasamplers)
        if modulo:
            self.current_dataset = self.current_dataset % n + 1
        else:
            d = self.datasamplers[self.current_dataset - 1]
            if d.current_batch == d.total_batches:
                self.current_dataset = self.current_dataset % n + 1
        d = self.datasamplers[self.current_dataset - 1]
        batch = d.sample_sequential_batch()
        self.current_sampled_id = d.current_sampled_id
        return batch, self.current_dataset
