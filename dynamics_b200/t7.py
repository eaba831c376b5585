"""Torch7 `.t7` binary serialisation (torch.save / torch.load), enough of it for the hot path's files:
batch files `batch<i>` = {FloatTensor(50,60,22), FloatTensor(50,N-1,60,22)} (reference data_process.lua:384-396,
data_sampler.lua:156-157) and checkpoints `epoch*_step*.t7` = {model={theta={params=FloatTensor(28222)},...}, mp=...}
(reference main.lua:284-295).  Format (upstream torch7 File:writeObject, little-endian): int type {0 nil, 1 number
(double), 2 string, 3 table, 4 torch object, 5 boolean}; tables and torch objects carry a file-global object index so
that shared storages are written once and referenced afterwards -- views into the flat parameter storage keep their
storageOffset, which is how the flat layout of a reference checkpoint is recovered.
Not verified against a real Torch7 file in this environment (none exists); see SURVEY.md appendix B.
"""
import struct

import numpy as np

TYPE_NIL, TYPE_NUMBER, TYPE_STRING, TYPE_TABLE, TYPE_TORCH, TYPE_BOOLEAN = 0, 1, 2, 3, 4, 5
_DTYPES = {"Float": np.float32, "Double": np.float64, "Long": np.int64, "Int": np.int32, "Byte": np.uint8,
           "Char": np.int8, "Short": np.int16}


class TorchObject(dict):
    """A torch class instance without a custom reader (nn.* modules ...): its field table plus the class name."""

    def __init__(self, typename, fields):
        super().__init__(fields or {})
        self.typename = typename


class ObjKey:
    """Wrapper for table keys that are themselves tables / torch objects (nngraph's mapindex maps node data -> index):
    hashed by identity, like a Lua table key."""

    def __init__(self, obj):
        self.obj = obj

    def __hash__(self):
        return id(self.obj)

    def __eq__(self, other):
        return isinstance(other, ObjKey) and other.obj is self.obj


class Reader:
    def __init__(self, f):
        self.f = f
        self.memo = {}

    def _i(self):
        return struct.unpack("<i", self.f.read(4))[0]

    def _l(self):
        return struct.unpack("<q", self.f.read(8))[0]

    def _s(self):
        n = self._i()
        return self.f.read(n).decode("latin1")

    def read(self):
        t = self._i()
        if t == TYPE_NIL:
            return None
        if t == TYPE_NUMBER:
            v = struct.unpack("<d", self.f.read(8))[0]
            return int(v) if v == int(v) and abs(v) < 2 ** 53 else v
        if t == TYPE_STRING:
            return self._s()
        if t == TYPE_BOOLEAN:
            return self._i() == 1
        if t == TYPE_TABLE:
            idx = self._i()
            if idx in self.memo:
                return self.memo[idx]
            out = {}
            self.memo[idx] = out
            for _ in range(self._i()):
                k = self.read()
                if isinstance(k, (dict, np.ndarray)):
                    k = ObjKey(k)
                out[k] = self.read()
            return out
        if t == TYPE_TORCH:
            idx = self._i()
            if idx in self.memo:
                return self.memo[idx]
            version = self._s()
            cls = self._s() if version.startswith("V ") else version
            obj = self._read_torch(cls, idx)
            self.memo[idx] = obj
            return obj
        raise ValueError(f"unsupported .t7 type tag {t}")

    def _read_torch(self, cls, idx):
        if cls.startswith("torch.") and cls.endswith("Storage"):
            dt = _DTYPES[cls[len("torch."):-len("Storage")]]
            n = self._l()
            return np.frombuffer(self.f.read(n * np.dtype(dt).itemsize), dtype=dt).copy()
        if cls.startswith("torch.") and cls.endswith("Tensor"):
            nd = self._i()
            size = [self._l() for _ in range(nd)]
            stride = [self._l() for _ in range(nd)]
            off = self._l() - 1
            storage = self.read()
            if storage is None or nd == 0:
                return np.zeros(0, _DTYPES.get(cls[len("torch."):-len("Tensor")], np.float32))
            view = np.lib.stride_tricks.as_strided(storage[off:], shape=size,
                                                   strides=[s * storage.itemsize for s in stride])
            view = view.view(TensorView)
            view.storage_offset = off
            view.storage = storage
            return view
        fields = self.read()
        return TorchObject(cls, fields if isinstance(fields, dict) else {"_": fields})


class TensorView(np.ndarray):
    """ndarray view that remembers its storage and storageOffset (flat-layout recovery)."""
    storage_offset = 0
    storage = None


def load(path):
    with open(path, "rb") as f:
        return Reader(f).read()


class Storage:
    """A torch.FloatStorage written once and referenced by every tensor view into it (shared parameters)."""

    def __init__(self, data):
        self.data = np.ascontiguousarray(data, np.float32).ravel()


class View:
    """A FloatTensor view (shape, row-major) at a 0-based storage offset of a shared Storage."""

    def __init__(self, storage, offset, shape):
        self.storage, self.offset, self.shape = storage, int(offset), tuple(int(x) for x in shape)


class Writer:
    def __init__(self, f):
        self.f = f
        self.next = 1
        self.memo = {}      # id(object) -> file-global index: tables / torch objects / storages are written once
        self.keep = []      # keeps memoised objects alive so ids stay unique

    def _ref(self, obj):
        """Writes the object index; True if the object was already written (a back-reference is all that follows)."""
        if id(obj) in self.memo:
            self._i(self.memo[id(obj)])
            return True
        self.memo[id(obj)] = self.next
        self.keep.append(obj)
        self._i(self.next); self.next += 1
        return False

    def _i(self, v):
        self.f.write(struct.pack("<i", v))

    def _l(self, v):
        self.f.write(struct.pack("<q", v))

    def _s(self, s):
        b = s.encode("latin1")
        self._i(len(b)); self.f.write(b)

    def write(self, obj):
        if obj is None:
            self._i(TYPE_NIL)
        elif isinstance(obj, bool):
            self._i(TYPE_BOOLEAN); self._i(1 if obj else 0)
        elif isinstance(obj, (int, float)):
            self._i(TYPE_NUMBER); self.f.write(struct.pack("<d", float(obj)))
        elif isinstance(obj, str):
            self._i(TYPE_STRING); self._s(obj)
        elif isinstance(obj, ObjKey):
            self.write(obj.obj)
        elif isinstance(obj, Storage):
            self._i(TYPE_TORCH)
            if not self._ref(obj):
                self._s("V 1"); self._s("torch.FloatStorage")
                self._l(obj.data.size); self.f.write(obj.data.tobytes())
        elif isinstance(obj, View):
            self._i(TYPE_TORCH)
            if not self._ref(obj):
                self._s("V 1"); self._s("torch.FloatTensor")
                self._i(len(obj.shape))
                for d in obj.shape:
                    self._l(d)
                stride = 1
                strides = []
                for d in reversed(obj.shape):
                    strides.append(stride); stride *= d
                for st in reversed(strides):
                    self._l(st)
                self._l(obj.offset + 1)
                self.write(obj.storage)
        elif isinstance(obj, TorchObject):
            self._i(TYPE_TORCH)
            if not self._ref(obj):
                self._s("V 1"); self._s(obj.typename)
                self.write(dict(obj))
        elif isinstance(obj, np.ndarray):
            name = {np.dtype(np.float32): "Float", np.dtype(np.float64): "Double", np.dtype(np.int64): "Long",
                    np.dtype(np.int32): "Int", np.dtype(np.uint8): "Byte"}[obj.dtype]
            a = np.ascontiguousarray(obj)
            self._i(TYPE_TORCH); self._i(self.next); self.next += 1
            self._s("V 1"); self._s(f"torch.{name}Tensor")
            self._i(a.ndim)
            for s in a.shape:
                self._l(s)
            for s in a.strides:
                self._l(s // a.itemsize)
            self._l(1)
            self._i(TYPE_TORCH); self._i(self.next); self.next += 1
            self._s("V 1"); self._s(f"torch.{name}Storage")
            self._l(a.size); self.f.write(a.tobytes())
        elif isinstance(obj, (dict, list, tuple)):
            items = obj.items() if isinstance(obj, dict) else enumerate(obj, 1)
            items = list(items)
            self._i(TYPE_TABLE)
            if self._ref(obj):
                return
            self._i(len(items))
            for k, v in items:
                self.write(k); self.write(v)
        else:
            raise TypeError(f"cannot serialise {type(obj)}")


def save(path, obj):
    with open(path, "wb") as f:
        Writer(f).write(obj)


def save_batch_file(path, focus, context):
    """batches/<set>/batch<i> (data_process.lua:384-396): {focus (B,T,22), context (B,N-1,T,22)}."""
    save(path, [np.asarray(focus, np.float32), np.asarray(context, np.float32)])


def load_batch_file(path):
    t = load(path)
    return np.asarray(t[1]), np.asarray(t[2])


def save_checkpoint(path, params, mp=None, iters=0, train_losses=(), val_losses=(), test_losses=()):
    """Checkpoint table of main.lua:288-295 / 338-345: {model, mp, train_losses, val_losses, test_losses, iters}.
    `model.network` is an nn.NPECuda object (class of dynamics_b200/lua/NPECuda.lua; its `write` stores {weight, mp}),
    `model.criterion` / `model.identitycriterion` are the criterion objects model.create clones on preload
    (npe.lua:82-84), and `model.theta.params` is a view of the SAME storage as the network's weight, as in a reference
    checkpoint -- so `checkpoint.model.network:clone()` followed by `:getParameters()` (npe.lua:81-98) and
    `checkpoint.model.theta.params` (eval.lua:523) both yield the flat 28 222-float vector."""
    params = np.asarray(params, np.float32).ravel()
    st = Storage(params)
    weight = View(st, 0, (params.size,))
    mp = dict(mp or {})
    model = {"network": TorchObject("nn.NPECuda", {"weight": weight, "mp": mp}),
             "criterion": TorchObject("nn.MSECriterion", {"sizeAverage": False, "gradInput": np.zeros(0, np.float32),
                                                          "output": 0}),
             "identitycriterion": TorchObject("nn.IdentityCriterion", {"gradInput": np.zeros(0, np.float32), "output": 0}),
             "theta": {"params": weight},
             "mp": mp}
    save(path, {"model": model, "mp": mp, "train_losses": list(train_losses), "val_losses": list(val_losses),
                "test_losses": list(test_losses), "iters": iters})


# canonical flat layout (include/npe_cuda.h): name -> (offset, (out, in) or (n,))
def _canonical_layout():
    ents, off = [], 0
    for name, shape in ([("enc.this.W", (25, 44)), ("enc.ctx.W", (25, 44))] + [(f"pair.L{l}.W", (50, 50)) for l in range(1, 6)] +
                        [("dec.L1.W", (50, 94)), ("dec.L1.b", (50,))] +
                        [x for i in range(2, 5) for x in ((f"dec.L{i}.W", (50, 50)), (f"dec.L{i}.b", (50,)))] +
                        [("dec.L5.W", (22, 50)), ("dec.L5.b", (22,))]):
        n = int(np.prod(shape))
        ents.append((name, off, shape)); off += n
    return ents, off


def _typename(o):
    return getattr(o, "typename", None)


def _linears_in_order(seq):
    """nn.Linear children of an nn.Sequential, in `modules` order."""
    mods = seq.get("modules", {})
    return [mods[i] for i in sorted(k for k in mods if isinstance(k, int)) if _typename(mods[i]) == "nn.Linear"]


def _find(o, pred, seen=None):
    """Depth-first search of the object graph (tables and torch objects) for the first object satisfying pred."""
    seen = set() if seen is None else seen
    if id(o) in seen or not isinstance(o, dict):
        return None
    seen.add(id(o))
    if pred(o):
        return o
    for k in sorted(o.keys(), key=lambda k: (not isinstance(k, int), str(k) if not isinstance(k, ObjKey) else "~")):
        r = _find(o[k], pred, seen)
        if r is not None:
            return r
    return None


def _encoder_linears(enc):
    """(this, context) Linears of init_object_encoder's gModule (modules.lua:10-26): the JoinTable node joins
    {thisp_out, contextp_out} in that order, so its data.mapindex[1] / [2] lead (through a ReLU node) to the Linear that
    encodes the focus / the context object -- independent of nngraph's topological module order."""
    def is_join_node(o):
        d = o.get("data") if isinstance(o, dict) else None
        return isinstance(d, dict) and _typename(d.get("module")) == "nn.JoinTable" and isinstance(d.get("mapindex"), dict)
    node = _find(enc, is_join_node)
    if node is not None:
        out = []
        for i in (1, 2):
            d = node["data"]["mapindex"].get(i)
            for _ in range(4):                      # ReLU node data -> Linear node data
                if d is None or _typename(d.get("module")) == "nn.Linear":
                    break
                d = d.get("mapindex", {}).get(1)
            if d is None or _typename(d.get("module")) != "nn.Linear":
                out = None
                break
            out.append(d["module"])
        if out:
            return out
    mods = enc.get("modules", {})                    # fallback: gModule.modules order (this before context)
    lin = [mods[i] for i in sorted(k for k in mods if isinstance(k, int)) if _typename(mods[i]) == "nn.Linear"]
    if len(lin) != 2:
        raise ValueError("encoder gModule with two nn.Linear not found")
    return lin


def params_from_network(network):
    """Flat parameter vector in the canonical layout from a reference `model.network` object graph (init_network,
    npe.lua:14-61): every nn.Linear's weight / bias is a VIEW into the storage getParameters() flattened; its
    storageOffset says where, so the layout is read from the file instead of assumed (nngraph's module order for the
    two encoder Linears is not fixed by the reference)."""
    is_enc = lambda o: _typename(o) == "nn.gModule" and sum(_typename(m) == "nn.Linear" for m in o.get("modules", {}).values()) == 2
    def is_step(o):
        if _typename(o) != "nn.Sequential":
            return False
        mods = o.get("modules", {})
        return any(is_enc(m) for m in mods.values() if isinstance(m, dict)) and len(_linears_in_order(o)) == 5
    step = _find(network, is_step)
    if step is None:
        raise ValueError("pairwise step (encoder + 5 Linear) not found in model.network")
    enc = next(m for m in step["modules"].values() if isinstance(m, dict) and is_enc(m))
    lin_this, lin_ctx = _encoder_linears(enc)
    pair = _linears_in_order(step)
    is_dec = lambda o: _typename(o) == "nn.Sequential" and len(_linears_in_order(o)) == 5 and o is not step
    mods = network.get("modules", {})
    dec = _find(mods.get(2, network), is_dec)
    if dec is None:
        raise ValueError("decoder (5 Linear) not found in model.network")
    linears = [lin_this, lin_ctx] + pair + _linears_in_order(dec)
    ents, n = _canonical_layout()
    flat = np.zeros(n, np.float32)
    it = iter(ents)
    for li, lin in enumerate(linears):
        name, off, shape = next(it)
        w = np.asarray(lin["weight"], np.float32)
        if w.shape != tuple(shape):
            raise ValueError(f"{name}: weight shape {w.shape} != {shape}")
        flat[off:off + w.size] = w.ravel()
        if li >= 7:                                   # decoder Linears carry a bias; with -nbrhd the others do not (npe.lua:17)
            name, off, shape = next(it)
            b = np.asarray(lin["bias"], np.float32)
            flat[off:off + b.size] = b.ravel()
        elif lin.get("bias") is not None and np.asarray(lin["bias"]).size:
            raise ValueError("checkpoint was trained without -nbrhd (biased pairwise layers): not the north-star path")
    return flat


def load_checkpoint_params(path):
    """Flat parameter vector (canonical layout) of a checkpoint.
    Ours: model.network is nn.NPECuda {weight}.  The reference's: model.network is the nn object graph -- the layout is
    derived from each nn.Linear view's storageOffset (params_from_network); model.theta.params is only the fallback
    (same storage, getParameters order)."""
    ck = load(path)
    model = ck["model"]
    net = model.get("network") if isinstance(model, dict) else None
    if _typename(net) == "nn.NPECuda":
        return np.asarray(net["weight"], np.float32).ravel().copy()
    if isinstance(net, dict):
        try:
            return params_from_network(net)
        except ValueError:
            pass
    theta = model.get("theta") if isinstance(model, dict) else None
    if theta and theta.get("params") is not None:
        return np.asarray(theta["params"], np.float32).ravel()
    found = []

    def walk(o, seen=set()):
        if id(o) in seen:
            return
        seen.add(id(o))
        if isinstance(o, TensorView) and o.storage is not None and o.storage.size == 28222:
            found.append(o.storage)
        elif isinstance(o, dict):
            for v in o.values():
                walk(v)
    walk(model)
    if not found:
        raise ValueError("no flat 28222-float parameter storage found in checkpoint")
    return np.asarray(found[0], np.float32)
