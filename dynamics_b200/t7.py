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
                k = self.read(); out[k] = self.read()
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


class Writer:
    def __init__(self, f):
        self.f = f
        self.next = 1

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
            self._i(TYPE_TABLE); self._i(self.next); self.next += 1
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


def save_checkpoint(path, params, mp=None, iters=0):
    """The subset of main.lua:288-295 the rollout needs: checkpoint.model.theta.params (flat) + mp."""
    save(path, {"model": {"theta": {"params": np.asarray(params, np.float32)}}, "mp": dict(mp or {}), "iters": iters})


def load_checkpoint_params(path):
    """Flat parameter vector of a checkpoint: model.theta.params when present (both ours and the reference's), else
    the storage shared by model.network's Linear weights (recovered through storageOffset)."""
    ck = load(path)
    model = ck["model"]
    theta = model.get("theta") if isinstance(model, dict) else None
    if theta and theta.get("params") is not None:
        return np.asarray(theta["params"], np.float32).ravel()
    found = []

    def walk(o):
        if isinstance(o, TensorView) and o.storage is not None and o.storage.size == 28222:
            found.append(o.storage)
        elif isinstance(o, dict):
            for v in o.values():
                walk(v)
    walk(model)
    if not found:
        raise ValueError("no flat 28222-float parameter storage found in checkpoint")
    return np.asarray(found[0], np.float32)
