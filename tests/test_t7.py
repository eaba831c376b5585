"""CPU tests of the .t7 reader/writer (checkpoint + batch-file interop)."""
import io
import struct

import numpy as np

from dynamics_b200 import t7


def test_batch_file_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    foc = rng.standard_normal((50, 60, 22)).astype(np.float32)
    ctx = rng.standard_normal((50, 3, 60, 22)).astype(np.float32)
    p = tmp_path / "batch1"
    t7.save_batch_file(p, foc, ctx)
    f2, c2 = t7.load_batch_file(p)
    assert np.array_equal(f2, foc) and np.array_equal(c2, ctx)
    raw = open(p, "rb").read()
    assert struct.unpack("<i", raw[:4])[0] == t7.TYPE_TABLE and b"torch.FloatTensor" in raw[:80]


def test_checkpoint_roundtrip_and_shared_storage(tmp_path):
    params = np.arange(28222, dtype=np.float32)
    p = tmp_path / "epoch1_step0_0.0100000.t7"
    t7.save_checkpoint(p, params, mp={"rnn_dim": 50, "layers": 5, "nbrhd": True, "name": "x"})
    assert np.array_equal(t7.load_checkpoint_params(p), params)
    ck = t7.load(p)
    assert ck["mp"]["rnn_dim"] == 50 and ck["mp"]["nbrhd"] is True and ck["mp"]["name"] == "x"
    # a reference-style checkpoint: no theta, Linear weights are VIEWS into one shared storage (back-reference)
    buf = io.BytesIO()
    w = t7.Writer(buf)
    w._i(t7.TYPE_TABLE); w._i(w.next); w.next += 1; w._i(1)
    w.write("model")
    w._i(t7.TYPE_TABLE); w._i(w.next); w.next += 1; w._i(2)
    storage_idx = None
    for key, (off, shape) in (("a", (0, (25, 44))), ("b", (1100, (25, 44)))):
        w.write(key)
        w._i(t7.TYPE_TORCH); w._i(w.next); w.next += 1; w._s("V 1"); w._s("torch.FloatTensor")
        w._i(2); w._l(shape[0]); w._l(shape[1]); w._l(shape[1]); w._l(1); w._l(off + 1)
        if storage_idx is None:
            storage_idx = w.next
            w._i(t7.TYPE_TORCH); w._i(w.next); w.next += 1; w._s("V 1"); w._s("torch.FloatStorage")
            w._l(28222); buf.write(params.tobytes())
        else:
            w._i(t7.TYPE_TORCH); w._i(storage_idx)
    q = tmp_path / "ref.t7"
    open(q, "wb").write(buf.getvalue())
    ck = t7.load(q)
    a, b = ck["model"]["a"], ck["model"]["b"]
    assert a.storage is b.storage and b.storage_offset == 1100 and b[0, 0] == 1100.0 and a[1, 0] == 44.0
    assert np.array_equal(t7.load_checkpoint_params(q), params)
