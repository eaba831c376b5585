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


def _reference_style_network(flat_ref, order):
    """Object graph shaped like init_network's result (npe.lua:14-61) as torch.save would write it: every nn.Linear holds
    VIEWS into one FloatStorage; `order` lists the Linears in the order getParameters() flattened them."""
    T = t7.TorchObject
    st = t7.Storage(flat_ref)
    shapes = {"this": (25, 44), "ctx": (25, 44), **{f"p{l}": (50, 50) for l in range(1, 6)},
              "d1": (50, 94), "d2": (50, 50), "d3": (50, 50), "d4": (50, 50), "d5": (22, 50)}
    lin, off = {}, 0
    for name in order:
        o, i = shapes[name]
        f = {"weight": t7.View(st, off, (o, i)), "gradWeight": np.zeros(0, np.float32)}
        off += o * i
        if name.startswith("d"):
            f["bias"] = t7.View(st, off, (o,)); off += o
        lin[name] = T("nn.Linear", f)
    assert off == 28222

    def node(module, parents=()):                       # nngraph.Node: data.mapindex maps index -> parent data and back
        d = {"module": module, "mapindex": {}}
        for i, p in enumerate(parents, 1):
            d["mapindex"][i] = p["data"]
            d["mapindex"][t7.ObjKey(p["data"])] = i
        return {"data": d, "children": {}}
    n_this, n_ctx = node(T("nn.Identity", {})), node(T("nn.Identity", {}))
    n_lt, n_lc = node(lin["this"], [n_this]), node(lin["ctx"], [n_ctx])
    n_rt, n_rc = node(T("nn.ReLU", {}), [n_lt]), node(T("nn.ReLU", {}), [n_lc])
    n_join = node(T("nn.JoinTable", {"dimension": 2}), [n_rt, n_rc])
    # gModule.modules in a topological order that puts the CONTEXT Linear first (the order is not fixed by the reference)
    encoder = T("nn.gModule", {"modules": {1: n_ctx["data"]["module"], 2: lin["ctx"], 3: n_this["data"]["module"], 4: lin["this"],
                                           5: n_rc["data"]["module"], 6: n_rt["data"]["module"], 7: n_join["data"]["module"]},
                               "forwardnodes": {1: n_this, 2: n_ctx, 3: n_lt, 4: n_lc, 5: n_rt, 6: n_rc, 7: n_join},
                               "outnode": n_join})
    step_mods = {1: encoder}
    for l in range(1, 6):
        step_mods[2 * l] = lin[f"p{l}"]; step_mods[2 * l + 1] = T("nn.ReLU", {})
    step = T("nn.Sequential", {"modules": step_mods})
    recursor = T("nn.Recursor", {"recurrentModule": step, "modules": {1: step}, "sharedClones": {1: step}})
    sequencer = T("nn.Sequencer", {"module": recursor, "modules": {1: recursor}})
    pairwise = T("nn.Sequential", {"modules": {1: sequencer, 2: T("nn.CAddTable", {})}})
    branches = T("nn.ParallelTable", {"modules": {1: pairwise, 2: T("nn.Identity", {})}})
    dec_mods = {}
    for i in range(1, 6):
        dec_mods[2 * i - 1] = lin[f"d{i}"]
        if i < 5:
            dec_mods[2 * i] = T("nn.ReLU", {})
    dec_net = T("nn.Sequential", {"modules": dec_mods})
    n_dec = node(dec_net)
    decoder = T("nn.gModule", {"modules": {1: T("nn.JoinTable", {}), 2: dec_net, 3: T("nn.Sigmoid", {})},
                               "forwardnodes": {1: n_dec}})
    net = T("nn.Sequential", {"modules": {1: branches, 2: decoder}})
    return net, st, lin


def test_reference_checkpoint_layout_is_derived_from_storage_offsets(tmp_path):
    """A reference checkpoint (main.lua:288-295) whose getParameters() order has the CONTEXT encoder first: the loader must
    walk model.network, identify each nn.Linear structurally (JoinTable input 1 = focus encoder) and place its weights by
    storageOffset into the canonical layout -- not take the shared storage as is."""
    from dynamics_b200.model import param_offsets
    rng = np.random.default_rng(1)
    canon = rng.standard_normal(28222).astype(np.float32)
    off = param_offsets()
    order = ["ctx", "this", "p1", "p2", "p3", "p4", "p5", "d1", "d2", "d3", "d4", "d5"]        # reference flattening
    names = {"this": "enc.this.W", "ctx": "enc.ctx.W", **{f"p{l}": f"pair.L{l}.W" for l in range(1, 6)}}
    flat_ref = []
    for nm in order:
        if nm.startswith("d"):
            i = nm[1]
            for key in (f"dec.L{i}.W", f"dec.L{i}.b"):
                o, shape = off[key]; flat_ref.append(canon[o:o + int(np.prod(shape))])
        else:
            o, shape = off[names[nm]]; flat_ref.append(canon[o:o + int(np.prod(shape))])
    flat_ref = np.concatenate(flat_ref)
    assert not np.array_equal(flat_ref, canon)
    net, st, lin = _reference_style_network(flat_ref, order)
    theta = {"params": t7.View(st, 0, (28222,)), "grad_params": np.zeros(28222, np.float32)}
    p = tmp_path / "epoch3_step100_0.0123456.t7"
    t7.save(p, {"model": {"network": net, "criterion": t7.TorchObject("nn.MSECriterion", {"sizeAverage": False}),
                          "identitycriterion": t7.TorchObject("nn.IdentityCriterion", {}), "theta": theta,
                          "mp": {"nbrhd": True}},
                "mp": {"nbrhd": True, "layers": 5}, "iters": 100})
    got = t7.load_checkpoint_params(p)
    assert np.array_equal(got, canon)
    ck = t7.load(p)                                    # shared storage survived: theta.params views the Linears' storage
    assert np.array_equal(np.asarray(ck["model"]["theta"]["params"]), flat_ref)


def test_our_checkpoint_has_the_reference_table_shape(tmp_path):
    """What train.py writes: the main.lua:288-295 table; model.network is an nn.NPECuda object whose weight shares its
    storage with model.theta.params, plus the criterion objects npe.lua:82-84 clones on preload."""
    params = np.linspace(-1, 1, 28222).astype(np.float32)
    p = tmp_path / "epoch1_step0_0.5000000.t7"
    t7.save_checkpoint(p, params, mp={"layers": 5, "nbrhd": True}, iters=7, train_losses=[0.5], val_losses=[0.4], test_losses=[0.3])
    ck = t7.load(p)
    m = ck["model"]
    assert m["network"].typename == "nn.NPECuda" and m["criterion"].typename == "nn.MSECriterion"
    assert m["identitycriterion"].typename == "nn.IdentityCriterion"
    assert m["network"]["weight"].storage is m["theta"]["params"].storage          # one storage, written once
    assert np.array_equal(np.asarray(m["theta"]["params"]), params)
    assert ck["iters"] == 7 and ck["val_losses"][1] == 0.4 and ck["mp"]["layers"] == 5
    assert np.array_equal(t7.load_checkpoint_params(p), params)
    assert open(p, "rb").read().count(b"torch.FloatStorage") == 3                   # params + the two empty gradInputs
