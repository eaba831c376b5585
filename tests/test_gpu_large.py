"""-m gpu parity tests of the THROUGHPUT regime (more than 592 foci): every kernel family that the large-batch training
step, the large forward and the rollout-by-kernels dispatch to is compared with the CPU oracle, at sizes where the
persistent tile loops of each kernel wrap (more tiles than resident CTAs / tile groups on a 148-SM part).

  configs[1]/[4]-style data: 5120 eight-ball scenes -> 40 960 foci, ~80 k active pairs
      64-row FFMA tiles: 640 focus tiles, > 1 200 pair tiles (148 CTAs)
      128-row tcgen05 tiles: 320 focus tiles, > 600 pair tiles (296 tile groups)
  configs[3]-style data: 64-ball scenes, 12-NN trim off (63 context slots, 32-lane pair builder), 4 096 foci

Dispatches of npe_train_step / npe_forward + npe_backward (include/npe_cuda.h NPE_OPT_*):
  default      tcgen05 3xTF32 forward + activation stash + tcgen05 backward
  ffma_bwd     tcgen05 3xTF32 forward + activation stash + FFMA <64> backward kernels
  ffma         FFMA <64> forward + stash + FFMA <64> backward
  recompute    FFMA <64> forward, no stash: the backward recomputes the pair chain
Tolerances: masks bit-exact, pred / loss 1e-4 relative, gradient 1e-4 of ||g||_inf, parameters after one
Adam step 1e-4 of ||x||_inf; single-pass TF32 forward 2e-3.  The achieved errors are printed (pytest -s shows them).
"""
import numpy as np
import pytest

from oracle import batcher, npe, optim, rollout, scenes
from tests.util import grad_err, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4
TOL_TF32 = 2e-3
DISPATCH = {
    "default": dict(train_fwd_tc=1, train_stash=1, bwd_tc=1),
    "ffma_bwd": dict(train_fwd_tc=1, train_stash=1, bwd_tc=0),
    "ffma": dict(train_fwd_tc=0, train_stash=1, bwd_tc=0),
    "recompute": dict(train_fwd_tc=0, train_stash=0, bwd_tc=0),
}


def scene_examples(st, t0):
    """(this (F,2,22), ctx (F,N-1,2,22), y (F,1,22)) of every object of every scene at window t0, scene-major /
    object-minor: the order of npe_select_windows."""
    S, N = st.shape[:2]
    others = np.array([[j for j in range(N) if j != o] for o in range(N)])
    this = st[:, :, t0:t0 + 2].reshape(S * N, 2, 22)
    ctx = st[:, others][:, :, :, t0:t0 + 2].reshape(S * N, N - 1, 2, 22)
    y = batcher.relative_pair(this, st[:, :, t0 + 2:t0 + 3].reshape(S * N, 1, 22))
    return this, ctx, y


def oracle_chunked(params, this, ctx, y, chunk=4096):
    """npe.fp_batch / npe.bp over chunks of foci (examples are independent; the gradient is a sum with the FULL batch
    as divisor, npe.lua:304-315)."""
    F = this.shape[0]
    preds, les, masks = [], [], []
    g = np.zeros(params.size, np.float64)
    for a in range(0, F, chunk):
        b = (this[a:a + chunk], ctx[a:a + chunk], y[a:a + chunk])
        gi, pred = npe.bp(params, b, divisor_batch=F)
        g += gi
        le, lv, lw = npe.losses_per_example(pred, b[2].reshape(len(pred), -1))
        preds.append(pred); les.append(np.stack([le, lv, lw], 1)); masks.append(npe.neighbor_mask(b[0], b[1]))
    pred = np.concatenate(preds); le = np.concatenate(les)
    lv = np.float64(2.0) * le[:, 1].astype(np.float64).sum(); lw = le[:, 2].astype(np.float64).sum()
    loss3 = np.array([(lv + lw) / (3 * F), lv / (2 * F), lw / F])
    return dict(pred=pred, le=le, loss3=loss3, grad=g.astype(np.float32), mask=np.concatenate(masks))


@pytest.fixture(scope="module")
def n8(trained_like_params):
    st = scenes.raw_to_state(scenes.gen_balls(5120, 8, 4, seed=88))
    this, ctx, y = scene_examples(st, 0)
    ref = oracle_chunked(trained_like_params, this, ctx, y)
    x = trained_like_params.copy()
    optim.adam_step(x, ref["grad"], optim.adam_init(x.size), 3e-4)
    ref["adam"] = x
    return st, ref


@pytest.fixture(scope="module")
def s64(trained_like_params):
    st = scenes.raw_to_state(scenes.gen_big(64, 64, 3, seed=66))
    this, ctx, y = scene_examples(st, 0)
    ref = oracle_chunked(trained_like_params, this, ctx, y, chunk=1024)
    x = trained_like_params.copy()
    optim.adam_step(x, ref["grad"], optim.adam_init(x.size), 3e-4)
    ref["adam"] = x
    return st, ref


def make_model(max_ctx, opts):
    from dynamics_b200 import ModelParams, NPEModel
    m = NPEModel(ModelParams(max_ctx=max_ctx))
    for k, v in opts.items():
        m.set_option(k, v)
    return m


def run_training_checks(m, st, ref, params, tag):
    S = st.shape[0]
    m.set_params(params); m.optim_reset(0.0)
    m.upload_scenes(st)
    m.upload_windows(np.arange(S), np.zeros(S, np.int32))
    F = m.select_windows(0, S)
    assert F == ref["pred"].shape[0] and F > 592
    idx, mask, cnt = m.pair_table()
    assert np.array_equal(mask[:, :ref["mask"].shape[1]], ref["mask"].astype(np.uint8))          # bit-exact
    f, p = m.batch_stats()
    assert p == int(ref["mask"].sum())
    # model:fp + model:bp as separate calls (training forward: FP32-grade kernels whatever the inference precision)
    pred, loss3 = m.forward_current()
    e_pred = rel_err(pred[:, :6], ref["pred"][:, :6]); e_loss = rel_err(loss3, ref["loss3"], 1e-6)
    g = m.bp()
    e_g = grad_err(g, ref["grad"])
    ents, _ = npe.param_layout()
    worst = max((grad_err(g[o:o + int(np.prod(s))], ref["grad"][o:o + int(np.prod(s))]), n) for n, o, s in ents)
    # run-to-run determinism of the whole path
    m.forward_current()
    assert np.array_equal(m.bp(), g), "gradient not run-to-run deterministic"
    # feval_train + optimiser in one call
    l3 = m.train_step("adam", 3e-4)
    e_l3 = rel_err(l3, ref["loss3"], 1e-6)
    e_gs = grad_err(m.get_grads(), ref["grad"])
    e_x = float(np.max(np.abs(m.get_params() - ref["adam"])) / np.max(np.abs(ref["adam"])))
    print(f"[{tag}] F={F} P={p}: pred {e_pred:.2e} loss {e_loss:.2e} grad {e_g:.2e} (worst block {worst[0]:.2e} {worst[1]}) "
          f"step: loss {e_l3:.2e} grad {e_gs:.2e} params {e_x:.2e}")
    assert e_pred < TOL and e_loss < TOL
    assert e_g < TOL and worst[0] < 5e-4, worst
    assert e_l3 < TOL and e_gs < TOL and e_x < TOL


@pytest.mark.parametrize("dispatch", list(DISPATCH))
def test_large_train_step_n8(n8, trained_like_params, dispatch):
    """40 960 foci / ~80 k active pairs: every persistent loop of the training kernels wraps at least twice."""
    st, ref = n8
    m = make_model(12, DISPATCH[dispatch])
    try:
        run_training_checks(m, st, ref, trained_like_params, f"n8/{dispatch}")
    finally:
        m.close()


@pytest.mark.parametrize("dispatch", list(DISPATCH))
def test_large_train_step_s64(s64, trained_like_params, dispatch):
    """configs[3] shape: 64-ball scenes, trim off, 4 096 foci (7x the small-batch switch)."""
    st, ref = s64
    m = make_model(0, DISPATCH[dispatch])
    try:
        run_training_checks(m, st, ref, trained_like_params, f"s64/{dispatch}")
    finally:
        m.close()


@pytest.mark.parametrize("dispatch", ["default", "ffma"])
def test_forced_large_path_on_reference_batch(model, trained_like_params, dispatch):
    """The reference's batch of 50 pushed through the large-batch kernels (NPE_OPT_SMALL_BATCH_MAX = 0): partial
    tiles, a single CTA, and the golden n4 fixture as the anchor."""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "n4_b50.npz"))
    batch = (z["this"], z["context"], z["y"])
    for k, v in DISPATCH[dispatch].items():
        model.set_option(k, v)
    model.set_option("small_batch_max", 0)
    loss, pred, lv, lw = model.fp(trained_like_params, batch)
    assert rel_err(pred[:, :6], z["pred"][:, :6]) < TOL and rel_err([loss, lv, lw], z["loss3"]) < TOL
    g = model.bp()
    g_ref, _ = npe.bp(trained_like_params, batch)
    assert grad_err(g, g_ref) < TOL
    assert np.max(np.abs(g[z["grad_idx"]] - z["grad_sel"])) <= TOL * float(z["grad_absmax"])
    model.set_params(trained_like_params); model.optim_reset(0.0)
    model._upload(batch)
    l3 = model.train_step("adam", 3e-4)
    assert rel_err(l3, z["loss3"]) < TOL and grad_err(model.get_grads(), g_ref) < TOL


@pytest.mark.parametrize("precision,tol", [("fp32", TOL), ("tf32x3", TOL), ("tf32", TOL_TF32)])
def test_large_forward_and_rollout_each_precision(n8, trained_like_params, precision, tol):
    """Inference forward on 40 960 foci and two rollout-by-kernels steps on 5 120 scenes for each arithmetic variant."""
    st, ref = n8
    S = st.shape[0]
    m = make_model(12, {})
    try:
        m.set_params(trained_like_params)
        m.upload_scenes(st)
        m.upload_windows(np.arange(S), np.zeros(S, np.int32))
        m.select_windows(0, S)
        m.set_precision(precision)
        pred, le = m.forward_per_example_current()
        scale = np.abs(ref["pred"][:, :6]).max()
        e = float(np.abs(pred[:, :6] - ref["pred"][:, :6]).max() / scale) if precision == "tf32" else rel_err(pred[:, :6], ref["pred"][:, :6])
        e_le = float(np.max(np.abs(le - ref["le"])))
        idx, mask, cnt = m.pair_table()
        assert np.array_equal(mask[:, :7], ref["mask"].astype(np.uint8))
        # rollout by kernels, teacher forced (one-step predictions at every step against the oracle's)
        m.set_rollout_mode("kernels")
        steps = 2
        traj, met = m.rollout(0, steps, use_gt=True, teacher=True)
        rtraj, rmet = rollout.rollout_scenes(trained_like_params, st[:, :, 0:2], np.full(S, 8), steps, gt=st[:, :, 2:],
                                             teacher=st[:, :, 2:])
        vs = np.abs(rtraj[..., 2:4]).max()
        e_roll = float(np.max(np.abs(traj[..., 2:4] - rtraj[..., 2:4])) / vs)
        print(f"[n8/{precision}] forward {e:.2e} per-example loss abs {e_le:.2e} rollout velocities {e_roll:.2e}")
        assert e < tol and e_roll < max(tol, 2e-4)
        assert e_le < (1e-6 if precision != "tf32" else 1e-3)
        assert np.array_equal(traj[..., 0:2], rtraj[..., 0:2])          # positions integrate the PREVIOUS velocity: exact
        assert met[:, 3].min() == S * 8
    finally:
        m.close()


def test_missing_peer_times_out_instead_of_hanging(model, params):
    """The data-parallel step waits for every peer's gradient cell with a bound: a rank whose peer never steps gets
    NPE_ERR_CUDA, not a wedged GPU.  Single process: rank 0 of 2 with its own region standing in for the absent peer."""
    import ctypes as C
    from dynamics_b200 import NPEError
    from tests.util import balls_batches
    _, batches = balls_batches(60, 4, seed=5)
    batch = batcher.make_batch(*batches[0], offset=1)
    buf = C.create_string_buffer(64)
    model._ck(model.lib.npe_peer_export(model.h, buf))
    both = C.create_string_buffer(buf.raw + buf.raw, 128)
    model._ck(model.lib.npe_peer_attach(model.h, 0, 2, both))
    model.set_option("peer_timeout_ms", 50)
    model.set_params(params); model.optim_reset(0.0)
    model._upload(batch)
    with pytest.raises(NPEError) as ei:
        model.train_step("adam", 3e-4, divisor_batch=100)
    assert "peer" in str(ei.value)


def test_tcgen05_backward_stage_by_stage(s64, trained_like_params):
    """The tcgen05 backward kernel by kernel against the oracle's intermediate gradients (npe.bp(inter=...)): decoder
    stash and input-gradient chain (dz5..dz1, ds), pair stash and chain (dz5..dz0 of every ACTIVE pair, list order =
    focus-major, context ascending), then the weight gradients.  Localises a fault to one kernel."""
    st, ref = s64
    this, ctx, y = scene_examples(st, 0)
    F = this.shape[0]
    inter = {}
    g_ref, _ = npe.bp(trained_like_params, (this, ctx, y), inter=inter)
    m = make_model(0, DISPATCH["default"])
    try:
        S = st.shape[0]
        m.set_params(trained_like_params)
        m.upload_scenes(st)
        m.upload_windows(np.arange(S), np.zeros(S, np.int32))
        m.select_windows(0, S)
        m.forward_current()
        g = m.bp()
        f, P = m.batch_stats()
        act = inter["m"][:, :, 0] > 0                                    # (F, C) active pairs, list order = row-major
        def err(a, b):
            return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))
        dact = m.debug_read("dact", (5, F, 56)).transpose(1, 0, 2)                # layer-major on the device
        e = {"z": err(dact[:, 0, :50], inter["us"][0][:, :50])}
        assert np.all(dact[:, :, 50] == 1) and np.all(dact[:, :, 51] == 0)
        for i in range(1, 5):
            e[f"u{i}"] = err(dact[:, i, :50], inter["us"][i])
        ddz = m.debug_read("ddz", (5, F, 56)).transpose(1, 0, 2)
        e["dz5_dec"] = err(ddz[:, 4, :22], inter["dz_dec"][5])
        for i in range(4, 0, -1):
            e[f"dz{i}_dec"] = err(ddz[:, i - 1, :50], inter["dz_dec"][i])
        ds = m.debug_read("ds", (F, 52))
        e["ds"] = err(ds[:, :50], inter["ds"])
        SR = (F * 16 + 127) // 128 * 128                                  # slot capacity: pair capacity (16 per focus) in 128-row tiles
        hact = m.debug_read("hact", (6, SR, 56))[:, :P].transpose(1, 0, 2)
        dzact = m.debug_read("dzact", (6, SR, 56))[:, :P].transpose(1, 0, 2)
        # A pre-activation within rounding noise of zero may be > 0 on one side and == 0 on the other; from that layer
        # down the two gradients of that row legitimately differ.  Rows with such a ReLU flip are counted and excluded.
        flip_p = np.zeros(P, bool)
        for l in range(6):
            flip_p |= ((hact[:, l, :50] > 0) != (inter["hs"][l][act] > 0)).any(axis=1)
        flip_f = np.zeros(F, bool)
        for i in range(1, 5):
            flip_f |= ((dact[:, i, :50] > 0) != (inter["us"][i] > 0)).any(axis=1)
        pf = np.repeat(np.arange(F), act.sum(axis=1))                    # focus of every active pair
        flip_p |= flip_f[pf]
        assert flip_p.mean() < 0.02 and flip_f.mean() < 0.02, (flip_p.mean(), flip_f.mean())
        okp, okf = ~flip_p, ~flip_f
        for i in range(4, 0, -1):
            e[f"dz{i}_dec"] = err(ddz[okf, i - 1, :50], inter["dz_dec"][i][okf])
        e["ds"] = err(ds[okf, :50], inter["ds"][okf])
        for l in range(6):
            e[f"h{l}"] = err(hact[:, l, :50], inter["hs"][l][act])
            e[f"dz{l}_pair"] = err(dzact[okp, l, :50], inter["dz_pair"][l][act][okp])
        e["flips"] = float(flip_p.sum() + flip_f.sum())
        ents, _ = npe.param_layout()
        for n, o, s in ents:
            k = int(np.prod(s))
            e["g:" + n] = grad_err(g[o:o + k], g_ref[o:o + k]) if np.max(np.abs(g_ref[o:o + k])) > 0 else 0.0
        print("[s64 stage by stage] " + " ".join(f"{k}={v:.1e}" for k, v in e.items()))
        bad = {k: v for k, v in e.items() if k != "flips" and not v < (5e-4 if k.startswith("g:") else 1e-4)}
        assert not bad, bad
    finally:
        m.close()


def test_resident_rollout_fused_postprocess_equals_separate_pass(trained_like_params):
    """npe_rollout_resident on ball scenes (every object a focus, no ground truth): the decoder kernel applies
    simulate_all_postprocess and shifts the window itself.  The window after 6 steps must be the bits of the separate
    post-process pass (k_rollout_post), which is the one compared with the oracle elsewhere; 2 048 scenes = 16 384 foci."""
    import os
    st = scenes.raw_to_state(scenes.gen_balls(2048, 8, 2, seed=31))
    st[::7, 3, :, 4] = 0.9; st[::7, 3, :, 5] = 0.3            # some angles that wrap (balls: zeroed by ball_zero_mode)
    m = make_model(12, {})
    try:
        m.set_params(trained_like_params)
        m.upload_scenes(st)
        m.set_precision("tf32x3"); m.set_rollout_mode("kernels")
        os.environ["NPE_NO_FUSED_POST"] = "1"
        m.rollout_resident(0, 6)
        w_sep = m.debug_read("roll", (2048, 8, 2, 22))
        del os.environ["NPE_NO_FUSED_POST"]
        m.rollout_resident(0, 6)
        w_fused = m.debug_read("roll", (2048, 8, 2, 22))
        assert np.array_equal(w_fused, w_sep)
        traj, _ = m.rollout(0, 6, use_gt=False, want_metrics=False)       # trajectory output: separate pass again
        assert np.array_equal(traj[:, :, 5], w_fused[:, :, 1]) and np.array_equal(traj[:, :, 4], w_fused[:, :, 0])
    finally:
        os.environ.pop("NPE_NO_FUSED_POST", None)
        m.close()


def test_weight_images_follow_the_optimiser(n8, trained_like_params):
    """Large-batch training keeps the hi / lo TF32 weight images of the tcgen05 kernels current from inside the optimiser
    kernel (no re-packing launches between steps).  Three steps in a row must give the SAME BITS as three steps with the
    images re-packed from the parameter vector before every step (npe_set_params marks them stale)."""
    st, _ = n8
    S = 1024
    out = []
    for repack in (False, True):
        m = make_model(12, DISPATCH["default"])
        try:
            m.set_params(trained_like_params); m.optim_reset(0.0)
            m.upload_scenes(st[:S])
            m.upload_windows(np.arange(S), np.zeros(S, np.int32))
            losses = []
            for step in range(3):
                if repack and step:
                    m.set_params(m.get_params())
                assert m.select_windows(0, S) == S * 8
                losses.append(m.train_step("adam", 1e-3))
            out.append((m.get_params(), np.array(losses)))
        finally:
            m.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
    assert not np.array_equal(out[0][0], trained_like_params)


@pytest.mark.parametrize("n_obj,n_scenes", [(33, 27), (14, 61), (64, 11)])
def test_throughput_pair_builder_trim_and_ties(n_obj, n_scenes):
    """k_build_pairs_scan (more than 64 foci: a thread per focus, chained scan across blocks) with the 12-nearest trim ACTIVE
    on an 80-px lattice (equidistant contexts everywhere, more than 12 inside many neighbourhoods): kept-context lists
    (the on-demand path of npe_build_pairs), neighbourhood masks, the active-pair count and the dense pair list that the
    forward consumes == the oracle.  Block counts: 891 / 854 / 704 foci = 4 / 4 / 3 blocks, the last one partial."""
    from oracle import batcher
    rng = np.random.default_rng(1000 + n_obj)
    pos = (rng.integers(0, 5, (n_scenes, n_obj, 2)) * 80).astype(np.float32) / np.float32(800)   # 5 x 5 lattice: dense, many exact ties
    st = np.zeros((n_scenes, n_obj, 3, 22), np.float32)
    st[:, :, :, 0:2] = pos[:, :, None, :]
    st[..., 2:4] = (rng.standard_normal((n_scenes, n_obj, 1, 2)) * 0.1).astype(np.float32)
    st[..., 6] = 1; st[..., 10] = 1; st[..., 14] = 1; st[..., 16] = 1; st[..., 18] = 1; st[..., 20] = 1
    this, ctx, y = scene_examples(st, 0)
    oidx, omask, tctx = None, None, None
    idx_o, d = batcher.k_nearest_indices(this, ctx, 12)
    tctx = np.take_along_axis(ctx, idx_o[:, :, None, None], axis=1)
    omask = npe.neighbor_mask(this, tctx, check_focus=False)
    params = npe.init_params(1)
    m = make_model(12, {})
    try:
        m.set_params(params)
        m.upload_scenes(st)
        m.upload_windows(np.arange(n_scenes), np.zeros(n_scenes, np.int32))
        F = m.select_windows(0, n_scenes)
        assert F == n_scenes * n_obj and F > 592
        # the hot path first: it only needs the neighbour bits (the kept-context list is completed on demand below)
        pred, loss3 = m.forward_current()
        f, p = m.batch_stats()
        assert p == int(omask.sum())
        pred_o = npe.forward(params, this.reshape(F, -1), tctx.reshape(F, tctx.shape[1], -1), omask)
        assert rel_err(pred[:, :6], pred_o[:, :6]) < TOL
        idx, mask, cnt = m.pair_table()
        k = idx_o.shape[1]
        assert np.all(cnt == k) and np.array_equal(idx[:, :k], idx_o.astype(np.int32))
        assert np.array_equal(mask[:, :k], omask.astype(np.uint8))
        # and the tables are the same after the list was completed (a refresh rewrites identical tables)
        pred2, _ = m.forward_current()
        assert np.array_equal(pred2, pred)
        # the trim really cut INSIDE neighbourhoods (more than 12 contexts within 210 px before the trim)
        assert npe.neighbor_mask(this, ctx, check_focus=False).sum(1).max() > 12
    finally:
        m.close()


def test_throughput_pair_builder_ragged_scenes(trained_like_params):
    """The same kernel on a padded scene tensor with variable object counts, obstacles that are contexts but never foci
    (wall scenes: 28 obstacles + 2 balls, trim to the 12 nearest active) and a focus count that is no multiple of the
    256-focus blocks: prediction and pair count per focus against the oracle, focus by focus."""
    nw, nb = 120, 70
    sw = scenes.raw_to_state(scenes.gen_walls(nw, 6, seed=5, variant="O"))        # N = 30, 2 foci per scene
    sb = scenes.raw_to_state(scenes.gen_balls(nb, 7, 6, seed=6))                   # N = 7, 7 foci per scene
    Nmax = 30
    S = nw + nb
    st = np.zeros((S, Nmax, 6, 22), np.float32)
    order = np.random.default_rng(3).permutation(S)                                # interleave the two kinds of scenes
    n_obj = np.zeros(S, np.int32)
    for dst, src in enumerate(order):
        if src < nw: st[dst] = sw[src]; n_obj[dst] = 30
        else: st[dst, :7] = sb[src - nw]; n_obj[dst] = 7
    m = make_model(12, {})
    try:
        m.set_params(trained_like_params)
        m.upload_scenes(st, n_obj)
        m.upload_windows(np.arange(S), np.ones(S, np.int32))
        F = m.select_windows(0, S)
        assert F == nw * 2 + nb * 7 and F > 592 and F % 256 != 0
        pred_g, _ = m.forward_current()
        idx_g, mask_g, cnt_g = m.pair_table()
        row, pairs = 0, 0
        for s in range(S):
            N = n_obj[s]
            for o in range(N):
                if st[s, o, 0, 11] == 1:
                    continue
                this = st[s:s + 1, o, 1:3]; ctx = np.delete(st[s:s + 1, :N], o, axis=1)[:, :, 1:3]
                idx, d = batcher.k_nearest_indices(this, ctx, 12)
                tctx = np.take_along_axis(ctx, idx[:, :, None, None], axis=1)
                mask = npe.neighbor_mask(this, tctx, check_focus=False)
                k = idx.shape[1]
                assert cnt_g[row] == k and np.array_equal(idx_g[row, :k], idx[0].astype(np.int32)), (s, o)
                assert np.array_equal(mask_g[row, :k], mask[0].astype(np.uint8)), (s, o)
                if row % 7 == 0:      # the forward on a sample of the foci (the oracle is a Python loop here)
                    pr = npe.forward(trained_like_params, this.reshape(1, -1), tctx.reshape(1, k, -1), mask)
                    assert rel_err(pred_g[row, :6], pr[0, :6]) < TOL, (s, o)
                pairs += int(mask.sum())
                row += 1
        assert row == F
        assert m.batch_stats() == (F, pairs)
    finally:
        m.close()


def test_stacked_decoder_forward(n8, trained_like_params):
    """k_dec_forward_s2 (hi and lo weight images stacked along N: 2 MMAs per K-step, layer-1 input in two K-halves) is what
    the rollout's steady state runs on; it is chosen automatically from 8 tiles per SM on and forced here
    (NPE_OPT_DEC_STACKED = 1).  Same three products per K-step as the plain 3xTF32 kernel: the FP32-grade tolerances hold
    for the inference forward on 40 960 foci (predictions, per-example losses) and for the fused rollout (window after 6
    steps against the plain kernel with the separate post-process pass, which other tests compare with the oracle:
    positions bit-exact, velocities within 1e-4)."""
    st, ref = n8
    S = st.shape[0]
    m = make_model(12, {})
    try:
        m.set_params(trained_like_params)
        m.upload_scenes(st)
        m.upload_windows(np.arange(S), np.zeros(S, np.int32))
        m.select_windows(0, S)
        m.set_precision("tf32x3")
        m.set_option("dec_stacked", 2)
        pred_plain, le_plain = m.forward_per_example_current()
        m.set_option("dec_stacked", 1)
        pred, le = m.forward_per_example_current()
        assert rel_err(pred[:, :6], ref["pred"][:, :6]) < TOL and float(np.max(np.abs(le - ref["le"]))) < 1e-6
        assert rel_err(pred[:, :6], pred_plain[:, :6]) < 2e-5
        assert np.max(np.abs(pred[:, 6:] - pred_plain[:, 6:])) < 1e-5      # sigmoid columns
        # fused rollout on ball scenes
        m.set_rollout_mode("kernels")
        sb = scenes.raw_to_state(scenes.gen_balls(2048, 8, 2, seed=31))
        m.upload_scenes(sb)
        m.set_option("dec_stacked", 2)
        m.rollout_resident(0, 6)
        w_plain = m.debug_read("roll", (2048, 8, 2, 22))
        m.set_option("dec_stacked", 1)
        m.rollout_resident(0, 6)
        w = m.debug_read("roll", (2048, 8, 2, 22))
        vs = np.abs(w_plain[..., 2:6]).max()
        assert float(np.max(np.abs(w[..., 2:6] - w_plain[..., 2:6])) / vs) < 1e-4
        assert float(np.max(np.abs(w[..., 0:2] - w_plain[..., 0:2]))) < 1e-5   # positions integrate velocities of earlier steps
        assert np.array_equal(w[..., 6:], w_plain[..., 6:])                     # properties are copied
    finally:
        m.close()


def test_overlapped_launches_do_not_change_results(n8, trained_like_params):
    """In the throughput regime a training loop overlaps three things: the decoder weight gradients run on a side stream
    beside the pair input-gradient chain, the NEXT batch's pair tables are built on the upload stream beside the reduce
    kernel (double-buffered tables), and the optimiser writes the weight images through.  With a kernel class selected for
    event bracketing (npe_profile_select) the first two are switched off (one stream, tables built lazily): four steps over
    alternating batches must give the same bits either way."""
    st, _ = n8
    S = 2048
    out = []
    for bracket in (False, True):
        m = make_model(12, DISPATCH["default"])
        try:
            m.set_params(trained_like_params); m.optim_reset(0.0)
            m.upload_scenes(st[:S])
            m.upload_windows(np.arange(S), np.zeros(S, np.int32))
            if bracket:
                m.profile_select("reduce")
            losses = []
            for step in range(4):
                assert m.select_windows((step % 2) * (S // 2), S // 2) == (S // 2) * 8
                losses.append(m.train_step("adam", 1e-3))
            out.append((m.get_params(), np.array(losses), m.get_grads()))
        finally:
            m.close()
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])
