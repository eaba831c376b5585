"""-m gpu tests of the callers either side of the hot path (SURVEY 8a a11/a12/a13, 8b loader boundary, 8f): the
device-resident loader, the main.lua training protocol and the eval.lua -mode sim driver, each against the oracle."""
import json
import math
import os

import numpy as np
import pytest

from dynamics_b200 import dataprep, evaluate, t7, train
from dynamics_b200.loader import GeneralDataSampler
from dynamics_b200.model import slot_metrics_to_batch_row
from oracle import batcher, npe, optim, rollout, scenes
from tests.test_loader import _args, _write_dataset
from tests.util import rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-4


def test_loader_batches_match_oracle(model, trained_like_params, tmp_path):
    files3 = _write_dataset(tmp_path, "balls_n3_t60_ex100_m", 3, 100, seed=1)["trainset"]
    files5 = _write_dataset(tmp_path, "balls_n5_t60_ex60_m", 5, 60, seed=2)["trainset"]
    g = GeneralDataSampler.create("trainset", _args(tmp_path, ["balls_n3_t60_ex100_m", "balls_n5_t60_ex60_m"], model=model))
    model.set_params(trained_like_params)
    for ds, files in ((1, files3), (2, files5)):
        d = g.datasamplers[ds - 1]
        for sid in (1, 20, 21, d.total_batches):
            b = d.load_subbatch_id(sid)
            fid, off = d.subbatch_location(sid)
            ref = batcher.make_batch(*files[fid - 1], offset=off - 1)
            l_ref, p_ref, _, _ = npe.fp(trained_like_params, ref)
            loss, pred, _, _ = model.fp(None, b)
            assert rel_err(pred[:, :6], p_ref[:, :6]) < TOL and abs(loss - l_ref) <= TOL * abs(l_ref)
            this, ctx, y, ctx_f, mask, orig, closest = b.host()
            assert np.array_equal(this, ref[0]) and np.array_equal(ctx, ref[1]) and np.array_equal(y, ref[2])
            assert np.array_equal(ctx_f, ref[3]) and np.array_equal(mask, ref[4]) and np.array_equal(closest, ref[6] + 1)


def test_loader_trims_walls_to_12_nearest(model, trained_like_params, tmp_path):
    st = scenes.raw_to_state(scenes.gen_walls(25, 60, seed=3, variant="U"))
    foc, ctx, _, _ = batcher.expand_for_each_object(st)             # 50 ball foci, 32 contexts each
    d = tmp_path / "walls_n2_t60_ex25_wU" / "batches" / "trainset"
    d.mkdir(parents=True)
    t7.save_batch_file(str(d / "batch1"), foc, ctx)
    g = GeneralDataSampler.create("trainset", _args(tmp_path, ["walls_n2_t60_ex25_wU"], model=model))
    b = g.datasamplers[0].load_subbatch_id(7)
    ref = batcher.make_batch(foc, ctx, offset=6)
    this, tctx, y, tctx_f, mask, orig, closest = b.host()
    assert tctx.shape[1] == 12 and np.array_equal(closest, ref[6] + 1) and np.array_equal(tctx, ref[1])
    assert np.array_equal(tctx_f, ref[3]) and mask[11] == 1 and mask.sum() == 1
    loss, pred, _, _ = model.fp(trained_like_params, b)
    l_ref, p_ref, _, _ = npe.fp(trained_like_params, ref)
    assert rel_err(pred[:, :6], p_ref[:, :6]) < TOL and abs(loss - l_ref) <= TOL * abs(l_ref)


def _experiment(tmp_path, tag, **kw):
    argv = ["-mode", "exp", "-nbrhd", "-rs", "-opt", "adam", "-name", tag, "-data_root", str(tmp_path),
            "-logs_root", str(tmp_path / "logs"), "-dataset_folders", "{'balls_n3_t60_ex100_m','balls_n4_t60_ex100_m'}",
            "-test_dataset_folders", "{'balls_n5_t60_ex60_m'}", "-max_iter", "23", "-print_every", "5", "-val_every", "10",
            "-save_every", "10", "-lrdecay_every", "4", "-lrdecayafter", "8", "-lrdecay", "0.5", "-seed", "3"]
    mp = train.parse_args(argv)
    for k, v in kw.items():
        setattr(mp, k, v)
    exp = train.Experiment(mp)
    exp.history = []
    return exp


def test_training_protocol_chunked_equals_stepwise_equals_oracle(tmp_path):
    f3 = _write_dataset(tmp_path, "balls_n3_t60_ex100_m", 3, 100, seed=1)
    f4 = _write_dataset(tmp_path, "balls_n4_t60_ex100_m", 4, 100, seed=2)
    _write_dataset(tmp_path, "balls_n5_t60_ex60_m", 5, 60, seed=5)
    a = _experiment(tmp_path, "chunked")
    a.run_experiment()
    b = _experiment(tmp_path, "stepwise")
    b.stepwise = True
    b.run_experiment()
    pa, pb = a.model.get_params(), b.model.get_params()
    assert np.array_equal(pa, pb) and a.history == b.history and len(a.history) == 23
    assert a.val_losses == b.val_losses and len(a.val_losses) == 3             # before training, after 10 and 20
    # learning rate: halves after iterations 8, 12, 16, 20 (t >= lrdecayafter and t % 4 == 0)
    lrs = [h[2] for h in a.history]
    assert lrs == [optim.lr_schedule(3e-4, t, 1, 0.5, 4, 8) for t in range(1, 24)]
    # oracle replay of the same sample sequence
    x = npe.init_params(3)
    x0 = x.copy()
    state = optim.adam_init(x.size)
    files = {1: f3["trainset"], 2: f4["trainset"]}
    for (ds, sid, lr, loss) in a.history:
        fid, off = a.train_loader.datasamplers[ds - 1].subbatch_location(sid)
        batch = batcher.make_batch(*files[ds][fid - 1], offset=off - 1)
        l_ref, _, _, _ = npe.fp(x, batch)
        g_ref, _ = npe.bp(x, batch)
        assert abs(loss - l_ref) <= 2e-4 * abs(l_ref)
        optim.adam_step(x, g_ref, state, lr)
    assert np.max(np.abs(pa - x)) / np.max(np.abs(x)) < TOL
    assert np.max(np.abs(pa - x0)) / np.max(np.abs(x)) > 50 * TOL              # and training moved the weights
    # artefacts of the protocol: logs, args, checkpoints readable by the rollout driver
    logdir = tmp_path / "logs" / "chunked"
    rows = open(logdir / "train.log").read().splitlines()
    assert rows[0].startswith("log MSE loss (train set)") and len(rows) == 24
    assert abs(float(rows[1]) - math.log(a.history[0][3])) < 1e-4
    assert len(open(logdir / "experiment.log").read().splitlines()) == 4
    snaps = sorted(os.listdir(logdir))
    assert sum(s.startswith("epoch") for s in snaps) == 3 and "args.t7" in snaps
    last = train.get_last_snapshot(str(logdir))
    assert "_step20_" in last
    ck = t7.load(last)
    assert int(ck["iters"]) == 20 and len(ck["val_losses"]) == 3
    # expload resumes at iteration 21 with the decayed learning rate and the checkpoint's weights
    c = train.Experiment(train.parse_args(["-mode", "expload", "-name", "chunked", "-data_root", str(tmp_path),
                                           "-logs_root", str(tmp_path / "logs")]))
    c.history = []
    c.run_experiment_load()
    assert len(c.history) == 3 and c.mp.mode == "expload"
    assert c.history[0][2] == pytest.approx(3e-4 * 0.5 ** 4)
    for m in (a.model, b.model, c.model):
        m.close()


def test_priority_sampling_runs_the_reference_protocol(tmp_path):
    _write_dataset(tmp_path, "balls_n3_t60_ex100_m", 3, 100, seed=1)
    _write_dataset(tmp_path, "balls_n4_t60_ex100_m", 4, 100, seed=2)
    _write_dataset(tmp_path, "balls_n5_t60_ex60_m", 5, 60, seed=5)
    e = _experiment(tmp_path, "prio", rs=False, max_iter=100, val_every=1000, save_every=1000, print_every=50)
    e.run_experiment()
    # sequential until each dataset's table is full (2 files x 20 offsets each), losses recorded as weights
    first = {}
    for ds, sid, lr, loss in e.history:
        first.setdefault(ds, []).append(sid)
    for ds, sampler in enumerate(e.train_loader.datasamplers, 1):
        n = min(40, len(first[ds]))
        assert first[ds][:n] == [int(i) for i in sampler.batch_idxs[:n]]
        assert (sampler.priority_sampler.batch_weights > 0).sum() >= n
    ds, sid, lr, loss = e.history[-1]
    assert e.train_loader.datasamplers[ds - 1].priority_sampler.batch_weights[sid - 1] == loss
    e.model.close()


def test_eval_sim_driver_matches_oracle(trained_like_params, tmp_path):
    files = _write_dataset(tmp_path, "balls_n4_t60_ex50_m", 4, 50, seed=8, sets=("testset",))["testset"]
    logdir = tmp_path / "logs" / "exp1"
    logdir.mkdir(parents=True)
    t7.save_checkpoint(str(logdir / "epoch1_step0_0.1000000.t7"), trained_like_params, {"batch_size": 50}, 0)
    steps = 9
    out = evaluate.main(["-mode", "sim", "-name", "exp1", "-test_dataset_folders", "{'balls_n4_t60_ex50_m'}", "-data_root",
                         str(tmp_path), "-logs_root", str(tmp_path / "logs"), "-ns", "1", "-steps", str(steps)])
    tables = out["balls_n4_t60_ex50_m"]
    assert tables["loss"].shape == (len(files), steps)
    refs, preds = [], []
    for foc, ctx in files:
        pred_sim, logs = rollout.simulate_all(trained_like_params, foc[:, 0:2], ctx[:, :, 0:2], foc[:, 2:], ctx[:, :, 2:], steps)
        refs.append(logs); preds.append(pred_sim)
    sub = logdir / "balls_n4_t60_ex50_m_predictions"
    log = [r.split("\t") for r in open(sub / "gt_divergence.log").read().splitlines()]
    assert log[0][:6] == ["Timesteps", "MSE Error", "Cosine Difference", "Magnitude Difference", "Velocity Error",
                          "Angular Velocity Error"]
    for col, key in ((1, "loss"), (2, "cos"), (3, "mag"), (4, "vel"), (5, "angvel")):
        want = np.mean([r[key] for r in refs], 0)
        got = np.array([float(r[col]) for r in log[1:]])
        assert np.allclose(got, want, rtol=2e-3, atol=1e-7), key
    # the one saved example file: past + prediction in raw units, readable by the json loader
    names = sorted(n for n in os.listdir(sub) if n.endswith(".json"))
    assert len([n for n in names if n.startswith("pred_")]) == 1 and len([n for n in names if n.startswith("gt_")]) == 1
    fid = int(names[-1].split("batch")[1].split(".")[0])
    raw = dataprep.load_data_json(json.load(open(sub / f"pred_batch{fid}.json")))
    want = scenes.state_to_raw(np.concatenate([np.concatenate([files[fid - 1][0][:, None, 0:2], files[fid - 1][1][:, :, 0:2]], 1),
                                               preds[fid - 1]], 2))
    assert raw.shape == want.shape == (50, 4, 2 + steps, 12)
    assert np.max(np.abs(raw[..., 0:2] - want[..., 0:2])) < 0.5 and np.max(np.abs(raw[..., 2:4] - want[..., 2:4])) < 0.1


def test_slot_metrics_reproduce_mixed_validity_average(model, trained_like_params):
    """Obstacles in a slot for some examples only: eval.lua sums per-slot means and divides by the valid fractions."""
    st = scenes.raw_to_state(scenes.gen_walls(12, 8, seed=12, variant="O", n_balls=3))
    N = st.shape[1]
    rng = np.random.default_rng(0)
    for e in range(0, 12, 2):                                                  # permute the objects of every other example
        st[e] = st[e, rng.permutation(N)]
    steps = 5
    model.set_params(trained_like_params)
    model.upload_scenes(st)
    traj, slots = model.rollout_slots(0, steps, use_gt=True)
    row = slot_metrics_to_batch_row(slots, 12)
    pred_sim, logs = rollout.simulate_all(trained_like_params, st[:, 0, 0:2], st[:, 1:, 0:2], st[:, 0, 2:], st[:, 1:, 2:], steps)
    for k in ("loss", "vel", "angvel", "cos", "mag"):
        assert np.allclose(row[k], logs[k], rtol=2e-3, atol=1e-7, equal_nan=True), k
    assert np.max(np.abs(traj[..., 0:4] - pred_sim[..., 0:4])) < 1e-3
    _, tot = model.rollout(0, steps, use_gt=True)
    assert np.allclose(slots.sum(1), tot, rtol=1e-12)


def test_grouped_validation_equals_per_batch_fp(model, trained_like_params, tmp_path):
    """Experiment.test (main.lua:398-414) evaluates VAL_GROUP reference batches per device call; the per-batch losses it
    derives from the per-example losses equal what model:fp returns for each batch alone (1e-5: a group of 1 600 examples
    runs on the 64-row kernels, a single batch of 50 on the 16-row ones), across two datasets with different object
    counts and a group boundary inside the walk."""
    _write_dataset(tmp_path, "balls_n3_t60_ex100_m", 3, 100, seed=1)
    _write_dataset(tmp_path, "balls_n5_t60_ex60_m", 5, 60, seed=2)
    g = GeneralDataSampler.create("trainset", _args(tmp_path, ["balls_n3_t60_ex100_m", "balls_n5_t60_ex60_m"], model=model))
    model.set_params(trained_like_params)
    n = 70
    single = []
    for _ in range(n):
        b, _ = g.sample_sequential_batch(False)
        single.append(model.fp(None, b)[0])
    for d in g.datasamplers:
        d.current_batch = 0
    g.current_dataset = 1

    class E:                                      # the two attributes Experiment.test / _eval_group read
        pass
    e = E(); e.model = model; e.mp = type("mp", (), {"fast": False})(); e.world = 1; e.rank = 0; e.dist = None; e.say = lambda *a: None
    e.VAL_GROUP = 32
    e._eval_group = lambda bs: train.Experiment._eval_group(e, bs)
    mean = train.Experiment.test(e, g, n)
    groups = []
    for d in g.datasamplers:
        d.current_batch = 0
    g.current_dataset = 1
    bs = [g.sample_sequential_batch(False)[0] for _ in range(n)]
    for a in range(0, n, 32):
        groups += train.Experiment._eval_group(e, bs[a:a + 32])
    assert np.allclose(groups, single, rtol=1e-5, atol=0)
    assert abs(mean - float(np.mean(np.asarray(groups, np.float64)))) < 1e-12
    # a group of ONE batch takes the same kernels as model:fp: the loss arithmetic on the host is the device's, bit for bit
    assert train.Experiment._eval_group(e, bs[:1]) == [float(single[0])]


def test_device_packer_raw_to_states_is_bit_exact(model):
    """f1 on the device (npe_scenes_upload_raw): raw matter-js states -> normalised one-hot states, bit for bit the oracle's
    data_process:normalize + properties2onehotall (oracle/scenes.py raw_to_state), on balls and on walls (static
    obstacles of mass 1e30), angles beyond pi included; obstacles are contexts, never foci; bad categories are refused."""
    from dynamics_b200 import NPEError
    rawb = scenes.gen_balls(12, 5, 6, seed=3)
    rawb[:, :, :, 4] = np.linspace(-3.5, 3.5, 12 * 5 * 6, dtype=np.float32).reshape(12, 5, 6)      # angles across +-pi
    rawb[:, :, :, 5] = 0.25
    raww = scenes.gen_walls(4, 6, seed=5, variant="O")
    for raw in (rawb, raww):
        model.upload_scenes_raw(raw)
        S, N, T, _ = raw.shape
        got = model.debug_read("states", (S, N, T, 22))
        assert np.array_equal(got, scenes.raw_to_state(raw))
    model.upload_windows(np.arange(4), np.zeros(4, np.int32))
    n_balls = int((raww[0, :, 0, 7] == 1).sum())
    assert model.select_windows(0, 4) == 4 * n_balls                     # foci = non-obstacle objects only
    bad = rawb.copy(); bad[0, 0, 0, 6] = 7.0                             # a mass outside {1, 5, 25, 1e30}
    with pytest.raises(NPEError):
        model.upload_scenes_raw(bad)


def test_device_priority_sampler(model):
    """priority_sampler.lua on the device: update_batch_weight / sample (uniform with probability alpha, else multinomial over
    weight^pow) / get_hardest_batch; frequencies of 40 000 draws against the exact probabilities; the reference's assert on
    a table that is not full; the loss of every step of npe_train_steps fed to the table without a host round trip."""
    from dynamics_b200 import NPEError
    w = np.array([1.0, 2.0, 3.0, 4.0, 10.0], np.float32)
    model.priority_init(0, 5, alpha=0.0)
    with pytest.raises(NPEError):
        model.priority_sample(0, 1.0, seed=1, n=4)                    # table not full: weights are all zero
    model.priority_update(0, np.arange(1, 6), w)
    assert model.priority_hardest(0) == [10.0, 5]
    for pw in (1.0, 2.0):
        ids = model.priority_sample(0, pw, seed=7, n=40000)
        assert ids.min() >= 1 and ids.max() <= 5
        freq = np.bincount(ids, minlength=6)[1:] / 40000.0
        p = w.astype(np.float64) ** pw; p /= p.sum()
        assert np.max(np.abs(freq - p)) < 0.01, (freq, p)
    assert np.array_equal(model.priority_sample(0, 1.0, seed=7, n=64), model.priority_sample(0, 1.0, seed=7, n=64))   # same seed, same draws
    model.priority_init(1, 5, alpha=1.0)                              # second table, always uniform: a zero table is fine
    ids = model.priority_sample(1, 1.0, seed=3, n=20000)
    assert np.max(np.abs(np.bincount(ids, minlength=6)[1:] / 20000.0 - 0.2)) < 0.015
    # losses of npe_train_steps -> table, on the device
    st = scenes.raw_to_state(scenes.gen_balls(100, 4, 12, seed=51))
    sc = np.arange(100); ob = np.tile(np.arange(4), 25); t0 = (np.arange(100) * 7) % 9
    model.upload_scenes(st); model.upload_foci(sc, ob, t0)
    model.set_params(npe.init_params(1234)); model.optim_reset(0.0)
    losses = model.train_steps(0, 50, 2, "adam", 3e-4, want_losses=True)
    model.priority_init(2, 4, alpha=0.0)
    model.priority_update(2, [3, 1])                                  # step 0 -> batch 3, step 1 -> batch 1
    model.priority_update(2, [2, 4], [1e-9, 1e-9])
    hw, hid = model.priority_hardest(2)
    assert hid == (3 if losses[0, 0] >= losses[1, 0] else 1) and abs(hw - max(losses[0, 0], losses[1, 0])) < 1e-12
