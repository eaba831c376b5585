"""CPU tests of the training protocol (dynamics_b200/train.py, mirror of main.lua:271-437,490-540) with a stand-in for the
device model: which iterations print / validate / save / decay the learning rate, chunked == stepwise, resume."""
import math
import os

import numpy as np
import pytest

from dynamics_b200 import t7, train
from tests.test_loader import _write_dataset


class FakeModel:
    """Deterministic stand-in for NPEModel: the 'loss' of a batch depends on its focus triples and on a scalar state
    that every optimiser step moves, so that the order and the learning rate of the steps are observable."""
    instances = []

    def __init__(self, params=None):
        self.p = np.zeros(28222, np.float32) if params is None else np.asarray(params, np.float32).copy()
        self.cur = None
        self.calls = []
        FakeModel.instances.append(self)

    @classmethod
    def create(cls, mp=None, preload=False, model_path=None, params=None):
        if preload:
            params = t7.load_checkpoint_params(model_path)
        return cls(params)

    def _loss(self):
        sc, t0 = self.cur
        return float(1.0 + 1e-3 * (int(sc.sum()) % 97) + 1e-2 * int(t0[0]) + abs(float(self.p[0])))

    def optim_reset(self, m0=0.0): pass
    def set_params(self, p): self.p = np.asarray(p, np.float32).copy()
    def get_params(self): return self.p.copy()
    def get_grads(self): return np.ones(28222, np.float32)
    def upload_scenes(self, states, n_obj=None): self.n_scenes = states.shape[0]
    def upload_foci(self, sc, ob, t0): self.foci = (np.asarray(sc), np.asarray(t0))
    def select_foci(self, first, count):
        self.cur = (self.foci[0][first:first + count], self.foci[1][first:first + count]); return count
    def fp(self, params_, batch, sim=False):
        batch.select(self); return self._loss(), None, 0.0, 0.0

    def train_step(self, opt="adam", lr=3e-4, divisor_batch=0, want_loss=True):
        loss = self._loss()
        self.p[0] -= np.float32(lr)                                  # the "optimiser step"
        self.calls.append(("step", lr))
        return np.array([loss, 0, 0], np.float32)

    def train_steps(self, first, batch, nsteps, opt="adam", lr=3e-4, divisor_batch=0, want_losses=False):
        out = np.zeros((nsteps, 3), np.float32)
        for s in range(nsteps):
            self.select_foci(first + s * batch, batch)
            out[s] = self.train_step(opt, lr)
        self.calls[-nsteps:] = [("chunk", nsteps, lr)]
        return out

    def close(self): pass


@pytest.fixture()
def fake_model(monkeypatch):
    FakeModel.instances = []
    monkeypatch.setattr(train, "NPEModel", FakeModel)
    return FakeModel


def _exp(tmp_path, tag, extra=()):
    argv = ["-mode", "exp", "-nbrhd", "-rs", "-opt", "adam", "-name", tag, "-data_root", str(tmp_path), "-logs_root",
            str(tmp_path / "logs"), "-dataset_folders", "{'balls_n3_t60_ex100_m'}", "-test_dataset_folders",
            "{'balls_n3_t60_ex100_m'}", "-max_iter", "31", "-print_every", "7", "-val_every", "10", "-save_every", "20",
            "-lrdecay_every", "5", "-lrdecayafter", "12", "-lrdecay", "0.5", "-seed", "1"] + list(extra)
    e = train.Experiment(train.parse_args(argv))
    e.history = []
    return e


def test_protocol_events_and_chunking(tmp_path, fake_model, capsys):
    _write_dataset(tmp_path, "balls_n3_t60_ex100_m", 3, 100, seed=1)
    a = _exp(tmp_path, "chunked"); a.run_experiment()
    b = _exp(tmp_path, "stepwise"); b.stepwise = True; b.run_experiment()
    assert a.history == b.history and len(a.history) == 31
    assert np.array_equal(a.model.get_params(), b.model.get_params())
    # chunks end exactly where main.lua prints (7,14,21,28), validates (10,20,30) or decays the lr (15,20,25,30)
    chunks = [c[1] for c in a.model.calls if c[0] == "chunk"]
    steps = [c for c in a.model.calls if c[0] == "step"]
    ends = np.cumsum([c[1] if c[0] == "chunk" else 1 for c in a.model.calls])
    assert sorted(set(ends.tolist())) == sorted({7, 14, 21, 28, 10, 20, 30, 5, 15, 25, 31})
    assert all(c[0] == "step" for c in b.model.calls) and len(b.model.calls) == 31 and len(chunks) + len(steps) == len(ends)
    # learning rate in force during iteration t: halves after t = 15, 20, 25, 30 (t >= 12 and t % 5 == 0)
    want = [3e-4 * 0.5 ** sum(1 for d in (15, 20, 25, 30) if t > d) for t in range(1, 32)]
    assert [h[2] for h in a.history] == pytest.approx(want)
    # validation before training and after iterations 10, 20, 30; checkpoints at 0 and 20 (save_every = 20)
    assert len(a.val_losses) == 4
    snaps = sorted(f for f in os.listdir(tmp_path / "logs" / "chunked") if f.startswith("epoch"))
    assert [s.split("_")[1] for s in snaps] == ["step0", "step20"]
    out = capsys.readouterr().out
    assert out.count("iteration") >= 8 and "Learning rate is now" in out
    rows = open(tmp_path / "logs" / "chunked" / "train.log").read().splitlines()
    assert len(rows) == 32 and float(rows[1]) == pytest.approx(math.log(a.history[0][3]), rel=1e-5)


def test_convergence_break_and_resume(tmp_path, fake_model):
    _write_dataset(tmp_path, "balls_n3_t60_ex100_m", 3, 100, seed=1)
    # a window of 2 validations whose spread is below val_eps stops the run at the second validation after the start one
    e = _exp(tmp_path, "conv", ["-val_window", "2", "-val_eps", "1e9"])
    e.run_experiment()
    assert len(e.history) == 10 and len(e.val_losses) == 2
    # resume: newest snapshot (step 20 of a full run), lr decayed as main.lua:517-519 does, iterations 21..31
    f = _exp(tmp_path, "full"); f.run_experiment()
    g = train.Experiment(train.parse_args(["-mode", "expload", "-name", "full", "-data_root", str(tmp_path), "-logs_root",
                                           str(tmp_path / "logs")]))
    g.history = []
    g.run_experiment_load()
    assert len(g.history) == 11 and g.mp.mode == "expload"
    assert g.history[0][2] == pytest.approx(3e-4 * 0.5 ** 2)          # decays at 15 and (on load) 20
    ck = t7.load(train.get_last_snapshot(str(tmp_path / "logs" / "full")))
    assert int(ck["iters"]) == 20 and np.asarray(ck["model"]["theta"]["params"]).size == 28222
    assert os.path.exists(tmp_path / "logs" / "full" / "train_21.log")


def test_slot_metrics_reduce_like_eval_lua():
    """model.slot_metrics_to_batch_row == the per-batch averages of eval.lua:337-358,385-391 computed the long way:
    per slot the mean over valid examples, summed over slots and divided by the summed valid fractions."""
    from dynamics_b200.model import slot_metrics_to_batch_row
    rng = np.random.default_rng(3)
    B, N, steps = 12, 7, 4
    valid = rng.random((steps, N, B)) < 0.7
    valid[:, 0] = True; valid[:, 1] = False                        # an all-valid slot and an all-obstacle slot
    amask = valid & (rng.random((steps, N, B)) < 0.8)
    amask[:, 2] = False                                             # valid foci but no usable angle in slot 2
    vals = rng.random((5, steps, N, B))
    slots = np.zeros((steps, N, 7))
    for k, (src, m) in enumerate(((0, valid), (1, valid), (2, valid))):
        slots[:, :, k] = (vals[src] * m).sum(-1)
    slots[:, :, 3] = valid.sum(-1)
    slots[:, :, 4] = (vals[3] * amask).sum(-1); slots[:, :, 5] = (vals[4] * amask).sum(-1); slots[:, :, 6] = amask.sum(-1)
    row = slot_metrics_to_batch_row(slots, B)
    for t in range(steps):
        acc = dict(loss=0.0, vel=0.0, angvel=0.0, cos=0.0, mag=0.0); cnt = 0.0; acnt = 0.0
        for j in range(N):
            v, am = valid[t, j], amask[t, j]
            if v.sum() == 0:
                continue                                            # every focus of the slot is an obstacle: copied, not scored
            for key, src in (("loss", 0), ("vel", 1), ("angvel", 2)):
                acc[key] += (vals[src, t, j] * v).sum() / v.sum()
            with np.errstate(invalid="ignore", divide="ignore"):
                acc["cos"] += (vals[3, t, j] * am).sum() / am.sum()      # 0/0 = nan, like apply_mask_avg
                acc["mag"] += (vals[4, t, j] * am).sum() / am.sum()
            cnt += v.sum() / B; acnt += am.sum() / B
        for key in ("loss", "vel", "angvel"):
            assert row[key][t] == pytest.approx(acc[key] / cnt)
        assert np.isnan(row["cos"][t]) and np.isnan(acc["cos"] / acnt)  # slot 2 poisons the angle columns, as in the reference
    # without the degenerate slot the angle columns are finite and agree
    amask[:, 2] = valid[:, 2]
    slots[:, :, 4] = (vals[3] * amask).sum(-1); slots[:, :, 5] = (vals[4] * amask).sum(-1); slots[:, :, 6] = amask.sum(-1)
    row = slot_metrics_to_batch_row(slots, B)
    for t in range(steps):
        num = sum((vals[3, t, j] * amask[t, j]).sum() / amask[t, j].sum() for j in range(N) if amask[t, j].sum() > 0)
        den = sum(amask[t, j].sum() / B for j in range(N))
        assert row["cos"][t] == pytest.approx(num / den)


def test_update_position_and_angle_mirror_the_reference():
    """model:update_position / update_angle (npe.lua:386-458; called by eval.lua:170-173) on the drop-in surface: the host
    mirror (same arithmetic as dynamics_b200/lua/npe.lua) against the oracle's restatement, bit for bit."""
    from oracle import rollout
    from dynamics_b200.model import NPEModel
    rng = np.random.default_rng(3)
    this = rng.uniform(-1, 1, (64, 2, 22)).astype(np.float32)
    pred = rng.uniform(-1, 1, (64, 1, 22)).astype(np.float32)
    this[:8, -1, 4] = 0.99; this[:8, -1, 5] = 0.5            # wraps past +pi
    this[8:16, -1, 4] = -0.99; this[8:16, -1, 5] = -0.5      # wraps past -pi
    m = NPEModel.__new__(NPEModel)                            # pure host arithmetic: no device handle needed
    assert np.array_equal(NPEModel.update_position(m, this, pred), rollout.update_position(this, pred))
    got = NPEModel.update_angle(m, this, pred)
    assert np.array_equal(got, rollout.update_angle(this, pred))
    assert np.all(np.abs(got[:16, 0, 4]) <= 1.0)
    # several future steps: step k integrates the velocity / angular velocity of step k - 1
    pred3 = rng.uniform(-1, 1, (64, 3, 22)).astype(np.float32)
    out = NPEModel.update_position(m, this, pred3)
    p = this[:, -1, 0:2] * np.float32(800); v = this[:, -1, 2:4] * np.float32(60)
    for k in range(3):
        p = p + v; v = pred3[:, k, 2:4] * np.float32(60)
        assert np.array_equal(out[:, k, 0:2], p / np.float32(800))
    assert np.array_equal(out[:, :, 2:], pred3[:, :, 2:])
    with pytest.raises(Exception):
        NPEModel.bp_input(m)
