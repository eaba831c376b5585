"""CPU tests of the oracle: internal cross-validation (the reference has no tests of its own -- parity unpinned,
SURVEY.md section 4/8c) and the committed golden fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import batcher, npe, optim, rollout, scenes
from oracle import config as OC
from tests.util import balls_batches, grad_err, rel_err

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_param_layout_matches_survey():
    ents, n = npe.param_layout()
    assert n == 28222
    offs = {name: off for name, off, _ in ents}
    assert offs["enc.this.W"] == 0 and offs["enc.ctx.W"] == 1100
    assert [offs[f"pair.L{l}.W"] for l in range(1, 6)] == [2200, 4700, 7200, 9700, 12200]
    assert offs["dec.L1.W"] == 14700 and offs["dec.L1.b"] == 19400
    assert offs["dec.L2.W"] == 19450 and offs["dec.L5.W"] == 27100 and offs["dec.L5.b"] == 28200
    p = npe.init_params(1234)
    v = npe.views(p)
    for l in range(2, 6):  # layer:clone() -> identical initial values (npe.lua:39)
        assert np.array_equal(v["pair.L1.W"], v[f"pair.L{l}.W"])
    assert np.abs(v["dec.L1.W"]).max() <= 1 / np.sqrt(94)


def _torch_objective(params64, this_past, ctx, y, mask, vlambda=1.0, lam=1.0):
    ents, _ = npe.param_layout()
    p = {n: params64[o:o + int(np.prod(s))].reshape(s) for n, o, s in ents}
    f = torch.tensor(this_past, dtype=torch.float64); c = torch.tensor(ctx, dtype=torch.float64)
    m = torch.tensor(mask, dtype=torch.float64)[:, :, None]
    et = torch.relu((f[:, None, :] * m) @ p["enc.this.W"].T); ec = torch.relu((c * m) @ p["enc.ctx.W"].T)
    h = torch.cat([et, ec], 2)
    for l in range(1, 6):
        h = torch.relu(h @ p[f"pair.L{l}.W"].T)
    u = torch.cat([h.sum(1), f], 1)
    for i in range(1, 6):
        u = u @ p[f"dec.L{i}.W"].T + p[f"dec.L{i}.b"]
        if i < 5:
            u = torch.relu(u)
    yt = torch.tensor(y, dtype=torch.float64)
    B = f.shape[0]
    obj = vlambda * ((u[:, 2:4] - yt[:, 2:4]) ** 2).sum() / (2 * B) + lam * ((u[:, 5] - yt[:, 5]) ** 2).sum() / B
    return obj, u


def test_manual_backward_matches_autograd(trained_like_params):
    _, batches = balls_batches(60, 5, seed=11)
    batch = batcher.make_batch(*batches[0], offset=7)
    g, pred = npe.bp(trained_like_params, batch, vlambda=0.7, lam=1.3)
    this_past, ctx, y, mask = npe.unpack_batch(batch)
    assert 0 < mask.mean() < 1
    tp = torch.tensor(trained_like_params, dtype=torch.float64, requires_grad=True)
    obj, u = _torch_objective(tp, this_past, ctx, y, mask, 0.7, 1.3)
    obj.backward()
    assert grad_err(g, tp.grad.numpy()) < 1e-6
    assert rel_err(pred[:, :6], u.detach().numpy()[:, :6]) < 1e-4


def test_reported_loss_is_not_the_backpropagated_objective(trained_like_params):
    """npe.lua:242 vs :306-307,314-315: loss = (Sv + Sw)/(3B); gradient is of Sv/(2B) + Sw/B."""
    _, batches = balls_batches(60, 4, seed=12)
    batch = batcher.make_batch(*batches[0], offset=1)
    l, pred, lv, lw = npe.fp(trained_like_params, batch)
    B = 50
    y = batch[2].reshape(B, -1)
    sv = ((pred[:, 2:4] - y[:, 2:4]) ** 2).sum(); sw = ((pred[:, 5] - y[:, 5]) ** 2).sum()
    assert np.isclose(l, (sv + sw) / (3 * B), rtol=1e-5)
    assert np.isclose(lv, sv / (2 * B), rtol=1e-5) and np.isclose(lw, sw / B, rtol=1e-5)


def test_masked_pair_equals_removed_pair(trained_like_params):
    _, batches = balls_batches(60, 6, seed=13)
    this, ctx, y, _, _, _, _ = batcher.make_batch(*batches[0], offset=2)
    B = this.shape[0]
    mask = npe.neighbor_mask(this, ctx)
    pred = npe.forward(trained_like_params, this.reshape(B, -1), ctx.reshape(B, 5, -1), mask)
    for b in range(0, B, 7):
        keep = mask[b] > 0
        pb = npe.forward(trained_like_params, this[b:b + 1].reshape(1, -1), ctx[b:b + 1, keep].reshape(1, int(keep.sum()), 44),
                         np.ones((1, int(keep.sum())), np.float32))
        assert np.allclose(pb, pred[b:b + 1], rtol=1e-5, atol=1e-6)


def test_context_permutation_invariance(trained_like_params):
    _, batches = balls_batches(60, 6, seed=14)
    this, ctx, y, *_ = batcher.make_batch(*batches[0], offset=0)
    perm = np.array([3, 0, 4, 1, 2])
    l1, p1, *_ = npe.fp(trained_like_params, (this, ctx, y))
    l2, p2, *_ = npe.fp(trained_like_params, (this, ctx[:, perm], y))
    assert np.allclose(p1, p2, rtol=1e-5, atol=1e-6)


def test_trim_is_identity_for_small_scenes_and_sorted_for_big():
    _, batches = balls_batches(60, 8, seed=15)
    foc, ctx = batches[0]
    this, tctx, y, tctx_f, mask12, orig, idx = batcher.make_batch(foc, ctx, offset=3)
    assert np.array_equal(tctx, orig[1]) and np.array_equal(idx, np.tile(np.arange(7), (50, 1)))   # data_sampler.lua:180-183
    assert mask12.shape == (12,) and mask12[6] == 1 and mask12.sum() == 1
    sw = scenes.raw_to_state(scenes.gen_walls(50, 4, seed=3, variant="U"))
    j = sw.shape[1] - 1
    idx, d = batcher.k_nearest_indices(sw[:, j, 0:2], np.delete(sw, j, 1)[:, :, 0:2])
    assert idx.shape == (50, 12) and np.all(np.diff(idx, axis=1) > 0)
    for b in range(50):   # kept set == 12 smallest distances with the lower-index tie rule
        order = sorted(range(d.shape[1]), key=lambda c: (d[b, c], c))[:12]
        assert sorted(order) == list(idx[b])


def test_relative_pair_roundtrip_and_window():
    _, batches = balls_batches(60, 3, seed=16)
    foc, ctx = batches[0]
    this, tctx, y, *_ = batcher.make_batch(foc, ctx, offset=19)
    assert np.array_equal(this, foc[:, 19:21]) and y.shape == (50, 1, 22)
    assert np.array_equal(y[:, 0, 6:], foc[:, 21, 6:])
    assert np.array_equal(y[:, 0, :6], foc[:, 21, :6] - foc[:, 20, :6])
    assert batcher.subbatch_id_to_file_offset(1) == (0, 0) and batcher.subbatch_id_to_file_offset(21) == (1, 0)
    assert batcher.subbatch_id_to_file_offset(40) == (1, 19)


def test_state_codec_roundtrip():
    raw = scenes.gen_walls(4, 5, seed=2, variant="I")
    st = scenes.raw_to_state(raw)
    assert st.shape[-1] == 22
    ball = st[0, -1, 0]
    assert list(ball[6:]) == [1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 1, 0, 1, 0, 1, 0]
    assert st[0, 0, 0, 9] == 1 and st[0, 0, 0, 11] == 1                              # obstacle: mass 1e30, oid obstacle
    back = scenes.state_to_raw(st)
    assert np.allclose(back[..., :4], raw[..., :4], atol=1e-3) and np.array_equal(back[..., 6:], raw[..., 6:])


def test_optimizers_against_torch():
    rng = np.random.default_rng(0)
    x0 = rng.standard_normal(300).astype(np.float32)
    xa = x0.copy(); sa = optim.adam_init(300)
    xt = torch.tensor(x0.astype(np.float64), requires_grad=True)
    topt = torch.optim.Adam([xt], lr=3e-4, betas=(0.9, 0.999), eps=1e-8)
    for t in range(20):
        g = rng.standard_normal(300).astype(np.float32)
        optim.adam_step(xa, g, sa, 3e-4)
        xt.grad = torch.tensor(g.astype(np.float64)); topt.step()
    # torch's eps placement differs from optim.adam's by O(eps); agreement to 1e-5 pins the update rule
    assert np.allclose(xa, xt.detach().numpy(), rtol=1e-5, atol=1e-6)
    xr = x0.copy(); sr = optim.rmsprop_init(300)
    xt = torch.tensor(x0.astype(np.float64), requires_grad=True)
    topt = torch.optim.RMSprop([xt], lr=3e-4, alpha=0.99, eps=1e-8)
    for t in range(20):
        g = rng.standard_normal(300).astype(np.float32)
        optim.rmsprop_step(xr, g, sr, 3e-4)
        xt.grad = torch.tensor(g.astype(np.float64)); topt.step()
    assert np.allclose(xr, xt.detach().numpy(), rtol=1e-5, atol=1e-6)
    assert optim.lr_schedule(3e-4, 50000) == 3e-4
    assert np.isclose(optim.lr_schedule(3e-4, 50001), 3e-4 * 0.99)
    assert np.isclose(optim.lr_schedule(3e-4, 55001), 3e-4 * 0.99 ** 3)


def test_rollout_literal_equals_scene_form(trained_like_params):
    s = scenes.raw_to_state(scenes.gen_balls(50, 3, 10, seed=6))
    ps, logs = rollout.simulate_all(trained_like_params, s[:, 0, :2], s[:, 1:, :2], s[:, 0, 2:], s[:, 1:, 2:], 8)
    tr, met = rollout.rollout_scenes(trained_like_params, s[:, :, :2], np.full(50, 3), 8, gt=s[:, :, 2:])
    assert np.array_equal(ps, tr)
    for k in logs:
        assert np.allclose(logs[k], met[k], rtol=1e-5), k
    # walls: obstacle foci are copied from ground truth, ball foci predicted
    sw = scenes.raw_to_state(scenes.gen_walls(50, 7, seed=1, variant="U"))
    N = sw.shape[1]; j = N - 1
    ps, logs = rollout.simulate_all(trained_like_params, sw[:, j, :2], np.delete(sw, j, 1)[:, :, :2], sw[:, j, 2:],
                                    np.delete(sw, j, 1)[:, :, 2:], 4)
    tr, met = rollout.rollout_scenes(trained_like_params, sw[:, :, :2], np.full(50, N), 4, gt=sw[:, :, 2:])
    perm = [j] + [i for i in range(N) if i != j]
    assert np.max(np.abs(ps - tr[:, perm])) < 1e-6
    assert np.allclose(logs["loss"], met["loss"], rtol=1e-4)


def test_rollout_position_uses_previous_velocity(trained_like_params):
    s = scenes.raw_to_state(scenes.gen_balls(5, 3, 6, seed=8))
    tr, _ = rollout.rollout_scenes(trained_like_params, s[:, :, :2], np.full(5, 3), 2)
    last = s[:, :, 1]
    exp = ((last[..., 0:2] * np.float32(800) + last[..., 2:4] * np.float32(60)) / np.float32(800)).astype(np.float32)
    assert np.array_equal(tr[:, :, 0, 0:2], exp)            # npe.lua:405-411
    assert np.all(tr[..., 4:6] == 0)                        # balls: eval.lua:176-178
    assert np.array_equal(tr[:, :, 0, 6:], last[..., 6:])   # eval.lua:167


# ---- golden fixtures ------------------------------------------------------------------------------------------
def _gold_params():
    p = npe.init_params(1234).copy()
    rng = np.random.default_rng(7)
    p += (0.05 * rng.standard_normal(p.shape)).astype(np.float32) * np.abs(p).mean()
    return p.astype(np.float32)


def test_golden_n4_b50():
    z = np.load(os.path.join(GOLD, "n4_b50.npz"))
    params = _gold_params()
    batch = (z["this"], z["context"], z["y"])
    l, pred, lv, lw = npe.fp(params, batch)
    assert np.allclose(pred, z["pred"], rtol=1e-5, atol=1e-6)
    assert np.allclose([l, lv, lw], z["loss3"], rtol=1e-5)
    assert np.array_equal(npe.neighbor_mask(z["this"], z["context"]).astype(np.uint8), z["nbr_mask"])
    g, _ = npe.bp(params, batch)
    assert np.allclose(g[z["grad_idx"]], z["grad_sel"], rtol=1e-4, atol=1e-7 * float(z["grad_absmax"]) * 100)
    x = params.copy(); sa = optim.adam_init(x.size); optim.adam_step(x, g, sa, 3e-4)
    assert np.allclose(x[z["grad_idx"]], z["adam_sel"], rtol=1e-6, atol=1e-7)


def test_golden_walls_trim_and_rollout():
    z = np.load(os.path.join(GOLD, "walls_u_trim.npz"))
    idx, _ = batcher.k_nearest_indices(z["this"], z["context_full"])
    assert np.array_equal(idx, z["closest_indices"])
    r = np.load(os.path.join(GOLD, "rollout_n3.npz"))
    tr, met = rollout.rollout_scenes(_gold_params(), r["states"][:, :, :2], np.full(8, 3), 10, gt=r["states"][:, :, 2:])
    assert np.allclose(tr, r["traj"], rtol=1e-4, atol=1e-5)
    assert np.allclose(met["loss"], r["m_loss"], rtol=1e-3)


# ---- property tests (hypothesis): the integer / bit outputs on lattices where ties are common -------------------------------
from hypothesis import given, settings, strategies as hst  # noqa: E402


@settings(max_examples=60, deadline=None)
@given(hst.integers(0, 2 ** 31 - 1), hst.integers(2, 30), hst.integers(1, 14))
def test_trim_and_mask_against_brute_force(seed, n_ctx, k_max):
    """k_nearest_indices == a per-focus Python sort by (float32 distance, index) on an 80-px lattice (equidistant contexts
    everywhere); the neighbourhood mask == d*800 <= 210 in float32 on the kept contexts."""
    rng = np.random.default_rng(seed)
    B = 6
    pos = (rng.integers(0, 10, (B, n_ctx + 1, 2)) * 80).astype(np.float32) / np.float32(800)
    st = np.zeros((B, n_ctx + 1, 2, 22), np.float32)
    st[:, :, :, 0:2] = pos[:, :, None, :]
    st[..., 10] = 1                                               # balls
    this, ctx = st[:, 0], st[:, 1:]
    idx, d = batcher.k_nearest_indices(this, ctx, k_max)
    k = min(k_max, n_ctx)
    for b in range(B):
        dist_b = [np.sqrt((ctx[b, c, -1, 0] - this[b, -1, 0]) ** 2 + (ctx[b, c, -1, 1] - this[b, -1, 1]) ** 2, dtype=np.float32)
                  for c in range(n_ctx)]
        order = sorted(range(n_ctx), key=lambda c: (dist_b[c], c))[:k]
        assert sorted(order) == idx[b].tolist()
        assert np.array_equal(np.asarray(dist_b, np.float32), d[b])
    tctx = np.take_along_axis(ctx, idx[:, :, None, None], axis=1)
    mask = npe.neighbor_mask(this, tctx, check_focus=False)
    want = (np.take_along_axis(d, idx, axis=1) * np.float32(800) <= np.float32(210))
    assert np.array_equal(mask.astype(bool), want)
