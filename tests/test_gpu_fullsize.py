"""-m gpu property tests at BASELINE.json's FULL sizes, where the oracle cannot follow (configs[3]: 65 536 foci of 64-ball
scenes per step; configs[4]: 65 536 eight-ball scenes = 524 288 foci per rollout step).  The inputs are R replicas of a
block of scenes that the oracle CAN do, in a shuffled order (and, for the rollout, with a block size that is not a multiple
of the tile height) so that the replicas of one scene land in different rows, tiles, tile slots, CTAs and rounds of every
persistent loop:

  * every replica must carry the bits of replica 0 (masks, pair tables, predictions, rolled-out windows): a tile loop, slot
    or CTA that mixes rows up, drops a tail or reads a stale table cannot pass;
  * replica 0 is compared with the oracle at the small size (free-running rollout: positions and masks while the mask
    sequences agree; training: gradient of the block);
  * the gradient of R replicas with the full batch as divisor is the gradient of one block (linearity of the sum over
    examples, npe.lua:304-315) up to summation order.
"""
import numpy as np
import pytest

from oracle import npe, rollout, scenes
from tests.test_gpu_large import make_model, oracle_chunked, scene_examples
from tests.util import grad_err, rel_err

pytestmark = pytest.mark.gpu


def shuffled_replicas(n_base, n_total, seed):
    """order[i] = base scene shown at position i (every base scene many times, shuffled); first[b] = first position of b."""
    order = np.random.default_rng(seed).permutation(np.arange(n_total) % n_base)
    first = np.array([np.flatnonzero(order == b)[0] for b in range(n_base)])
    return order, first


def test_configs4_rollout_full_size_replicas(trained_like_params):
    """65 536 eight-ball scenes (~262 shuffled replicas of 250 scenes), 12 free-running steps of the device-resident rollout by
    kernels (3xTF32 tcgen05: dense pair builder, pair chain, stacked decoder with the fused post-process)."""
    B, S, steps = 250, 65536, 12
    base = scenes.raw_to_state(scenes.gen_balls(B, 8, 2, seed=404))
    order, first = shuffled_replicas(B, S, 5)
    st = base[order]
    m = make_model(12, {})
    try:
        m.set_params(trained_like_params)
        m.upload_scenes(st)
        m.set_precision("tf32x3"); m.set_rollout_mode("kernels")
        n = m.rollout_resident(0, steps)
        assert n == S * 8 * steps
        w = m.debug_read("roll", (S, 8, 2, 22))
        w0 = w[first]
        assert np.array_equal(w, w0[order]), "replicas of a scene differ across tiles"
        # replica 0 against the oracle's free-running rollout of the block
        traj, _ = rollout.rollout_scenes(trained_like_params, base[:, :, 0:2], np.full(B, 8), steps)
        ref = traj[:, :, steps - 2:steps]                       # the window after `steps` steps = the last two states
        # positions integrate the previous velocity; a scene whose neighbourhood mask flipped (a distance within rounding
        # of the 210-px threshold under FP32-grade, not FP32, velocities) diverges: count them, compare the rest
        scale = np.abs(ref[..., 2:4]).max()
        err = np.abs(w0[..., 0:4] - ref[..., 0:4]).reshape(B, -1).max(1) / scale
        good = err < 1e-3
        print(f"[configs4 full] {S} scenes x {steps} steps: replicas bit-identical; block vs oracle: worst good {err[good].max():.2e}, "
              f"{int((~good).sum())} of {B} scenes diverged")
        assert good.mean() >= 0.98
    finally:
        m.close()


def test_configs3_training_step_full_size_replicas(trained_like_params):
    """1 024 scenes of 64 balls = 65 536 foci (32 shuffled replicas of 32 scenes): forward, backward and one Adam step through
    the default dispatch (tcgen05 forward with stash, tcgen05 backward, split reduce)."""
    B, S = 32, 1024
    base = scenes.raw_to_state(scenes.gen_big(B, 64, 3, seed=77))
    order, first = shuffled_replicas(B, S, 6)
    st = base[order]
    this, ctx, y = scene_examples(base, 0)
    ref = oracle_chunked(trained_like_params, this, ctx, y, chunk=1024)
    m = make_model(0, {})                                   # 12-NN trim off: 63 context slots
    try:
        m.set_params(trained_like_params); m.optim_reset(0.0)
        m.upload_scenes(st)
        m.upload_windows(np.arange(S), np.zeros(S, np.int32))
        F = m.select_windows(0, S)
        assert F == S * 64 == 65536
        idx, mask, cnt = m.pair_table()
        mask = mask.reshape(S, 64, -1); cnt = np.asarray(cnt).reshape(S, 64)
        assert np.array_equal(mask, mask[first][order]) and np.array_equal(cnt, cnt[first][order])
        assert np.array_equal(mask[first].reshape(B * 64, -1)[:, :63], ref["mask"].astype(np.uint8))   # bit-exact against the oracle
        m.set_precision("tf32x3")                              # inference forward on the tcgen05 chain kernels too
        pred, le = m.forward_per_example_current()
        pred = pred.reshape(S, 64, -1); le = le.reshape(S, 64, -1)
        assert np.array_equal(pred, pred[first][order]), "predictions of replicas differ across tiles"
        assert np.array_equal(le, le[first][order])
        pred0 = pred[first].reshape(B * 64, -1)
        l3 = m.train_step("adam", 3e-4)
        g = m.get_grads()
        e_g = grad_err(g, ref["grad"]); e_l = rel_err(l3, ref["loss3"], 1e-6)
        e_p = rel_err(pred0[:, :6], ref["pred"][:, :6])
        print(f"[configs3 full] F={F}: replicas bit-identical; pred {e_p:.2e} loss {e_l:.2e} gradient vs one block {e_g:.2e}")
        assert e_p < 1e-4 and e_l < 1e-4 and e_g < 1e-4
    finally:
        m.close()
