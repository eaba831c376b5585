"""Generates the committed golden fixtures from the CPU oracle (run from the repo root):
    python tests/golden/make_golden.py
The reference has no golden vectors and cannot run here (no Torch7); these freeze the oracle's
outputs so that oracle regressions and the CUDA path are both checked against fixed numbers.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import batcher, npe, optim, rollout, scenes  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def perturbed_params():
    p = npe.init_params(1234).copy()
    rng = np.random.default_rng(7)
    p += (0.05 * rng.standard_normal(p.shape)).astype(np.float32) * np.abs(p).mean()
    return p.astype(np.float32)


def main():
    params = perturbed_params()
    # --- n4, B = 50: one training example batch (config 1 shape) ---
    st = scenes.raw_to_state(scenes.gen_balls(50, 4, 12, seed=0))
    foc, ctx, _, _ = batcher.expand_for_each_object(st)
    this, tctx, y, _, mask12, _, idx = batcher.make_batch(foc[:50], ctx[:50], offset=4)
    l, pred, lv, lw = npe.fp(params, (this, tctx, y))
    le, _, lve, lwe = npe.fp_batch(params, (this, tctx, y))
    g, _ = npe.bp(params, (this, tctx, y))
    nmask = npe.neighbor_mask(this, tctx)
    x = params.copy(); sa = optim.adam_init(x.size); optim.adam_step(x, g, sa, 3e-4)
    xr = params.copy(); sr = optim.rmsprop_init(xr.size); optim.rmsprop_step(xr, g, sr, 3e-4)
    gsel = np.concatenate([np.arange(0, 28222, 97), [14699, 14700, 19400, 28200, 28221]])
    np.savez_compressed(os.path.join(OUT, "n4_b50.npz"), this=this, context=tctx, y=y, closest_indices=idx,
                        nbr_mask=nmask.astype(np.uint8), pred=pred, loss3=np.array([l, lv, lw], np.float32),
                        loss_per_example=np.stack([le, lve, lwe], 1), grad_idx=gsel, grad_sel=g[gsel],
                        grad_absmax=np.float32(np.abs(g).max()), grad_sum=np.float64(g.astype(np.float64).sum()),
                        adam_sel=x[gsel], rmsprop_sel=xr[gsel])
    # --- walls U: 12-NN trim fixture ---
    sw = scenes.raw_to_state(scenes.gen_walls(50, 4, seed=3, variant="U"))
    j = sw.shape[1] - 1
    wthis = sw[:, j, 0:2]; wctx = np.delete(sw, j, axis=1)[:, :, 0:2]
    widx, wd = batcher.k_nearest_indices(wthis, wctx)
    wt = np.take_along_axis(wctx, widx[:, :, None, None], axis=1)
    wmask = npe.neighbor_mask(wthis, wt, check_focus=False)
    wpred = npe.forward(params, wthis.reshape(50, -1), wt.reshape(50, 12, -1), wmask)
    np.savez_compressed(os.path.join(OUT, "walls_u_trim.npz"), this=wthis, context_xy=wctx[:, :, -1, 0:2].copy(),
                        context_full=wctx.astype(np.float32), closest_indices=widx, nbr_mask=wmask.astype(np.uint8),
                        pred=wpred)
    # --- rollout n3, 10 steps ---
    s3 = scenes.raw_to_state(scenes.gen_balls(8, 3, 12, seed=33))
    traj, met = rollout.rollout_scenes(params, s3[:, :, 0:2], np.full(8, 3), 10, gt=s3[:, :, 2:])
    np.savez_compressed(os.path.join(OUT, "rollout_n3.npz"), states=s3, traj=traj,
                        **{f"m_{k}": v for k, v in met.items()})
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
