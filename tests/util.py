import numpy as np


def rel_err(a, ref, floor=1e-2):
    """Elementwise relative error max |a - ref| / max(|ref|, floor).  States are normalised (velocity 1.0 = 60 px per
    frame), so with the default floor 1e-2 the 1e-4 tolerance means: 1e-4 relative for |value| >= 0.01 and 1e-6
    absolute (6e-5 px/frame) below that -- two float32 evaluations of the 11-layer network with different summation
    orders already differ by ~2e-7 absolute, so a smaller floor would test rounding noise, not parity."""
    a = np.asarray(a, np.float64); ref = np.asarray(ref, np.float64)
    return float(np.max(np.abs(a - ref) / np.maximum(np.abs(ref), floor)))


def grad_err(g, ref):
    """infinity-norm error relative to ||ref||_inf."""
    g = np.asarray(g, np.float64); ref = np.asarray(ref, np.float64)
    return float(np.max(np.abs(g - ref)) / np.max(np.abs(ref)))


def balls_batches(n_scenes, n_balls, seed, T=60):
    from oracle import scenes, batcher
    st = scenes.raw_to_state(scenes.gen_balls(n_scenes, n_balls, T, seed=seed))
    foc, ctx, si, oi = batcher.expand_for_each_object(st)
    return st, batcher.cut_batches(foc, ctx)
