import numpy as np


def rel_err(a, ref, floor=1e-6):
    """max |a - ref| / max(|ref|, floor) elementwise (SURVEY section 8d tolerance definition)."""
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
