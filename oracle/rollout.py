"""Autoregressive rollout oracle.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
  /root/reference/src/lua/eval.lua:122-153   simulate_all_preprocess
  /root/reference/src/lua/eval.lua:155-182   simulate_all_postprocess
  /root/reference/src/lua/eval.lua:185-234   make_invalid_dummy / replace_invalid_dummy / find_valid_focus_mask
  /root/reference/src/lua/eval.lua:237-454   simulate_all
  /root/reference/src/lua/npe.lua:386-458    update_position / update_angle
  /root/reference/src/lua/infer.lua:9-74     relative_error / angle_magnitude
Two forms: `simulate_all` is the literal batch form (every example carries its whole scene and every object is made
the focus in turn); `rollout_scenes` simulates each scene once with identical per-object arithmetic -- it is what
the CUDA rollout kernel is compared against.  They agree whenever each focus slot of a batch is type-homogeneous
(balls and walls datasets), which tests/ checks.
"""
import numpy as np

from . import config as C
from . import npe
from .batcher import k_nearest_context, relative_pair

F32 = np.float32
PI32 = F32(np.pi)


def update_position(this, pred):
    """npe.lua:386-419 with num_future=1: pos_new = (pos_last*pnc + vel_last*vnc)/pnc -- the PREVIOUS velocity."""
    pred = pred.copy()
    lastpos = (this[:, -1, 0:2] * F32(C.PNC)).astype(F32)
    lastvel = (this[:, -1, 2:4] * F32(C.VNC)).astype(F32)
    pos = (lastpos + lastvel).astype(F32)
    pred[:, 0, 0:2] = (pos / F32(C.PNC)).astype(F32)
    return pred


def update_angle(this, pred):
    """npe.lua:422-458 with num_future=1."""
    pred = pred.copy()
    la = (this[:, -1, 4] * F32(C.ANC)).astype(F32)
    lav = (this[:, -1, 5] * F32(C.ANC)).astype(F32)
    ang = (la + lav).astype(F32)
    gt = ang > PI32
    le = ang <= -PI32
    ang = (ang + F32(-2 * np.pi) * gt.astype(F32)).astype(F32)
    ang = (ang + F32(2 * np.pi) * le.astype(F32)).astype(F32)
    pred[:, 0, 4] = (ang / F32(C.ANC)).astype(F32)
    return pred


def postprocess(pred_flat, this, batchwide_ball_test=True):
    """simulate_all_postprocess (eval.lua:155-182).  pred_flat (B,22) relative -> (B,1,22) absolute state."""
    B = pred_flat.shape[0]
    pred = pred_flat.reshape(B, 1, C.OBJ_DIM).astype(F32)
    pred = relative_pair(this, pred, True)                      # eval.lua:163
    pred[:, :, C.OSSI:] = this[:, -1:, C.OSSI:]                 # eval.lua:167
    pred = update_position(this, pred)
    pred = update_angle(this, pred)
    is_ball = pred[:, 0, C.OID0] == 1
    if batchwide_ball_test:
        if is_ball.all():                                       # eval.lua:176-178
            pred[:, :, 4:6] = 0
    else:
        pred[is_ball, :, 4:6] = 0
    return pred


def angle_magnitude(pred_flat, this, y_rel):
    """infer.lua:32-74 with within_batch=true -> (cos (B), rel_mag_err (B), mask (B))."""
    B = pred_flat.shape[0]
    pred = relative_pair(this, pred_flat.reshape(B, 1, -1), True)
    fut = relative_pair(this, y_rel.reshape(B, 1, -1), True)
    pv = (pred[:, 0, 2:4] * F32(C.VNC)).astype(F32)
    gv = (fut[:, 0, 2:4] * F32(C.VNC)).astype(F32)
    pm = np.sqrt((pv * pv).sum(1, dtype=F32)).astype(F32)
    gm = np.sqrt((gv * gv).sum(1, dtype=F32)).astype(F32)
    mask = gm != 0
    gm_safe = np.where(mask, gm, F32(1)).astype(F32)
    rel = np.abs(F32(1) - pm / gm_safe).astype(F32)
    rel = np.where(mask, rel, F32(0)).astype(F32)
    with np.errstate(divide="ignore", invalid="ignore"):
        cos = ((pv * gv).sum(1, dtype=F32) / (pm * gm).astype(F32)).astype(F32)
    cos = np.where(mask, cos, F32(0)).astype(F32)
    return cos, rel, mask


def _masked_avg(x, mask):
    n = mask.sum()
    return (x * mask).sum() / n if n > 0 else np.nan


def simulate_all(params, this_orig, context_orig, y_orig, context_future_orig, numsteps, nbrhdsize=C.NBRHDSIZE):
    """Literal batch form of simulate_all for ONE batch (eval.lua:263-393).
    this_orig (B,2,22), context_orig (B,N-1,2,22) untrimmed, y_orig (B,>=numsteps,22) absolute gt,
    context_future_orig (B,N-1,>=numsteps,22).  Returns pred_sim (B,N,numsteps,22) and a dict of per-step logs."""
    B = this_orig.shape[0]
    past = np.concatenate([this_orig[:, None], context_orig], 1).astype(F32)          # eval.lua:283
    future = np.concatenate([y_orig[:, None, :numsteps], context_future_orig[:, :, :numsteps]], 1).astype(F32)
    N = past.shape[1]
    pred_sim = np.zeros((B, N, numsteps, C.OBJ_DIM), F32)
    logs = {k: np.zeros(numsteps) for k in ("loss", "vel", "angvel", "cos", "mag")}
    for t in range(numsteps):
        acc = dict(loss=0.0, vel=0.0, angvel=0.0, cos=0.0, mag=0.0)
        counter = 0.0; am_counter = 0.0
        for j in range(N):
            this = past[:, j]
            y_abs = future[:, j, t:t + 1]
            y = relative_pair(this, y_abs, False)                                     # eval.lua:131-133
            context = np.delete(past, j, axis=1)
            tctx, _ = k_nearest_context(this, context)                                # eval.lua:149
            obstacle = np.zeros(3, F32); obstacle[C.OID_OBSTACLE - 1] = 1
            invalid = (this[:, -1, C.OID0:C.OID1] == obstacle).all(axis=1)            # eval.lua:226-234
            valid = ~invalid
            if invalid.sum() < B:
                this_in = this
                if invalid.any():                                                      # make_invalid_dummy
                    this_in = this.copy()
                    this_in[invalid] = this[np.nonzero(valid)[0][0]]
                l, pred, lv, lw = npe.fp_batch(params, (this_in, tctx, y), nbrhdsize=nbrhdsize)
                cos, mag, amask = angle_magnitude(pred, this_in, y[:, 0])
                vam = valid & amask
                acc["loss"] += _masked_avg(l, valid); acc["vel"] += _masked_avg(lv, valid)
                acc["angvel"] += _masked_avg(lw, valid)
                acc["cos"] += _masked_avg(cos, vam); acc["mag"] += _masked_avg(mag, vam)
                counter += valid.sum() / B; am_counter += vam.sum() / B
                out = postprocess(pred, this_in)
                if invalid.any():
                    out[invalid] = y_abs[invalid]                                      # replace_invalid_dummy
                pred_sim[:, j, t] = out[:, 0]
            else:
                assert np.array_equal(this[:, -1], y_abs[:, 0])                        # eval.lua:373
                pred_sim[:, j, t] = y_abs[:, 0]
        past = np.concatenate([past[:, :, 1:], pred_sim[:, :, t:t + 1]], 2)           # eval.lua:380
        for k in ("loss", "vel", "angvel"):
            logs[k][t] = acc[k] / counter if counter > 0 else np.nan
        logs["cos"][t] = acc["cos"] / am_counter if am_counter > 0 else np.nan
        logs["mag"][t] = acc["mag"] / am_counter if am_counter > 0 else np.nan
    return pred_sim, logs


def rollout_scenes(params, past0, n_obj, numsteps, gt=None, nbrhdsize=C.NBRHDSIZE, k_max=C.MAX_CTX,
                   teacher=None, return_masks=False):
    """Scene form.  past0 (S,Nmax,2,22), n_obj (S,) -> traj (S,Nmax,numsteps,22).
    gt (S,Nmax,numsteps,22) absolute ground truth: required when obstacles are present (they are copied from gt,
    eval.lua:362-375) and for the per-step metrics.  `teacher`: optional (S,Nmax,numsteps,22) states fed back
    instead of the predictions (teacher forcing) -- the returned traj still holds the one-step predictions.
    Scenes are grouped by n_obj and processed as batches; the ball a/av zeroing is per object."""
    S, Nmax = past0.shape[:2]
    traj = np.zeros((S, Nmax, numsteps, C.OBJ_DIM), F32)
    metrics = {k: np.zeros(numsteps) for k in ("loss", "vel", "angvel", "cos", "mag")}
    cnt = np.zeros(numsteps); am_cnt = np.zeros(numsteps)
    all_masks = []
    for N in np.unique(n_obj):
        sel = np.nonzero(n_obj == N)[0]
        past = past0[sel, :N].astype(F32).copy()
        B = sel.size
        for t in range(numsteps):
            new = np.zeros((B, N, C.OBJ_DIM), F32)
            for j in range(N):
                this = past[:, j]
                is_obst = this[:, -1, C.OID0 + C.OID_OBSTACLE - 1] == 1
                if gt is not None:
                    y_abs = gt[sel, j, t:t + 1].astype(F32)
                else:
                    y_abs = this[:, -1:].copy()
                y = relative_pair(this, y_abs, False)
                context = np.delete(past, j, axis=1)
                tctx, idx = k_nearest_context(this, context, k_max)
                mask = npe.neighbor_mask(this, tctx, nbrhdsize, check_focus=False)
                pred = npe.forward(params, this.reshape(B, -1), tctx.reshape(B, tctx.shape[1], -1), mask)
                if return_masks:
                    all_masks.append((int(N), t, j, idx.copy(), mask.copy()))
                out = postprocess(pred, this, batchwide_ball_test=False)[:, 0]
                out[is_obst] = y_abs[is_obst, 0]
                new[:, j] = out
                v = ~is_obst
                if gt is not None and v.any():
                    l, lv, lw = npe.losses_per_example(pred, y[:, 0])
                    cos, mag, amask = angle_magnitude(pred, this, y[:, 0])
                    metrics["loss"][t] += l[v].sum(dtype=np.float64); metrics["vel"][t] += lv[v].sum(dtype=np.float64)
                    metrics["angvel"][t] += lw[v].sum(dtype=np.float64); cnt[t] += v.sum()
                    va = v & amask
                    metrics["cos"][t] += cos[va].sum(dtype=np.float64); metrics["mag"][t] += mag[va].sum(dtype=np.float64)
                    am_cnt[t] += va.sum()
            traj[sel, :N, t] = new
            nxt = new if teacher is None else teacher[sel, :N, t].astype(F32)
            past = np.concatenate([past[:, :, 1:], nxt[:, :, None]], 2)
    with np.errstate(divide="ignore", invalid="ignore"):
        for k in ("loss", "vel", "angvel"):
            metrics[k] = metrics[k] / cnt
        metrics["cos"] = metrics["cos"] / am_cnt
        metrics["mag"] = metrics["mag"] / am_cnt
    if return_masks:
        return traj, metrics, all_masks
    return traj, metrics
