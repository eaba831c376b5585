"""Trajectory batcher oracle: per-focus expansion, time window, relative targets, 12-NN context trim.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
  /root/reference/src/lua/data_process.lua:234-298  expand_for_each_object
  /root/reference/src/lua/data_sampler.lua:83-101   split_time
  /root/reference/src/lua/data_process.lua:87-96    relative_pair
  /root/reference/src/lua/data_process.lua:51-83    k_nearest_context / get_euc_dist
  /root/reference/src/lua/data_sampler.lua:153-203  load_subbatch_id_any_offset / load_subbatch_id
torch.topk's tie order is unspecified upstream; the rule fixed here (and in the CUDA kernel) is: smaller distance
first, then lower context index.
"""
import numpy as np

from . import config as C
from .npe import euc_dist

F32 = np.float32


def expand_for_each_object(scenes, focus_oid=C.OID_BALL):
    """scenes (S,N,T,22) -> focus (E,T,22), context (E,N-1,T,22), (scene_idx, obj_idx) int arrays.
    Focus-index-major concatenation: for i in objects: every scene where object i is of the focus type
    (data_process.lua:254-295)."""
    S, N, T, D = scenes.shape
    foc, ctx, sidx, oidx = [], [], [], []
    col = C.OID0 + focus_oid - 1
    for i in range(N):
        sel = np.nonzero(scenes[:, i, 0, col] == 1)[0]
        if sel.size == 0:
            continue
        foc.append(scenes[sel, i])
        ctx.append(np.delete(scenes[sel], i, axis=1))
        sidx.append(sel); oidx.append(np.full(sel.size, i))
    return (np.concatenate(foc, 0), np.concatenate(ctx, 0),
            np.concatenate(sidx).astype(np.int32), np.concatenate(oidx).astype(np.int32))


def split_time(focus, context, offset, sim=False, num_past=C.NUM_PAST, num_future=C.NUM_FUTURE):
    """data_sampler.lua:83-101; `offset` is 0-based here (Lua offset-1)."""
    a, b = offset, offset + num_past
    fp_, cp_ = focus[:, a:b], context[:, :, a:b]
    if sim:
        ff, cf = focus[:, b:], context[:, :, b:]
    else:
        ff, cf = focus[:, b:b + num_future], context[:, :, b:b + num_future]
    return fp_, cp_, ff, cf


def relative_pair(past, future, relative_to_absolute=False):
    """data_process.lua:87-96 (returns a new array)."""
    fut = future.astype(F32).copy()
    ref = past[:, -1:, 0:6].astype(F32)
    if relative_to_absolute:
        fut[:, :, 0:6] = (fut[:, :, 0:6] + ref).astype(F32)
    else:
        fut[:, :, 0:6] = (fut[:, :, 0:6] - ref).astype(F32)
    return fut


def k_nearest_indices(focus_past, context_past, k_max=C.MAX_CTX):
    """data_process.lua:51-63: distances on the last past frame; k smallest (ties: lower index); indices re-sorted
    ascending.  Returns (B,k) int64, and the (B,Cn) float32 distances."""
    d = euc_dist(focus_past[:, None, -1, 0:2], context_past[:, :, -1, 0:2])   # (B,Cn)
    Cn = d.shape[1]
    k = min(k_max, Cn) if k_max > 0 else Cn
    order = np.argsort(d, axis=1, kind="stable")[:, :k]      # stable: ties -> lower index first
    return np.sort(order, axis=1).astype(np.int64), d


def k_nearest_context(focus_past, context_past, k_max=C.MAX_CTX):
    idx, _ = k_nearest_indices(focus_past, context_past, k_max)
    return np.take_along_axis(context_past, idx[:, :, None, None], axis=1), idx


def make_batch(focus, context, offset, sim=False, relative=True, k_max=C.MAX_CTX):
    """load_subbatch_id_any_offset (data_sampler.lua:153-196) on an in-memory batch 'file' {focus (B,T,22),
    context (B,N-1,T,22)}.  Returns the reference 7-tuple
    (this, trimmed_context, y, trimmed_context_future, mask12, original_batch, closest_indices)."""
    this, ctx, y, ctx_f = split_time(focus, context, offset, sim)
    if relative and not sim:
        y = relative_pair(this, y, False)
    tctx, idx = k_nearest_context(this, ctx, k_max)
    tctx_f = np.take_along_axis(ctx_f, idx[:, :, None, None], axis=1)
    mask = np.zeros(max(C.MAX_CTX, tctx.shape[1]), F32)
    mask[tctx.shape[1] - 1] = 1
    return (this.astype(F32), tctx.astype(F32), y.astype(F32), tctx_f.astype(F32), mask,
            (this, ctx, y, ctx_f), idx)


def subbatch_id_to_file_offset(sid, maxwinsize=C.MAXWINSIZE, winsize=C.WINSIZE):
    """load_subbatch_id (data_sampler.lua:198-203); sid is 1-based like the Lua id; returns (file 0-based, offset
    0-based).  num_subbatches_per_batch = floor(60/3) = 20 consecutive offsets."""
    per = maxwinsize // winsize
    b = (sid - 1) // per
    return b, (sid - 1) - per * b


def cut_batches(focus, context, batch_size=C.BATCH_SIZE):
    """Consecutive batches of `batch_size` examples; the tail shorter than a batch is dropped here (the reference
    re-queues leftovers across json files, data_process.lua:500-527)."""
    nb = focus.shape[0] // batch_size
    return [(focus[i * batch_size:(i + 1) * batch_size], context[i * batch_size:(i + 1) * batch_size])
            for i in range(nb)]
