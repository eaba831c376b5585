"""NPE model oracle (float32 numpy): parameter layout, neighbourhood mask, forward, loss, manual backward.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows, op for op:
  /root/reference/src/lua/npe.lua:14-61      init_network
  /root/reference/src/lua/modules.lua:10-26  init_object_encoder
  /root/reference/src/lua/modules.lua:46-88  init_object_decoder_with_identity
  /root/reference/src/lua/npe.lua:120-222    unpack_batch / select_neighbors / apply_mask
  /root/reference/src/lua/npe.lua:225-336    fp / fp_batch / bp
  /root/reference/src/lua/data_utils.lua:276-285  compute_euc_dist
Upstream-rock semantics (nn.Linear (out,in) weights, Threshold ReLU, MSECriterion(false), CAddTable backward =
copy, getParameters flattening order) are restated from the rocks' published behaviour.
"""
import numpy as np

from . import config as C

F32 = np.float32


# ----------------------------------------------------------------------------------------------------------------
# flat parameter layout  (nn.Module:getParameters order: weight then bias per Linear,
# modules in container order: encoder {this, context}, 5 pairwise layers, decoder 5 Linears)
# ----------------------------------------------------------------------------------------------------------------
def param_layout(rnn_dim=C.RNN_DIM, layers=C.LAYERS, input_dim=C.INPUT_DIM, out_dim=C.OUT_DIM, nbrhd=True):
    """Returns list of (name, offset, shape). With -nbrhd the encoder and pairwise Linears have no bias
    (npe.lua:17,21,28); decoder Linears are always biased (modules.lua:59-69)."""
    H, h2 = rnn_dim, rnn_dim // 2
    ents = []
    off = 0

    def add(name, shape):
        nonlocal off
        n = int(np.prod(shape))
        ents.append((name, off, tuple(shape)))
        off += n

    add("enc.this.W", (h2, input_dim))
    if not nbrhd:
        add("enc.this.b", (h2,))
    add("enc.ctx.W", (h2, input_dim))
    if not nbrhd:
        add("enc.ctx.b", (h2,))
    for l in range(1, layers + 1):
        add(f"pair.L{l}.W", (H, H))
        if not nbrhd:
            add(f"pair.L{l}.b", (H,))
    dec_in = H + input_dim  # JoinTable({rnn_out, orig_state})  modules.lua:54-55
    if layers <= 1:
        add("dec.L1.W", (out_dim, dec_in))
        add("dec.L1.b", (out_dim,))
    else:
        for i in range(1, layers + 1):
            if i == 1:
                add("dec.L1.W", (H, dec_in)); add("dec.L1.b", (H,))
            elif i == layers:
                add(f"dec.L{i}.W", (out_dim, H)); add(f"dec.L{i}.b", (out_dim,))
            else:
                add(f"dec.L{i}.W", (H, H)); add(f"dec.L{i}.b", (H,))
    return ents, off


def num_params(**kw):
    return param_layout(**kw)[1]


def views(flat, **kw):
    """dict name -> ndarray view into the flat vector."""
    ents, n = param_layout(**kw)
    assert flat.shape == (n,), (flat.shape, n)
    return {name: flat[off:off + int(np.prod(shape))].reshape(shape) for name, off, shape in ents}


def init_params(seed=1234, **kw):
    """nn.Linear:reset(): U(-1/sqrt(in), 1/sqrt(in)) for weight and bias (upstream nn rock).  The `layers`
    pairwise Linears are layer:clone() of ONE Linear (npe.lua:21,39): identical initial values."""
    rng = np.random.default_rng(seed)
    ents, n = param_layout(**kw)
    flat = np.zeros(n, F32)
    first_pair = None
    for name, off, shape in ents:
        fan_in = shape[1] if len(shape) == 2 else None
        if len(shape) == 1:  # bias follows its weight: fan_in of that Linear
            fan_in = prev_in
        else:
            prev_in = shape[1]
        stdv = 1.0 / np.sqrt(fan_in)
        vals = rng.uniform(-stdv, stdv, size=shape).astype(F32)
        if name.startswith("pair.L") and name.endswith(".W"):
            if first_pair is None:
                first_pair = vals
            vals = first_pair
        flat[off:off + vals.size] = vals.ravel()
    return flat


# ----------------------------------------------------------------------------------------------------------------
# distance + neighbourhood mask
# ----------------------------------------------------------------------------------------------------------------
def euc_dist(focus_xy, ctx_xy):
    """compute_euc_dist (data_utils.lua:276-285): diff = b - a; pow(diff,2); sqrt(dx2 + dy2); every op one float32
    rounding.  focus_xy (..., 2), ctx_xy (..., 2) broadcastable."""
    diff = (ctx_xy.astype(F32) - focus_xy.astype(F32)).astype(F32)
    sq = (diff * diff).astype(F32)
    return np.sqrt((sq[..., 0] + sq[..., 1]).astype(F32)).astype(F32)


def neighbor_mask(this_past, context, nbrhdsize=C.NBRHDSIZE, check_focus=True):
    """select_neighbors (npe.lua:170-206).  this_past (B,2,22), context (B,Cn,2,22) -> mask (B,Cn) float32 {0,1}.
    The look-ahead position is computed and then overwritten by the current position (npe.lua:189-193), so the
    distance uses the LAST past frame's positions.  Threshold = nbrhdsize * base_size(ball|block) in pixels
    (npe.lua:180-186), compared after d*position_normalize_constant (npe.lua:197-201)."""
    B = this_past.shape[0]
    oid = this_past[:, -1, C.OID0:C.OID1]
    if check_focus:
        tb = np.zeros_like(oid); tb[:, C.OID_BALL - 1] = 1
        tk = np.zeros_like(oid); tk[:, C.OID_BLOCK - 1] = 1
        if np.array_equal(oid, tb):
            thr = nbrhdsize * C.BASE_SIZE["ball"]
        elif np.array_equal(oid, tk):
            thr = nbrhdsize * C.BASE_SIZE["block"]
        else:
            raise AssertionError("Unknown object id")  # npe.lua:185
    else:
        thr = nbrhdsize * C.BASE_SIZE["ball"]
    d = euc_dist(this_past[:, None, -1, 0:2], context[:, :, -1, 0:2])  # (B,Cn)
    d_px = (d * F32(C.PNC)).astype(F32)
    return (d_px <= F32(thr)).astype(F32)


# ----------------------------------------------------------------------------------------------------------------
# forward
# ----------------------------------------------------------------------------------------------------------------
def relu(x):
    return np.maximum(x, F32(0))


def sigmoid(x):
    # THNN Sigmoid: 1/(1+exp(-x)) evaluated in double by the C library then stored as float
    return (1.0 / (1.0 + np.exp(-x.astype(np.float64)))).astype(F32)


def forward(params, this_past, context, mask, layers=C.LAYERS, keep=False):
    """network:forward (npe.lua:231; topology npe.lua:36-59, modules.lua:10-26,46-88) in the -nbrhd configuration.
    this_past (B,44), context (B,Cn,44), mask (B,Cn) -> pred (B,22).
    apply_mask multiplies BOTH inputs of a pair by its mask (npe.lua:209-222); with no bias a masked pair yields
    exactly zero through every layer."""
    p = views(params, layers=layers)
    B, Cn, _ = context.shape
    m = mask.astype(F32)[:, :, None]
    f_in = (this_past[:, None, :] * m).astype(F32)          # (B,Cn,44)
    c_in = (context * m).astype(F32)
    e_t = relu(f_in @ p["enc.this.W"].T)
    e_c = relu(c_in @ p["enc.ctx.W"].T)
    h = np.concatenate([e_t, e_c], axis=2).astype(F32)      # JoinTable(2)  modules.lua:23
    hs = [h]
    for l in range(1, layers + 1):
        h = relu(h @ p[f"pair.L{l}.W"].T).astype(F32)
        hs.append(h)
    s = h.sum(axis=1, dtype=F32)                             # CAddTable (npe.lua:55)
    z = np.concatenate([s, this_past], axis=1).astype(F32)  # JoinTable({rnn_out, orig_state}) modules.lua:55
    us = [z]
    u = z
    for i in range(1, layers + 1):
        u = (u @ p[f"dec.L{i}.W"].T + p[f"dec.L{i}.b"]).astype(F32)
        if i < layers:
            u = relu(u)
        us.append(u)
    o = u
    pred = o.copy()
    pred[:, 6:] = sigmoid(o[:, 6:])                          # modules.lua:80-86
    if keep:
        return pred, dict(f_in=f_in, c_in=c_in, hs=hs, us=us, m=m)
    return pred


def losses(pred, y):
    """model:fp loss (npe.lua:233-246): MSECriterion(false) sums; loss = (sum_vel + sum_angvel)/(3B)."""
    B = pred.shape[0]
    ev = (pred[:, 2:4] - y[:, 2:4]).astype(F32)
    ew = (pred[:, 5] - y[:, 5]).astype(F32)
    lv = F32((ev.astype(np.float64) ** 2).sum())
    lw = F32((ew.astype(np.float64) ** 2).sum())
    return F32((lv + lw) / (3 * B)), F32(lv / (2 * B)), F32(lw / B)


def losses_per_example(pred, y):
    """model:fp_batch (npe.lua:268-280)."""
    ev = (pred[:, 2:4] - y[:, 2:4]).astype(F32)
    ew = (pred[:, 5] - y[:, 5]).astype(F32)
    lv = (ev * ev).sum(axis=1, dtype=F32)
    lw = (ew * ew).astype(F32)
    return ((lv + lw) / F32(3)).astype(F32), (lv / F32(2)).astype(F32), lw


def unpack_batch(batch, nbrhd=True, nbrhdsize=C.NBRHDSIZE):
    """model:unpack_batch (npe.lua:120-166): flatten time, build masks."""
    this, context, y = batch[0], batch[1], batch[2]
    B = this.shape[0]
    this_past = this.reshape(B, -1).astype(F32)
    ctx = context.reshape(B, context.shape[1], -1).astype(F32)
    yy = y.reshape(B, -1).astype(F32)
    if nbrhd:
        mask = neighbor_mask(this, context, nbrhdsize)
    else:
        mask = np.ones((B, context.shape[1]), F32)
    return this_past, ctx, yy, mask


def fp(params, batch, **kw):
    """model:fp (npe.lua:225-247) -> loss, pred, vel_loss, angvel_loss"""
    this_past, ctx, y, mask = unpack_batch(batch, **kw)
    pred = forward(params, this_past, ctx, mask)
    l, lv, lw = losses(pred, y)
    return l, pred, lv, lw


def fp_batch(params, batch, **kw):
    """model:fp_batch (npe.lua:250-284) -> loss(B), pred, vel_loss(B), angvel_loss(B)"""
    this_past, ctx, y, mask = unpack_batch(batch, **kw)
    pred = forward(params, this_past, ctx, mask)
    l, lv, lw = losses_per_example(pred, y)
    return l, pred, lv, lw


def d_pred(pred, y, vlambda=1.0, lam=1.0, divisor_batch=None):
    """Gradient wrt the prediction that the reference back-propagates (npe.lua:301-320): MSECriterion(false) grad
    2(p - y), times vlambda (lambda), divided by nElement (2B for vel, B for ang vel); zero on every other column.
    `divisor_batch` overrides B (data-parallel: the GLOBAL batch)."""
    B = pred.shape[0] if divisor_batch is None else divisor_batch
    g = np.zeros_like(pred, dtype=F32)
    g[:, 2:4] = (F32(2) * (pred[:, 2:4] - y[:, 2:4]) * F32(vlambda) / F32(2 * B)).astype(F32)
    g[:, 5] = (F32(2) * (pred[:, 5] - y[:, 5]) * F32(lam) / F32(B)).astype(F32)
    return g


def bp(params, batch, vlambda=1.0, lam=1.0, layers=C.LAYERS, divisor_batch=None, inter=None, **kw):
    """model:bp (npe.lua:290-336): manual chain rule -> grad_params (flat).  Float32 storage, float64 accumulation
    inside each matrix product (numpy matmul on float32 uses float32 BLAS; we promote the reductions that sum over
    the batch to keep the oracle the more accurate side of the comparison)."""
    this_past, ctx, y, mask = unpack_batch(batch, **kw)
    pred, st = forward(params, this_past, ctx, mask, layers=layers, keep=True)
    p = views(params, layers=layers)
    g = np.zeros_like(params, dtype=F32)
    gv = views(g, layers=layers)
    dp = d_pred(pred, y, vlambda, lam, divisor_batch).astype(np.float64)
    # sigmoid columns receive zero gradient (IdentityCriterion); cols 0..5 are linear
    du = dp.copy()
    du[:, 6:] = 0.0
    us = st["us"]
    if inter is not None:   # intermediate gradients for kernel-by-kernel parity checks: dz of every layer, ds
        inter.update(us=us, hs=st["hs"], m=st["m"], dz_dec={}, dz_pair={})
    for i in range(layers, 0, -1):
        u_in = us[i - 1].astype(np.float64)
        if i < layers:
            du = du * (us[i] > 0)
        if inter is not None:
            inter["dz_dec"][i] = du.astype(F32)
        gv[f"dec.L{i}.W"][...] = (du.T @ u_in).astype(F32)
        gv[f"dec.L{i}.b"][...] = du.sum(axis=0).astype(F32)
        du = du @ p[f"dec.L{i}.W"].astype(np.float64)
    H = p["pair.L1.W"].shape[0]
    ds = du[:, :H]                                   # gradient into the pairwise sum
    B, Cn, _ = ctx.shape
    dh = np.broadcast_to(ds[:, None, :], (B, Cn, H)) * st["m"]   # CAddTable bwd = copy; apply_mask (npe.lua:326-327)
    hs = st["hs"]
    if inter is not None:
        inter["ds"] = ds.astype(F32)
    for l in range(layers, 0, -1):
        dh = dh * (hs[l] > 0)
        if inter is not None:
            inter["dz_pair"][l] = dh.astype(F32)
        gv[f"pair.L{l}.W"][...] = np.einsum("bcj,bck->jk", dh, hs[l - 1].astype(np.float64)).astype(F32)
        dh = dh @ p[f"pair.L{l}.W"].astype(np.float64)
    dh = dh * (hs[0] > 0)
    if inter is not None:
        inter["dz_pair"][0] = dh.astype(F32)
    h2 = H // 2
    gv["enc.this.W"][...] = np.einsum("bcj,bck->jk", dh[:, :, :h2], st["f_in"].astype(np.float64)).astype(F32)
    gv["enc.ctx.W"][...] = np.einsum("bcj,bck->jk", dh[:, :, h2:], st["c_in"].astype(np.float64)).astype(F32)
    return g, pred
