"""Steps either side of the hot path: matter-js json <-> scene tensor <-> batch files.

f1  json -> (ex, obj, t, 12) raw tensor -> normalised one-hot states (ex, obj, t, 22) -> scene tensor for the device
    batcher, or the reference's per-focus batch files.  Follows json_interface.lua:26-63 (load_data_json),
    data_process.lua:119-139 (normalize), :184-207 (properties2onehotall), :234-298 (expand_for_each_object),
    :311-330,:543-550 (batches of 50).
f2  predicted / ground-truth scene tensors -> json for render.html.  Follows data_process.lua:556-562
    (record_trajectories: condense -> onehot2propertiesall -> unnormalize -> data2table, json_interface.lua:67-94).
Host-side data plumbing (the reference does it in Lua once per dataset); no part of the timed hot path.
"""
import json
import math
import os
import re

import numpy as np

PNC, VNC, ANC = 800.0, 60.0, math.pi
MASSES = (1.0, 5.0, 25.0, 1e30)
OIDS = {"ball": 1, "obstacle": 2, "block": 3}
ROIDS = {v: k for k, v in OIDS.items()}
SIZES = (0.5, 1.0, 2.0)
F32 = np.float32


def global_params_from_name(jsonfile):
    """get_global_params (json_interface.lua:6-23): a flag is on when its marker occurs exactly once in the path
    ('_g' gravity -- also forced on for tower datasets --, '_f' friction, '_p' pairwise)."""
    s = str(jsonfile)
    g = 1 if s.count("_g") == 1 or "tower" in s else 0
    f = 1 if s.count("_f") == 1 else 0
    p = 1 if s.count("_p") == 1 else 0
    return g, f, p


def load_data_json(jsonfile_or_obj, gfp=(0, 0, 0)):
    """matter-js dump {trajectories[ex][obj][t]{position{x,y}, velocity{x,y}, mass, angle, angularVelocity, objtype,
    sizemul}} -> raw (ex, obj, t, 12) float32 in rsi order px,py,vx,vy,a,av,m,oid,os,g,f,p."""
    data = jsonfile_or_obj
    if isinstance(jsonfile_or_obj, (str, os.PathLike)):
        gfp = global_params_from_name(jsonfile_or_obj)
        with open(jsonfile_or_obj) as fh:
            data = json.load(fh)
    tr = data["trajectories"]
    E, N, T = len(tr), len(tr[0]), len(tr[0][0])
    # one flat pass over the records and ONE array conversion (the nested assignment loop it replaces spent its time in
    # per-record numpy row writes); ragged input (an object with fewer frames) fails the reshape loudly
    g, f, p = (float(v) for v in gfp)
    flat = [(s["position"]["x"], s["position"]["y"], s["velocity"]["x"], s["velocity"]["y"], s["angle"], s["angularVelocity"],
             s["mass"], OIDS[s["objtype"]], s["sizemul"], g, f, p) for ex in tr for ob in ex for s in ob]
    return np.asarray(flat, np.float64).astype(F32).reshape(E, N, T, 12)


def raw_to_states(raw, sizes=SIZES):
    """normalize + properties2onehotall: (.., 12) -> (.., 22)."""
    raw = np.asarray(raw, F32)
    out = np.zeros(raw.shape[:-1] + (22,), F32)
    out[..., 0:2] = raw[..., 0:2] / F32(PNC)
    out[..., 2:4] = raw[..., 2:4] / F32(VNC)
    ang = raw[..., 4:6].copy()
    ang = (ang + F32(-2 * math.pi) * (ang > F32(math.pi)).astype(F32)).astype(F32)      # wrap_pi
    out[..., 4:6] = ang / F32(ANC)
    for col, cats, lo in ((6, MASSES, 6), (7, (1, 2, 3), 10), (8, sizes, 13), (9, (0, 1), 16), (10, (0, 1), 18), (11, (0, 1), 20)):
        hit = np.zeros(raw.shape[:-1], bool)
        for k, c in enumerate(cats):
            m = raw[..., col] == F32(c)
            out[..., lo + k] = m
            hit |= m
        if not hit.all():
            raise ValueError(f"raw column {col} holds a value outside {cats}")
    return out


def states_to_raw(states, sizes=SIZES):
    """onehot2propertiesall + unnormalize (data_process.lua:143-162,209-232)."""
    st = np.asarray(states, F32)
    raw = np.zeros(st.shape[:-1] + (12,), F32)
    raw[..., 0:2] = st[..., 0:2] * F32(PNC)
    raw[..., 2:4] = st[..., 2:4] * F32(VNC)
    ang = (st[..., 4:6] * F32(ANC)).astype(F32)
    a = ang[..., 0]
    ang[..., 0] = (a + F32(2 * math.pi) * (a < 0).astype(F32)).astype(F32)                # wrap_2pi on the angle only
    raw[..., 4:6] = ang
    raw[..., 6] = np.asarray(MASSES, F32)[st[..., 6:10].argmax(-1)]
    raw[..., 7] = st[..., 10:13].argmax(-1) + 1
    raw[..., 8] = np.asarray(sizes, F32)[st[..., 13:16].argmax(-1)]
    raw[..., 9] = st[..., 16:18].argmax(-1)
    raw[..., 10] = st[..., 18:20].argmax(-1)
    raw[..., 11] = st[..., 20:22].argmax(-1)
    return raw


def data2table(raw):
    """json_interface.lua:67-94."""
    E, N, T, _ = raw.shape
    return [[[{"position": {"x": float(s[0]), "y": float(s[1])}, "velocity": {"x": float(s[2]), "y": float(s[3])},
               "angle": float(s[4]), "angularVelocity": float(s[5]), "mass": float(s[6]), "objtype": ROIDS[int(s[7])],
               "sizemul": float(s[8])} for s in raw[e, i]] for i in range(N)] for e in range(E)]


def record_trajectories(states, jsonfile, config=None):
    """data_process:record_trajectories (data_process.lua:556-562) for a scene tensor (ex, obj, t, 22): writes the
    gt_*.json / pred_*.json that src/js/demo/render.html consumes (eval.lua:468-497)."""
    obj = {"trajectories": data2table(states_to_raw(states)), "config": config or {}}
    with open(jsonfile, "w") as fh:
        json.dump(obj, fh)
    return obj


def dataset_flags(folder):
    """Dataset properties parsed out of the folder name (data_process.lua:451-468): balls_n4_t60_ex200_m, walls_n2_..._wU."""
    name = os.path.basename(os.path.normpath(str(folder)).replace("/jsons", ""))
    out = {"env": name.split("_")[0]}
    for key, pat in (("n", r"_n(\d+)"), ("t", r"_t(\d+)"), ("ex", r"_ex(\d+)")):
        m = re.search(pat, name)
        if m:
            out[key] = int(m.group(1))
    m = re.search(r"_w([OLUI])", name)
    if m:
        out["walls"] = m.group(1)
        out["num_obj"] = {"O": 30, "L": 30, "U": 33, "I": 32}[m.group(1)]
    else:
        out["num_obj"] = out.get("n")
    return out


def expand_for_each_object(states, focus_oid=1):
    """data_process.lua:234-298: focus-index-major list of (focus (E,T,22), context (E,N-1,T,22))."""
    foc, ctx = [], []
    for i in range(states.shape[1]):
        sel = np.nonzero(states[:, i, 0, 10 + focus_oid - 1] == 1)[0]
        if sel.size:
            foc.append(states[sel, i]); ctx.append(np.delete(states[sel], i, axis=1))
    return np.concatenate(foc, 0), np.concatenate(ctx, 0)


def json_to_batch_files(jsonfiles, outfolder, batch_size=50, seed=0):
    """create_datasets_batches (data_process.lua:449-528) in memory: per-focus examples of every json file are cut into
    batches of `batch_size` (short tails are queued and re-cut at the end), each batch goes to trainset / valset /
    testset by uniform sampling against 70/15/15 quotas counted in batches, files are Torch7 `.t7` batch<i>."""
    from . import t7
    full, left_f, left_c = [], [], []
    for jf in jsonfiles:
        f, c = expand_for_each_object(raw_to_states(load_data_json(jf)))
        nb = f.shape[0] // batch_size
        for b in range(nb):
            full.append((f[b * batch_size:(b + 1) * batch_size], c[b * batch_size:(b + 1) * batch_size]))
        if f.shape[0] % batch_size:
            left_f.append(f[nb * batch_size:]); left_c.append(c[nb * batch_size:])
    if left_f:
        f, c = np.concatenate(left_f, 0), np.concatenate(left_c, 0)
        for b in range(f.shape[0] // batch_size):
            full.append((f[b * batch_size:(b + 1) * batch_size], c[b * batch_size:(b + 1) * batch_size]))
    n = len(full)
    n_test = int(math.floor(n * 0.15)); n_val = n_test; n_train = n - 2 * n_test
    limits = {"trainset": n_train, "valset": n_val, "testset": n_test}
    counters = {k: 0 for k in limits}
    rng = np.random.default_rng(seed)
    for f, c in full:
        while True:
            k = ("trainset", "valset", "testset")[int(rng.integers(3))]
            if counters[k] < limits[k]:
                break
        counters[k] += 1
        d = os.path.join(outfolder, "batches", k)
        os.makedirs(d, exist_ok=True)
        t7.save_batch_file(os.path.join(d, f"batch{counters[k]}"), f, c)
    return counters


def scenes_from_json(jsonfiles):
    """All scenes of a dataset as ONE padded scene tensor for npe_scenes_upload: (S, Nmax, T, 22), n_obj (S)."""
    parts = [raw_to_states(load_data_json(jf)) for jf in jsonfiles]
    Nmax = max(p.shape[1] for p in parts); T = parts[0].shape[2]
    S = sum(p.shape[0] for p in parts)
    out = np.zeros((S, Nmax, T, 22), F32); n_obj = np.zeros(S, np.int32)
    a = 0
    for p in parts:
        out[a:a + p.shape[0], :p.shape[1]] = p; n_obj[a:a + p.shape[0]] = p.shape[1]
        a += p.shape[0]
    return out, n_obj
