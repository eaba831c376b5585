"""Synthetic scene generators + the raw(12) -> one-hot(22) state codec.

TEST INFRASTRUCTURE (see oracle/__init__.py).  matter-js / node are absent, so trajectories come from a small
seeded elastic-disc simulator whose *distribution* follows the reference's generator:
  world 800x600, ball radius 60, initial positions uniform in [61,739)x[61,539) rejected at distance < 150
    (/root/reference/src/js/common.js:8-44, src/js/demo/js/Examples.js:1987-2003)
  integer initial velocities in [-20,20]^2 (common.js:122-132), masses uniform over {1,5,25} with -m
    (common.js:146-155), no gravity/friction, restitution 1, no rotation (Examples.js:1993,2040-2052)
  walls: 80-px static obstacles on the 9x7 ring, centres x in {80..720}, y in {60..540} (Examples.js:2775-2800)
The codec follows /root/reference/src/lua/data_process.lua:119-139 (normalize), :184-207 (properties2onehotall)
and data_utils.lua:287-293 (num2onehot).
"""
import numpy as np

from . import config as C

F32 = np.float32
BALL_R = 60.0


# ------------------------------------------------------------------------------------------------------------
# codec
# ------------------------------------------------------------------------------------------------------------
def wrap_pi(a):
    """data_process.lua:102-107: a - 2pi where a > pi (float32)."""
    a = a.astype(F32)
    return (a + F32(-2 * np.pi) * (a > F32(np.pi)).astype(F32)).astype(F32)


def raw_to_state(raw, obj_sizes=C.OBJECT_SIZES):
    """raw (...,12) float32 -> state (...,22): normalize (data_process.lua:119-139) then one-hot the six property
    columns (data_process.lua:184-207).  booleans 0 -> [1,0], 1 -> [0,1] (config.lua:44)."""
    raw = np.asarray(raw, F32)
    out = np.zeros(raw.shape[:-1] + (C.OBJ_DIM,), F32)
    out[..., 0:2] = (raw[..., 0:2] / F32(C.PNC)).astype(F32)
    out[..., 2:4] = (raw[..., 2:4] / F32(C.VNC)).astype(F32)
    out[..., 4:6] = (wrap_pi(raw[..., 4:6]) / F32(C.ANC)).astype(F32)

    def onehot(col, cats, lo):
        vals = raw[..., col]
        hit = np.zeros(vals.shape, bool)
        for i, c in enumerate(cats):
            m = vals == F32(c)
            out[..., lo + i] = m
            hit |= m
        assert hit.all(), f"value outside categories for raw column {col}"

    onehot(C.R_M, C.MASSES, C.M0)
    onehot(C.R_OID, (1, 2, 3), C.OID0)
    onehot(C.R_OS, obj_sizes, C.OS0)
    onehot(C.R_G, (0, 1), C.G0)
    onehot(C.R_F, (0, 1), C.FR0)
    onehot(C.R_P, (0, 1), C.P0)
    return out


def state_to_raw(state, obj_sizes=C.OBJECT_SIZES):
    """Inverse codec: onehot2propertiesall (data_process.lua:209-232) + unnormalize (:143-162, incl. wrap_2pi of
    the angle column only)."""
    state = np.asarray(state, F32)
    raw = np.zeros(state.shape[:-1] + (C.RAW_DIM,), F32)
    raw[..., 0:2] = (state[..., 0:2] * F32(C.PNC)).astype(F32)
    raw[..., 2:4] = (state[..., 2:4] * F32(C.VNC)).astype(F32)
    ang = (state[..., 4:6] * F32(C.ANC)).astype(F32)
    a = ang[..., 0]
    ang[..., 0] = (a + F32(2 * np.pi) * (a < 0).astype(F32)).astype(F32)
    raw[..., 4:6] = ang
    raw[..., C.R_M] = np.asarray(C.MASSES, F32)[state[..., C.M0:C.M1].argmax(-1)]
    raw[..., C.R_OID] = state[..., C.OID0:C.OID1].argmax(-1) + 1
    raw[..., C.R_OS] = np.asarray(obj_sizes, F32)[state[..., C.OS0:C.OS1].argmax(-1)]
    raw[..., C.R_G] = state[..., C.G0:C.G1].argmax(-1)
    raw[..., C.R_F] = state[..., C.FR0:C.FR1].argmax(-1)
    raw[..., C.R_P] = state[..., C.P0:C.P1].argmax(-1)
    return raw


# ------------------------------------------------------------------------------------------------------------
# elastic-disc simulator (vectorised over scenes)
# ------------------------------------------------------------------------------------------------------------
def _initial_positions(rng, S, N, world, min_dist, margin):
    W, H = world
    pos = np.zeros((S, N, 2))
    for i in range(N):
        todo = np.arange(S)
        for _ in range(10000):
            if todo.size == 0:
                break
            cand = np.stack([rng.integers(margin, W - margin, todo.size),
                             rng.integers(margin, H - margin, todo.size)], -1).astype(np.float64)
            if i == 0:
                ok = np.ones(todo.size, bool)
            else:
                d = np.linalg.norm(pos[todo, :i] - cand[:, None, :], axis=-1)
                ok = (d >= min_dist).all(axis=1)
            pos[todo[ok], i] = cand[ok]
            todo = todo[~ok]
        assert todo.size == 0, "rejection sampling did not converge"
    return pos


def _step(pos, vel, mass, movable, world, radius):
    """One frame: free flight, border reflection, pairwise elastic impulses (restitution 1)."""
    W, H = world
    pos = pos + vel * movable[..., None]
    for ax, lim in ((0, W), (1, H)):
        lo = (pos[..., ax] < radius) & movable
        hi = (pos[..., ax] > lim - radius) & movable
        vel[..., ax] = np.where(lo, np.abs(vel[..., ax]), vel[..., ax])
        vel[..., ax] = np.where(hi, -np.abs(vel[..., ax]), vel[..., ax])
    S, N, _ = pos.shape
    for i in range(N):
        for j in range(i + 1, N):
            rsum = radius[:, i] + radius[:, j]
            dvec = pos[:, j] - pos[:, i]
            dist = np.linalg.norm(dvec, axis=-1)
            n = dvec / np.maximum(dist, 1e-9)[:, None]
            rel = ((vel[:, j] - vel[:, i]) * n).sum(-1)
            hit = (dist < rsum) & (rel < 0) & (movable[:, i] | movable[:, j])
            if not hit.any():
                continue
            im_i = np.where(movable[:, i], 1.0 / mass[:, i], 0.0)
            im_j = np.where(movable[:, j], 1.0 / mass[:, j], 0.0)
            jimp = np.where(hit, -2.0 * rel / np.maximum(im_i + im_j, 1e-30), 0.0)
            vel[:, i] -= (jimp * im_i)[:, None] * n
            vel[:, j] += (jimp * im_j)[:, None] * n
    return pos, vel


def _simulate(pos, vel, mass, movable, radius, T, world):
    S, N, _ = pos.shape
    P = np.zeros((S, N, T, 2)); V = np.zeros((S, N, T, 2))
    for t in range(T):
        P[:, :, t], V[:, :, t] = pos, vel
        pos, vel = _step(pos.copy(), vel.copy(), mass, movable, world, radius)
    return P, V


def gen_balls(n_scenes, n_balls, T=C.MAXWINSIZE, seed=0, world=(800, 600), variable_mass=True, max_v0=20):
    """S-n{k}: raw trajectories (S,N,T,12) float32.  Trajectories with |v| > 60 are clipped by re-drawing the seed
    stream is avoided: elastic collisions among masses {1,5,25} with |v0|<=20*sqrt2 can exceed 60 for light balls;
    such scenes are regenerated (generate.js:496-508 rejects them)."""
    rng = np.random.default_rng(seed)
    out = np.zeros((n_scenes, n_balls, T, C.RAW_DIM), F32)
    todo = np.arange(n_scenes)
    for _ in range(50):
        if todo.size == 0:
            break
        S = todo.size
        pos = _initial_positions(rng, S, n_balls, world, 150.0, 61)
        vel = rng.integers(-max_v0, max_v0 + 1, (S, n_balls, 2)).astype(np.float64)
        if variable_mass:
            mass = rng.choice(np.array([1.0, 5.0, 25.0]), (S, n_balls))
        else:
            mass = np.ones((S, n_balls))
        movable = np.ones((S, n_balls), bool)
        radius = np.full((S, n_balls), BALL_R)
        P, V = _simulate(pos, vel, mass, movable, radius, T, world)
        speed_ok = (np.abs(V) <= C.VNC).all(axis=(1, 2, 3))
        inb = ((P[..., 0] >= 0) & (P[..., 0] <= world[0]) & (P[..., 1] >= 0) & (P[..., 1] <= world[1])).all(axis=(1, 2))
        ok = speed_ok & inb
        raw = np.zeros((S, n_balls, T, C.RAW_DIM), F32)
        raw[..., 0:2] = P; raw[..., 2:4] = V
        raw[..., C.R_M] = mass[:, :, None]
        raw[..., C.R_OID] = C.OID_BALL
        raw[..., C.R_OS] = 1.0
        out[todo[ok]] = raw[ok]
        todo = todo[~ok]
    assert todo.size == 0
    return out


def wall_ring():
    """28 obstacle centres of the O geometry: the border cells of the 9x7 lattice x in {80..720 step 80},
    y in {60..540 step 80} (Examples.js:2775-2800)."""
    xs = [80 + 80 * i for i in range(9)]
    ys = [60 + 80 * j for j in range(7)]
    cells = [(x, y) for j, y in enumerate(ys) for i, x in enumerate(xs) if i in (0, 8) or j in (0, 6)]
    return np.asarray(cells, np.float64)


def gen_walls(n_scenes, T=C.MAXWINSIZE, seed=0, variant="O", n_balls=2, max_v0=20):
    """S-walls: obstacles first (mass 1e30, oid obstacle, static), then balls.  Variants add inner blocks:
    O: 28 (N=30), L: 28 (N=30, ring with one corner block moved inward), U: +3 (N=33), I: +2 (N=32)."""
    rng = np.random.default_rng(seed)
    ring = wall_ring()
    extra = {"O": np.zeros((0, 2)), "L": np.zeros((0, 2)),
             "U": np.asarray([(320.0, 220.0), (400.0, 220.0), (480.0, 220.0)]),
             "I": np.asarray([(400.0, 220.0), (400.0, 300.0)])}[variant]
    walls = np.concatenate([ring, extra], 0)
    if variant == "L":
        walls = walls.copy(); walls[0] = (160.0, 140.0)
    nw = walls.shape[0]
    N = nw + n_balls
    out = np.zeros((n_scenes, N, T, C.RAW_DIM), F32)
    todo = np.arange(n_scenes)
    world = (800, 600)
    for _ in range(200):
        if todo.size == 0:
            break
        S = todo.size
        pos = np.zeros((S, N, 2)); pos[:, :nw] = walls
        # balls inside the ring, away from every wall (circumscribed radius 40*sqrt2) and from each other
        for b in range(n_balls):
            left = np.arange(S)
            for _ in range(10000):
                if left.size == 0:
                    break
                cand = np.stack([rng.integers(180, 621, left.size), rng.integers(160, 441, left.size)], -1).astype(float)
                dw = np.linalg.norm(pos[left, :nw + b] - cand[:, None, :], axis=-1)
                lim = np.concatenate([np.full(nw, BALL_R + 57.0), np.full(b, 150.0)])
                ok = (dw >= lim).all(axis=1)
                pos[left[ok], nw + b] = cand[ok]
                left = left[~ok]
            assert left.size == 0
        vel = np.zeros((S, N, 2)); vel[:, nw:] = rng.integers(-max_v0, max_v0 + 1, (S, n_balls, 2))
        mass = np.full((S, N), 1e30); mass[:, nw:] = 1.0
        movable = np.zeros((S, N), bool); movable[:, nw:] = True
        radius = np.full((S, N), 40.0 * np.sqrt(2.0)); radius[:, nw:] = BALL_R
        P, V = _simulate(pos, vel, mass, movable, radius, T, world)
        ok = (np.abs(V) <= C.VNC).all(axis=(1, 2, 3))
        raw = np.zeros((S, N, T, C.RAW_DIM), F32)
        raw[..., 0:2] = P; raw[..., 2:4] = V
        raw[..., C.R_M] = mass[:, :, None].astype(F32)
        raw[:, :nw, :, C.R_OID] = C.OID_OBSTACLE
        raw[:, nw:, :, C.R_OID] = C.OID_BALL
        raw[..., C.R_OS] = 1.0
        out[todo[ok]] = raw[ok]
        todo = todo[~ok]
    assert todo.size == 0
    return out


def gen_big(n_scenes, n_balls=64, T=3, seed=64, world=(3200, 2400)):
    """S-64: 64 balls in a 3200x2400 world (areal density of 4 balls in 800x600); pnc stays 800."""
    return gen_balls(n_scenes, n_balls, T, seed, world=world, variable_mass=True)
