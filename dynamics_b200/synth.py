"""Synthetic scenes for benchmarks and examples (no node/matter-js, no datasets in this environment).

Distribution facts follow the reference's generator (src/js/common.js:8-44,122-155; src/js/demo/js/Examples.js:1987-2003):
world 800x600, ball radius 60, integer initial positions in [61,739)x[61,539) at pairwise distance >= 150, integer
velocities in [-20,20]^2, masses in {1,5,25}.  Frames advance by free flight with elastic border reflection and
equal-time elastic disc collisions.  States are emitted directly in the 22-float one-hot layout (config.lua:25).
"""
import numpy as np

PNC, VNC = 800.0, 60.0
MASS_COL = {1.0: 6, 5.0: 7, 25.0: 8}


def _positions(rng, S, N, world, min_dist=150.0, margin=61):
    W, H = world
    pos = np.zeros((S, N, 2))
    for i in range(N):
        todo = np.arange(S)
        while todo.size:
            cand = np.stack([rng.integers(margin, W - margin, todo.size), rng.integers(margin, H - margin, todo.size)], -1)
            ok = np.ones(todo.size, bool) if i == 0 else \
                (np.linalg.norm(pos[todo, :i] - cand[:, None, :], axis=-1) >= min_dist).all(1)
            pos[todo[ok], i] = cand[ok]
            todo = todo[~ok]
    return pos


def ball_scenes(n_scenes, n_balls, T, seed, world=(800, 600), collide=True):
    """-> states (S, N, T, 22) float32 of ball-only scenes."""
    rng = np.random.default_rng(seed)
    S, N = n_scenes, n_balls
    pos = _positions(rng, S, N, world)
    vel = rng.integers(-20, 21, (S, N, 2)).astype(np.float64)
    mass = rng.choice(np.array([1.0, 5.0, 25.0]), (S, N))
    out = np.zeros((S, N, T, 22), np.float32)
    for m, col in MASS_COL.items():
        out[..., col] = (mass == m)[:, :, None]
    out[..., 10] = 1                       # oid: ball
    out[..., 14] = 1                       # size 1
    out[..., 16] = out[..., 18] = out[..., 20] = 1   # gravity/friction/pairwise off -> [1,0]
    R = 60.0
    for t in range(T):
        out[:, :, t, 0:2] = pos / PNC
        out[:, :, t, 2:4] = vel / VNC
        pos = pos + vel
        for ax, lim in ((0, world[0]), (1, world[1])):
            vel[..., ax] = np.where(pos[..., ax] < R, np.abs(vel[..., ax]), vel[..., ax])
            vel[..., ax] = np.where(pos[..., ax] > lim - R, -np.abs(vel[..., ax]), vel[..., ax])
        if collide and N <= 16:
            for i in range(N):
                for j in range(i + 1, N):
                    d = pos[:, j] - pos[:, i]
                    dist = np.linalg.norm(d, axis=-1)
                    n = d / np.maximum(dist, 1e-9)[:, None]
                    rel = ((vel[:, j] - vel[:, i]) * n).sum(-1)
                    hit = (dist < 2 * R) & (rel < 0)
                    if hit.any():
                        imp = np.where(hit, -2.0 * rel / (1.0 / mass[:, i] + 1.0 / mass[:, j]), 0.0)
                        vel[:, i] -= (imp / mass[:, i])[:, None] * n
                        vel[:, j] += (imp / mass[:, j])[:, None] * n
            np.clip(vel, -VNC, VNC, out=vel)
    return out


def reference_batch_schedule(n_scenes_per_set, set_sizes, steps, batch_size, seed, scene_offsets, max_offset=20):
    """Example schedule with the reference's batch composition: every step draws a dataset
    uniformly, then a batch = `batch_size` consecutive scenes, ONE focus slot and ONE window offset in [0, 20).
    Returns (scene_ids, obj_ids, t0) int32 arrays of steps * batch_size examples."""
    rng = np.random.default_rng(seed)
    sc, ob, tt = [], [], []
    for _ in range(steps):
        d = int(rng.integers(len(set_sizes)))
        nb = n_scenes_per_set // batch_size
        b = int(rng.integers(nb))
        o = int(rng.integers(set_sizes[d]))
        t0 = int(rng.integers(max_offset))
        sc.append(scene_offsets[d] + b * batch_size + np.arange(batch_size))
        ob.append(np.full(batch_size, o)); tt.append(np.full(batch_size, t0))
    return (np.concatenate(sc).astype(np.int32), np.concatenate(ob).astype(np.int32), np.concatenate(tt).astype(np.int32))


def wall_scenes(n_scenes, T, seed, variant="O", n_balls=2):
    """configs[2] shape: a ring of static 80-px obstacles (9 x 7 border = 28 blocks, centres x in {80..720}, y in {60..540};
    Examples.js:2775-2800) plus the variant's inner blocks (L: +1 corner piece of 2, U / I not generated here), then
    `n_balls` balls bouncing inside.  Object order = walls first, then balls (Examples.js:3124-3200).
    -> states (S, N, T, 22) float32; obstacles: mass one-hot 1e30 (col 9), oid obstacle (col 11), size 1."""
    rng = np.random.default_rng(seed)
    xs = np.arange(80, 721, 80); ys = np.arange(60, 541, 80)
    ring = [(x, y) for x in xs for y in (ys[0], ys[-1])] + [(x, y) for y in ys[1:-1] for x in (xs[0], xs[-1])]
    if variant == "L":
        ring += [(240, 220), (240, 300)]
    walls = np.asarray(ring, np.float64)
    W = len(walls)
    S, N = n_scenes, W + n_balls
    out = np.zeros((S, N, T, 22), np.float32)
    out[:, :W, :, 0:2] = (walls / PNC)[None, :, None, :]
    out[:, :W, :, 9] = 1; out[:, :W, :, 11] = 1
    out[:, W:, :, 10] = 1
    mass = rng.choice(np.array([1.0, 5.0, 25.0]), (S, n_balls))
    for m, col in MASS_COL.items():
        out[:, W:, :, col] = (mass == m)[:, :, None]
    out[..., 14] = 1
    out[..., 16] = out[..., 18] = out[..., 20] = 1
    lo = np.array([180.0, 160.0]); hi = np.array([620.0, 440.0])          # interior kept clear of the ring
    pos = lo + rng.random((S, n_balls, 2)) * (hi - lo)
    vel = rng.integers(-20, 21, (S, n_balls, 2)).astype(np.float64)
    for t in range(T):
        out[:, W:, t, 0:2] = pos / PNC
        out[:, W:, t, 2:4] = vel / VNC
        pos = pos + vel
        for ax in (0, 1):
            vel[..., ax] = np.where(pos[..., ax] < lo[ax], np.abs(vel[..., ax]), vel[..., ax])
            vel[..., ax] = np.where(pos[..., ax] > hi[ax], -np.abs(vel[..., ax]), vel[..., ax])
    return out
