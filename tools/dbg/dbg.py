import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from oracle import batcher, npe
from tests.util import balls_batches
from dynamics_b200 import NPEModel
p = npe.init_params(1234).copy(); rng = np.random.default_rng(7)
p += (0.05 * rng.standard_normal(p.shape)).astype(np.float32) * np.abs(p).mean(); p = p.astype(np.float32)
_, batches = balls_batches(60, int(sys.argv[1]), seed=int(sys.argv[1]))
batch = batcher.make_batch(*batches[0], offset=5)
l, pred, lv, lw = npe.fp(p, batch)
m = NPEModel()
gl, gpred, _, _ = m.fp(p, batch)
idx, mask, cnt = m.pair_table()
err = np.abs(gpred[:, :6] - pred[:, :6]).max(1)
act = mask.sum(1)
cum = np.cumsum(act) - act
bad=[b for b in range(50) if err[b]>1e-6]; print("P", act.sum(), "bad foci", bad, "their first pair rows", [int(cum[b]) for b in bad])
for b in range(0):
    print(b, act[b], cum[b], f"{err[b]:.2e}")
