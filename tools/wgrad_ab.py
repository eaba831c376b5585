"""A/B of the weight-gradient kernels on the golden batch forced through the large-batch path: prints per-block errors
against the oracle.  usage: python tools/wgrad_ab.py [out.npy]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamics_b200 import NPEModel
from oracle import npe
z = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "n4_b50.npz"))
batch = (z["this"], z["context"], z["y"])
params = npe.init_params(1234).copy()
params += (0.05 * np.random.default_rng(7).standard_normal(params.shape)).astype(np.float32) * np.abs(params).mean()
params = params.astype(np.float32)
m = NPEModel()
m.set_option("small_batch_max", 0)
m.fp(params, batch)
g = m.bp()
g_ref, _ = npe.bp(params, batch)
gm = np.abs(g_ref).max()
blocks = [("enc_T", 0, 1100), ("enc_C", 1100, 2200)] + [(f"pair{l}", 2200 + 2500 * l, 2200 + 2500 * (l + 1)) for l in range(5)]
o = 2200 + 12500
blocks += [("dec1_w", o, o + 4700), ("dec1_b", o + 4700, o + 4750)]
o += 4750
for l in range(3):
    blocks += [(f"dec{l+2}", o, o + 2550)]; o += 2550
blocks += [("dec5", o, o + 1122)]
for name, a, b in blocks:
    print(f"{name:8s} max|ref| {np.abs(g_ref[a:b]).max():.3e}  err/max|g| {np.abs(g[a:b] - g_ref[a:b]).max() / gm:.3e}  nonzero {np.count_nonzero(g[a:b])}/{b - a}")
if len(sys.argv) > 1: np.save(sys.argv[1], g)
m.close()
