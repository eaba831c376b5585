"""Per-kernel-class times of the configs[4] rollout step by kernels (event-bracketed, one class at a time)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamics_b200 import ModelParams, NPEModel, synth
from dynamics_b200.model import init_params
m = NPEModel(ModelParams())
st = synth.ball_scenes(65536, 8, 2, seed=5)
m.set_params(init_params()); m.upload_scenes(st)
m.set_precision("tf32x3"); m.set_rollout_mode("kernels")
m.rollout_resident(0, 400); m.sync()   # past the dense early steps
out = {}
for k in ("build_pairs", "forward", "rollout"):
    m.profile_select(k)
    m.rollout_resident(0, 1000)
    n, tot = m.profile_read()
    out[k] = (n, round(1e3 * tot / max(n, 1), 1))
m.profile_select(None)
m.sync()
print("kernel us (launch brackets, mean)", out)
m.close()
