import sys; sys.path.insert(0,'/root/repo')
import numpy as np
from dynamics_b200 import NPEModel, ModelParams, synth
from dynamics_b200.model import init_params
m = NPEModel(); m.set_params(init_params())
st = synth.ball_scenes(16384, 8, 2, seed=5)
m.upload_scenes(st)
m.set_precision("tf32x3")
m.rollout_resident(0, 3); m.sync()
for k in ("build_pairs", "forward", "rollout"):
    m.profile_select(k); m.rollout_resident(0, 10); n, tot = m.profile_read(); print(k, n, tot / max(n,1) * 1e3, "us per bracket")
m.profile_select(None)
f, p = 0, 0
m.close()
