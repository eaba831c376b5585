import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamics_b200 import ModelParams, NPEModel, synth
from dynamics_b200.model import init_params
m = NPEModel(ModelParams())
st = synth.ball_scenes(8192, 8, 2, seed=5)
m.set_params(init_params())
m.set_precision("tf32x3"); m.set_rollout_mode("kernels")
m.upload_scenes(st); m.rollout(0, 2, use_gt=False, want_metrics=False)
for rep in range(3):
    t0 = time.perf_counter(); m.upload_scenes(st); t1 = time.perf_counter()
    traj, _ = m.rollout(0, 58, use_gt=False, want_metrics=False); t2 = time.perf_counter()
    print(f"upload {1e3*(t1-t0):.1f} ms, rollout+traj {1e3*(t2-t1):.1f} ms")
t0 = time.perf_counter(); n = m.rollout_resident(0, 58); m.sync(); print(f"resident 58 steps {1e3*(time.perf_counter()-t0):.1f} ms")
m.close()
