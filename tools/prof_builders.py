"""Driver for ncu captures of the pair builder in both throughput regimes: 2 S-64 training steps, then 3 steps of the configs[4] rollout."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamics_b200 import ModelParams, NPEModel, synth
from dynamics_b200.model import init_params
m = NPEModel(ModelParams(max_ctx=0))
st = synth.ball_scenes(1024, 64, 3, seed=64, world=(3200, 2400), collide=False)
m.set_params(init_params()); m.optim_reset(0.0)
m.upload_scenes(st); m.upload_windows(np.arange(1024), np.zeros(1024, np.int32))
for s in range(2):
    m.select_windows(0, 1024)
    m.train_step("adam", 3e-4, want_loss=False)
m.sync(); m.close()
m = NPEModel(ModelParams())
st = synth.ball_scenes(65536, 8, 2, seed=5)
m.set_params(init_params()); m.upload_scenes(st)
m.set_precision("tf32x3"); m.set_rollout_mode("kernels")
m.rollout_resident(0, 3); m.sync()
print("done")
m.close()
