"""Driver for an ncu launch list of LATE rollout steps (steady state): 310 steps of the configs[4] rollout by kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamics_b200 import ModelParams, NPEModel, synth
from dynamics_b200.model import init_params
m = NPEModel(ModelParams())
st = synth.ball_scenes(65536, 8, 2, seed=5)
m.set_params(init_params()); m.upload_scenes(st)
m.set_precision("tf32x3"); m.set_rollout_mode("kernels")
m.rollout_resident(0, 310); m.sync()
print("done", m.launch_count())
m.close()
