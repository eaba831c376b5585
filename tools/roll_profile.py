import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamics_b200 import ModelParams, NPEModel, synth
from dynamics_b200.model import init_params
m = NPEModel(ModelParams())
st = synth.ball_scenes(65536, 8, 2, seed=5)
m.set_params(init_params()); m.upload_scenes(st)
m.set_precision("tf32x3"); m.set_rollout_mode("kernels")
m.rollout_resident(0, 3); m.sync()
prev_n, prev_ms = 0, 0.0
for n in (25, 50, 100, 200, 400, 700, 1000):
    m.event_record(0); m.rollout_resident(0, n); m.event_record(1); m.sync()
    ms = m.event_elapsed_ms(0, 1)
    print(f"steps {prev_n:4d}..{n:4d}: {1e3 * (ms - prev_ms) / (n - prev_n):7.1f} us per step")
    prev_n, prev_ms = n, ms
m.close()
