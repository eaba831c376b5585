"""configs[4] rollout step: device-resident rollout by kernels on 65 536 eight-ball scenes (usage: roll_time.py [steps] [precision])."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamics_b200 import ModelParams, NPEModel, synth
from dynamics_b200.model import init_params
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 50
prec = sys.argv[2] if len(sys.argv) > 2 else "tf32x3"
m = NPEModel(ModelParams())
st = synth.ball_scenes(65536, 8, 2, seed=5)
m.set_params(init_params()); m.upload_scenes(st)
m.set_precision(prec); m.set_rollout_mode("kernels")
m.rollout_resident(0, 3); m.sync()
m.event_record(0)
n = m.rollout_resident(0, steps)
m.event_record(1); m.sync()
ms = m.event_elapsed_ms(0, 1)
print(f"{prec}: {n / (ms * 1e-3) / 1e9:.3f} G object-steps/s, {1e3 * ms / steps:.1f} us per step")
m.close()
