"""Small driver for ncu captures: runs a few training steps (and optionally a short rollout) of one workload.
usage: python tools/prof_workload.py {b50|stress64|rollout} [steps]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamics_b200 import ModelParams, NPEModel, synth  # noqa: E402
from dynamics_b200.model import init_params  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "stress64"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
if which == "stress64":
    m = NPEModel(ModelParams(max_ctx=0))
    st = synth.ball_scenes(1024, 64, 3, seed=64, world=(3200, 2400), collide=False)
    m.set_params(init_params()); m.optim_reset(0.0)
    m.upload_scenes(st); m.upload_windows(np.arange(1024), np.zeros(1024, np.int32))
    for s in range(steps):
        m.select_windows(0, 1024)
        m.train_step("adam", 3e-4, want_loss=False)
    m.sync()
elif which in ("fwd_tf32", "fwd_tf32x3", "fwd_fp32"):
    m = NPEModel(ModelParams(max_ctx=0))
    st = synth.ball_scenes(1024, 64, 3, seed=64, world=(3200, 2400), collide=False)
    m.set_params(init_params())
    m.upload_scenes(st); m.upload_windows(np.arange(1024), np.zeros(1024, np.int32))
    m.select_windows(0, 1024)
    m.set_precision(which.split("_")[1])
    for s in range(steps):
        m._ck(m.lib.npe_forward_per_example(m.h, None, None))
    m.sync()
elif which == "b50":
    m = NPEModel()
    st = synth.ball_scenes(200, 4, 10, seed=4)
    m.set_params(init_params()); m.optim_reset(0.0)
    m.upload_scenes(st)
    m.upload_foci(np.arange(200), np.zeros(200, np.int32), np.full(200, 3, np.int32))
    for s in range(steps):
        m.select_foci((s % 4) * 50, 50)
        m.train_step("adam", 3e-4, want_loss=False)
    m.sync()
else:
    m = NPEModel()
    st = synth.ball_scenes(148 * 8 * 2, 8, 2, seed=5)
    m.set_params(init_params())
    m.upload_scenes(st)
    m.rollout_resident(0, steps * 10)
    m.sync()
print("done", which, m.launch_count())
m.close()
