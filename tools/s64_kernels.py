"""Per-kernel-class times of the S-64 training step (event-bracketed, one class at a time)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamics_b200 import ModelParams, NPEModel, synth
from dynamics_b200.model import init_params
m = NPEModel(ModelParams(max_ctx=0))
st = synth.ball_scenes(2048, 64, 3, seed=64, world=(3200, 2400), collide=False)
m.set_params(init_params()); m.optim_reset(0.0)
m.upload_scenes(st); m.upload_windows(np.arange(2048), np.zeros(2048, np.int32))
def step(s):
    m.select_windows((s % 2) * 1024, 1024)
    m.train_step("adam", 3e-4, want_loss=False)
for s in range(5): step(s)
out = {}
for k in ("build_pairs", "forward", "dec_backward", "pair_backward", "wgrad", "reduce"):
    m.profile_select(k)
    for s in range(8): step(s)
    n, tot = m.profile_read()
    out[k] = round(1e3 * tot / max(n, 1), 1)
m.profile_select(None)
m.sync()
m.event_record(0)
for s in range(20): step(s)
m.event_record(1); m.sync()
print("kernel us", out, "sum", round(sum(out.values()), 1), "step us", round(1e3 * m.event_elapsed_ms(0, 1) / 20, 1))
m.close()
