import sys, os, time, numpy as np
sys.path.insert(0, os.getcwd())
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
m.sync()
for rep in range(3):
    m.event_record(0)
    for s in range(20): step(s)
    m.event_record(1); m.sync()
    print("S-64 step ms", m.event_elapsed_ms(0, 1) / 20)
m.close()
