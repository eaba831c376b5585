import sys, time; sys.path.insert(0,'/root/repo')
import numpy as np
from dynamics_b200 import NPEModel, synth, lib as L
from dynamics_b200.model import init_params
m = NPEModel(); m.set_params(init_params()); m.optim_reset(0.0)
st = synth.ball_scenes(50, 4, 4, seed=1)
this = np.ascontiguousarray(st[:,0,0:2]); ctx = np.ascontiguousarray(st[:,1:,0:2]); y = np.ascontiguousarray(st[:,0,2]); 
pins = [L.pinned_array(m.lib, a.shape)[0] for a in (this, ctx, y)]
for p,a in zip(pins,(this,ctx,y)): p[...] = a
loss3 = np.zeros(3, np.float32)
def t(fn, n=2000):
    for _ in range(50): fn()
    m.sync(); t0=time.perf_counter()
    for _ in range(n): fn()
    m.sync(); return (time.perf_counter()-t0)/n*1e6
up = lambda: m._ck(m.lib.npe_batch_upload(m.h, L.fptr(pins[0]), L.fptr(pins[1]), L.fptr(pins[2]), 50, 3))
print('upload only (pinned)      %.1f us' % t(up))
up2 = lambda: m._ck(m.lib.npe_batch_upload(m.h, L.fptr(this), L.fptr(ctx), L.fptr(y), 50, 3))
print('upload only (pageable)    %.1f us' % t(up2))
def step_noloss(): up(); m._ck(m.lib.npe_train_step(m.h, 0, 3e-4, 50, None))
print('upload+train (no loss)    %.1f us' % t(step_noloss))
def step_loss(): up(); m._ck(m.lib.npe_train_step(m.h, 0, 3e-4, 50, L.fptr(loss3)))
print('upload+train+loss readback %.1f us' % t(step_loss))
def ts(): m._ck(m.lib.npe_train_step(m.h, 0, 3e-4, 50, None))
up(); print('train only (same batch, pairs cached) %.1f us' % t(ts))
def fwd(): up(); m._ck(m.lib.npe_forward_per_example(m.h, None, None))
print('upload+forward %.1f us' % t(fwd))
m.close()
