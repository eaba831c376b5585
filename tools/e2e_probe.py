"""Where does the end-to-end step go?  Host-side time of each API call in the pipelined loop of bench.py."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from dynamics_b200 import ModelParams, NPEModel, synth, lib as L
from dynamics_b200.model import init_params

states, n_obj, offs = bench.make_training_data(0)
total = 2050
sc, ob, tt = synth.reference_batch_schedule(bench.SCENES_PER_SET, bench.SETS, total, bench.BATCH, 11, offs)
m = NPEModel(ModelParams())
m.set_params(init_params(1234)); m.optim_reset(0.0)
ring = 64
hb = bench.host_batches(states, sc, ob, tt, ring)
pinned = []
for this, ctx, y in hb:
    bufs = []
    for a in (this, ctx, y):
        arr, ptr = L.pinned_array(m.lib, a.shape); arr[...] = a; bufs.append(arr)
    pinned.append(bufs)
loss3 = np.zeros(3, np.float32)
tu = ts = tw = 0.0
N = 4000
SKIP_UPLOAD = bool(os.environ.get('PROBE_SKIP_UPLOAD'))
t00 = time.perf_counter()
for s in range(N):
    this, ctx, y = pinned[s % ring]
    t0 = time.perf_counter()
    if not SKIP_UPLOAD or s == 0:
        m._ck(m.lib.npe_batch_upload(m.h, L.fptr(this), L.fptr(ctx), L.fptr(y), 50, ctx.shape[1]))
    t1 = time.perf_counter()
    m._ck(m.lib.npe_train_step_async(m.h, 0, 3e-4, 50, s % 64))
    t2 = time.perf_counter()
    if s:
        m._ck(m.lib.npe_loss_wait(m.h, (s - 1) % 64, L.fptr(loss3)))
    t3 = time.perf_counter()
    tu += t1 - t0; ts += t2 - t1; tw += t3 - t2
m.sync()
tt_ = time.perf_counter() - t00
print(f"per step: total {tt_ / N * 1e6:.1f} us | upload {tu / N * 1e6:.1f}  step_async {ts / N * 1e6:.1f}  loss_wait {tw / N * 1e6:.1f}")
m.close()
