"""Where does the end-to-end batch-50 loop spend its time?  Host enqueue cost per iteration (no waits) against the
pipelined loop with the loss of step s-1 / s-2 read after step s is enqueued."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dynamics_b200 import NPEModel, synth, lib as L
from dynamics_b200.model import init_params
m = NPEModel()
st = synth.ball_scenes(64 * 50, 4, 3, seed=4)
this = st[:, 0, 0:2].reshape(-1, 50, 2, 22); ctx = st[:, 1:, 0:2].reshape(-1, 50, 3, 2, 22)
y = (st[:, 0, 2:3] - 0).reshape(-1, 50, 1, 22).copy()
m.set_params(init_params()); m.optim_reset(0.0)
pinned = []
for i in range(64):
    bufs = []
    for a in (this[i], ctx[i], y[i]):
        arr, ptr = L.pinned_array(m.lib, a.shape); arr[...] = a; bufs.append(arr)
    pinned.append(bufs)
argp = [(L.fptr(a), L.fptr(b), L.fptr(c), b.shape[1]) for a, b, c in pinned]
loss3 = np.zeros(3, np.float32); loss_p = L.fptr(loss3)
lib, hdl = m.lib, m.h
def run(n, lag):
    for s in range(n):
        a, b, c, cn = argp[s % 64]
        rc = lib.npe_train_batch_async(hdl, a, b, c, 50, cn, 0, 3e-4, 50, s % 64)
        if lag and s >= lag: rc = rc or lib.npe_loss_wait(hdl, (s - lag) % 64, loss_p)
        assert rc == 0
    m.sync()
for lag in (1, 2, 4, 0):
    run(200, lag)
    t0 = time.perf_counter(); run(2000 if lag else 400, lag); t1 = time.perf_counter()
    n = 2000 if lag else 400
    print(f"loss read {lag} steps late: {1e6 * (t1 - t0) / n:.1f} us per iteration (wall clock incl. final sync)")
# host cost of the calls alone: time the enqueue loop without the final sync
m.sync(); t0 = time.perf_counter()
for s in range(300):
    a, b, c, cn = argp[s % 64]
    lib.npe_train_batch_async(hdl, a, b, c, 50, cn, 0, 3e-4, 50, s % 64)
t1 = time.perf_counter(); m.sync()
print(f"enqueue only: {1e6 * (t1 - t0) / 300:.1f} us per npe_train_batch_async call")
m.close()
