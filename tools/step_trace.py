"""Timeline of k_step_fused (phase stamps per CTA) for a batch-50 step in the middle of a free-running loop."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from dynamics_b200 import ModelParams, NPEModel, synth
from dynamics_b200.model import init_params

states, n_obj, offs = bench.make_training_data(0)
sc, ob, tt = synth.reference_batch_schedule(bench.SCENES_PER_SET, bench.SETS, 400, bench.BATCH, 11, offs)
m = NPEModel(ModelParams())
m.set_params(init_params(1234)); m.optim_reset(0.0)
m.upload_scenes(states, n_obj); m.upload_foci(sc, ob, tt)
m.train_steps(0, 50, 300)
m._ck(m.lib.npe_step_trace(m.h, None, 0))
m.train_steps(300 * 50, 50, 60)
buf = np.zeros((64, 8), np.uint64)
m._ck(m.lib.npe_step_trace(m.h, buf.ctypes.data_as(C.POINTER(C.c_uint64)), 64))
rows = [r for r in buf if r[0] > 0]
t0 = min(int(r[1]) for r in rows)
print("CTA  loads_done  wait_done  handover1  handover2  grads_done  stored   (us relative to the first wait_done)")
for i, r in enumerate(rows):
    print(f"{i:3d} " + " ".join(f"{(int(x) - t0) / 1e3:9.2f}" if x else "     -   " for x in r[:6]))
m.close()
