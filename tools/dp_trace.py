"""Time line of the data-parallel batch-50 step (torchrun, 2+ GPUs; NPE_CHAIN_PROF=1): the last two k_dp_step launches and the
last k_step_fused of a free-running npe_train_steps loop, absolute globaltimer stamps printed by npe_sync on stderr.
NPE_DP_DEBUG (one digit per rank: 0 normal, 1 push without waiting, 2 neither) separates the cost of the push."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from dynamics_b200 import ModelParams, NPEModel, synth
from dynamics_b200.model import init_params
import torch.distributed as dist

rank, local_rank, world = bench.dist_env()
dist.init_process_group("gloo", rank=rank, world_size=world)
states, n_obj, offs = bench.make_training_data(0)
sc, ob, tt = synth.reference_batch_schedule(bench.SCENES_PER_SET, bench.SETS, 1400, bench.BATCH, 11, offs)
m = NPEModel(ModelParams(device=local_rank))
m.set_params(init_params(1234)); m.optim_reset(0.0)
bench.attach_dp(m, dist, rank, world, "p2p")
m.upload_scenes(states, n_obj); m.upload_foci(sc, ob, tt)
gB = bench.BATCH * world
m.train_steps(0, bench.BATCH, 1000, "adam", bench.LR, divisor_batch=gB)
m._ck(m.lib.npe_step_trace(m.h, None, 0))
dist.barrier()
m.train_steps(1000 * bench.BATCH, bench.BATCH, 301, "adam", bench.LR, divisor_batch=gB)
m.sync()
dist.barrier()
m.close()
