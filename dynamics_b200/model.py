"""Host-side mirror of the reference's Lua ``model`` object (src/lua/npe.lua:69-526) over libnpe_cuda.so.

Same method names, argument meaning and return values as the table returned by ``require 'npe'``:
``create(mp, preload, model_path)``, ``fp``, ``fp_batch``, ``bp``, ``theta.params`` / ``theta.grad_params``.
Tensors are numpy float32 arrays where the reference has torch.FloatTensors.  All arithmetic happens in the CUDA
library; this file only marshals buffers (it is what lua/npe.lua does through LuaJIT FFI).
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import lib as L


@dataclass
class ModelParams:
    """The ``mp`` fields npe.lua reads (main.lua:36-57,119-123)."""
    batch_size: int = 50
    rnn_dim: int = 50
    layers: int = 5
    num_past: int = 2
    num_future: int = 1
    object_dim: int = 22
    nbrhd: bool = True
    nbrhdsize: float = 3.5
    vlambda: float = 1.0
    lambda_: float = 1.0
    max_ctx: int = 12
    model: str = "npe"
    device: int = 0
    ball_zero_mode: int = 1

    @property
    def input_dim(self):
        return self.object_dim * self.num_past

    @property
    def out_dim(self):
        return self.object_dim * self.num_future


# Flat getParameters() layout of init_network with -nbrhd (npe.lua:14-61, modules.lua:10-26,46-88): (name, shape)
PARAM_LAYOUT = ([("enc.this.W", (25, 44)), ("enc.ctx.W", (25, 44))] + [(f"pair.L{l}.W", (50, 50)) for l in range(1, 6)] +
                [("dec.L1.W", (50, 94)), ("dec.L1.b", (50,))] +
                [x for i in range(2, 5) for x in ((f"dec.L{i}.W", (50, 50)), (f"dec.L{i}.b", (50,)))] +
                [("dec.L5.W", (22, 50)), ("dec.L5.b", (22,))])


def param_offsets():
    off, out = 0, {}
    for name, shape in PARAM_LAYOUT:
        n = int(np.prod(shape))
        out[name] = (off, shape)
        off += n
    assert off == L.NPE_NUM_PARAMS
    return out


def init_params(seed=1234):
    """model.create without preload: nn.Linear:reset() draws weight and bias from U(-1/sqrt(in), 1/sqrt(in)); the five
    pairwise Linears are layer:clone() of one Linear, i.e. start identical (npe.lua:21,39)."""
    rng = np.random.default_rng(seed)
    flat = np.zeros(L.NPE_NUM_PARAMS, np.float32)
    first_pair, fan_in = None, None
    for name, (off, shape) in param_offsets().items():
        if len(shape) == 2:
            fan_in = shape[1]
        vals = rng.uniform(-1.0 / np.sqrt(fan_in), 1.0 / np.sqrt(fan_in), size=shape).astype(np.float32)
        if name.startswith("pair."):
            if first_pair is None:
                first_pair = vals
            vals = first_pair
        flat[off:off + vals.size] = vals.ravel()
    return flat


class _Theta:
    """model.theta: params / grad_params are host mirrors that read through to the device vectors."""

    def __init__(self, model):
        self._m = model

    @property
    def params(self):
        return self._m.get_params()

    @property
    def grad_params(self):
        return self._m.get_grads()


class NPEModel:
    def __init__(self, mp=None, library=None):
        self.mp = mp or ModelParams()
        self.lib = library or L.load_library()
        if self.mp.model != "npe":
            raise L.NPEError(-3, "only -model npe is built (np / lstm baselines are out of scope)")
        cfg = self.lib.default_config(device=self.mp.device, rnn_dim=self.mp.rnn_dim, layers=self.mp.layers,
                                      num_past=self.mp.num_past, num_future=self.mp.num_future,
                                      object_dim=self.mp.object_dim, nbrhd=int(self.mp.nbrhd),
                                      nbrhdsize=self.mp.nbrhdsize, max_ctx=self.mp.max_ctx, vlambda=self.mp.vlambda,
                                      ball_zero_mode=self.mp.ball_zero_mode, **{"lambda": self.mp.lambda_})
        h = C.c_void_p()
        rc = self.lib.npe_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise L.NPEError(rc, self.lib.npe_last_error(None).decode())
        self.h = h
        self.theta = _Theta(self)
        self.n_params = self.lib.npe_param_count(self.h)
        self._batch_id = None

    # model.create(mp_, preload, model_path)  (npe.lua:72-102)
    @classmethod
    def create(cls, mp=None, preload=False, model_path=None, params=None):
        m = cls(mp)
        if preload:
            from .t7 import load_checkpoint_params
            params = load_checkpoint_params(model_path)
        if params is not None:
            m.set_params(params)
        return m

    def close(self):
        if getattr(self, "h", None):
            self.lib.npe_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise L.NPEError(rc, self.lib.npe_last_error(self.h).decode())

    # ---- parameters -------------------------------------------------------------------------------------------
    def set_params(self, flat):
        flat = L.f32(flat).ravel()
        self._ck(self.lib.npe_params_set(self.h, L.fptr(flat), flat.size))

    def get_params(self):
        out = np.empty(self.n_params, np.float32)
        self._ck(self.lib.npe_params_get(self.h, L.fptr(out), out.size))
        return out

    def get_grads(self):
        out = np.empty(self.n_params, np.float32)
        self._ck(self.lib.npe_grads_get(self.h, L.fptr(out), out.size))
        return out

    def set_grads(self, flat):
        flat = L.f32(flat).ravel()
        self._ck(self.lib.npe_grads_set(self.h, L.fptr(flat), flat.size))

    # ---- batch form (reference batch tuple) -------------------------------------------------------------------
    def _upload(self, batch):
        if hasattr(batch, "select"):            # loader.DeviceBatch: examples already resident, only the triples travel
            return batch.select(self)
        this, context, y = batch[0], batch[1], batch[2]
        this = L.f32(this); context = L.f32(context)
        B = this.shape[0]
        Cn = context.shape[1] if context.ndim >= 2 else 0
        yy = None if y is None else L.f32(np.asarray(y).reshape(B, -1))
        self._ck(self.lib.npe_batch_upload(self.h, L.fptr(this), L.fptr(context), L.fptr(yy), B, Cn))
        return B

    def fp(self, params_, batch, sim=False):
        """model:fp (npe.lua:225-247) -> loss, prediction (B,22), vel_loss, ang_vel_loss"""
        if params_ is not None:
            self.set_params(params_)
        B = self._upload(batch)
        pred = np.empty((B, L.OBJ_DIM), np.float32)
        loss3 = np.empty(3, np.float32)
        self._ck(self.lib.npe_forward(self.h, L.fptr(pred), L.fptr(loss3)))
        return float(loss3[0]), pred, float(loss3[1]), float(loss3[2])

    def fp_batch(self, params_, batch, sim=False):
        """model:fp_batch (npe.lua:250-284) -> loss (B), prediction, vel_loss (B), ang_vel_loss (B)"""
        if params_ is not None:
            self.set_params(params_)
        B = self._upload(batch)
        pred = np.empty((B, L.OBJ_DIM), np.float32)
        le = np.empty((B, 3), np.float32)
        self._ck(self.lib.npe_forward_per_example(self.h, L.fptr(pred), L.fptr(le)))
        return le[:, 0].copy(), pred, le[:, 1].copy(), le[:, 2].copy()

    def bp(self, batch=None, prediction=None, sim=False, divisor_batch=0):
        """model:bp (npe.lua:290-336) -> grad_params.  Like the reference it relies on the state of the immediately
        preceding fp on the same batch (`batch`/`prediction` are accepted for signature compatibility)."""
        self._ck(self.lib.npe_backward(self.h, divisor_batch))
        return self.get_grads()

    # ---- simulate_all_postprocess helpers (eval.lua:170-173): host tensors, mirrored from lua/npe.lua --------------------
    PNC, VNC, ANC = np.float32(800.0), np.float32(60.0), np.float32(np.pi)       # config.lua:16-17,51

    def update_position(self, this, pred):
        """model:update_position (npe.lua:386-419): this (B,num_past,22), pred (B,num_future,22) -> copy of pred whose
        positions integrate the PREVIOUS step's velocity, starting from the last past frame."""
        this = L.f32(this); out = L.f32(pred).copy()
        pos = this[:, -1, 0:2] * self.PNC
        vel = this[:, -1, 2:4] * self.VNC
        for k in range(out.shape[1]):
            pos = (pos + vel).astype(np.float32)
            vel = out[:, k, 2:4] * self.VNC
            out[:, k, 0:2] = pos / self.PNC
        return out

    def update_angle(self, this, pred):
        """model:update_angle (npe.lua:422-458): running angle sum wrapped back into (-pi, pi]."""
        this = L.f32(this); out = L.f32(pred).copy()
        nf = out.shape[1]
        ang = np.empty((out.shape[0], nf + 1), np.float32)
        ang[:, 0] = this[:, -1, 4] * self.ANC
        av = this[:, -1, 5] * self.ANC
        for k in range(nf):
            ang[:, k + 1] = ang[:, k] + av
            av = out[:, k, 5] * self.ANC
        big = (ang > self.ANC).astype(np.float32); small = (ang <= -self.ANC).astype(np.float32)
        ang = (ang + np.float32(-2 * np.pi) * big).astype(np.float32)
        ang = (ang + np.float32(2 * np.pi) * small).astype(np.float32)
        out[:, :, 4] = ang[:, 1:] / self.ANC
        return out

    def bp_input(self, batch=None, prediction=None, sim=False):
        """model:bp_input (npe.lua:341-384) is not on the training / rollout path (and returns an undefined value in the
        reference): explicit error, as in lua/npe.lua."""
        raise L.NPEError(-3, "bp_input is not supported: input gradients are not part of the training / rollout path")

    def pair_table(self):
        """(closest_indices (F,K) int32, neighbour mask (F,K) uint8, kept count (F)) of the current batch."""
        F, K = self.lib.npe_num_foci(self.h), self.lib.npe_ctx_slots(self.h)
        idx = np.empty((F, K), np.int32); mask = np.empty((F, K), np.uint8); cnt = np.empty(F, np.int32)
        self._ck(self.lib.npe_build_pairs(self.h, L.iptr(idx), mask.ctypes.data_as(C.POINTER(C.c_uint8)), L.iptr(cnt)))
        return idx, mask, cnt

    # ---- optimiser boundary (main.lua:135-144,301-302) ---------------------------------------------------------
    def optim_reset(self, rmsprop_m_init=0.0):
        self._ck(self.lib.npe_optim_reset(self.h, rmsprop_m_init))

    def adam_step(self, lr, beta1=0.9, beta2=0.999, eps=1e-8):
        self._ck(self.lib.npe_adam_step(self.h, lr, beta1, beta2, eps))

    def rmsprop_step(self, lr, alpha=0.99, eps=1e-8):
        self._ck(self.lib.npe_rmsprop_step(self.h, lr, alpha, eps))

    def train_step(self, opt="adam", lr=3e-4, divisor_batch=0, want_loss=True):
        loss3 = np.empty(3, np.float32) if want_loss else None
        self._ck(self.lib.npe_train_step(self.h, 0 if opt == "adam" else 1, lr, divisor_batch, L.fptr(loss3)))
        return loss3

    def train_step_async(self, slot, opt="adam", lr=3e-4, divisor_batch=0):
        """train_step without the host wait; the loss lands in ring slot `slot` (0..63), see loss_wait."""
        self._ck(self.lib.npe_train_step_async(self.h, 0 if opt == "adam" else 1, lr, divisor_batch, slot))

    def train_batch_async(self, batch, slot, opt="adam", lr=3e-4, divisor_batch=0):
        """Upload a host batch {this, context, y} and enqueue its training step in one call (loss: loss_wait(slot))."""
        this, context, y = L.f32(batch[0]), L.f32(batch[1]), L.f32(np.asarray(batch[2]).reshape(len(batch[0]), -1))
        self._ck(self.lib.npe_train_batch_async(self.h, L.fptr(this), L.fptr(context), L.fptr(y), this.shape[0],
                                                context.shape[1] if context.ndim >= 2 else 0,
                                                0 if opt == "adam" else 1, lr, divisor_batch, slot))

    def loss_wait(self, slot):
        loss3 = np.empty(3, np.float32)
        self._ck(self.lib.npe_loss_wait(self.h, slot, L.fptr(loss3)))
        return loss3

    def train_steps(self, first, batch, nsteps, opt="adam", lr=3e-4, divisor_batch=0, want_losses=False):
        """The inner loop of train() (main.lua:299-304) over the uploaded schedule, one host call for all steps."""
        losses = np.empty((nsteps, 3), np.float32) if want_losses else None
        self._ck(self.lib.npe_train_steps(self.h, first, batch, nsteps, 0 if opt == "adam" else 1, lr, divisor_batch,
                                          L.fptr(losses)))
        return losses

    # ---- scene form (device-resident trajectory batcher) -------------------------------------------------------
    def upload_scenes(self, states, n_obj=None):
        states = L.f32(states)
        S, Nmax, T, D = states.shape
        assert D == L.OBJ_DIM
        n_obj = L.i32(np.full(S, Nmax) if n_obj is None else n_obj)
        self._ck(self.lib.npe_scenes_upload(self.h, L.fptr(states), L.iptr(n_obj), S, Nmax, T))
        self.scene_shape = (S, Nmax, T)

    def upload_scenes_raw(self, raw, n_obj=None):
        """Raw matter-js states (S,Nmax,T,12) (dataprep.load_data_json): normalisation and one-hot coding happen on the
        device (data_process.lua:119-139,184-207)."""
        raw = L.f32(raw)
        S, Nmax, T, K = raw.shape
        assert K == 12
        n_obj = L.i32(np.full(S, Nmax) if n_obj is None else n_obj)
        self._ck(self.lib.npe_scenes_upload_raw(self.h, L.fptr(raw), L.iptr(n_obj), S, Nmax, T))
        self.scene_shape = (S, Nmax, T)

    def upload_windows(self, scene_ids, t0):
        scene_ids, t0 = L.i32(scene_ids), L.i32(t0)
        self._ck(self.lib.npe_windows_upload(self.h, L.iptr(scene_ids), L.iptr(t0), scene_ids.size))

    def upload_foci(self, scene_ids, obj_ids, t0):
        scene_ids, obj_ids, t0 = L.i32(scene_ids), L.i32(obj_ids), L.i32(t0)
        self._ck(self.lib.npe_foci_upload(self.h, L.iptr(scene_ids), L.iptr(obj_ids), L.iptr(t0), scene_ids.size))

    def select_foci(self, first, count):
        self._ck(self.lib.npe_select_foci(self.h, first, count))
        return self.lib.npe_num_foci(self.h)

    def select_windows(self, first, count):
        self._ck(self.lib.npe_select_windows(self.h, first, count))
        return self.lib.npe_num_foci(self.h)

    def download_batch(self):
        F = self.lib.npe_num_foci(self.h)
        this = np.empty((F, L.INPUT_DIM), np.float32); y = np.empty((F, L.OBJ_DIM), np.float32)
        self._ck(self.lib.npe_batch_download(self.h, L.fptr(this), L.fptr(y)))
        return this, y

    def forward_current(self, want_loss=True):
        F = self.lib.npe_num_foci(self.h)
        pred = np.empty((F, L.OBJ_DIM), np.float32)
        loss3 = np.empty(3, np.float32) if want_loss else None
        self._ck(self.lib.npe_forward(self.h, L.fptr(pred), L.fptr(loss3)))
        return pred, loss3

    def forward_per_example_current(self):
        F = self.lib.npe_num_foci(self.h)
        pred = np.empty((F, L.OBJ_DIM), np.float32); le = np.empty((F, 3), np.float32)
        self._ck(self.lib.npe_forward_per_example(self.h, L.fptr(pred), L.fptr(le)))
        return pred, le

    def batch_stats(self):
        f, p = C.c_int64(), C.c_int64()
        self._ck(self.lib.npe_batch_stats(self.h, C.byref(f), C.byref(p)))
        return f.value, p.value

    # ---- rollout (eval.lua:237-454) ---------------------------------------------------------------------------
    def rollout(self, t0, steps, use_gt=True, teacher=False, want_traj=True, want_metrics=True, traj_out=None):
        """traj_out: optional caller-owned (S, Nmax, steps, 22) float32 array for the trajectories, e.g. page-locked memory
        from lib.pinned_array (the device-to-host copy then runs at link speed instead of through pageable staging)."""
        S, Nmax, T = self.scene_shape
        if traj_out is not None:
            assert want_traj and traj_out.shape == (S, Nmax, steps, L.OBJ_DIM) and traj_out.dtype == np.float32 and traj_out.flags.c_contiguous
            traj = traj_out
        else:
            traj = np.empty((S, Nmax, steps, L.OBJ_DIM), np.float32) if want_traj else None
        met = np.zeros((steps, 7), np.float64) if want_metrics else None
        mp = met.ctypes.data_as(C.POINTER(C.c_double)) if met is not None else None
        self._ck(self.lib.npe_rollout(self.h, t0, steps, int(use_gt), int(teacher), L.fptr(traj), mp))
        return traj, met

    def rollout_slots(self, t0, steps, use_gt=True, teacher=False, want_traj=True):
        """rollout() with the metric sums kept per object slot: (traj, slot_metrics (steps,Nmax,7))."""
        S, Nmax, T = self.scene_shape
        traj = np.empty((S, Nmax, steps, L.OBJ_DIM), np.float32) if want_traj else None
        met = np.zeros((steps, Nmax, 7), np.float64)
        self._ck(self.lib.npe_rollout_slots(self.h, t0, steps, int(use_gt), int(teacher), L.fptr(traj),
                                            met.ctypes.data_as(C.POINTER(C.c_double))))
        return traj, met

    def rollout_resident(self, t0, steps):
        n = C.c_int64()
        self._ck(self.lib.npe_rollout_resident(self.h, t0, steps, C.byref(n)))
        return n.value

    def set_precision(self, mode):
        """'fp32' (FFMA kernels, default), 'tf32' (tcgen05, single pass) or 'tf32x3' (tcgen05, FP32-grade split)."""
        self._ck(self.lib.npe_set_precision(self.h, {"fp32": 0, "tf32": 1, "tf32x3": 2}[mode]))

    def set_option(self, name, value):
        """Kernel-dispatch switches (include/npe_cuda.h NPE_OPT_*): 'train_fwd_tc', 'train_stash', 'bwd_tc',
        'peer_timeout_ms', 'small_batch_max', 'dec_stacked'."""
        self._ck(self.lib.npe_set_option(self.h, L.OPTIONS[name], int(value)))

    def debug_read(self, which, shape):
        """Intermediate device buffer of the current batch (tests): 'hact', 'dzact', 'dact', 'ddz', 'ds', 'hp'."""
        out = np.empty(shape, np.float32)
        self._ck(self.lib.npe_debug_read(self.h, {"hact": 0, "dzact": 1, "dact": 2, "ddz": 3, "ds": 4, "hp": 5, "states": 6, "roll": 7}[which],
                                         L.fptr(out), out.size))
        return out

    def set_rollout_mode(self, mode):
        """'persistent' (one kernel, window on chip) or 'kernels' (window in HBM, 5 launches per step)."""
        self._ck(self.lib.npe_set_rollout_mode(self.h, {"persistent": 0, "kernels": 1}[mode]))

    # ---- device-side priority sampler (priority_sampler.lua) ---------------------------------------------------
    def priority_init(self, table, num_batches, alpha=0.1):
        self._ck(self.lib.npe_priority_init(self.h, table, num_batches, alpha))

    def priority_update(self, table, ids, weights=None):
        """weights None: the per-step losses the last train_steps call left on the device (step i -> ids[i])."""
        ids = L.i32(np.atleast_1d(ids))
        w = None if weights is None else L.f32(np.atleast_1d(weights))
        self._ck(self.lib.npe_priority_update(self.h, table, L.iptr(ids), L.fptr(w), ids.size))

    def priority_sample(self, table, pow_=1.0, seed=0, n=1):
        out = np.empty(n, np.int32)
        self._ck(self.lib.npe_priority_sample(self.h, table, pow_, int(seed) & (2 ** 64 - 1), n, L.iptr(out)))
        return out

    def priority_hardest(self, table):
        w = C.c_float(); i = C.c_int32()
        self._ck(self.lib.npe_priority_hardest(self.h, table, C.byref(w), C.byref(i)))
        return [float(w.value), int(i.value)]

    # ---- misc -------------------------------------------------------------------------------------------------
    def sync(self):
        self._ck(self.lib.npe_sync(self.h))

    def event_record(self, slot):
        self._ck(self.lib.npe_event_record(self.h, slot))

    def event_elapsed_ms(self, a, b):
        ms = C.c_float()
        self._ck(self.lib.npe_event_elapsed_ms(self.h, a, b, C.byref(ms)))
        return ms.value

    def profile_select(self, kernel, period=1):
        self._ck(self.lib.npe_profile_period(self.h, period))
        self._ck(self.lib.npe_profile_select(self.h, -1 if kernel is None else L.KERNEL_IDS[kernel]))

    def profile_read(self):
        n, ms = C.c_int32(), C.c_double()
        self._ck(self.lib.npe_profile_read(self.h, C.byref(n), C.byref(ms)))
        return n.value, ms.value

    def launch_count(self):
        return int(self.lib.npe_launch_count(self.h))

    def comm_init(self, rank, nranks, id_bytes):
        buf = (C.c_char * 128).from_buffer_copy(bytes(id_bytes))
        self._ck(self.lib.npe_comm_init(self.h, rank, nranks, C.cast(buf, C.c_void_p)))

    def unique_id(self):
        buf = (C.c_char * 128)()
        rc = self.lib.npe_comm_unique_id(C.cast(buf, C.c_void_p))
        if rc != 0:
            raise L.NPEError(rc, self.lib.npe_last_error(None).decode())
        return bytes(buf)

    def peer_export(self):
        buf = (C.c_char * 64)()
        self._ck(self.lib.npe_peer_export(self.h, C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def peer_attach(self, rank, nranks, all_handles):
        blob = bytes(all_handles)
        assert len(blob) == 64 * nranks
        buf = (C.c_char * len(blob)).from_buffer_copy(blob)
        self._ck(self.lib.npe_peer_attach(self.h, rank, nranks, C.cast(buf, C.c_void_p)))

    def attach_data_parallel(self, dist, rank, world, mode="p2p"):
        """Joins the data-parallel group of a one-process-per-GPU launch.  `dist` is torch.distributed with any CPU
        backend (gloo): it only carries the 64-byte CUDA-IPC handles (mode 'p2p': gradient exchange over NVLink peer
        memory inside the reduce + optimiser kernel) or the 128-byte NCCL id (mode 'nccl')."""
        if world <= 1:
            return
        import torch
        if mode == "nccl":
            idt = torch.zeros(128, dtype=torch.uint8)
            if rank == 0:
                idt = torch.tensor(list(self.unique_id()), dtype=torch.uint8)
            dist.broadcast(idt, 0)
            self.comm_init(rank, world, bytes(idt.tolist()))
        else:
            mine = torch.tensor(list(self.peer_export()), dtype=torch.uint8)
            allh = [torch.zeros(64, dtype=torch.uint8) for _ in range(world)]
            dist.all_gather(allh, mine)
            self.peer_attach(rank, world, bytes(torch.cat(allh).tolist()))

    def allreduce_grads(self):
        self._ck(self.lib.npe_allreduce_grads(self.h))


def slot_metrics_to_batch_row(slot_met, n_examples):
    """One batch's row of the *_through_time_all_batches tables (eval.lua:337-358,385-391) from per-slot sums
    (steps,Nmax,7): sum over slots of the mean over valid examples, divided by the summed valid fractions."""
    with np.errstate(divide="ignore", invalid="ignore"):
        nv, na = slot_met[:, :, 3], slot_met[:, :, 6]
        def col(k, n):
            means = np.where(n > 0, slot_met[:, :, k] / n, 0.0)
            # a slot with valid foci but no valid angle mask contributes 0/0 = nan in the reference as well
            means = np.where((n == 0) & (nv > 0), np.nan, means) if n is na else means
            return means.sum(1) / (n.sum(1) / n_examples)
        return {"loss": col(0, nv), "vel": col(1, nv), "angvel": col(2, nv), "cos": col(4, na), "mag": col(5, na)}


def rollout_metrics_to_logs(met):
    """(steps,7) sums -> the per-step columns of gt_divergence.log (eval.lua:437-452)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return {"loss": met[:, 0] / met[:, 3], "vel": met[:, 1] / met[:, 3], "angvel": met[:, 2] / met[:, 3],
                "cos": met[:, 4] / met[:, 6], "mag": met[:, 5] / met[:, 6]}
