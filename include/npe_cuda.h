/*
 * libnpe_cuda.so -- C ABI of the B200-native Neural Physics Engine hot path.
 *
 * This is the drop-in boundary for the NPE model that `th main.lua -model npe -nbrhd -rs` and
 * `th eval.lua -mode sim` build in mbchang/dynamics.  Every entry point names the reference interface it replaces
 * (file:line under /root/reference/src/lua).  Plain C: opaque handle, plain pointers and sizes; callable from LuaJIT
 * `ffi.cdef`, Python `ctypes` and C.  See INTEGRATION.md for the Lua-side binding.
 *
 * Ownership: the caller owns every host buffer; the library owns all device memory behind the handle.
 * Errors:    0 = NPE_OK, negative npe_status otherwise; message via npe_last_error().  Never aborts/throws.
 * Threading: a handle is bound to one device and one CUDA stream and is not thread-safe; handles are independent.
 * Async:     calls enqueue work on the handle's stream; any call that returns host data synchronises the stream.
 *
 * All tensors are float32, C-contiguous.  The object state vector has 22 floats (config.lua:25):
 *   [px/800, py/800, vx/60, vy/60, a/pi, av/pi, mass1hot(4), oid1hot(3: ball,obstacle,block), size1hot(3), g(2), f(2), p(2)]
 */
#ifndef NPE_CUDA_H
#define NPE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPE_ABI_VERSION 1

typedef enum npe_status {
    NPE_OK = 0,
    NPE_ERR_INVALID = -1,     /* bad argument */
    NPE_ERR_CUDA = -2,        /* CUDA runtime error (message has the cudaError string) */
    NPE_ERR_UNSUPPORTED = -3, /* configuration outside the north-star path (e.g. rnn_dim != 50) */
    NPE_ERR_STATE = -4,       /* call order violated (e.g. backward before forward) */
    NPE_ERR_NCCL = -5,        /* NCCL error or libnccl not loadable */
    NPE_ERR_NOMEM = -6
} npe_status;

#define NPE_OBJ_DIM 22
#define NPE_NUM_PAST 2
#define NPE_INPUT_DIM 44
#define NPE_RNN_DIM 50
#define NPE_LAYERS 5
#define NPE_NUM_PARAMS 28222 /* flattened getParameters() length, layout below */
#define NPE_MAX_OBJECTS 64   /* objects per scene (contexts per focus <= 63) */

/*
 * Flat parameter layout == nn.Module:getParameters() of init_network (npe.lua:14-61, modules.lua:10-26,46-88) with
 * -nbrhd (no bias on encoder / pairwise layers), weights (out,in) row-major, offsets in floats:
 *       0 enc.this.W[25x44]   1100 enc.ctx.W[25x44]
 *    2200 pair.L1.W[50x50]  4700 L2  7200 L3  9700 L4  12200 L5
 *   14700 dec.L1.W[50x94] 19400 b[50]   (input = [pair_sum(50) | this_past(44)], modules.lua:55)
 *   19450 dec.L2.W[50x50] 21950 b   22000 dec.L3.W 24500 b   24550 dec.L4.W 27050 b
 *   27100 dec.L5.W[22x50] 28200 b[22]
 */

typedef struct npe_config {
    int32_t struct_size; /* = sizeof(npe_config), ABI guard */
    int32_t device;      /* CUDA device ordinal */
    /* model hyper-parameters (mp.* in main.lua:36-44); only the north-star values are supported */
    int32_t rnn_dim;    /* 50  (main.lua:37) */
    int32_t layers;     /* 5   (main.lua:40) */
    int32_t num_past;   /* 2   (main.lua:43) */
    int32_t num_future; /* 1   (main.lua:44) */
    int32_t object_dim; /* 22  (main.lua:120) */
    int32_t nbrhd;      /* 1   (-nbrhd, main.lua:38) */
    float nbrhdsize;    /* 3.5 (main.lua:39); threshold = nbrhdsize * 60 px (config.lua:34, npe.lua:180-186) */
    int32_t max_ctx;    /* 12 nearest contexts kept (data_sampler.lua:168, data_process.lua:59); <= 0: no trim */
    float vlambda;      /* 1 (main.lua:56) */
    float lambda;       /* 1 (main.lua:57) */
    int32_t ball_zero_mode; /* rollout: 1 = zero a/av of ball foci (eval.lua:176-178, per object), 0 = never */
    int32_t reserved[4];
} npe_config;

typedef struct npe_handle npe_handle;

/* Fills *cfg with the north-star defaults (main.lua:36-57). */
void npe_default_config(npe_config* cfg);
int npe_abi_version(void);

/* model.create (npe.lua:72-102): allocates parameters (zero), gradient, optimiser state on cfg->device. */
int npe_create(const npe_config* cfg, npe_handle** out);
int npe_destroy(npe_handle* h);
/* Last error message of this handle (or of npe_create when h == NULL). */
const char* npe_last_error(const npe_handle* h);

/* model.theta.params / grad_params (npe.lua:97-98): flat float32 vectors of npe_param_count() elements. */
int npe_param_count(const npe_handle* h);
int npe_params_set(npe_handle* h, const float* host, int32_t n);  /* fp's `params_` copy-in (npe.lua:226) */
int npe_params_get(npe_handle* h, float* host, int32_t n);
int npe_grads_get(npe_handle* h, float* host, int32_t n);         /* return value of model:bp (npe.lua:335) */
int npe_grads_set(npe_handle* h, const float* host, int32_t n);   /* optimizer(feval, x, state) boundary: external grads */

/*
 * Batch form == the reference batch tuple {this, context, y, _, mask} (data_sampler.lua:190; npe.lua:120-132).
 *   this_past (B,2,22), context (B,C,2,22), y (B,22) relative target or NULL (forward only).
 * C may exceed max_ctx: the pair builder then trims to the max_ctx nearest (data_process.lua:51-70).
 */
int npe_batch_upload(npe_handle* h, const float* this_past, const float* context, const float* y, int32_t B, int32_t C);

/*
 * Scene form (the new trajectory batcher): padded device tensor states (S,Nmax,T,22), n_obj (S) objects per scene.
 * Replaces the per-iteration torch.load of a batch file (data_sampler.lua:156-157).
 */
int npe_scenes_upload(npe_handle* h, const float* states, const int32_t* n_obj, int32_t S, int32_t Nmax, int32_t T);
/*
 * The same from RAW matter-js states (S,Nmax,T,12) = {px, py, vx, vy, angle, angularVelocity, mass, objtype (1 ball, 2
 * obstacle, 3 block), sizemul, gravity, friction, pairwise} as json_interface.lua:26-63 reads them: the device does
 * data_process:normalize (data_process.lua:119-139) and properties2onehotall (:184-207) -- the data-prep step before the
 * path -- bit-identically (one rounding per operation).  A value outside its category set returns NPE_ERR_INVALID.
 */
int npe_scenes_upload_raw(npe_handle* h, const float* raw, const int32_t* n_obj, int32_t S, int32_t Nmax, int32_t T);
/*
 * Window schedule: W (scene, t0) windows, uploaded once; past = frames t0,t0+1, target = frame t0+2
 * (split_time, data_sampler.lua:83-101).  npe_select_windows makes windows [first, first+count) the current batch:
 * every non-obstacle object of each window becomes a focus (expand_for_each_object, data_process.lua:254-295),
 * in window-major, object-minor order; targets are made relative (relative_pair, data_process.lua:87-96).
 */
int npe_windows_upload(npe_handle* h, const int32_t* scene_ids, const int32_t* t0, int32_t W);
int npe_select_windows(npe_handle* h, int32_t first, int32_t count);
/*
 * Explicit focus schedule: n (scene, object, t0) examples uploaded once; npe_select_foci makes examples
 * [first, first+count) the current batch.  This reproduces the reference's batch composition exactly (50 scenes, the
 * same focus slot and window offset: data_process.lua:254-295, data_sampler.lua:198-203) with device-resident data.
 */
int npe_foci_upload(npe_handle* h, const int32_t* scene_ids, const int32_t* obj_ids, const int32_t* t0, int32_t n);
int npe_select_foci(npe_handle* h, int32_t first, int32_t count);
/* Number of foci (examples) and context slots per focus of the current batch. */
int npe_num_foci(const npe_handle* h);
int npe_ctx_slots(const npe_handle* h);
/* Copies the current batch's gathered focus rows (F,44) and relative targets (F,22) to the host (parity checks). */
int npe_batch_download(npe_handle* h, float* this_past_out, float* y_out);

/*
 * Pair table + neighbourhood mask (k_nearest_context data_process.lua:51-70; select_neighbors npe.lua:170-206).
 * Outputs (each may be NULL): ctx_idx (F,K) int32 kept context indices ascending, -1 padded (closest_indices);
 * mask (F,K) uint8 neighbourhood mask of each kept slot; cnt (F) kept count.  K = npe_ctx_slots().
 * Bit-exact with the reference (no FMA contraction; ties: smaller distance, then lower index).
 */
int npe_build_pairs(npe_handle* h, int32_t* ctx_idx_out, uint8_t* mask_out, int32_t* cnt_out);

/*
 * model:fp (npe.lua:225-247).  pred_out (F,22) or NULL; loss3_out = {loss, vel_loss, angvel_loss} or NULL.
 * Builds the pair table if the batch changed.
 */
int npe_forward(npe_handle* h, float* pred_out, float* loss3_out);
/* model:fp_batch (npe.lua:250-284): per-example losses loss_out (F,3) = {loss_i, vel_i, angvel_i}. */
int npe_forward_per_example(npe_handle* h, float* pred_out, float* loss_out);
/*
 * model:bp (npe.lua:290-336): gradient of the reference-scaled objective into grad_params.  divisor_batch = the B of
 * nElement() (npe.lua:307,315); 0 = this batch's F; data-parallel callers pass the GLOBAL batch.
 * Must follow npe_forward on the same batch.
 */
int npe_backward(npe_handle* h, int32_t divisor_batch);

/* optim.adam / optim.rmsprop on the flat vector (main.lua:135-144,301-302); state lives on the device. */
int npe_adam_step(npe_handle* h, float lr, float beta1, float beta2, float eps);
int npe_rmsprop_step(npe_handle* h, float lr, float alpha, float eps);
/* Clears optimiser state; rmsprop_m_init: initial mean-square (0 = early optim rock, 1 = later versions). */
int npe_optim_reset(npe_handle* h, float rmsprop_m_init);
/* feval_train + optimizer in one call (main.lua:247-269,301-302) on the current batch: forward, backward,
 * (all-reduce when a communicator is attached), Adam (opt=0) or RMSprop (opt=1).  loss3_out may be NULL (no sync). */
int npe_train_step(npe_handle* h, int32_t opt, float lr, int32_t divisor_batch, float* loss3_out);

/*
 * The inner loop of train() (main.lua:299-304) over a device-resident schedule: step s trains on examples
 * [first + s*batch, first + (s+1)*batch) of the uploaded focus/window schedule.  One host call enqueues all `nsteps`
 * steps (no host round trip between steps).  losses_out (nsteps,3) or NULL: per-step {loss, vel_loss, angvel_loss},
 * copied back once at the end.  With the peer-memory exchange attached (npe_peer_attach) the steps are captured into CUDA
 * graphs of 8 steps each: same launches, same results (a stream launch that wrote into a peer's memory pays a
 * system-scope flush at its end, a graph node does not).
 */
int npe_train_steps(npe_handle* h, int32_t first, int32_t batch, int32_t nsteps, int32_t opt, float lr,
                    int32_t divisor_batch, float* losses_out);

/*
 * npe_train_step without the host wait: the step is enqueued, its {loss, vel_loss, angvel_loss} land in ring slot `slot`
 * (0..63) of host-mapped memory and npe_loss_wait(slot) returns them once that step has finished.  A caller that reads
 * the loss of step s-1 after enqueueing step s (trainLogger / update_batch_weight one iteration late, main.lua:306)
 * keeps the GPU queue non-empty across iterations.
 */
int npe_train_step_async(npe_handle* h, int32_t opt, float lr, int32_t divisor_batch, int32_t slot);
int npe_loss_wait(npe_handle* h, int32_t slot, float* loss3_out);
/* feval_train + optimiser (main.lua:247-269,301-302) of one host batch in one call: npe_batch_upload followed by
 * npe_train_step_async. */
int npe_train_batch_async(npe_handle* h, const float* this_past, const float* context, const float* y, int32_t B, int32_t C,
                          int32_t opt, float lr, int32_t divisor_batch, int32_t slot);

/*
 * simulate_all (eval.lua:237-454) in scene form on the uploaded scenes: past = frames t0,t0+1; `steps` autoregressive
 * steps.  use_gt != 0: frames t0+2.. are ground truth (obstacles copied from it, metrics computed against it);
 * otherwise obstacles stay static and metrics are zero.
 *   traj_out (S,Nmax,steps,22) or NULL;  metrics_out (steps,7) or NULL:
 *   {sum loss_i, sum vel_i, sum angvel_i, n_valid, sum cos, sum relmag, n_anglemask}  (eval.lua:337-356, infer.lua:32-74)
 * teacher != 0: feed ground truth back instead of the predictions (teacher forcing; traj_out holds one-step preds).
 */
int npe_rollout(npe_handle* h, int32_t t0, int32_t steps, int32_t use_gt, int32_t teacher, float* traj_out,
                double* metrics_out);
/*
 * npe_rollout with the metric sums kept per object slot: slot_metrics_out (steps,Nmax,7), same 7 columns.  This is what
 * simulate_all's per-batch average needs when some slots hold obstacles in only some examples: it sums the per-slot
 * MEANS over valid examples and divides by the summed valid FRACTIONS (eval.lua:337-358,385-391).
 */
int npe_rollout_slots(npe_handle* h, int32_t t0, int32_t steps, int32_t use_gt, int32_t teacher, float* traj_out,
                      double* slot_metrics_out);
/* Device-resident rollout for throughput runs: no host copies; object-steps advanced are returned. */
int npe_rollout_resident(npe_handle* h, int32_t t0, int32_t steps, int64_t* object_steps_out);

/*
 * Arithmetic variant of the forward path (north star: FP32 FFMA variant and tf32 tcgen05 variant):
 *   NPE_PRECISION_FP32 (default): FP32 FFMA kernels everywhere (the parity path, 1e-4).
 *   NPE_PRECISION_TF32_TC: npe_forward / npe_forward_per_example / npe_rollout run the tcgen05 kernels (TF32 inputs,
 *   FP32 accumulation in TMEM; ~1e-3 relative).  Backward, optimiser, pair table and masks are unaffected.
 * Rollout mode: NPE_ROLLOUT_PERSISTENT (default for FP32: one persistent kernel, window on chip) or
 * NPE_ROLLOUT_BY_KERNELS (window in HBM, 5 launches per step; always used by the tcgen05 variant).
 */
#define NPE_PRECISION_FP32 0
#define NPE_PRECISION_TF32_TC 1
#define NPE_PRECISION_TF32X3_TC 2 /* tcgen05 with the 3xTF32 split, activations in tensor memory: FP32-grade (~1e-6) */
#define NPE_ROLLOUT_PERSISTENT 0
#define NPE_ROLLOUT_BY_KERNELS 1
int npe_set_precision(npe_handle* h, int32_t mode);
int npe_set_rollout_mode(npe_handle* h, int32_t mode);
/*
 * Kernel-dispatch switches of the training step (model:fp + model:bp, npe.lua:225-336).  Every combination computes the
 * same function; they exist so that each kernel family can be compared with the oracle and with the others.
 *   NPE_OPT_TRAIN_FWD_TC     1 (default): large batches run the pair forward on the FP32-grade 3xTF32 tcgen05 kernel; 0: FFMA
 *   NPE_OPT_TRAIN_STASH      1 (default): the forward writes the pair activations for the backward; 0: the backward recomputes
 *   NPE_OPT_BWD_TC           1 (default): large batches run input- and weight-gradients on tcgen05 (3xTF32); 0: FFMA kernels
 *   NPE_OPT_PEER_TIMEOUT_MS  how long the data-parallel step waits for a peer's gradient before failing (default 10000)
 *   NPE_OPT_SMALL_BATCH_MAX  foci up to which the 16-row latency mapping is used (-1 = default, 4 x SM count)
 *   NPE_OPT_DEC_STACKED      decoder forward of inference / rollout with stacked hi | lo weight images (2 MMAs per K-step):
 *                            0 = automatic (from 8 tiles per SM on), 1 = always, 2 = never
 */
#define NPE_OPT_TRAIN_FWD_TC 0
#define NPE_OPT_TRAIN_STASH 1
#define NPE_OPT_BWD_TC 2
#define NPE_OPT_PEER_TIMEOUT_MS 3
#define NPE_OPT_SMALL_BATCH_MAX 4
#define NPE_OPT_DEC_STACKED 5
int npe_set_option(npe_handle* h, int32_t option, int32_t value);

/* Data-parallel: one handle per GPU/process.  id = 128-byte ncclUniqueId from npe_comm_unique_id on rank 0. */
int npe_comm_unique_id(void* id128_out);
int npe_comm_init(npe_handle* h, int32_t rank, int32_t nranks, const void* id128);
int npe_allreduce_grads(npe_handle* h); /* sum of flat gradients (+3 loss sums) over ranks, in place */
/*
 * Data-parallel step over NVLink peer memory (replaces reduce + ncclAllReduce + optimiser with ONE kernel): every rank
 * owns an exchange region of 8-byte {gradient, step stamp} cells that the peers map through CUDA IPC; the reduce +
 * optimiser kernel of each rank pushes its reduced gradient into every peer's region (posted NVLink stores), waits for
 * the peers' cells in its own region, sums them in rank order (replicas stay bit-identical) and applies Adam / RMSprop.
 *   npe_peer_export: writes this rank's 64-byte IPC handle.  The launcher all-gathers the handles (nranks * 64 bytes).
 *   npe_peer_attach: opens every peer's region.  After it, npe_train_step uses the fused path (NCCL stays available
 *   through npe_allreduce_grads).
 */
#define NPE_PEER_HANDLE_BYTES 64
int npe_peer_export(npe_handle* h, void* handle64_out);
int npe_peer_attach(npe_handle* h, int32_t rank, int32_t nranks, const void* all_handles);

/*
 * Device-side priority sampler (priority_sampler.lua:4-54; main.lua:249-253 when -rs is off): the table of the last loss
 * per batch lives on the device, one table (0..15) per data sampler / dataset folder.  init: num_batches zero weights, `alpha` = probability of a uniform draw (0.1 in the
 * reference).  update: weight[ids[i]] = weights[i] (ids 1-based; update_batch_weight, :18-22); weights == NULL takes the
 * per-step losses the last npe_train_steps call left on the device (step i -> ids[i]).  sample: n draws (uniform with
 * probability alpha, else multinomial over weight^pow, :37-48) from counter-based random numbers of `seed`; fails with
 * NPE_ERR_STATE when a multinomial draw meets a table that is not full (the reference asserts).  hardest: {max, argmax}
 * (get_hardest_batch, :50-54).
 */
int npe_priority_init(npe_handle* h, int32_t table, int32_t num_batches, float alpha);
int npe_priority_update(npe_handle* h, int32_t table, const int32_t* ids, const float* weights, int32_t n);
int npe_priority_sample(npe_handle* h, int32_t table, float pow_, uint64_t seed, int32_t n, int32_t* ids_out);
int npe_priority_hardest(npe_handle* h, int32_t table, float* weight_out, int32_t* id_out);

int npe_sync(npe_handle* h);
/* Page-locked host buffers for callers that stream batches (cudaHostAlloc / cudaFreeHost). */
void* npe_host_alloc(size_t bytes);
void npe_host_free(void* p);
/*
 * Kernel-level timing for roofline reports: kernel_id selects one kernel class (NPE_K_*) whose launches are
 * bracketed with CUDA events on the handle's stream (-1 = off).  npe_profile_read returns the number of bracketed
 * launches and their summed duration since the last npe_profile_select.
 */
#define NPE_K_BUILD_PAIRS 0
#define NPE_K_FORWARD 1
#define NPE_K_DEC_BACKWARD 2
#define NPE_K_PAIR_BACKWARD 3
#define NPE_K_REDUCE 4
#define NPE_K_OPTIM 5
#define NPE_K_ROLLOUT 6
#define NPE_K_TARGETS 7
#define NPE_K_STEP_FUSED 8   /* forward + backward of a small batch as one grid of tiles (k_step_fused) */
#define NPE_K_WGRAD 9        /* tcgen05 weight-gradient kernels of the large-batch backward (k_wgrad_tc, two launches) */
#define NPE_K_COUNT 10
/* Tools: the first call switches on phase time stamps (globaltimer, ns) in k_step_fused; stamps_out (max_ctas, 8) receives
 * the stamps of the last launch: {static loads done, wait done, first hand-over, second hand-over, gradients done,
 * partial stored, 0, 0} per CTA (pair tiles first, then decoder tiles). */
int npe_step_trace(npe_handle* h, uint64_t* stamps_out, int32_t max_ctas);
int npe_profile_select(npe_handle* h, int32_t kernel_id);
/* Bracket only every `period`-th launch of the selected class (default 1): an event between two kernels disables their
 * programmatic-dependent-launch overlap, so a long timed region samples instead of bracketing every launch. */
int npe_profile_period(npe_handle* h, int32_t period);
int npe_profile_read(npe_handle* h, int32_t* launches_out, double* total_ms_out);
/* CUDA-event timing on the handle's stream (bench/roofline): record marks, read elapsed ms between two marks. */
int npe_event_record(npe_handle* h, int32_t slot);
int npe_event_elapsed_ms(npe_handle* h, int32_t slot_a, int32_t slot_b, float* ms_out);
/*
 * Parity checks of intermediate results (tests only): copies the first nfloats of an internal device buffer of the
 * current batch.  Layouts (layer-major, rows of 56 floats, 52 used; csrc/npe_kernels.cuh): HACT / DZACT [6][rows][56] =
 * h0..h5 / dz0..dz5 of every active pair in list order, rows = pair capacity rounded up to 128; DACT [5][F][56] =
 * {pair sum | 1 | 0, u1..u4}; DDZ [5][F][56] = dz1..dz5; DS [F][52] = gradient into the pair sum; HP [pair][52] = h5.
 */
#define NPE_DBG_HACT 0
#define NPE_DBG_DZACT 1
#define NPE_DBG_DACT 2
#define NPE_DBG_DDZ 3
#define NPE_DBG_DS 4
#define NPE_DBG_HP 5
#define NPE_DBG_STATES 6 /* the resident scene tensor (S,Nmax,T,22) */
#define NPE_DBG_ROLL 7   /* the two-frame window (S,Nmax,2,22) after a rollout by kernels */
int npe_debug_read(npe_handle* h, int32_t which, float* out, int64_t nfloats);

/* The pair builder's form of the neighbourhood test.  The reference keeps context c of a focus when
 * fl(800 * fl(sqrt(s))) <= thr_px on the squared (normalised) distance s (npe.lua:170-206 over data_utils.lua:276-285);
 * both roundings are monotone, so that set is s <= s* for one float s*, which this returns (negative: nothing passes).
 * Pure host arithmetic, no GPU: exported so that the equivalence can be tested exhaustively around the threshold. */
float npe_nbr_threshold_sq(float thr_px);
/* Counters: kernels launched by this handle since creation; active pairs / foci of the current batch. */
int64_t npe_launch_count(const npe_handle* h);
int npe_batch_stats(npe_handle* h, int64_t* foci_out, int64_t* active_pairs_out);

#ifdef __cplusplus
}
#endif
#endif /* NPE_CUDA_H */
