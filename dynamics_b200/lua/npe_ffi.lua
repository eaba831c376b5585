-- LuaJIT FFI binding of libnpe_cuda.so (include/npe_cuda.h).  Authored against the header; there is no LuaJIT/Torch7
-- in the build container, so every acceptance test drives the identical C ABI through Python ctypes
-- (dynamics_b200/lib.py mirrors this cdef one to one).
local ffi = require 'ffi'

ffi.cdef[[
typedef struct npe_config {
    int32_t struct_size; int32_t device;
    int32_t rnn_dim; int32_t layers; int32_t num_past; int32_t num_future; int32_t object_dim; int32_t nbrhd;
    float nbrhdsize; int32_t max_ctx; float vlambda; float lambda; int32_t ball_zero_mode; int32_t reserved[4];
} npe_config;
typedef struct npe_handle npe_handle;
void npe_default_config(npe_config* cfg);
int npe_abi_version(void);
int npe_create(const npe_config* cfg, npe_handle** out);
int npe_destroy(npe_handle* h);
const char* npe_last_error(const npe_handle* h);
int npe_param_count(const npe_handle* h);
int npe_params_set(npe_handle* h, const float* host, int32_t n);
int npe_params_get(npe_handle* h, float* host, int32_t n);
int npe_grads_get(npe_handle* h, float* host, int32_t n);
int npe_grads_set(npe_handle* h, const float* host, int32_t n);
int npe_batch_upload(npe_handle* h, const float* this_past, const float* context, const float* y, int32_t B, int32_t C);
int npe_scenes_upload(npe_handle* h, const float* states, const int32_t* n_obj, int32_t S, int32_t Nmax, int32_t T);
int npe_windows_upload(npe_handle* h, const int32_t* scene_ids, const int32_t* t0, int32_t W);
int npe_select_windows(npe_handle* h, int32_t first, int32_t count);
int npe_foci_upload(npe_handle* h, const int32_t* scene_ids, const int32_t* obj_ids, const int32_t* t0, int32_t n);
int npe_select_foci(npe_handle* h, int32_t first, int32_t count);
int npe_num_foci(const npe_handle* h);
int npe_ctx_slots(const npe_handle* h);
int npe_batch_download(npe_handle* h, float* this_past_out, float* y_out);
int npe_build_pairs(npe_handle* h, int32_t* ctx_idx_out, uint8_t* mask_out, int32_t* cnt_out);
int npe_forward(npe_handle* h, float* pred_out, float* loss3_out);
int npe_forward_per_example(npe_handle* h, float* pred_out, float* loss_out);
int npe_backward(npe_handle* h, int32_t divisor_batch);
int npe_adam_step(npe_handle* h, float lr, float beta1, float beta2, float eps);
int npe_rmsprop_step(npe_handle* h, float lr, float alpha, float eps);
int npe_optim_reset(npe_handle* h, float rmsprop_m_init);
int npe_train_step(npe_handle* h, int32_t opt, float lr, int32_t divisor_batch, float* loss3_out);
int npe_train_step_async(npe_handle* h, int32_t opt, float lr, int32_t divisor_batch, int32_t slot);
int npe_loss_wait(npe_handle* h, int32_t slot, float* loss3_out);
int npe_train_batch_async(npe_handle* h, const float* this_past, const float* context, const float* y, int32_t B, int32_t C, int32_t opt, float lr, int32_t divisor_batch, int32_t slot);
int npe_train_steps(npe_handle* h, int32_t first, int32_t batch, int32_t nsteps, int32_t opt, float lr, int32_t divisor_batch, float* losses_out);
int npe_rollout(npe_handle* h, int32_t t0, int32_t steps, int32_t use_gt, int32_t teacher, float* traj_out, double* metrics_out);
int npe_rollout_slots(npe_handle* h, int32_t t0, int32_t steps, int32_t use_gt, int32_t teacher, float* traj_out, double* slot_metrics_out);
int npe_rollout_resident(npe_handle* h, int32_t t0, int32_t steps, int64_t* object_steps_out);
int npe_set_precision(npe_handle* h, int32_t mode);
int npe_set_rollout_mode(npe_handle* h, int32_t mode);
int npe_comm_unique_id(void* id128_out);
int npe_comm_init(npe_handle* h, int32_t rank, int32_t nranks, const void* id128);
int npe_allreduce_grads(npe_handle* h);
int npe_peer_export(npe_handle* h, void* handle64_out);
int npe_peer_attach(npe_handle* h, int32_t rank, int32_t nranks, const void* all_handles);
int npe_sync(npe_handle* h);
void* npe_host_alloc(size_t bytes);
void npe_host_free(void* p);
int npe_step_trace(npe_handle* h, uint64_t* stamps_out, int32_t max_ctas);
int npe_profile_select(npe_handle* h, int32_t kernel_id);
int npe_profile_period(npe_handle* h, int32_t period);
int npe_profile_read(npe_handle* h, int32_t* launches_out, double* total_ms_out);
int npe_event_record(npe_handle* h, int32_t slot);
int npe_event_elapsed_ms(npe_handle* h, int32_t slot_a, int32_t slot_b, float* ms_out);
int64_t npe_launch_count(const npe_handle* h);
int npe_batch_stats(npe_handle* h, int64_t* foci_out, int64_t* active_pairs_out);
]]

local C = ffi.load(os.getenv('NPE_CUDA_LIB') or 'npe_cuda')   -- no fallback: a missing library is an error

local M = {C = C, ffi = ffi}

-- converts a negative npe_status into a Lua error carrying npe_last_error (never swallow: there is no CPU path)
function M.check(h, rc)
    if rc ~= 0 then
        error(string.format('libnpe_cuda error %d: %s', rc, ffi.string(C.npe_last_error(h))), 2)
    end
end

-- FloatTensor -> const float* (the tensor must be contiguous; callers :contiguous() first)
function M.fptr(t)
    assert(t:type() == 'torch.FloatTensor' and t:isContiguous(), 'expected a contiguous FloatTensor')
    return ffi.cast('float*', torch.data(t))
end

return M
