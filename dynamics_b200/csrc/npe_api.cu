// libnpe_cuda.so: C ABI over the sm_100a NPE kernels (see include/npe_cuda.h).
// Host side: handle/state management, buffer growth, launch sequencing on one stream.  No CPU fallback anywhere:
// every compute entry point launches CUDA kernels or returns an error.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <tuple>
#include <vector>

#include "../../include/npe_cuda.h"
#include "npe_kernels.cuh"
#include "npe_step16.cuh"
#include "npe_rollout.cuh"
#include "npe_tc.cuh"
#include "npe_tcbwd.cuh"
#include "npe_chain.cuh"
#include "npe_wgrad.cuh"

using namespace npe;

namespace {

thread_local std::string g_create_error;

// ---- NCCL through dlopen (libnccl.so.2; the torch-bundled copy is reused when already loaded) ----
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) { err = std::string("dlopen libnccl.so.2 failed: ") + dlerror(); return false; }
        GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
        CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
        CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
        AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
        GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
        if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllReduce) { err = "libnccl symbols missing"; return false; }
        return true;
    }
};
NcclApi g_nccl;
constexpr int NCCL_FLOAT32 = 7, NCCL_SUM = 0;

template <typename T>
struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t ensure(size_t n) {
        if (n <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = n + n / 4 + 64;
        cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

constexpr int LOSS_RING = 64;

struct npe_handle {
    npe_config cfg;
    int device = 0, num_sms = 0;
    cudaStream_t stream = nullptr;
    std::string err;
    int64_t launches = 0;
    cudaEvent_t events[16] = {};

    // model + optimiser state
    float *params = nullptr, *grads = nullptr, *adam_m = nullptr, *adam_v = nullptr, *rms_m = nullptr;
    float* allreduce_buf = nullptr;   // NPARAM + 4 floats: gradient + {sum_v, sum_w, F}
    int adam_t = 0;
    float rms_m_init = 0.f;
    float* partial = nullptr;         // [num_sms][PT_TOTAL], tile-major (npe_layout.h)
    float* wpack = nullptr;           // packed (shared-memory layout) weight image, kept in sync with params
    float* wtc = nullptr;             // tensor-core (K-major canonical layout, TF32-rounded) weight image
    bool wtc_dirty = true;
    float* wtc3 = nullptr;            // hi | lo images for the 3xTF32 variant (2 * TC_TOTAL floats)
    bool wtc3_dirty = true;
    float* wenc = nullptr;            // 64-row encoder images of the chain kernels (hi T | hi C | lo T | lo C), refreshed with wtc3
    float* wtt3 = nullptr;            // hi | lo TRANSPOSED images for the tcgen05 input-gradient chains (2 * TT_TOTAL floats)
    bool wtt3_dirty = true;
    DevBuf<float> dzact, dact, ddz;   // tcgen05 backward: pair gradient stash, decoder activation / gradient stashes
    // device-side priority sampler (priority_sampler.lua): last loss per batch, scan scratch, draw output
    static constexpr int PRIO_TABLES = 16;                 // one table per data_sampler (dataset folder)
    DevBuf<float> prio_w[PRIO_TABLES], prio_vals, prio_out2;
    DevBuf<double> prio_cum;
    DevBuf<int> prio_ids;
    int prio_nb[PRIO_TABLES] = {};
    float prio_alpha[PRIO_TABLES] = {};
    DevBuf<long long> pair_frow;      // focus-window offset of every active pair (encoder weight gradient)
    DevBuf<u64> hmask, umask;         // ReLU bits of h0..h5 per active pair [6][pair_cap] / of u1..u4 per focus [4][F]
    bool dact_valid = false;          // the decoder stash holds the current batch (forward ran with the tcgen05 backward in mind)
    long long* chain_prof_dev = nullptr;
    long long* dp_prof_dev = nullptr;    // tools only: k_dp_step stamps (NPE_CHAIN_PROF)
    cudaGraphExec_t step_exec[4] = {nullptr, nullptr, nullptr, nullptr};   // npe_train_steps: ring of executable graphs (updated in place)
    cudaEvent_t step_exec_ev[4] = {nullptr, nullptr, nullptr, nullptr};    // ... and the end of each one's last launch
    int step_exec_next = 0;
    cudaGraph_t step_graph[4] = {nullptr, nullptr, nullptr, nullptr};      // the captured graphs the executables came from (own the node handles)
    std::vector<cudaGraphNode_t> step_nodes[4];                            // their kernel nodes in launch order
    int step_exec_opt[4] = {-1, -1, -1, -1};                               // optimiser the slot was captured with
    // launch_pdl modes inside npe_train_steps: 0 = launch, 1 = launch into a capturing stream and note the node, 2 = no launch:
    // write the kernel's parameters into node sink_idx of an executable graph (cudaGraphExecKernelNodeSetParams)
    int sink_mode = 0, sink_slot = 0, sink_idx = 0;
    bool sink_ok = true;
    long long graphs_captured = 0, graphs_set = 0, graphs_replayed = 0;      // statistics (NPE_CHAIN_PROF)
    double graph_fill_us = 0.0, graph_launch_us = 0.0;                        // host time of the re-parameterised groups: steps, launch
    int* sched_host = nullptr;             // page-locked staging of a chunk's {first focus, foci} per step (sched_build)
    bool graph_warm[2] = {false, false};   // npe_train_steps: a step with this optimiser has run outside a stream capture
    long long stash_rows = 0;         // row capacity of one slot of the layer-major pair stashes (pair_cap rounded up to 128)
    int precision = 0, rollout_mode = 0;
    int train_fwd_tc = 1;             // large training batches: pair forward on the 3xTF32 tcgen05 kernel
    int train_stash = 1;              // forward writes the pair activations for the backward kernels (0: recompute)
    int bwd_tc = 1;                   // large training batches: backward (input + weight gradients) on tcgen05
    int small_batch_max = -1;         // foci up to which the 16-row latency mapping is used (-1: 4 * SM count)
    int dec_stacked = 0;              // NPE_OPT_DEC_STACKED: 0 automatic, 1 always, 2 never
    int pdl_max_foci = 592;           // batches up to this many foci chain their kernels with programmatic dependent launch
    DevBuf<float> roll;               // rollout-by-kernels window (S, Nmax, 2, 22)
    DevBuf<int> roll_scene, roll_obj, roll_t0, foc_of_obj;
    DevBuf<float2> roll_pos;          // fused rollout: last-frame positions per focus, written by the decoder kernel for the next builder
    int roll_F = 0;
    float* loss3_dev = nullptr;
    double* loss_sums_dev = nullptr;

    // scenes / foci
    DevBuf<float> states, raw_stage;
    DevBuf<int> n_obj;
    int S = 0, Nmax = 0, T = 0;
    std::vector<std::vector<int>> scene_foci;   // non-obstacle objects per scene
    int64_t scene_pred_objects = 0;
    DevBuf<int> foc_scene, foc_obj, foc_t0;
    std::vector<int> win_start;                 // prefix of foci per window (size W + 1)
    int foc_first = 0, F = 0;
    bool batch_form = false;
    bool have_y = false;
    const float* y_view = nullptr;    // batch form: the targets sit behind the packed scenes in `states` (no copy)
    const float* yptr() const { return y_view ? y_view : y.p; }

    // per-batch buffers
    DevBuf<float> y, pred, loss_ex, ds, this_rows, Hp, hact;
    bool hact_valid = false;          // the activation stash holds the current batch's pair activations
    DevBuf<u64> keep, nbr;
    DevBuf<int> cnt, block_sums, pair_start, pair_foc;   // cnt = CTA-local exclusive prefix of active counts
    DevBuf<long long> foc_row, pair_crow;
    int pair_cap = 0;
    unsigned* step_flags = nullptr;   // k_step_fused: per-tile arrival flags (2 x SM count), stamped per launch
    unsigned step_stamp = 0;
    unsigned long long* step_trace = nullptr;   // npe_step_trace: device buffer of phase time stamps (tools only)
    int* info_dev = nullptr;          // {P, overflow}
    // Batch form is double-buffered: npe_batch_upload copies the batch and builds its pair table on `stream2` into the
    // slot that is NOT in use, so an upload overlaps the training kernels of the previous batch (npe_train_step_async).
    // `alt` holds the other slot's buffers; slot_ready is signalled on stream2, slot_free on the main stream.
    struct BatchSlot {
        DevBuf<float> states;
        DevBuf<u64> keep, nbr;
        DevBuf<int> cnt, block_sums, pair_start, pair_foc;
        DevBuf<long long> foc_row, pair_crow;
        DevBuf<int> n_obj, foc_scene, foc_obj, foc_t0;   // per slot: batches of different object counts alternate
        int foci_B = -1, foci_N = -1;
        int* info_dev = nullptr;
        u64* active_total = nullptr;
    } alt;
    // Scene form, throughput regime: the pair tables are double-buffered.  npe_select_windows builds the tables of the batch
    // it selects at once, on `stream2`, into the table set that the step in flight does NOT use; the builder is gated by an
    // event recorded just before that step's reduce + optimiser launch, so it runs BESIDE that (light) kernel instead of
    // after it -- and not under the chain / weight-gradient kernels, which own every SM (tried: it only delays them).
    struct TableSet {
        DevBuf<u64> keep, nbr;
        DevBuf<int> cnt, block_sums, pair_start, pair_foc;
        DevBuf<long long> foc_row, pair_crow;
        DevBuf<float> y;
        int* info_dev = nullptr;
        u64* active_total = nullptr;
    } tab_alt;
    cudaEvent_t tab_free[2] = {nullptr, nullptr}, tab_ready = nullptr, pre_reduce = nullptr;
    bool tab_free_valid[2] = {false, false}, pre_reduce_valid = false;
    int tab_cur = 0;
    cudaStream_t stream2 = nullptr;
    cudaEvent_t slot_ready = nullptr;          // upload + pair table of the current slot complete (stream2)
    cudaEvent_t wg_fork = nullptr, wg_join = nullptr;   // large-batch backward: decoder weight gradients on a side stream
    cudaStream_t stream3 = nullptr;
    cudaEvent_t slot_free[2] = {nullptr, nullptr};   // main-stream work on that slot enqueued so far is complete
    int slot = 0;
    bool slot_free_valid[2] = {false, false};
    // batch-form staging: the three host arrays are packed into the scene layout in page-locked memory and shipped
    // with ONE async copy (two buffers alternate; an event guards reuse)
    float* stage[2] = {nullptr, nullptr};
    size_t stage_cap = 0;
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    int stage_idx = 0;
    int foci_B = -1, foci_N = -1;     // shape the iota focus arrays / n_obj were last built for
    DevBuf<float> loss_hist;          // per-step {loss, vel, angvel, overflow} of npe_train_steps
    // pair tables of a whole chunk of scheduled steps (k_build_pairs_sched), sliced per step by npe_train_steps
    struct Sched {
        DevBuf<u64> keep, nbr, total;
        DevBuf<int> cnt, block_sums, pair_start, pair_foc, info, step_first, step_F;
        DevBuf<long long> foc_row, pair_crow;
        DevBuf<float> y;
    } sched;
    float* loss4_override = nullptr;  // when set, the fused loss reduction writes here instead of the mapped buffer
    float* loss4_host = nullptr;      // host-mapped {loss, vel, angvel, overflow}
    float* loss4_devptr = nullptr;
    cudaEvent_t ring_ev[64] = {};     // npe_train_step_async / npe_loss_wait: one event per ring slot (slot i at loss4_host + 16 + 4 i)
    u64* active_total = nullptr;
    bool pairs_valid = false, fwd_valid = false;
    DevBuf<unsigned> scan_count[2];   // throughput pair builder (k_build_pairs_scan): block counts and grid-barrier words,
    unsigned* scan_bar[2] = {nullptr, nullptr};   // one set per stream the builder may run on (main, upload)
    int scan_ctas_per_sm[2] = {0, 0}; // resident CTAs per SM of its two variants (occupancy query, cached)
    unsigned scan_epoch = 0;
    bool keep_valid = false;          // the kept-context bits of the current pair table are complete (API view; see warp_pairs)
    // The last launch on the main stream was a pair builder: the next kernel must not be a programmatic dependent of it
    // (it would read the pair tables before griddepcontrol.wait, and only the wait guarantees the builder's writes are
    // visible).  launch_pdl drops the attribute once and clears the flag.
    bool after_builder = false;
    DevBuf<int> ctx_idx_out, cnt_out;
    DevBuf<unsigned char> mask_out;

    // rollout
    DevBuf<float> traj;
    DevBuf<double> metrics;
    DevBuf<double> slot_metrics;
    bool want_slots = false;   // set by npe_rollout_slots around one rollout

    // kernel-class profiling (roofline): event pairs around launches of one selected kernel class
    int prof_kernel = -1;
    int prof_period = 1;              // bracket every prof_period-th launch of the selected class
    long long prof_seen = 0;
    std::vector<cudaEvent_t> prof_events;   // pairs
    size_t prof_used = 0;

    // comm
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    // NVLink peer exchange (fused DP step)
    float* xregion = nullptr;         // exchange region of k_dp_step ({value, stamp} cells of every source rank)
    void* peer_ptr[MAX_PEERS] = {};
    bool peer_opened[MAX_PEERS] = {};
    bool peer_ready = false;
    unsigned peer_step = 0;
    int peer_timeout_ms = 10000;      // k_dp_step gives up on a peer's gradient cell after this long (NPE_OPT_PEER_TIMEOUT_MS)

    FocusSrc src() const {
        FocusSrc s;
        s.states = states.p; s.n_obj = n_obj.p; s.Nmax = Nmax; s.T = T;
        s.foc_scene = foc_scene.p + foc_first; s.foc_obj = foc_obj.p + foc_first; s.foc_t0 = foc_t0.p + foc_first;
        s.F = F;
        return s;
    }
};

namespace {

int fail(npe_handle* h, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (h) h->err = buf; else g_create_error = buf;
    return code;
}

#define CK(call)                                                                                          \
    do {                                                                                                  \
        cudaError_t e__ = (call);                                                                         \
        if (e__ != cudaSuccess)                                                                           \
            return fail(h, NPE_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define CKL()                                                                                             \
    do {                                                                                                  \
        h->launches++;                                                                                    \
        cudaError_t e__ = cudaGetLastError();                                                             \
        if (e__ != cudaSuccess)                                                                           \
            return fail(h, NPE_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
// Brackets one launch with events when its class is selected.
struct ProfScope {
    npe_handle* h; bool on;
    ProfScope(npe_handle* hh, int kid) : h(hh), on(hh->prof_kernel == kid) {
        if (on) on = (h->prof_seen++ % h->prof_period) == 0;
        if (!on) return;
        if (h->prof_used + 2 > h->prof_events.size()) {
            for (int i = 0; i < 256; ++i) { cudaEvent_t e; cudaEventCreate(&e); h->prof_events.push_back(e); }
        }
        cudaEventRecord(h->prof_events[h->prof_used], h->stream);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(h->prof_events[h->prof_used + 1], h->stream);
        h->prof_used += 2;
    }
};
#define NEED(h)                                      \
    do {                                             \
        if (!(h)) return NPE_ERR_INVALID;            \
        cudaError_t e0__ = cudaSetDevice((h)->device); \
        if (e0__ != cudaSuccess) return fail(h, NPE_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e0__)); \
    } while (0)

__global__ void k_batch_foci(int* scene, int* obj, int* t0, int* n_obj, int B, int N) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) { scene[i] = i; obj[i] = 0; t0[i] = 0; n_obj[i] = N; }
}

// Launch with programmatic stream serialization (PDL): the kernel may start while its predecessor drains; it calls
// griddepcontrol.wait before touching predecessor outputs.
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(npe_handle* h, void (*kernel)(KArgs...), int grid, int block, size_t smem, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = h->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    // directly after a pair builder: a plain (fully serialised) launch, because the kernels below read the builder's
    // outputs in their prologue, before griddepcontrol.wait
    // Programmatic launches pay off in the latency regime (a step of ~25 us made of kernels of a few us).  In the throughput
    // regime every kernel is a persistent grid that fills the machine: measured on the S-64 step, dependents that become
    // resident early and sit in griddepcontrol.wait cost 0.420 ms against 0.335 ms with plain launches (tools/s64_time.py).
    static const bool no_pdl = getenv("NPE_NO_PDL") != nullptr;    // A/B measurements only
    cfg.numAttrs = (h->after_builder || no_pdl || h->F > h->pdl_max_foci) ? 0 : 1;
    h->after_builder = false;
    if (h->sink_mode == 2) {
        // the launch becomes the parameters of an existing kernel node (same place in the chain, same kind of edge)
        auto& nodes = h->step_nodes[h->sink_slot];
        if (h->sink_idx >= (int)nodes.size() || cfg.numAttrs != 1) { h->sink_ok = false; return cudaSuccess; }
        std::tuple<KArgs...> held(static_cast<KArgs>(args)...);
        void* ptrs[sizeof...(KArgs) > 0 ? sizeof...(KArgs) : 1];
        int k = 0;
        std::apply([&](auto&... a) { ((ptrs[k++] = (void*)&a), ...); }, held);
        cudaKernelNodeParams kp = {};
        kp.func = (void*)kernel; kp.gridDim = cfg.gridDim; kp.blockDim = cfg.blockDim; kp.sharedMemBytes = (unsigned)smem;
        kp.kernelParams = ptrs; kp.extra = nullptr;
        if (cudaGraphExecKernelNodeSetParams(h->step_exec[h->sink_slot], nodes[(size_t)h->sink_idx], &kp) != cudaSuccess) {
            cudaGetLastError();
            h->sink_ok = false;
        }
        h->sink_idx++;
        return cudaSuccess;
    }
    const cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
    if (h->sink_mode == 1 && e == cudaSuccess) {
        // capturing: the node just added is the stream's current dependency
        cudaStreamCaptureStatus st; const cudaGraphNode_t* deps = nullptr; const cudaGraphEdgeData* ed = nullptr; size_t nd = 0;
        if (cudaStreamGetCaptureInfo_v3(h->stream, &st, nullptr, nullptr, &deps, &ed, &nd) == cudaSuccess && nd == 1 && cfg.numAttrs == 1)
            h->step_nodes[h->sink_slot].push_back(deps[0]);
        else { cudaGetLastError(); h->sink_ok = false; }
    }
    return e;
}

int ctx_slots(const npe_handle* h) {
    const int cn = h->Nmax - 1;
    return (h->cfg.max_ctx > 0 && h->cfg.max_ctx < cn) ? h->cfg.max_ctx : cn;
}

int ensure_batch_buffers(npe_handle* h, int F) {
    CK(h->y.ensure((size_t)F * OUT));
    CK(h->pred.ensure((size_t)F * OUT));
    CK(h->loss_ex.ensure((size_t)F * 3));
    CK(h->ds.ensure((size_t)F * 52));
    CK(h->keep.ensure(F));
    CK(h->nbr.ensure(F));
    CK(h->cnt.ensure(F));
    CK(h->block_sums.ensure((size_t)(F + FPB - 1) / FPB));
    CK(h->foc_row.ensure(F));
    CK(h->pair_start.ensure((size_t)F + 1));
    // dense active-pair list: capacity = F * min(context slots, 16).  Non-overlapping radius-60 discs cannot have more
    // than ~15 centres within 210 px, so 16 per focus is a physical bound for ball scenes; overflow is flagged.
    int per = ctx_slots(h);
    if (per > 16) per = 16;
    if (per < 1) per = 1;
    const size_t cap = (size_t)F * per;
    CK(h->pair_foc.ensure(cap));
    CK(h->pair_crow.ensure(cap));
    CK(h->Hp.ensure(cap * 52 + 4));
    h->pair_cap = (int)cap;
    return NPE_OK;
}

// Pair builder launch: sub-warp variants when every scene has at most 9 / 17 objects.  One CTA (at most 64 foci): the
// latency-tuned kernel; more: the single-launch chained-scan kernel.
// Largest squared distance s whose reference test  fl(800 * fl(sqrt(s))) <= thr  passes (see warp_pairs_core): bisection over
// the bit patterns of the non-negative floats (the test is monotone in s).  Negative: nothing passes.
float nbr_threshold_sq(float thr) {
    auto passes = [thr](uint32_t bits) {
        float s; memcpy(&s, &bits, 4);
        volatile float d = sqrtf(s);                      // correctly rounded, like __fsqrt_rn
        volatile float m = d * 800.f;                     // one rounding, like __fmul_rn
        return m <= thr;
    };
    if (!passes(0u)) return -1.f;
    uint32_t lo = 0u, hi = 0x7f800000u;                   // passes(lo), !passes(hi) (+inf never passes a finite threshold)
    if (passes(hi)) hi = 0x7f800001u;
    while (hi - lo > 1u) {
        const uint32_t mid = lo + (hi - lo) / 2u;
        if (passes(mid)) lo = mid; else hi = mid;
    }
    float s; memcpy(&s, &lo, 4);
    return s;
}

long long* chain_prof(npe_handle* h);
struct PairOut { u64* keep; u64* nbr; int* cnt; int* block_sums; long long* foc_row; float* y; };
int launch_build_pairs(npe_handle* h, const FocusSrc& src, float thr, const PairOut& o, bool want_keep) {
    const int nmax = h->Nmax, kmax = h->cfg.max_ctx;
    if (src.F <= FPB) {
        if (nmax - 1 <= 8)
            k_build_pairs<8><<<1, 512, 0, h->stream>>>(src, kmax, thr, o.keep, o.nbr, o.cnt, o.block_sums, o.foc_row, o.y, h->active_total,
                                                       h->pair_start.p, h->pair_foc.p, h->pair_crow.p, h->pair_cap, h->info_dev, (int)want_keep);
        else if (nmax - 1 <= 16)
            k_build_pairs<16><<<1, 1024, 0, h->stream>>>(src, kmax, thr, o.keep, o.nbr, o.cnt, o.block_sums, o.foc_row, o.y, h->active_total,
                                                         h->pair_start.p, h->pair_foc.p, h->pair_crow.p, h->pair_cap, h->info_dev, (int)want_keep);
        else
            k_build_pairs<32><<<1, NT_PAIRS, 0, h->stream>>>(src, kmax, thr, o.keep, o.nbr, o.cnt, o.block_sums, o.foc_row, o.y, h->active_total,
                                                            h->pair_start.p, h->pair_foc.p, h->pair_crow.p, h->pair_cap, h->info_dev, (int)want_keep);
    } else {
        const int nblk = (src.F + NT_SCAN - 1) / NT_SCAN;
        const int w = h->stream == h->stream2;
        DevBuf<unsigned>& cnt = h->scan_count[w];
        CK(cnt.ensure((size_t)nblk));
        if (!h->scan_bar[w]) {
            CK(cudaMalloc((void**)&h->scan_bar[w], 128 * sizeof(unsigned)));
            CK(cudaMemsetAsync(h->scan_bar[w], 0, 128 * sizeof(unsigned), h->stream));
        }
        ScanArgs A;
        A.src = src; A.kmax = kmax; A.thr_s = thr; A.want_keep = (int)want_keep;
        A.keep = o.keep; A.nbr = o.nbr; A.foc_row = o.foc_row; A.y = o.y; A.active_total = h->active_total;
        A.pair_start = h->pair_start.p; A.pair_foc = h->pair_foc.p; A.pair_crow = h->pair_crow.p; A.cap = h->pair_cap; A.info = h->info_dev;
        A.sc.count = cnt.p; A.sc.bar = h->scan_bar[w];
        A.sc.epoch = ++h->scan_epoch;
        if (A.sc.epoch == 0) A.sc.epoch = ++h->scan_epoch;                      // 0 is what fresh barrier words hold
        A.sc.nblk = nblk;
        A.sc.prof = chain_prof(h) ? chain_prof(h) + 64 : nullptr;
        // cooperative launch: the grid never exceeds what is resident at once (occupancy x SMs), see the kernel
        const void* fn = o.y ? (const void*)k_build_pairs_scan<true> : (const void*)k_build_pairs_scan<false>;
        int& per_sm = h->scan_ctas_per_sm[o.y ? 1 : 0];
        if (per_sm == 0) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, NT_SCAN, 0));
        if (per_sm < 1) return fail(h, NPE_ERR_CUDA, "pair builder: kernel does not fit an SM");
        const int grid = nblk < per_sm * h->num_sms ? nblk : per_sm * h->num_sms;
        void* args[] = {(void*)&A};
        CK(cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(NT_SCAN), args, 0, h->stream));
    }
    h->launches++;
    h->after_builder = true;
    return NPE_OK;
}

// want_keep: the caller reads the trimmed context LIST (npe_build_pairs outputs), not only the neighbour bits.  The hot
// paths never do, and on scenes of more than 17 objects the complete list costs O(N^2) shuffles per focus (warp_pairs).
int build_pairs(npe_handle* h, bool want_keep = false) {
    if (h->pairs_valid && (h->keep_valid || !want_keep)) return NPE_OK;
    const bool refresh = h->pairs_valid;          // tables are current; only the kept-context bits are incomplete
    if (h->F <= 0) return fail(h, NPE_ERR_STATE, "no batch selected");
    const float thr = nbr_threshold_sq(h->cfg.nbrhdsize * 60.f);   // nbrhdsize * object_base_size.ball|block (npe.lua:180-186)
    {
        ProfScope ps__(h, NPE_K_BUILD_PAIRS);
        const PairOut o = {h->keep.p, h->nbr.p, h->cnt.p, h->block_sums.p, h->foc_row.p, h->batch_form ? nullptr : h->y.p};
        const int rc = launch_build_pairs(h, h->src(), thr, o, want_keep);
        if (rc) return rc;
    }
    CK(cudaGetLastError());
    h->pairs_valid = true;
    h->keep_valid = want_keep || (h->F <= FPB && h->Nmax - 1 <= 16);   // the sub-warp single-CTA variants rank every context
    if (!refresh) { h->hact_valid = false; h->dact_valid = false; }   // a refresh rewrites identical tables
    return NPE_OK;
}

// Called at points that synchronise anyway: surfaces a dropped-pair condition instead of returning wrong sums.
int check_pair_overflow(npe_handle* h) {
    int info[2] = {0, 0};
    CK(cudaMemcpyAsync(info, h->info_dev, sizeof info, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (info[1]) return fail(h, NPE_ERR_NOMEM, "active-pair list overflow (capacity %d pairs); results are invalid", h->pair_cap);
    return NPE_OK;
}

// After a synchronisation point in peer mode: k_dp_step sets info[1] = 3 when a peer's gradient never arrived.
int check_peer_timeout(npe_handle* h) {
    int info[2] = {0, 0};
    CK(cudaMemcpy(info, h->info_dev, sizeof info, cudaMemcpyDeviceToHost));
    if (info[1] == 3) {
        cudaMemset(h->info_dev + 1, 0, sizeof(int));
        return fail(h, NPE_ERR_CUDA, "data-parallel step: a peer rank did not deliver its gradient within %d ms "
                                     "(rank fault, or ranks issued different numbers of steps); results are invalid", h->peer_timeout_ms);
    }
    return NPE_OK;
}

// Small batches take the 16-row (latency) mapping: more CTAs, each pass ~3x shorter; large ones the 64-row mapping.
bool small_batch(const npe_handle* h, int F) { return F <= (h->small_batch_max >= 0 ? h->small_batch_max : 16 * h->num_sms / 4); }

int pair_grid(const npe_handle* h, int rows) {
    const long long tiles = ((long long)h->pair_cap + rows - 1) / rows;
    return (int)(tiles < h->num_sms ? (tiles > 0 ? tiles : 1) : h->num_sms);
}
int focus_grid(const npe_handle* h, int F, int rows) {
    const int tiles = (F + rows - 1) / rows;
    return tiles < h->num_sms ? tiles : h->num_sms;
}

int refresh_wtt(npe_handle* h);

// tools only (NPE_CHAIN_PROF=1): cycle counters of CTA 0 / thread 0 of the pair forward chain, printed at npe_sync
long long* chain_prof(npe_handle* h) {
    static const bool want = getenv("NPE_CHAIN_PROF") != nullptr;
    if (!want) return nullptr;
    if (!h->chain_prof_dev) { cudaMalloc((void**)&h->chain_prof_dev, 96 * sizeof(long long)); cudaMemset(h->chain_prof_dev, 0, 96 * sizeof(long long)); }
    return h->chain_prof_dev;
}

int refresh_wtc(npe_handle* h, int mode) {
    if (mode == NPE_PRECISION_TF32_TC && h->wtc_dirty) {
        k_pack_tc_scatter<<<(NPARAM + 255) / 256, 256, 0, h->stream>>>(h->params, h->wtc);
        CKL();
        h->wtc_dirty = false;
    } else if (mode == NPE_PRECISION_TF32X3_TC && h->wtc3_dirty) {
        k_pack_tc3_scatter<<<(NPARAM + 255) / 256, 256, 0, h->stream>>>(h->params, h->wtc3, h->wtc3 + TC_TOTAL);
        CKL();
        k_pack_enc64<<<(OFF_PAIR + 255) / 256, 256, 0, h->stream>>>(h->params, h->wenc);
        CKL();
        h->wtc3_dirty = false;
    }
    return NPE_OK;
}

// Activation stash for training (h0..h5 per active pair); skipped (recompute instead) when it would exceed 2 GB.
float* training_stash(npe_handle* h) {
    h->stash_rows = ((long long)h->pair_cap + 127) / 128 * 128;
    const size_t need = (size_t)(NL + 1) * h->stash_rows * HSLOT + 128 * HSLOT;   // + slack: tile-granular bulk reads
    if (need * sizeof(float) > (size_t)4 << 30) return nullptr;
    if (h->hact.ensure(need) != cudaSuccess) return nullptr;
    return h->hact.p;
}

// pair MLP + decoder forward on the current pair list; `st` = scene tensor the row offsets point into.
// stash: activation stash for a following backward (training) or nullptr; skip_decoder: the fused training step lets
// k_dec_backward produce pred / losses.
int launch_forward_kernels(npe_handle* h, const float* st, int F, const float* y, float* loss_ex, int tc,
                           float* stash = nullptr, bool skip_decoder = false, float* dact = nullptr, float* roll = nullptr,
                           float2* roll_pos = nullptr) {
    ProfScope ps__(h, NPE_K_FORWARD);
    if (tc) {
        int rc = refresh_wtc(h, tc);
        if (rc) return rc;
        const int tiles_per_cta = (tc == NPE_PRECISION_TF32X3_TC) ? 2 : 1;   // the 3xTF32 kernels run two tile groups per CTA
        const long long ptiles = (((long long)h->pair_cap + 127) / 128 + tiles_per_cta - 1) / tiles_per_cta;
        const int pg = (int)(ptiles < h->num_sms ? (ptiles > 0 ? ptiles : 1) : h->num_sms);
        const int ft = ((F + 127) / 128 + tiles_per_cta - 1) / tiles_per_cta;
        const int fg = ft < h->num_sms ? ft : h->num_sms;
        if (tc == NPE_PRECISION_TF32_TC) {
            k_pair_forward_tc<<<pg, NT, sizeof(PairTcSmem), h->stream>>>(st, h->pair_foc.p, h->pair_crow.p, h->foc_row.p,
                                                                       h->info_dev, h->wtc, h->Hp.p);
            k_dec_forward_tc<<<fg, NT, sizeof(DecTcSmem), h->stream>>>(st, h->foc_row.p, F, h->pair_start.p, h->Hp.p, h->wtc, y,
                                                                       h->pred.p, loss_ex);
        } else {
            // warp-specialised chain kernels (npe_chain.cuh): tile slots of 128 row-owner threads + one MMA warp per CTA; the
            // pair kernel runs three slots, the decoder kernel two
            // training batches take the three-slot variant of the pair kernel (fewer ROUNDS of the per-tile latency chain when
            // an SM sees only a few tiles); inference and the rollout (many tiles per SM) the two-slot one
            const long long pt3 = (((long long)h->pair_cap + 127) / 128 + P3_SLOTS - 1) / P3_SLOTS;
            const int pg3 = (int)(pt3 < h->num_sms ? (pt3 > 0 ? pt3 : 1) : h->num_sms);
            if (stash)
                CK(launch_pdl(h, k_pair_forward_c3, pg3, P3_NT, sizeof(Pair3ChainSmem), st, h->pair_foc.p, h->pair_crow.p, h->foc_row.p,
                              h->info_dev, h->wtc3, h->wenc, h->Hp.p, stash, dact ? h->hmask.p : nullptr, h->stash_rows, chain_prof(h)));
            else
            CK(launch_pdl(h, k_pair_forward_c2, pg, CH_NT, sizeof(PairChainSmem), st, h->pair_foc.p, h->pair_crow.p, h->foc_row.p,
                          h->info_dev, h->wtc3, h->wenc, h->Hp.p, stash, dact ? h->hmask.p : nullptr, h->stash_rows, chain_prof(h)));
            // stacked images pay in the steady state (many tiles per tile slot: the MMA issue rate counts); with a round or two
            // of tiles the per-tile latency chain counts, and the two K-halves of layer 1 lengthen it.  Training keeps the plain
            // kernel (it also writes the stashes).
            const bool many_tiles = (F + 127) / 128 >= 4 * 2 * h->num_sms;
            const bool stacked = !dact && (h->dec_stacked == 1 || (h->dec_stacked == 0 && many_tiles));
            if (!skip_decoder && stacked)
                CK(launch_pdl(h, k_dec_forward_s2, fg, CH_NT, sizeof(DecS2Smem), st, h->foc_row.p, F, h->pair_start.p, h->Hp.p,
                              h->wtc3, y, h->pred.p, loss_ex, dact, dact ? h->umask.p : nullptr, roll, h->cfg.ball_zero_mode, roll_pos));
            else if (!skip_decoder)
                CK(launch_pdl(h, k_dec_forward_c2, fg, CH_NT, sizeof(DecChainSmem), st, h->foc_row.p, F, h->pair_start.p, h->Hp.p,
                              h->wtc3, y, h->pred.p, loss_ex, dact, dact ? h->umask.p : nullptr, roll, h->cfg.ball_zero_mode, roll_pos));
        }
        h->launches += (skip_decoder && tc == NPE_PRECISION_TF32X3_TC) ? 1 : 2;
    } else {
        if (small_batch(h, F)) {
            CK(launch_pdl(h, k_pair_forward<16>, pair_grid(h, 16), NT, sizeof(PairFwdSmem<16>),
                          st, h->pair_foc.p, h->pair_crow.p, h->foc_row.p, h->info_dev, h->wpack, h->Hp.p, stash, h->stash_rows));
            if (!skip_decoder)
                CK(launch_pdl(h, k_dec_forward<16>, focus_grid(h, F, 16), NT, sizeof(DecFwdSmem<16>),
                              st, h->foc_row.p, F, h->pair_start.p, h->Hp.p, h->wpack, y, h->pred.p, loss_ex));
        } else {
            CK(launch_pdl(h, k_pair_forward<64>, pair_grid(h, 64), NT, sizeof(PairFwdSmem<64>),
                          st, h->pair_foc.p, h->pair_crow.p, h->foc_row.p, h->info_dev, h->wpack, h->Hp.p, stash, h->stash_rows));
            if (!skip_decoder)
                CK(launch_pdl(h, k_dec_forward<64>, focus_grid(h, F, 64), NT, sizeof(DecFwdSmem<64>),
                              st, h->foc_row.p, F, h->pair_start.p, h->Hp.p, h->wpack, y, h->pred.p, loss_ex));
        }
        h->launches += skip_decoder ? 1 : 2;
    }
    CK(cudaGetLastError());
    return NPE_OK;
}

// training == true: the FP32 kernels always, with the activation stash for the backward kernels; fused == true
// (npe_train_step): the decoder forward is left to k_dec_backward, which recomputes it anyway.
int run_forward(npe_handle* h, bool training, bool fused = false) {
    int rc = build_pairs(h);
    if (rc) return rc;
    // training forward: FP32 FFMA kernels, or (train_fwd_tc) the FP32-grade tcgen05 pair kernel; never single-pass TF32
    const int tc = training ? (h->train_fwd_tc && !small_batch(h, h->F) ? NPE_PRECISION_TF32X3_TC : 0) : h->precision;
    float* stash = (training && h->have_y && h->train_stash) ? training_stash(h) : nullptr;
    // tcgen05 backward (large batches): it consumes the pair stash AND a decoder stash, so the decoder forward runs here
    // (with the stash) instead of inside k_dec_backward
    float* dact = nullptr;
    if (training && h->bwd_tc && tc == NPE_PRECISION_TF32X3_TC && stash &&
        h->dact.ensure((size_t)5 * h->F * HSLOT + 128 * HSLOT) == cudaSuccess && h->umask.ensure((size_t)4 * h->F) == cudaSuccess &&
        h->hmask.ensure((size_t)(NL + 1) * h->stash_rows) == cudaSuccess)
        dact = h->dact.p;
    // the transposed images of the input-gradient chains are refreshed here, before the forward chain, so that no plain
    // launch sits between the programmatically chained kernels of the step
    if (dact && (rc = refresh_wtt(h))) return rc;
    rc = launch_forward_kernels(h, h->states.p, h->F, h->have_y ? h->yptr() : nullptr, h->have_y ? h->loss_ex.p : nullptr, tc,
                                stash, fused && !dact, dact);
    if (rc) return rc;
    h->fwd_valid = (tc != NPE_PRECISION_TF32_TC);
    h->hact_valid = stash != nullptr;
    h->dact_valid = dact != nullptr;
    return NPE_OK;
}

int run_loss_reduce(npe_handle* h) {
    k_loss_reduce<<<1, 256, 0, h->stream>>>(h->loss_ex.p, h->F, h->loss3_dev, h->loss_sums_dev);
    CKL();
    return NPE_OK;
}

// Arguments of the reduction + optimiser (+ peer exchange) stage; advances the optimiser / exchange step counters.
struct ReducePlan {
    OptArgs o; LossArgs L; DpArgs P; int buf = 0; unsigned stamp = 0; int opt = 0; bool peer = false;
};
ReducePlan prepare_reduce(npe_handle* h, int fuse_opt, float lr, bool want_loss, int peer_opt) {
    ReducePlan r;
    OptArgs& o = r.o;
    o.x = h->params; o.m = h->adam_m; o.v = h->adam_v; o.wpack = h->wpack;
    o.b1 = 0.9f; o.omb1 = (float)(1.0 - 0.9); o.b2 = 0.999f; o.omb2 = (float)(1.0 - 0.999); o.eps = 1e-8f; o.step = lr;
    r.L.loss_ex = h->loss_ex.p; r.L.F = h->F;
    r.L.loss4_out = want_loss ? (h->loss4_override ? h->loss4_override : h->loss4_devptr) : nullptr; r.L.info = h->info_dev;
    r.opt = peer_opt ? peer_opt : fuse_opt;
    r.peer = peer_opt != 0;
    r.P.nranks = h->nranks; r.P.rank = h->rank;
    r.P.abort = h->info_dev + 1; r.P.timeout_ns = (unsigned long long)h->peer_timeout_ms * 1000000ull;
    for (int q = 0; q < MAX_PEERS; ++q) r.P.region[q] = (r.peer && q < h->nranks) ? (uint2*)h->peer_ptr[q] : nullptr;
    r.P.prof = nullptr;
    {   // timing experiments only: NPE_DP_DEBUG = one digit per rank ("12": rank 0 pushes without waiting, rank 1 does neither)
        const char* dbg = getenv("NPE_DP_DEBUG");
        r.P.debug_mode = (dbg && (int)strlen(dbg) > h->rank) ? dbg[h->rank] - '0' : (dbg && *dbg ? dbg[0] - '0' : 0);
    }
    if (r.peer && chain_prof(h)) {
        if (!h->dp_prof_dev) { cudaMalloc((void**)&h->dp_prof_dev, 8 * DP_NBLK * sizeof(long long)); cudaMemset(h->dp_prof_dev, 0, 8 * DP_NBLK * sizeof(long long)); }
        r.P.prof = h->dp_prof_dev;
    }
    if (r.peer) {
        r.buf = (int)(h->peer_step & 1u);
        r.stamp = h->peer_step + 1u;
        h->peer_step += 1u;
    }
    if (r.opt == 1) {
        h->adam_t += 1;
        const double t = (double)h->adam_t;
        o.step = (float)((double)lr * std::sqrt(1.0 - std::pow(0.999, t)) / (1.0 - std::pow(0.9, t)));
    } else if (r.opt == 2) {
        o.m = h->rms_m; o.b1 = 0.99f; o.omb1 = (float)(1.0 - 0.99);
    }
    o.wtc3_hi = o.wtc3_lo = o.wtt3_hi = o.wtt3_lo = o.wenc_hi = o.wenc_lo = nullptr;
    if (r.opt) {
        // the step that ends here ran on the tcgen05 images (both are current): the optimiser keeps them current
        const bool through = !h->wtc3_dirty && !h->wtt3_dirty && h->wtc3 && h->wtt3 && h->wenc;
        if (through) {
            o.wtc3_hi = h->wtc3; o.wtc3_lo = h->wtc3 + TC_TOTAL;
            o.wtt3_hi = h->wtt3; o.wtt3_lo = h->wtt3 + TT_TOTAL;
            o.wenc_hi = h->wenc; o.wenc_lo = h->wenc + 2 * ENC64_SZ;
        }
        h->wtc_dirty = true;
        if (!through) { h->wtc3_dirty = true; h->wtt3_dirty = true; }
    }
    return r;
}

// Fixed-order reduction of the CTA partials (+ optimiser / peer exchange / batch loss, see run_backward).
int launch_reduce(npe_handle* h, int gp, int gd, int fuse_opt, float lr, bool want_loss, int peer_opt) {
    ProfScope ps__(h, NPE_K_REDUCE);
    const ReducePlan r = prepare_reduce(h, fuse_opt, lr, want_loss, peer_opt);
    // many partials (throughput regime: one per SM and part): four lanes per parameter, see ordered_partial_sum
    const bool split = (gp > gd ? gp : gd) > 32;
    const int nb = ((split ? 4 : 1) * NPARAM + 255) / 256 + (want_loss ? 1 : 0);
#define NPE_REDUCE(RS)                                                                                                              \
    do {                                                                                                                            \
        if (r.peer && r.opt == 1) CK(launch_pdl(h, k_dp_step<1, RS>, nb, 256, 0, h->partial, gp, gd, r.P, r.buf, r.stamp, h->grads, r.o, r.L)); \
        else if (r.peer) CK(launch_pdl(h, k_dp_step<2, RS>, nb, 256, 0, h->partial, gp, gd, r.P, r.buf, r.stamp, h->grads, r.o, r.L));          \
        else if (r.opt == 1) CK(launch_pdl(h, k_reduce_grads<1, RS>, nb, 256, 0, h->partial, gp, gd, h->grads, r.o, r.L));          \
        else if (r.opt == 2) CK(launch_pdl(h, k_reduce_grads<2, RS>, nb, 256, 0, h->partial, gp, gd, h->grads, r.o, r.L));          \
        else CK(launch_pdl(h, k_reduce_grads<0, RS>, nb, 256, 0, h->partial, gp, gd, h->grads, r.o, r.L));                          \
    } while (0)
    if (split) NPE_REDUCE(4); else NPE_REDUCE(1);
#undef NPE_REDUCE
    h->launches++;
    CK(cudaGetLastError());
    return NPE_OK;
}

int refresh_wtt(npe_handle* h) {
    if (!h->wtt3_dirty) return NPE_OK;
    k_pack_tt3_scatter<<<(NPARAM + 255) / 256, 256, 0, h->stream>>>(h->params, h->wtt3, h->wtt3 + TT_TOTAL);
    CKL();
    h->wtt3_dirty = false;
    return NPE_OK;
}

WJob wjob_layer(const float* G, int gcols, const float* X, int tmem_col, int pt_base, int n_valid) {
    WJob J = {};
    J.G = G; J.g_cols = gcols;
    J.X0 = X; J.x0_valid = 52; J.x0_store = 56;
    J.ngs = 0; J.nb = 56; J.tmem_col = tmem_col; J.kind = 0; J.pt_base = pt_base; J.kt = 13; J.n_valid = n_valid;
    return J;
}

// model:bp for large batches on tcgen05 (npe_chain.cuh, npe_tcbwd.cuh): decoder input-gradient chain -> ds, pair
// input-gradient chain, then the weight gradients of the pair part and of the decoder part (accumulators in tensor memory
// per CTA) into the per-CTA partials that the reduce / optimiser / exchange kernel consumes.
int run_backward_tc(npe_handle* h, float sv, float sw, int* gp_out, int* gd_out) {
    int rc = refresh_wtt(h);
    if (rc) return rc;
    const int F = h->F;
    const size_t SR = (size_t)h->stash_rows;
    CK(h->ddz.ensure((size_t)5 * F * HSLOT + 128 * HSLOT));
    CK(h->dzact.ensure((size_t)(NL + 1) * SR * HSLOT + 128 * HSLOT));
    CK(h->pair_frow.ensure((size_t)h->pair_cap));
    const int ftiles = (F + 127) / 128;
    const long long ptiles = ((long long)h->pair_cap + 127) / 128;
    const int fg = (ftiles + 1) / 2 < h->num_sms ? (ftiles + 1) / 2 : h->num_sms;
    const int pg = (int)((ptiles + P3_SLOTS - 1) / P3_SLOTS < h->num_sms ? ((ptiles + P3_SLOTS - 1) / P3_SLOTS > 0 ? (ptiles + P3_SLOTS - 1) / P3_SLOTS : 1) : h->num_sms);   // three tile slots per SM
    const int gd = ftiles < h->num_sms ? ftiles : h->num_sms;
    const int gp = (int)(ptiles < h->num_sms ? (ptiles > 0 ? ptiles : 1) : h->num_sms);
    {
        ProfScope ps__(h, NPE_K_DEC_BACKWARD);
        CK(launch_pdl(h, k_dec_dgrad_c2, fg, CH_NT, sizeof(DecDgradChainSmem), F, h->pred.p, h->yptr(), h->umask.p, h->wtt3, sv, sw,
                      h->ddz.p, h->ds.p));
        h->launches++;
    }
    auto stacked = [](const WArgs& W, const int (&ia)[3], const int (&ib)[3]) {
        WmArgs M = {};
        for (int p = 0; p < 3; ++p) {
            WPair& P = M.pair[p];
            P.a = W.job[ia[p]];
            P.has_b = ib[p] >= 0;
            if (P.has_b) P.b = W.job[ib[p]];
            P.ca = P.a.ngs ? 3 : 2;
            P.cb = !P.has_b ? 0 : (P.b.ngs ? 3 : 2);
            P.tmem_col = 128 * p;
        }
        M.npairs = 3; M.nrows_dev = W.nrows_dev; M.nrows = W.nrows; M.partial = W.partial; M.prof = W.prof;
        return M;
    };
    // decoder part of the weight gradients.  Layer 5: G = dz5 (22 of 24 columns), X = u4; layers 4..2: G = dz_l, X = u_{l-1};
    // column 50 of X is the constant 1, i.e. accumulator column 50 is the bias gradient
    WArgs B = {};
    const size_t FS = (size_t)F * HSLOT;
    B.job[0] = wjob_layer(h->ddz.p + 4 * FS, 24, h->dact.p + 4 * FS, 0, PT_DEC5, OUT);
    for (int l = 4; l >= 2; --l)
        B.job[5 - l] = wjob_layer(h->ddz.p + (l - 1) * FS, 52, h->dact.p + (l - 1) * FS, 64 * (5 - l), PT_DEC2 + (l - 2) * PT_LAYER, H);
    WJob& D1 = B.job[4];            // layer 1: G = dz1, X = z = [pair sum (50) | 1 | 0 | focus window (44)]
    D1 = WJob{};
    D1.G = h->ddz.p; D1.g_cols = 52;
    D1.X0 = h->dact.p; D1.x0_valid = 52; D1.x0_store = 52;
    D1.gs[0].base = h->states.p; D1.gs[0].offs = h->foc_row.p; D1.gs[0].cols_valid = IN; D1.gs[0].cols_store = IN; D1.gs[0].col0 = ZCOL_F;
    D1.ngs = 1; D1.nb = 96; D1.tmem_col = 256; D1.kind = 1; D1.kt = 24; D1.n_valid = H;
    B.njobs = 5; B.nrows_dev = nullptr; B.nrows = F; B.partial = h->partial; B.prof = chain_prof(h) ? chain_prof(h) + 32 : nullptr;
    // That part only needs what exists now (dz of the decoder chain, forward stashes): in the throughput regime it goes to a
    // low-priority side stream and fills the SMs that the pair input-gradient chain leaves idle in its last round of tiles
    // (S-64: 320 pair tiles on 296 tile slots -- the second round keeps 12 of 148 SMs busy), then runs beside the pair part.
    static const bool side_ok = getenv("NPE_NO_WGRAD_SIDE_STREAM") == nullptr;   // A/B measurements
    const bool fork = side_ok && h->prof_kernel < 0 && h->F > h->pdl_max_foci;   // kernel brackets (tools) time ONE stream
    {
        ProfScope ps__(h, NPE_K_PAIR_BACKWARD);
        if (fork) CK(cudaEventRecord(h->wg_fork, h->stream));
        CK(launch_pdl(h, k_pair_dgrad_c2, pg, P3_NT, sizeof(PairDgradChainSmem), h->pair_foc.p, h->foc_row.p, h->info_dev, h->hmask.p,
                      h->stash_rows, h->ds.p, h->wtt3, h->dzact.p, h->pair_frow.p));
        h->launches++;
    }
    if (fork) {
        CK(cudaStreamWaitEvent(h->stream3, h->wg_fork, 0));
        cudaStream_t main_stream = h->stream;
        h->stream = h->stream3;
        const cudaError_t e = launch_pdl(h, k_wgrad_mn<1>, gd, WM_NT, sizeof(WmSmem) + 1024, stacked(B, {0, 2, 4}, {1, 3, -1}));   // (D5, D4) (D3, D2) (D1)
        h->stream = main_stream;
        CK(e);
        CK(cudaEventRecord(h->wg_join, h->stream3));
        h->launches++;
    }
    {
        ProfScope ps__(h, NPE_K_WGRAD);
        WArgs A = {};
        for (int l = NL; l >= 1; --l)   // pairwise layer l: G = dz_l, X = h_{l-1}
            A.job[NL - l] = wjob_layer(h->dzact.p + (size_t)l * SR * HSLOT, 52, h->hact.p + (size_t)(l - 1) * SR * HSLOT,
                                       64 * (NL - l), PT_PAIR + (l - 1) * PT_LAYER, H);
        WJob& E = A.job[NL];            // both encoders: G = dz0, X = [focus window (48) | context window (48)]
        E = WJob{};
        E.G = h->dzact.p; E.g_cols = 52; E.X0 = nullptr;
        E.gs[0].base = h->states.p; E.gs[0].offs = h->pair_frow.p; E.gs[0].cols_valid = IN; E.gs[0].cols_store = 48; E.gs[0].col0 = 0;
        E.gs[1].base = h->states.p; E.gs[1].offs = h->pair_crow.p; E.gs[1].cols_valid = IN; E.gs[1].cols_store = 48; E.gs[1].col0 = 48;
        E.ngs = 2; E.nb = 96; E.tmem_col = 64 * NL; E.kind = 2; E.n_valid = H;
        A.njobs = NL + 1; A.nrows_dev = h->info_dev; A.nrows = 0; A.partial = h->partial; A.prof = chain_prof(h);
        CK(launch_pdl(h, k_wgrad_mn<0>, gp, WM_NT, sizeof(WmSmem) + 1024, stacked(A, {0, 2, 4}, {1, 3, 5})));   // (L5, L4) (L3, L2) (L1, encoders)
        h->launches++;
        if (fork) {
            CK(cudaStreamWaitEvent(h->stream, h->wg_join, 0));      // the reduce kernel needs both parts
        } else {
            CK(launch_pdl(h, k_wgrad_mn<1>, gd, WM_NT, sizeof(WmSmem) + 1024, stacked(B, {0, 2, 4}, {1, 3, -1})));
            h->launches++;
        }
    }
    CK(cudaGetLastError());
    *gp_out = gp; *gd_out = gd;
    return NPE_OK;
}

// fuse_opt: 0 = reduction only; 1 = + Adam; 2 = + RMSprop (single-GPU train step)
// peer_opt: 1 / 2 = fused NVLink all-reduce + Adam / RMSprop (requires npe_peer_attach)
int run_backward(npe_handle* h, int divisor_batch, int fuse_opt = 0, float lr = 0.f, bool want_loss = false, int peer_opt = 0,
                 bool fused_fwd = false) {
    if (!h->fwd_valid) return fail(h, NPE_ERR_STATE, "npe_backward must follow npe_forward on the same batch");
    if (!h->have_y) return fail(h, NPE_ERR_STATE, "batch has no targets");
    const int B = divisor_batch > 0 ? divisor_batch : h->F;
    // d_vel = 2 (p - y) * vlambda / (2B); d_ang_vel = 2 (p - y) * lambda / B   (npe.lua:304-315)
    const float sv = 2.f * h->cfg.vlambda / (float)(2 * B);
    const float sw = 2.f * h->cfg.lambda / (float)B;
    const bool small = small_batch(h, h->F);
    if (h->dact_valid && h->hact_valid && !small) {
        int gpt = 0, gdt = 0;
        const int rc = run_backward_tc(h, sv, sw, &gpt, &gdt);
        if (rc) return rc;
        if (h->prof_kernel < 0) { CK(cudaEventRecord(h->pre_reduce, h->stream)); h->pre_reduce_valid = true; }   // gate of the next pair builder
        return launch_reduce(h, gpt, gdt, fuse_opt, lr, want_loss, peer_opt);
    }
    const int rows = small ? 16 : 64;
    const int gd = focus_grid(h, h->F, rows);
    const int gp = pair_grid(h, rows);
    {
        ProfScope ps__(h, NPE_K_DEC_BACKWARD);
        if (small)
            CK(launch_pdl(h, k_dec_backward<16>, gd, NT, sizeof(DecBwdSmem<16>), h->states.p, h->foc_row.p, h->F,
                          h->pair_start.p, h->Hp.p, h->wpack, h->yptr(), h->ds.p, h->partial, sv, sw,
                          fused_fwd ? h->pred.p : nullptr, fused_fwd ? h->loss_ex.p : nullptr));
        else
            CK(launch_pdl(h, k_dec_backward<64>, gd, NT, sizeof(DecBwdSmem<64>), h->states.p, h->foc_row.p, h->F,
                          h->pair_start.p, h->Hp.p, h->wpack, h->yptr(), h->ds.p, h->partial, sv, sw,
                          fused_fwd ? h->pred.p : nullptr, fused_fwd ? h->loss_ex.p : nullptr));
        h->launches++;
    }
    {
        ProfScope ps__(h, NPE_K_PAIR_BACKWARD);
        if (small)
            CK(launch_pdl(h, k_pair_backward<16>, gp, NT, sizeof(PairBwdSmem<16>), h->states.p, h->pair_foc.p,
                          h->pair_crow.p, h->foc_row.p, h->info_dev, h->wpack, h->ds.p, h->partial,
                          h->hact_valid ? h->hact.p : nullptr, h->stash_rows));
        else
            CK(launch_pdl(h, k_pair_backward<64>, gp, NT, sizeof(PairBwdSmem<64>), h->states.p, h->pair_foc.p,
                          h->pair_crow.p, h->foc_row.p, h->info_dev, h->wpack, h->ds.p, h->partial,
                          h->hact_valid ? h->hact.p : nullptr, h->stash_rows));
        h->launches++;
    }
    CK(cudaGetLastError());
    return launch_reduce(h, gp, gd, fuse_opt, lr, want_loss, peer_opt);
}

// Forward + backward of a small batch as ONE grid of tiles (npe_step16.cuh) followed by the reduce kernel.  Taken by
// npe_train_step when every pair tile and decoder tile gets its own SM; same results as run_forward + run_backward.
// npe_train_steps runs its steps as graphs when the data-parallel exchange writes into peer memory (see there); NPE_STEP_GRAPH=1
// forces graphs for a single GPU as well (A/B measurements).
bool step_graph_wanted(const npe_handle* h) {
    static const bool force = getenv("NPE_STEP_GRAPH") != nullptr;
    return (h->peer_ready || force) && (h->peer_ready || !h->comm);      // not with ncclAllReduce in the loop
}
bool step_fused_ok(const npe_handle* h) {
    if (!small_batch(h, h->F) || !h->have_y) return false;
    static const bool off = getenv("NPE_NO_FUSED_STEP") != nullptr;   // A/B measurements only
    if (off) return false;
    const int gp = (h->pair_cap + SR - 1) / SR, gd = (h->F + SR - 1) / SR;
    return gp >= 1 && gp + gd <= h->num_sms && gp + gd <= 500;
}
int run_step_fused(npe_handle* h, int divisor_batch, int fuse_opt, float lr, bool want_loss, int peer_opt) {
    int rc = build_pairs(h);
    if (rc) return rc;
    const int B = divisor_batch > 0 ? divisor_batch : h->F;
    const int gp = (h->pair_cap + SR - 1) / SR, gd = (h->F + SR - 1) / SR;
    StepArgs A;
    A.states = h->states.p; A.pair_foc = h->pair_foc.p; A.pair_crow = h->pair_crow.p; A.foc_row = h->foc_row.p;
    A.pair_start = h->pair_start.p; A.info = h->info_dev; A.wpack = h->wpack; A.y = h->yptr();
    A.Hp = h->Hp.p; A.ds = h->ds.p; A.partial = h->partial; A.pred = h->pred.p; A.loss_ex = h->loss_ex.p;
    A.F = h->F; A.gp = gp;
    A.scale_v = 2.f * h->cfg.vlambda / (float)(2 * B);      // npe.lua:304-315, as in run_backward
    A.scale_w = 2.f * h->cfg.lambda / (float)B;
    A.flags = h->step_flags; A.stamp = ++h->step_stamp;
    A.abort = h->info_dev + 1;
    A.trace = h->step_trace;
    {
        ProfScope ps__(h, NPE_K_STEP_FUSED);
        CK(launch_pdl(h, k_step_fused, gp + gd, NT, sizeof(StepSmem), A));
        h->launches++;
    }
    CK(cudaGetLastError());
    h->hact_valid = false;
    return launch_reduce(h, gp, gd, fuse_opt, lr, want_loss, peer_opt);
}

int run_adam(npe_handle* h, float lr, float b1, float b2, float eps) {
    h->adam_t += 1;
    h->wtc_dirty = true; h->wtc3_dirty = true; h->wtt3_dirty = true;
    h->fwd_valid = false; h->hact_valid = false; h->dact_valid = false;   // parameters change: forward state is stale
    const double t = (double)h->adam_t;
    const float step = (float)((double)lr * std::sqrt(1.0 - std::pow((double)b2, t)) / (1.0 - std::pow((double)b1, t)));
    { ProfScope ps__(h, NPE_K_OPTIM); k_adam<<<(NPARAM + 255) / 256, 256, 0, h->stream>>>(h->params, h->grads, h->adam_m, h->adam_v, h->wpack, NPARAM, b1, (float)(1.0 - (double)b1), b2, (float)(1.0 - (double)b2), eps, step);
    h->launches++; }
    { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return fail(h, NPE_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); }
    return NPE_OK;
}

int run_rmsprop(npe_handle* h, float lr, float alpha, float eps) {
    h->wtc_dirty = true; h->wtc3_dirty = true; h->wtt3_dirty = true;
    h->fwd_valid = false; h->hact_valid = false; h->dact_valid = false;
    { ProfScope ps__(h, NPE_K_OPTIM); k_rmsprop<<<(NPARAM + 255) / 256, 256, 0, h->stream>>>(h->params, h->grads, h->rms_m, h->wpack, NPARAM, alpha, (float)(1.0 - (double)alpha), eps, lr);
    h->launches++; }
    { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return fail(h, NPE_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); }
    return NPE_OK;
}

int run_allreduce(npe_handle* h) {
    if (!h->comm) return NPE_OK;
    ncclResult_t r = g_nccl.AllReduce(h->grads, h->grads, NPARAM, NCCL_FLOAT32, NCCL_SUM, h->comm, h->stream);
    if (r != 0) return fail(h, NPE_ERR_NCCL, "ncclAllReduce: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    return NPE_OK;
}

}  // namespace

// ================================================================================================================
extern "C" {

int npe_abi_version(void) { return NPE_ABI_VERSION; }

void npe_default_config(npe_config* cfg) {
    if (!cfg) return;
    memset(cfg, 0, sizeof *cfg);
    cfg->struct_size = (int32_t)sizeof(npe_config);
    cfg->device = 0;
    cfg->rnn_dim = 50; cfg->layers = 5; cfg->num_past = 2; cfg->num_future = 1; cfg->object_dim = 22;
    cfg->nbrhd = 1; cfg->nbrhdsize = 3.5f; cfg->max_ctx = 12; cfg->vlambda = 1.f; cfg->lambda = 1.f;
    cfg->ball_zero_mode = 1;
}

const char* npe_last_error(const npe_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int npe_create(const npe_config* cfg, npe_handle** out) {
    npe_handle* h = nullptr;
    if (!cfg || !out) return fail(h, NPE_ERR_INVALID, "null argument");
    if (cfg->struct_size != (int32_t)sizeof(npe_config)) return fail(h, NPE_ERR_INVALID, "npe_config size mismatch");
    if (cfg->rnn_dim != H || cfg->layers != NL || cfg->num_past != 2 || cfg->num_future != 1 || cfg->object_dim != D)
        return fail(h, NPE_ERR_UNSUPPORTED,
                    "only the north-star configuration is built: rnn_dim 50, layers 5, num_past 2, num_future 1, object_dim 22");
    if (!cfg->nbrhd) return fail(h, NPE_ERR_UNSUPPORTED, "only -nbrhd (bias-free pairwise layers) is built");
    if (cfg->max_ctx > 63) return fail(h, NPE_ERR_INVALID, "max_ctx must be <= 63");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(h, NPE_ERR_CUDA, "no CUDA device: %s (libnpe_cuda has no CPU path)", cudaGetErrorString(e));
    if (cfg->device < 0 || cfg->device >= ndev) return fail(h, NPE_ERR_INVALID, "device %d out of range", cfg->device);
    npe_handle* hh = new npe_handle();
    hh->cfg = *cfg;
    hh->device = cfg->device;
    h = hh;
    auto bail = [&](const char* what, cudaError_t ce) {
        g_create_error = std::string(what) + ": " + cudaGetErrorString(ce);
        delete hh;
        return NPE_ERR_CUDA;
    };
    if ((e = cudaSetDevice(h->device)) != cudaSuccess) return bail("cudaSetDevice", e);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, h->device)) != cudaSuccess) return bail("cudaGetDeviceProperties", e);
    h->num_sms = prop.multiProcessorCount;
    if ((size_t)prop.sharedMemPerBlockOptin < sizeof(PairBwdSmem<64>)) {
        g_create_error = "device shared memory per block too small (needs an sm_100-class GPU)";
        delete hh;
        return NPE_ERR_UNSUPPORTED;
    }
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    for (auto& ev : h->events)
        if ((e = cudaEventCreate(&ev)) != cudaSuccess) return bail("cudaEventCreate", e);
    float** bufs[] = {&h->params, &h->grads, &h->adam_m, &h->adam_v, &h->rms_m};
    for (float** b : bufs) {
        if ((e = cudaMalloc((void**)b, NPARAM * sizeof(float))) != cudaSuccess) return bail("cudaMalloc", e);
        cudaMemset(*b, 0, NPARAM * sizeof(float));
    }
    if ((e = cudaMalloc((void**)&h->partial, (size_t)h->num_sms * PT_TOTAL * sizeof(float))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void**)&h->wpack, WPACK_TOTAL * sizeof(float))) != cudaSuccess) return bail("cudaMalloc", e);
    k_pack_zero<<<(WPACK_TOTAL + 255) / 256, 256, 0, h->stream>>>(h->wpack);
    if ((e = cudaMalloc((void**)&h->wtc, TC_TOTAL * sizeof(float))) != cudaSuccess) return bail("cudaMalloc", e);
    k_pack_tc_zero<<<(TC_TOTAL + 255) / 256, 256, 0, h->stream>>>(h->wtc);
    if ((e = cudaMalloc((void**)&h->wtc3, 2 * (size_t)TC_TOTAL * sizeof(float))) != cudaSuccess) return bail("cudaMalloc", e);
    k_pack_tc_zero<<<(TC_TOTAL + 255) / 256, 256, 0, h->stream>>>(h->wtc3);                 // hi image: unit entries
    cudaMemsetAsync(h->wtc3 + TC_TOTAL, 0, (size_t)TC_TOTAL * sizeof(float), h->stream);   // lo image: zero
    if ((e = cudaMalloc((void**)&h->wenc, 4 * (size_t)ENC64_SZ * sizeof(float))) != cudaSuccess) return bail("cudaMalloc", e);
    cudaMemsetAsync(h->wenc, 0, 4 * (size_t)ENC64_SZ * sizeof(float), h->stream);
    if ((e = cudaMalloc((void**)&h->wtt3, 2 * (size_t)TT_TOTAL * sizeof(float))) != cudaSuccess) return bail("cudaMalloc", e);
    cudaMemsetAsync(h->wtt3, 0, 2 * (size_t)TT_TOTAL * sizeof(float), h->stream);
    if ((e = cudaMalloc((void**)&h->loss3_dev, 4 * sizeof(float))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void**)&h->loss_sums_dev, 4 * sizeof(double))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void**)&h->active_total, sizeof(u64))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void**)&h->info_dev, 4 * sizeof(int))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void**)&h->alt.active_total, sizeof(u64))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void**)&h->step_flags, 512 * sizeof(unsigned))) != cudaSuccess) return bail("cudaMalloc", e);
    cudaMemset(h->step_flags, 0, 512 * sizeof(unsigned));
    if ((e = cudaMalloc((void**)&h->alt.info_dev, 4 * sizeof(int))) != cudaSuccess) return bail("cudaMalloc", e);
    cudaMemset(h->alt.info_dev, 0, 4 * sizeof(int));
    cudaMemset(h->alt.active_total, 0, sizeof(u64));
    if ((e = cudaMalloc((void**)&h->tab_alt.active_total, sizeof(u64))) != cudaSuccess) return bail("cudaMalloc", e);
    if ((e = cudaMalloc((void**)&h->tab_alt.info_dev, 4 * sizeof(int))) != cudaSuccess) return bail("cudaMalloc", e);
    cudaMemset(h->tab_alt.info_dev, 0, 4 * sizeof(int));
    cudaMemset(h->tab_alt.active_total, 0, sizeof(u64));
    for (int i = 0; i < 2; ++i)
        if ((e = cudaEventCreateWithFlags(&h->tab_free[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&h->tab_ready, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&h->pre_reduce, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaStreamCreateWithFlags(&h->stream2, cudaStreamNonBlocking)) != cudaSuccess) return bail("cudaStreamCreate", e);
    if ((e = cudaEventCreateWithFlags(&h->slot_ready, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&h->wg_fork, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaEventCreateWithFlags(&h->wg_join, cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    { int lo_pri = 0, hi_pri = 0; cudaDeviceGetStreamPriorityRange(&lo_pri, &hi_pri);   // numerically greatest = lowest priority
      if ((e = cudaStreamCreateWithPriority(&h->stream3, cudaStreamNonBlocking, lo_pri)) != cudaSuccess) return bail("cudaStreamCreate", e); }
    for (int i = 0; i < 2; ++i)
        if ((e = cudaEventCreateWithFlags(&h->slot_free[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    if ((e = cudaHostAlloc((void**)&h->loss4_host, (16 + 4 * LOSS_RING) * sizeof(float), cudaHostAllocMapped)) != cudaSuccess) return bail("cudaHostAlloc", e);
    if ((e = cudaHostGetDevicePointer((void**)&h->loss4_devptr, h->loss4_host, 0)) != cudaSuccess) return bail("cudaHostGetDevicePointer", e);
    for (int i = 0; i < 2; ++i)
        if ((e = cudaEventCreateWithFlags(&h->stage_ev[i], cudaEventDisableTiming)) != cudaSuccess) return bail("cudaEventCreate", e);
    cudaMemset(h->info_dev, 0, 4 * sizeof(int));
    cudaMemset(h->active_total, 0, sizeof(u64));
    if ((e = cudaFuncSetAttribute(k_pair_forward<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PairFwdSmem<64>))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_dec_forward<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecFwdSmem<64>))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_dec_backward<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecBwdSmem<64>))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_pair_backward<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PairBwdSmem<64>))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_pair_forward<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PairFwdSmem<16>))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_dec_forward<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecFwdSmem<16>))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_dec_backward<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecBwdSmem<16>))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_pair_backward<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PairBwdSmem<16>))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_step_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(StepSmem))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_rollout, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RollSmem))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_pair_forward_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PairTcSmem))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_dec_forward_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecTcSmem))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_wgrad_mn<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WmSmem) + 1024)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_wgrad_mn<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(WmSmem) + 1024)) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_pair_forward_c2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PairChainSmem))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_pair_forward_c3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Pair3ChainSmem))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_dec_forward_c2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecChainSmem))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_dec_forward_s2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecS2Smem))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_dec_dgrad_c2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(DecDgradChainSmem))) != cudaSuccess ||
        (e = cudaFuncSetAttribute(k_pair_dgrad_c2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(PairDgradChainSmem))) != cudaSuccess)
        return bail("cudaFuncSetAttribute(MaxDynamicSharedMemorySize)", e);
    *out = h;
    return NPE_OK;
}

int npe_destroy(npe_handle* h) {
    if (!h) return NPE_OK;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
    for (int r = 0; r < MAX_PEERS; ++r) if (h->peer_opened[r]) cudaIpcCloseMemHandle(h->peer_ptr[r]);
    if (h->xregion) cudaFree(h->xregion);
    float* bufs[] = {h->params, h->grads, h->adam_m, h->adam_v, h->rms_m, h->partial, h->loss3_dev, h->allreduce_buf, h->wpack, h->wtc, h->wtc3, h->wtt3, h->wenc};
    for (float* b : bufs) if (b) cudaFree(b);
    if (h->loss_sums_dev) cudaFree(h->loss_sums_dev);
    if (h->active_total) cudaFree(h->active_total);
    for (int i = 0; i < 2; ++i) { h->scan_count[i].release(); if (h->scan_bar[i]) cudaFree(h->scan_bar[i]); }
    if (h->info_dev) cudaFree(h->info_dev);
    if (h->alt.active_total) cudaFree(h->alt.active_total);
    if (h->step_flags) cudaFree(h->step_flags);
    if (h->sched_host) cudaFreeHost(h->sched_host);
    for (int k = 0; k < 4; ++k) { if (h->step_exec[k]) cudaGraphExecDestroy(h->step_exec[k]); if (h->step_graph[k]) cudaGraphDestroy(h->step_graph[k]); if (h->step_exec_ev[k]) cudaEventDestroy(h->step_exec_ev[k]); }
    if (h->step_trace) cudaFree(h->step_trace);
    if (h->alt.info_dev) cudaFree(h->alt.info_dev);
    if (h->tab_alt.info_dev) cudaFree(h->tab_alt.info_dev);
    if (h->tab_alt.active_total) cudaFree(h->tab_alt.active_total);
    for (int i = 0; i < 2; ++i) if (h->tab_free[i]) cudaEventDestroy(h->tab_free[i]);
    if (h->tab_ready) cudaEventDestroy(h->tab_ready);
    if (h->pre_reduce) cudaEventDestroy(h->pre_reduce);
    { auto& a = h->tab_alt; a.keep.release(); a.nbr.release(); a.cnt.release(); a.block_sums.release(); a.pair_start.release(); a.pair_foc.release(); a.foc_row.release(); a.pair_crow.release(); a.y.release(); }
    { auto& a = h->alt; a.states.release(); a.keep.release(); a.nbr.release(); a.cnt.release(); a.block_sums.release();
      a.pair_start.release(); a.pair_foc.release(); a.foc_row.release(); a.pair_crow.release();
      a.n_obj.release(); a.foc_scene.release(); a.foc_obj.release(); a.foc_t0.release(); }
    if (h->slot_ready) cudaEventDestroy(h->slot_ready);
    if (h->wg_fork) cudaEventDestroy(h->wg_fork);
    if (h->wg_join) cudaEventDestroy(h->wg_join);
    if (h->stream3) cudaStreamDestroy(h->stream3);
    for (int i = 0; i < 2; ++i) if (h->slot_free[i]) cudaEventDestroy(h->slot_free[i]);
    if (h->stream2) cudaStreamDestroy(h->stream2);
    for (int i = 0; i < LOSS_RING; ++i) if (h->ring_ev[i]) cudaEventDestroy(h->ring_ev[i]);
    if (h->loss4_host) cudaFreeHost(h->loss4_host);
    for (int i = 0; i < 2; ++i) { if (h->stage[i]) cudaFreeHost(h->stage[i]); if (h->stage_ev[i]) cudaEventDestroy(h->stage_ev[i]); }
    h->cnt.release(); h->block_sums.release(); h->pair_start.release(); h->pair_foc.release(); h->foc_row.release(); h->pair_crow.release(); h->Hp.release();
    h->states.release(); h->raw_stage.release(); h->n_obj.release(); h->foc_scene.release(); h->foc_obj.release(); h->foc_t0.release();
    h->y.release(); h->pred.release(); h->loss_ex.release(); h->ds.release(); h->hact.release(); h->this_rows.release();
    for (auto& w : h->prio_w) w.release(); h->prio_vals.release(); h->prio_out2.release(); h->prio_cum.release(); h->prio_ids.release();
    h->dzact.release(); h->dact.release(); h->ddz.release(); h->pair_frow.release(); h->hmask.release(); h->umask.release();
    h->keep.release(); h->nbr.release(); h->ctx_idx_out.release(); h->cnt_out.release(); h->mask_out.release();
    h->traj.release(); h->metrics.release(); h->slot_metrics.release();
    { auto& q = h->sched; q.keep.release(); q.nbr.release(); q.total.release(); q.cnt.release(); q.block_sums.release(); q.pair_start.release(); q.pair_foc.release(); q.info.release(); q.step_first.release(); q.step_F.release(); q.foc_row.release(); q.pair_crow.release(); q.y.release(); }
    h->roll.release(); h->roll_pos.release(); h->roll_scene.release(); h->roll_obj.release(); h->roll_t0.release(); h->foc_of_obj.release();
    for (auto& ev : h->events) if (ev) cudaEventDestroy(ev);
    for (auto& ev : h->prof_events) cudaEventDestroy(ev);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return NPE_OK;
}

int npe_param_count(const npe_handle* h) { (void)h; return NPARAM; }

int npe_params_set(npe_handle* h, const float* host, int32_t n) {
    NEED(h);
    if (!host || n != NPARAM) return fail(h, NPE_ERR_INVALID, "params_set: expected %d floats", NPARAM);
    CK(cudaMemcpyAsync(h->params, host, NPARAM * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    k_pack_scatter<<<(NPARAM + 255) / 256, 256, 0, h->stream>>>(h->params, h->wpack);
    CKL();
    CK(cudaStreamSynchronize(h->stream));
    h->fwd_valid = false; h->hact_valid = false; h->dact_valid = false;
    h->wtc_dirty = true; h->wtc3_dirty = true; h->wtt3_dirty = true;
    return NPE_OK;
}
int npe_params_get(npe_handle* h, float* host, int32_t n) {
    NEED(h);
    if (!host || n != NPARAM) return fail(h, NPE_ERR_INVALID, "params_get: expected %d floats", NPARAM);
    CK(cudaMemcpyAsync(host, h->params, NPARAM * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NPE_OK;
}
int npe_grads_get(npe_handle* h, float* host, int32_t n) {
    NEED(h);
    if (!host || n != NPARAM) return fail(h, NPE_ERR_INVALID, "grads_get: expected %d floats", NPARAM);
    CK(cudaMemcpyAsync(host, h->grads, NPARAM * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NPE_OK;
}
int npe_grads_set(npe_handle* h, const float* host, int32_t n) {
    NEED(h);
    if (!host || n != NPARAM) return fail(h, NPE_ERR_INVALID, "grads_set: expected %d floats", NPARAM);
    CK(cudaMemcpyAsync(h->grads, host, NPARAM * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NPE_OK;
}

// Exchanges the current batch slot with `alt` (buffers and their capacities travel together).
static void swap_batch_slot(npe_handle* h) {
    auto& a = h->alt;
    std::swap(h->states, a.states); std::swap(h->keep, a.keep); std::swap(h->nbr, a.nbr); std::swap(h->cnt, a.cnt);
    std::swap(h->block_sums, a.block_sums); std::swap(h->pair_start, a.pair_start); std::swap(h->pair_foc, a.pair_foc);
    std::swap(h->foc_row, a.foc_row); std::swap(h->pair_crow, a.pair_crow);
    std::swap(h->info_dev, a.info_dev); std::swap(h->active_total, a.active_total);
    std::swap(h->n_obj, a.n_obj); std::swap(h->foc_scene, a.foc_scene); std::swap(h->foc_obj, a.foc_obj); std::swap(h->foc_t0, a.foc_t0);
    std::swap(h->foci_B, a.foci_B); std::swap(h->foci_N, a.foci_N);
    h->slot ^= 1;
}

int npe_batch_upload(npe_handle* h, const float* this_past, const float* context, const float* y, int32_t B, int32_t C) {
    NEED(h);
    if (!this_past || B <= 0 || C < 0 || (C > 0 && !context)) return fail(h, NPE_ERR_INVALID, "batch_upload: bad arguments");
    if (C + 1 > MAXOBJ) return fail(h, NPE_ERR_INVALID, "batch_upload: at most %d contexts", MAXOBJ - 1);
    const int N = C + 1;
    // everything enqueued so far works on the current slot: mark its release point, then move to the other slot
    CK(cudaEventRecord(h->slot_free[h->slot], h->stream));
    h->slot_free_valid[h->slot] = true;
    swap_batch_slot(h);
    if (h->slot_free_valid[h->slot]) CK(cudaStreamWaitEvent(h->stream2, h->slot_free[h->slot], 0));
    const bool reshape = h->foci_B != B || h->foci_N != N;   // this slot's focus / object-count arrays are rewritten below
    CK(h->n_obj.ensure(B));
    CK(h->foc_scene.ensure(B)); CK(h->foc_obj.ensure(B)); CK(h->foc_t0.ensure(B));
    h->S = B; h->Nmax = N; h->T = 2;
    int rc = ensure_batch_buffers(h, B);
    if (rc) return rc;
    // scene b = [this_b | context_b,0 | ... | context_b,C-1], T = 2  (== `past` of eval.lua:283), followed by y (B,22):
    // packed on the host into page-locked staging, ONE H2D copy
    const size_t n_state = (size_t)B * N * IN, n_y = y ? (size_t)B * OUT : 0;
    if (n_state + n_y > h->stage_cap) {
        for (int i = 0; i < 2; ++i) {
            if (h->stage[i]) { cudaEventSynchronize(h->stage_ev[i]); cudaFreeHost(h->stage[i]); h->stage[i] = nullptr; }
            CK(cudaHostAlloc((void**)&h->stage[i], (n_state + (size_t)B * OUT) * 2 * sizeof(float), cudaHostAllocDefault));
        }
        h->stage_cap = (n_state + (size_t)B * OUT) * 2;
    }
    const int si = h->stage_idx;
    h->stage_idx ^= 1;
    CK(cudaEventSynchronize(h->stage_ev[si]));
    float* sg = h->stage[si];
    for (int b = 0; b < B; ++b) {
        memcpy(sg + (size_t)b * N * IN, this_past + (size_t)b * IN, IN * sizeof(float));
        if (C > 0) memcpy(sg + (size_t)b * N * IN + IN, context + (size_t)b * C * IN, (size_t)C * IN * sizeof(float));
    }
    CK(h->states.ensure(n_state + (size_t)B * OUT));
    if (y) memcpy(sg + n_state, y, n_y * sizeof(float));
    cudaStream_t main_stream = h->stream;
    h->stream = h->stream2;            // the copy and the pair builder of this slot run on the upload stream
    cudaError_t e = cudaMemcpyAsync(h->states.p, sg, (n_state + n_y) * sizeof(float), cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaEventRecord(h->stage_ev[si], h->stream);
    h->y_view = y ? h->states.p + n_state : nullptr;
    if (e == cudaSuccess && reshape) {
        k_batch_foci<<<(B + 255) / 256, 256, 0, h->stream>>>(h->foc_scene.p, h->foc_obj.p, h->foc_t0.p, h->n_obj.p, B, N);
        e = cudaGetLastError();
        h->foci_B = B; h->foci_N = N;
    }
    h->foc_first = 0; h->F = B;
    h->batch_form = true; h->have_y = (y != nullptr);
    h->pairs_valid = false; h->fwd_valid = false;
    h->win_start.clear();
    h->scene_foci.clear();
    rc = (e == cudaSuccess) ? build_pairs(h) : NPE_OK;
    if (e == cudaSuccess && rc == NPE_OK) e = cudaEventRecord(h->slot_ready, h->stream);
    h->stream = main_stream;
    h->after_builder = false;   // the builder ran on the upload stream; the event wait below is a full dependency
    if (e != cudaSuccess) return fail(h, NPE_ERR_CUDA, "batch_upload: %s", cudaGetErrorString(e));
    if (rc) return rc;
    CK(cudaStreamWaitEvent(h->stream, h->slot_ready, 0));   // later work on the main stream sees the new batch
    return NPE_OK;
}

int npe_scenes_upload(npe_handle* h, const float* states, const int32_t* n_obj, int32_t S, int32_t Nmax, int32_t T) {
    NEED(h);
    if (!states || !n_obj || S <= 0 || Nmax <= 0 || T < 2) return fail(h, NPE_ERR_INVALID, "scenes_upload: bad arguments");
    if (Nmax > MAXOBJ) return fail(h, NPE_ERR_INVALID, "scenes_upload: at most %d objects per scene", MAXOBJ);
    CK(cudaStreamSynchronize(h->stream2));
    const size_t n = (size_t)S * Nmax * T * D;
    CK(h->states.ensure(n));
    CK(h->n_obj.ensure(S));
    CK(cudaMemcpyAsync(h->states.p, states, n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->n_obj.p, n_obj, (size_t)S * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    h->foci_B = h->foci_N = -1;
    h->scene_foci.assign(S, {});
    h->scene_pred_objects = 0;
    for (int s = 0; s < S; ++s) {
        if (n_obj[s] < 1 || n_obj[s] > Nmax) return fail(h, NPE_ERR_INVALID, "scenes_upload: n_obj[%d] = %d out of range", s, n_obj[s]);
        for (int o = 0; o < n_obj[s]; ++o) {
            const float* st = states + (((size_t)s * Nmax + o) * T) * D;
            if (st[11] != 1.f) h->scene_foci[s].push_back(o);   // oid one-hot col 11 = obstacle: never a focus
        }
        h->scene_pred_objects += (int64_t)h->scene_foci[s].size();
    }
    CK(cudaStreamSynchronize(h->stream));
    h->S = S; h->Nmax = Nmax; h->T = T;
    h->F = 0; h->foc_first = 0;
    h->batch_form = false; h->have_y = false; h->y_view = nullptr;
    h->pairs_valid = false; h->fwd_valid = false;
    h->win_start.clear();
    return NPE_OK;
}

int npe_scenes_upload_raw(npe_handle* h, const float* raw, const int32_t* n_obj, int32_t S, int32_t Nmax, int32_t T) {
    NEED(h);
    if (!raw || !n_obj || S <= 0 || Nmax <= 0 || T < 2) return fail(h, NPE_ERR_INVALID, "scenes_upload_raw: bad arguments");
    if (Nmax > MAXOBJ) return fail(h, NPE_ERR_INVALID, "scenes_upload_raw: at most %d objects per scene", MAXOBJ);
    CK(cudaStreamSynchronize(h->stream2));
    const size_t nframes = (size_t)S * Nmax * T;
    CK(h->states.ensure(nframes * D));
    CK(h->n_obj.ensure(S));
    CK(h->raw_stage.ensure(nframes * 12));
    CK(cudaMemcpyAsync(h->raw_stage.p, raw, nframes * 12 * sizeof(float), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->n_obj.p, n_obj, (size_t)S * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(h->info_dev + 2, 0, sizeof(int), h->stream));
    k_raw_to_state<<<(unsigned)((nframes + 255) / 256), 256, 0, h->stream>>>(h->raw_stage.p, h->states.p, (long long)nframes, h->info_dev + 2);
    CKL();
    h->foci_B = h->foci_N = -1;
    h->scene_foci.assign(S, {});
    h->scene_pred_objects = 0;
    for (int s = 0; s < S; ++s) {
        if (n_obj[s] < 1 || n_obj[s] > Nmax) return fail(h, NPE_ERR_INVALID, "scenes_upload_raw: n_obj[%d] = %d out of range", s, n_obj[s]);
        for (int o = 0; o < n_obj[s]; ++o)
            if (raw[(((size_t)s * Nmax + o) * T) * 12 + 7] != 2.f) h->scene_foci[s].push_back(o);   // objtype 2 = obstacle: never a focus
        h->scene_pred_objects += (int64_t)h->scene_foci[s].size();
    }
    int bad = 0;
    CK(cudaMemcpyAsync(&bad, h->info_dev + 2, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->S = S; h->Nmax = Nmax; h->T = T;
    h->F = 0; h->foc_first = 0;
    h->batch_form = false; h->have_y = false; h->y_view = nullptr;
    h->pairs_valid = false; h->fwd_valid = false;
    h->win_start.clear();
    if (bad) return fail(h, NPE_ERR_INVALID, "scenes_upload_raw: a mass / objtype / sizemul / flag value is outside its category set");
    return NPE_OK;
}

int npe_windows_upload(npe_handle* h, const int32_t* scene_ids, const int32_t* t0, int32_t W) {
    NEED(h);
    if (h->batch_form || h->S == 0 || h->scene_foci.empty()) return fail(h, NPE_ERR_STATE, "windows_upload: upload scenes first");
    if (!scene_ids || !t0 || W <= 0) return fail(h, NPE_ERR_INVALID, "windows_upload: bad arguments");
    std::vector<int> fs, fo, ft;
    h->win_start.assign(1, 0);
    for (int w = 0; w < W; ++w) {
        const int s = scene_ids[w];
        if (s < 0 || s >= h->S || t0[w] < 0 || t0[w] + 3 > h->T)
            return fail(h, NPE_ERR_INVALID, "windows_upload: window %d (scene %d, t0 %d) out of range", w, s, t0[w]);
        for (int o : h->scene_foci[s]) { fs.push_back(s); fo.push_back(o); ft.push_back(t0[w]); }
        h->win_start.push_back((int)fs.size());
    }
    const size_t n = fs.size();
    if (n == 0) return fail(h, NPE_ERR_INVALID, "windows_upload: no focus objects");
    CK(h->foc_scene.ensure(n)); CK(h->foc_obj.ensure(n)); CK(h->foc_t0.ensure(n));
    CK(cudaMemcpyAsync(h->foc_scene.p, fs.data(), n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->foc_obj.p, fo.data(), n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->foc_t0.p, ft.data(), n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->F = 0;
    h->pairs_valid = false; h->fwd_valid = false;
    return NPE_OK;
}

int npe_select_windows(npe_handle* h, int32_t first, int32_t count) {
    NEED(h);
    const int W = (int)h->win_start.size() - 1;
    if (W <= 0) return fail(h, NPE_ERR_STATE, "select_windows: no window schedule uploaded");
    if (first < 0 || count <= 0 || first + count > W) return fail(h, NPE_ERR_INVALID, "select_windows: range out of bounds");
    h->foc_first = h->win_start[first];
    h->F = h->win_start[first + count] - h->foc_first;
    if (h->F <= 0) return fail(h, NPE_ERR_INVALID, "select_windows: empty batch");
    h->have_y = true;   // relative targets are written by the pair builder (K1 + K7)
    h->pairs_valid = false; h->fwd_valid = false;
    static const bool eager_ok = getenv("NPE_NO_EAGER_PAIRS") == nullptr;    // A/B measurements
    if (eager_ok && h->F > h->pdl_max_foci && h->pre_reduce_valid && h->prof_kernel < 0) {
        // a large-batch step is in flight: build the tables of this batch now, on the upload stream, beside its reduce kernel
        auto& a = h->tab_alt;
        CK(cudaEventRecord(h->tab_free[h->tab_cur], h->stream));            // everything enqueued so far used the current set
        h->tab_free_valid[h->tab_cur] = true;
        std::swap(h->keep, a.keep); std::swap(h->nbr, a.nbr); std::swap(h->cnt, a.cnt); std::swap(h->block_sums, a.block_sums);
        std::swap(h->pair_start, a.pair_start); std::swap(h->pair_foc, a.pair_foc); std::swap(h->foc_row, a.foc_row);
        std::swap(h->pair_crow, a.pair_crow); std::swap(h->y, a.y);
        std::swap(h->info_dev, a.info_dev); std::swap(h->active_total, a.active_total);
        h->tab_cur ^= 1;
        int rc = ensure_batch_buffers(h, h->F);
        if (rc) return rc;
        CK(cudaStreamWaitEvent(h->stream2, h->pre_reduce, 0));
        h->pre_reduce_valid = false;
        if (h->tab_free_valid[h->tab_cur]) CK(cudaStreamWaitEvent(h->stream2, h->tab_free[h->tab_cur], 0));
        cudaStream_t main_stream = h->stream;
        h->stream = h->stream2;
        rc = build_pairs(h);
        cudaError_t e = rc == NPE_OK ? cudaEventRecord(h->tab_ready, h->stream) : cudaSuccess;
        h->stream = main_stream;
        h->after_builder = false;                                            // the event wait below is a full dependency
        if (rc) return rc;
        CK(e);
        CK(cudaStreamWaitEvent(h->stream, h->tab_ready, 0));                 // later work on the main stream sees the tables
        return NPE_OK;
    }
    int rc = ensure_batch_buffers(h, h->F);
    if (rc) return rc;
    return NPE_OK;
}

int npe_foci_upload(npe_handle* h, const int32_t* scene_ids, const int32_t* obj_ids, const int32_t* t0, int32_t n) {
    NEED(h);
    if (h->batch_form || h->S == 0) return fail(h, NPE_ERR_STATE, "foci_upload: upload scenes first");
    if (!scene_ids || !obj_ids || !t0 || n <= 0) return fail(h, NPE_ERR_INVALID, "foci_upload: bad arguments");
    for (int i = 0; i < n; ++i)
        if (scene_ids[i] < 0 || scene_ids[i] >= h->S || obj_ids[i] < 0 || obj_ids[i] >= h->Nmax || t0[i] < 0 || t0[i] + 3 > h->T)
            return fail(h, NPE_ERR_INVALID, "foci_upload: example %d out of range", i);
    CK(h->foc_scene.ensure(n)); CK(h->foc_obj.ensure(n)); CK(h->foc_t0.ensure(n));
    CK(cudaMemcpyAsync(h->foc_scene.p, scene_ids, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->foc_obj.p, obj_ids, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->foc_t0.p, t0, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->win_start.resize((size_t)n + 1);
    for (int i = 0; i <= n; ++i) h->win_start[i] = i;   // one-example "windows": select_foci == select_windows
    h->F = 0;
    h->pairs_valid = false; h->fwd_valid = false;
    return NPE_OK;
}

int npe_select_foci(npe_handle* h, int32_t first, int32_t count) { return npe_select_windows(h, first, count); }

int npe_num_foci(const npe_handle* h) { return h ? h->F : 0; }
int npe_ctx_slots(const npe_handle* h) { return h ? ctx_slots(h) : 0; }

int npe_batch_download(npe_handle* h, float* this_past_out, float* y_out) {
    NEED(h);
    if (h->F <= 0) return fail(h, NPE_ERR_STATE, "no batch selected");
    if (this_past_out) {
        CK(h->this_rows.ensure((size_t)h->F * IN));
        k_gather_this<<<((size_t)h->F * IN + 255) / 256, 256, 0, h->stream>>>(h->src(), h->this_rows.p);
        CKL();
        CK(cudaMemcpyAsync(this_past_out, h->this_rows.p, (size_t)h->F * IN * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    }
    if (y_out) {
        if (!h->have_y) return fail(h, NPE_ERR_STATE, "batch has no targets");
        { int rc = build_pairs(h); if (rc) return rc; }
        CK(cudaMemcpyAsync(y_out, h->yptr(), (size_t)h->F * OUT * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    return NPE_OK;
}

int npe_build_pairs(npe_handle* h, int32_t* ctx_idx_out, uint8_t* mask_out, int32_t* cnt_out) {
    NEED(h);
    int rc = build_pairs(h, ctx_idx_out || mask_out || cnt_out);
    if (rc) return rc;
    if (ctx_idx_out || mask_out || cnt_out) {
        const int K = npe_ctx_slots(h);
        const size_t n = (size_t)h->F * (K > 0 ? K : 1);
        CK(h->ctx_idx_out.ensure(n)); CK(h->mask_out.ensure(n)); CK(h->cnt_out.ensure(h->F));
        k_expand_pairs<<<(h->F + 255) / 256, 256, 0, h->stream>>>(h->keep.p, h->nbr.p, h->F, K, h->ctx_idx_out.p,
                                                                  h->mask_out.p, h->cnt_out.p);
        CKL();
        if (ctx_idx_out && K > 0) CK(cudaMemcpyAsync(ctx_idx_out, h->ctx_idx_out.p, (size_t)h->F * K * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        if (mask_out && K > 0) CK(cudaMemcpyAsync(mask_out, h->mask_out.p, (size_t)h->F * K, cudaMemcpyDeviceToHost, h->stream));
        if (cnt_out) CK(cudaMemcpyAsync(cnt_out, h->cnt_out.p, (size_t)h->F * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    return NPE_OK;
}

int npe_forward(npe_handle* h, float* pred_out, float* loss3_out) {
    NEED(h);
    if (h->F <= 0) return fail(h, NPE_ERR_STATE, "no batch selected");
    int rc = run_forward(h, h->precision == NPE_PRECISION_FP32);
    if (rc) return rc;
    if (loss3_out) {
        if (!h->have_y) return fail(h, NPE_ERR_STATE, "batch has no targets");
        if ((rc = run_loss_reduce(h))) return rc;
        CK(cudaMemcpyAsync(loss3_out, h->loss3_dev, 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    }
    if (pred_out) CK(cudaMemcpyAsync(pred_out, h->pred.p, (size_t)h->F * OUT * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (pred_out || loss3_out) return check_pair_overflow(h);
    return NPE_OK;
}

int npe_forward_per_example(npe_handle* h, float* pred_out, float* loss_out) {
    NEED(h);
    if (h->F <= 0) return fail(h, NPE_ERR_STATE, "no batch selected");
    int rc = run_forward(h, false);
    if (rc) return rc;
    if (loss_out) {
        if (!h->have_y) return fail(h, NPE_ERR_STATE, "batch has no targets");
        CK(cudaMemcpyAsync(loss_out, h->loss_ex.p, (size_t)h->F * 3 * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    }
    if (pred_out) CK(cudaMemcpyAsync(pred_out, h->pred.p, (size_t)h->F * OUT * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (pred_out || loss_out) return check_pair_overflow(h);
    return NPE_OK;
}

int npe_backward(npe_handle* h, int32_t divisor_batch) {
    NEED(h);
    return run_backward(h, divisor_batch);
}

int npe_adam_step(npe_handle* h, float lr, float beta1, float beta2, float eps) {
    NEED(h);
    return run_adam(h, lr, beta1, beta2, eps);
}
int npe_rmsprop_step(npe_handle* h, float lr, float alpha, float eps) {
    NEED(h);
    return run_rmsprop(h, lr, alpha, eps);
}
int npe_optim_reset(npe_handle* h, float rmsprop_m_init) {
    NEED(h);
    h->adam_t = 0;
    h->rms_m_init = rmsprop_m_init;
    CK(cudaMemsetAsync(h->adam_m, 0, NPARAM * sizeof(float), h->stream));
    CK(cudaMemsetAsync(h->adam_v, 0, NPARAM * sizeof(float), h->stream));
    k_fill<<<(NPARAM + 255) / 256, 256, 0, h->stream>>>(h->rms_m, NPARAM, rmsprop_m_init);
    CKL();
    return NPE_OK;
}

int npe_train_step(npe_handle* h, int32_t opt, float lr, int32_t divisor_batch, float* loss3_out) {
    NEED(h);
    if (h->F <= 0) return fail(h, NPE_ERR_STATE, "no batch selected");
    int rc;
    if (!h->have_y) return fail(h, NPE_ERR_STATE, "batch has no targets");
    if (opt != 0 && opt != 1) return fail(h, NPE_ERR_INVALID, "train_step: opt must be 0 (adam) or 1 (rmsprop)");
    const bool want_loss = loss3_out != nullptr || h->loss4_override != nullptr;
    const bool one_grid = (h->peer_ready || !h->comm) && step_fused_ok(h);
    if (one_grid) {
        // small batch: every tile of the forward and the backward is a CTA of one grid, then reduce (+ exchange) + optimiser
        if ((rc = run_step_fused(h, divisor_batch, h->peer_ready ? 0 : (opt == 0 ? 1 : 2), lr, want_loss,
                                 h->peer_ready ? (opt == 0 ? 1 : 2) : 0))) return rc;
    } else if ((rc = run_forward(h, true, /*fused=*/true))) {
        return rc;
    } else if (h->peer_ready) {
        // data parallel over NVLink peer memory: local reduce + publish, then all-reduce fused with the optimiser
        if ((rc = run_backward(h, divisor_batch, 0, lr, want_loss, opt == 0 ? 1 : 2, true))) return rc;
    } else if (!h->comm) {
        // single GPU: reduction of the CTA partials, optimiser step, weight-image write-through and the batch loss in ONE kernel
        if ((rc = run_backward(h, divisor_batch, opt == 0 ? 1 : 2, lr, want_loss, 0, true))) return rc;
    } else {
        if ((rc = run_backward(h, divisor_batch, 0, 0.f, want_loss, 0, true))) return rc;
        if ((rc = run_allreduce(h))) return rc;
        rc = (opt == 0) ? run_adam(h, lr, 0.9f, 0.999f, 1e-8f) : run_rmsprop(h, lr, 0.99f, 1e-8f);
        if (rc) return rc;
    }
    h->fwd_valid = false;   // parameters changed
    if (want_loss && !h->loss4_override) {
        // the loss kernel wrote {loss, vel, angvel, overflow} straight into host-mapped memory: one sync, no copy
        CK(cudaStreamSynchronize(h->stream));
        if (h->peer_ready) { int rcp = check_peer_timeout(h); if (rcp) return rcp; }
        if (h->loss4_host[3] == 2.f) return fail(h, NPE_ERR_CUDA, "k_step_fused: a tile waited too long for its producer; results are invalid");
        if (h->loss4_host[3] != 0.f)
            return fail(h, NPE_ERR_NOMEM, "active-pair list overflow (capacity %d pairs); results are invalid", h->pair_cap);
        loss3_out[0] = h->loss4_host[0]; loss3_out[1] = h->loss4_host[1]; loss3_out[2] = h->loss4_host[2];
    }
    return NPE_OK;
}

// Pair tables of steps [s0, s0 + n) of the schedule in one launch (CTA s = step s); false when a step does not fit one
// builder CTA (more than 64 foci), in which case the caller builds the tables step by step.
constexpr int SCHED_CHUNK = 1024;
constexpr int STEP_GRAPH = 8;         // npe_train_steps: steps per captured graph,
constexpr int STEP_PLAIN = 4;         // plain launches at the start of a chunk (the GPU works while the first graph is captured),
constexpr int STEP_EXECS = 4;         // executable graphs in the ring,
constexpr int STEP_GRAPH_MIN = STEP_PLAIN + STEP_GRAPH;    // shortest chunk that gets a graph.  A slot's FIRST use captures and
                                      // instantiates (~0.2 ms, once per handle and optimiser); after that a 20-step call at N = 2
                                      // runs 30.0-30.7 us per step against 32.9 with plain launches (profiles/r2_dp_exchange.md)
static int sched_build(npe_handle* h, int first, int batch, int s0, int n, int* fmax_out, int* cap_out) {
    // page-locked staging (one allocation per handle): the copies below are asynchronous and nothing waits for them -- a
    // pageable source plus a stream synchronisation here was a ~20-us bubble at the head of every npe_train_steps call.
    // The previous user of the staging area (the previous chunk) ended with a stream synchronisation.
    if (!h->sched_host) CK(cudaHostAlloc((void**)&h->sched_host, (size_t)2 * SCHED_CHUNK * sizeof(int), cudaHostAllocDefault));
    int* const sf = h->sched_host; int* const sF = h->sched_host + SCHED_CHUNK;
    int fmax = 0;
    for (int i = 0; i < n; ++i) {
        const int w = first + (s0 + i) * batch;
        sf[i] = h->win_start[(size_t)w];
        sF[i] = h->win_start[(size_t)w + batch] - sf[i];
        if (sF[i] <= 0) return fail(h, NPE_ERR_INVALID, "train_steps: empty batch");
        fmax = sF[i] > fmax ? sF[i] : fmax;
    }
    int per = ctx_slots(h);
    per = per > 16 ? 16 : (per < 1 ? 1 : per);
    const int cap = fmax * per;
    auto& q = h->sched;
    const size_t nf = (size_t)n * fmax, np = (size_t)n * cap;
    CK(q.keep.ensure(nf)); CK(q.nbr.ensure(nf)); CK(q.cnt.ensure(nf)); CK(q.foc_row.ensure(nf)); CK(q.y.ensure(nf * D));
    CK(q.block_sums.ensure(n)); CK(q.total.ensure(n)); CK(q.info.ensure((size_t)2 * n));
    CK(q.pair_start.ensure((size_t)n * (fmax + 1))); CK(q.pair_foc.ensure(np)); CK(q.pair_crow.ensure(np));
    CK(q.step_first.ensure(n)); CK(q.step_F.ensure(n));
    CK(cudaMemcpyAsync(q.step_first.p, sf, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(q.step_F.p, sF, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemsetAsync(q.total.p, 0, (size_t)n * sizeof(u64), h->stream));
    SchedOut o;
    o.keep = q.keep.p; o.nbr = q.nbr.p; o.cnt = q.cnt.p; o.block_sums = q.block_sums.p; o.foc_row = q.foc_row.p; o.y = q.y.p;
    o.total = q.total.p; o.pair_start = q.pair_start.p; o.pair_foc = q.pair_foc.p; o.pair_crow = q.pair_crow.p; o.info = q.info.p;
    FocusSrc src = h->src();
    src.foc_scene = h->foc_scene.p; src.foc_obj = h->foc_obj.p; src.foc_t0 = h->foc_t0.p;
    const float thr = nbr_threshold_sq(h->cfg.nbrhdsize * 60.f);
    {
        ProfScope ps__(h, NPE_K_BUILD_PAIRS);
        if (h->Nmax - 1 <= 8) k_build_pairs_sched<8><<<n, 512, 0, h->stream>>>(src, q.step_first.p, q.step_F.p, fmax, cap, h->cfg.max_ctx, thr, o);
        else if (h->Nmax - 1 <= 16) k_build_pairs_sched<16><<<n, 1024, 0, h->stream>>>(src, q.step_first.p, q.step_F.p, fmax, cap, h->cfg.max_ctx, thr, o);
        else k_build_pairs_sched<32><<<n, NT_PAIRS, 0, h->stream>>>(src, q.step_first.p, q.step_F.p, fmax, cap, h->cfg.max_ctx, thr, o);
        h->launches++;
        h->after_builder = true;
    }
    CKL();
    *fmax_out = fmax; *cap_out = cap;
    return NPE_OK;
}

int npe_train_step_async(npe_handle* h, int32_t opt, float lr, int32_t divisor_batch, int32_t slot) {
    NEED(h);
    if (slot < 0 || slot >= LOSS_RING) return fail(h, NPE_ERR_INVALID, "train_step_async: slot must be in [0, %d)", LOSS_RING);
    if (!h->ring_ev[slot]) CK(cudaEventCreateWithFlags(&h->ring_ev[slot], cudaEventDisableTiming));
    h->loss4_override = h->loss4_devptr + 16 + 4 * slot;
    const int rc = npe_train_step(h, opt, lr, divisor_batch, nullptr);
    h->loss4_override = nullptr;
    if (rc) return rc;
    CK(cudaEventRecord(h->ring_ev[slot], h->stream));
    return NPE_OK;
}

int npe_train_batch_async(npe_handle* h, const float* this_past, const float* context, const float* y, int32_t B, int32_t C,
                          int32_t opt, float lr, int32_t divisor_batch, int32_t slot) {
    const int rc = npe_batch_upload(h, this_past, context, y, B, C);
    return rc ? rc : npe_train_step_async(h, opt, lr, divisor_batch, slot);
}

int npe_loss_wait(npe_handle* h, int32_t slot, float* loss3_out) {
    NEED(h);
    if (slot < 0 || slot >= LOSS_RING || !h->ring_ev[slot]) return fail(h, NPE_ERR_INVALID, "loss_wait: slot %d was never issued", slot);
    CK(cudaEventSynchronize(h->ring_ev[slot]));
    const volatile float* r = h->loss4_host + 16 + 4 * slot;
    if (r[3] != 0.f) return fail(h, NPE_ERR_NOMEM, "active-pair list overflow (capacity %d pairs); results are invalid", h->pair_cap);
    if (loss3_out) { loss3_out[0] = r[0]; loss3_out[1] = r[1]; loss3_out[2] = r[2]; }
    return NPE_OK;
}

int npe_train_steps(npe_handle* h, int32_t first, int32_t batch, int32_t nsteps, int32_t opt, float lr,
                    int32_t divisor_batch, float* losses_out) {
    NEED(h);
    if (batch <= 0 || nsteps <= 0 || first < 0) return fail(h, NPE_ERR_INVALID, "train_steps: bad arguments");
    const int W = (int)h->win_start.size() - 1;
    if (W <= 0) return fail(h, NPE_ERR_STATE, "train_steps: no focus / window schedule uploaded");
    if ((long long)first + (long long)nsteps * batch > W) return fail(h, NPE_ERR_INVALID, "train_steps: schedule has %d entries", W);
    if (losses_out) CK(h->loss_hist.ensure((size_t)nsteps * 4));
    int rc = NPE_OK;
    // The batcher runs one chunk of steps ahead of the training kernels: ONE launch builds the pair tables of up to
    // SCHED_CHUNK steps, each step then works on its slice.
    int Fmax = 0;
    for (int s = 0; s < nsteps; ++s) {
        const int w = first + s * batch;
        const int Fs = h->win_start[(size_t)w + batch] - h->win_start[(size_t)w];
        Fmax = Fs > Fmax ? Fs : Fmax;
    }
    const bool sched = Fmax <= FPB && !getenv("NPE_NO_SCHED");   // env switch: A/B measurements only
    if ((rc = ensure_batch_buffers(h, Fmax))) return rc;   // no reallocation while the per-step slices are installed
    for (int s0 = 0; s0 < nsteps && rc == NPE_OK; s0 += SCHED_CHUNK) {
        const int n = nsteps - s0 < SCHED_CHUNK ? nsteps - s0 : SCHED_CHUNK;
        int fmax = 0, cap = 0;
        if (sched && (rc = sched_build(h, first, batch, s0, n, &fmax, &cap))) break;
        // base pointers of the per-batch buffers; restored after the chunk
        u64* const b_keep = h->keep.p; u64* const b_nbr = h->nbr.p; int* const b_cnt = h->cnt.p; int* const b_bs = h->block_sums.p;
        long long* const b_row = h->foc_row.p; float* const b_y = h->y.p; int* const b_ps = h->pair_start.p;
        int* const b_pf = h->pair_foc.p; long long* const b_pc = h->pair_crow.p; int* const b_info = h->info_dev;
        u64* const b_tot = h->active_total;
        // The steps of a chunk run as CUDA graphs of STEP_GRAPH steps each: a grid that has written into a PEER's memory pays
        // a system-scope flush when it completes as a stream launch -- +2.2 us per grid, +3.8 us when its dependent is a
        // programmatic launch (tools/microbench/nvlink_ipc_store.cu) -- and none as a node of a graph, whose edges stay on the
        // device.  A ring of STEP_EXECS executable graphs: a slot is CAPTURED once (the very same launches into a capturing
        // stream; launch_pdl notes the kernel nodes) and from then on its nodes just get the parameters of the launches
        // (launch_pdl in sink mode: cudaGraphExecKernelNodeSetParams, ~1-2 us per node against ~13 us for capture +
        // cudaGraphExecUpdate), after an event says the slot's previous launch has finished.  The first STEP_PLAIN steps of
        // a chunk are plain launches, so that the GPU has work while the host prepares the first graph.
        static const bool graph_ok = getenv("NPE_NO_STEP_GRAPH") == nullptr;    // A/B measurements
        static const int graph_min = getenv("NPE_STEP_GRAPH_MIN") ? atoi(getenv("NPE_STEP_GRAPH_MIN")) : STEP_GRAPH_MIN;    // experiments
        // only for chunks whose every step is the one-grid step (k_step_fused + the reduce / exchange kernel, both through
        // launch_pdl): other paths enqueue work that the sink mode of launch_pdl does not see
        static const bool fused_off = getenv("NPE_NO_FUSED_STEP") != nullptr;
        const int gp_c = (cap + SR - 1) / SR, gd_c = (fmax + SR - 1) / SR;
        const bool fused_chunk = !fused_off && small_batch(h, fmax) && (h->peer_ready || !h->comm) && gp_c >= 1 &&
                                 gp_c + gd_c <= h->num_sms && gp_c + gd_c <= 500;
        const bool use_graph = graph_ok && sched && fused_chunk && h->prof_kernel < 0 && step_graph_wanted(h) && n >= graph_min;
        int group_first = -1, group_end = -1, plain_until = -1;      // the group being built; steps replayed as plain launches
        std::chrono::steady_clock::time_point group_t0;
        struct { unsigned step_stamp, peer_step; int adam_t; int64_t launches; } snap = {0, 0, 0, 0};
        auto leave_capture = [&]() {          // error path: the stream must not stay in capture mode
            cudaGraph_t g = nullptr;
            cudaStreamEndCapture(h->stream, &g);
            if (g) cudaGraphDestroy(g);
            cudaGetLastError();
            h->step_nodes[h->sink_slot].clear();
        };
        // finishes the group in slot k: 0 = launched, 1 = the sink did not fit (replay the group as plain launches), < 0 = error
        auto finish_group = [&]() -> int {
            const int k = h->sink_slot, mode = h->sink_mode;
            h->sink_mode = 0;
            cudaError_t e = cudaSuccess;
            if (mode == 1) {
                cudaGraph_t g = nullptr;
                e = cudaStreamEndCapture(h->stream, &g);
                if (e != cudaSuccess || !g) { cudaGetLastError(); return fail(h, NPE_ERR_CUDA, "train_steps: stream capture failed: %s", cudaGetErrorString(e)); }
                size_t nn = 0;
                cudaGraphGetNodes(g, nullptr, &nn);
                if (!h->sink_ok || nn != h->step_nodes[k].size()) h->step_nodes[k].clear();     // not a pure chain of programmatic launches: capture it every time
                h->step_graph[k] = g;
                h->step_exec_opt[k] = opt;
                e = cudaGraphInstantiate(&h->step_exec[k], g, 0);
                h->graphs_captured++;
            } else if (!h->sink_ok || h->sink_idx != (int)h->step_nodes[k].size()) {
                h->step_nodes[k].clear();         // next time this slot is captured again
                h->graphs_replayed++;
                return 1;
            } else h->graphs_set++;
            if (e == cudaSuccess) e = cudaGraphLaunch(h->step_exec[k], h->stream);
            if (e == cudaSuccess && !h->step_exec_ev[k]) e = cudaEventCreateWithFlags(&h->step_exec_ev[k], cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventRecord(h->step_exec_ev[k], h->stream);
            if (e != cudaSuccess) return fail(h, NPE_ERR_CUDA, "train_steps: graph launch failed: %s", cudaGetErrorString(e));
            h->step_exec_next = (k + 1) % STEP_EXECS;
            return 0;
        };
        for (int i = 0; i < n && rc == NPE_OK; ++i) {
            const int s = s0 + i;
            if (use_graph && h->sink_mode == 0 && h->graph_warm[opt & 1] && i >= STEP_PLAIN && i > plain_until && n - i >= STEP_GRAPH) {
                const int k = h->step_exec_next;
                cudaError_t e = cudaSuccess;
                if (h->step_exec[k] && h->step_exec_ev[k]) e = cudaEventSynchronize(h->step_exec_ev[k]);    // its previous launch has finished
                const bool reuse = e == cudaSuccess && h->step_exec[k] && !h->step_nodes[k].empty() && h->step_exec_opt[k] == opt;
                if (e == cudaSuccess && !reuse) {
                    if (h->step_exec[k]) { cudaGraphExecDestroy(h->step_exec[k]); h->step_exec[k] = nullptr; }
                    if (h->step_graph[k]) { cudaGraphDestroy(h->step_graph[k]); h->step_graph[k] = nullptr; }
                    h->step_nodes[k].clear();
                    // relaxed mode: this thread's other runtime calls (tools-only allocations, lazy module loads) stay legal
                    e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed);
                }
                if (e != cudaSuccess) { rc = fail(h, NPE_ERR_CUDA, "train_steps: graph slot: %s", cudaGetErrorString(e)); break; }
                h->sink_mode = reuse ? 2 : 1; h->sink_slot = k; h->sink_idx = 0; h->sink_ok = true;
                group_first = i; group_end = i + STEP_GRAPH - 1;
                snap = {h->step_stamp, h->peer_step, h->adam_t, h->launches};
                group_t0 = std::chrono::steady_clock::now();
            }
            if ((rc = npe_select_windows(h, first + s * batch, batch))) break;
            if (sched) {
                auto& q = h->sched;
                const size_t fo = (size_t)i * fmax, po = (size_t)i * cap;
                h->keep.p = q.keep.p + fo; h->nbr.p = q.nbr.p + fo; h->cnt.p = q.cnt.p + fo; h->block_sums.p = q.block_sums.p + i;
                h->foc_row.p = q.foc_row.p + fo; h->y.p = q.y.p + fo * D; h->pair_start.p = q.pair_start.p + (size_t)i * (fmax + 1);
                h->pair_foc.p = q.pair_foc.p + po; h->pair_crow.p = q.pair_crow.p + po; h->info_dev = q.info.p + 2 * i;
                h->active_total = q.total.p + i;
                h->pair_cap = cap;
                h->pairs_valid = true; h->keep_valid = true; h->hact_valid = false; h->dact_valid = false;
            }
            h->loss4_override = losses_out ? h->loss_hist.p + (size_t)s * 4 : nullptr;
            rc = npe_train_step(h, opt, lr, divisor_batch, nullptr);
            if (h->sink_mode != 0 && rc == NPE_OK && i == group_end) {
                const bool set_mode = h->sink_mode == 2;
                const auto t1 = std::chrono::steady_clock::now();
                const int r = finish_group();
                if (set_mode && r == 0) {
                    h->graph_fill_us += std::chrono::duration<double, std::micro>(t1 - group_t0).count();
                    h->graph_launch_us += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t1).count();
                }
                if (r < 0) rc = r;
                else if (r == 1) {
                    // the launches of these steps were not what the slot's nodes hold: nothing has been enqueued for them;
                    // put the step counters back and run the group again as plain launches
                    h->step_stamp = snap.step_stamp; h->peer_step = snap.peer_step; h->adam_t = snap.adam_t; h->launches = snap.launches;
                    plain_until = group_end;
                    i = group_first - 1;
                }
            } else if (h->sink_mode == 0 && rc == NPE_OK) {
                h->graph_warm[opt & 1] = true;      // every kernel of a step has been launched (and loaded) once
            }
        }
        if (h->sink_mode == 1) leave_capture();      // a step failed while the stream was capturing
        h->sink_mode = 0;
        h->loss4_override = nullptr;   // only the npe_train_step calls above may see it (every return path below is clean)
        h->keep.p = b_keep; h->nbr.p = b_nbr; h->cnt.p = b_cnt; h->block_sums.p = b_bs; h->foc_row.p = b_row; h->y.p = b_y;
        h->pair_start.p = b_ps; h->pair_foc.p = b_pf; h->pair_crow.p = b_pc; h->info_dev = b_info; h->active_total = b_tot;
        if (sched) { h->pairs_valid = false; h->fwd_valid = false; h->F = 0; }
        if (sched && rc == NPE_OK) {
            // one look at the chunk's {P, flag} words: a dropped pair (1) or a timed-out hand-over (2) must not pass silently
            std::vector<int> info((size_t)2 * n);
            CK(cudaMemcpyAsync(info.data(), h->sched.info.p, info.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            for (int i = 0; i < n; ++i) {
                if (info[(size_t)2 * i + 1] == 0) continue;
                if (info[(size_t)2 * i + 1] == 2) return fail(h, NPE_ERR_CUDA, "train_steps: step %d: a tile waited too long for its producer; results are invalid", s0 + i);
                if (info[(size_t)2 * i + 1] == 3) return fail(h, NPE_ERR_CUDA, "train_steps: step %d: a peer rank did not deliver its gradient in time; results are invalid", s0 + i);
                return fail(h, NPE_ERR_NOMEM, "train_steps: active-pair list overflow in step %d (capacity %d pairs)", s0 + i, cap);
            }
        }
    }
    h->loss4_override = nullptr;
    if (rc) return rc;
    if (losses_out) {
        std::vector<float> tmp((size_t)nsteps * 4);
        CK(cudaMemcpyAsync(tmp.data(), h->loss_hist.p, tmp.size() * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        for (int s = 0; s < nsteps; ++s) {
            if (tmp[(size_t)s * 4 + 3] != 0.f) return fail(h, NPE_ERR_NOMEM, "active-pair list overflow in step %d", s);
            for (int k = 0; k < 3; ++k) losses_out[(size_t)s * 3 + k] = tmp[(size_t)s * 4 + k];
        }
    }
    return NPE_OK;
}

static int rollout_impl(npe_handle* h, int t0, int steps, int use_gt, int teacher, bool want_traj, bool want_metrics) {
    if (h->batch_form || h->S == 0) return fail(h, NPE_ERR_STATE, "rollout: upload scenes first");
    if (steps <= 0 || t0 < 0 || t0 + 2 > h->T) return fail(h, NPE_ERR_INVALID, "rollout: bad t0/steps");
    if (use_gt && t0 + 2 + steps > h->T)
        return fail(h, NPE_ERR_INVALID, "rollout: %d steps need %d frames of ground truth, scenes have %d", steps, t0 + 2 + steps, h->T);
    RollArgs A;
    A.states = h->states.p; A.n_obj = h->n_obj.p; A.S = h->S; A.Nmax = h->Nmax; A.T = h->T;
    A.t0 = t0; A.steps = steps; A.use_gt = use_gt; A.teacher = teacher;
    A.kmax = h->cfg.max_ctx; A.ball_zero = h->cfg.ball_zero_mode;
    A.thr_s = nbr_threshold_sq(h->cfg.nbrhdsize * 60.f);
    A.wpack = h->wpack;
    A.traj = nullptr; A.metrics = nullptr;
    if (want_traj) {
        CK(h->traj.ensure((size_t)h->S * h->Nmax * steps * D));
        CK(cudaMemsetAsync(h->traj.p, 0, (size_t)h->S * h->Nmax * steps * D * sizeof(float), h->stream));
        A.traj = h->traj.p;
    }
    if (want_metrics) {
        CK(h->metrics.ensure((size_t)steps * 7));
        CK(cudaMemsetAsync(h->metrics.p, 0, (size_t)steps * 7 * sizeof(double), h->stream));
        A.metrics = h->metrics.p;
    }
    A.slot_metrics = nullptr;
    if (want_metrics && h->want_slots) {
        CK(h->slot_metrics.ensure((size_t)steps * h->Nmax * 7));
        CK(cudaMemsetAsync(h->slot_metrics.p, 0, (size_t)steps * h->Nmax * 7 * sizeof(double), h->stream));
        A.slot_metrics = h->slot_metrics.p;
    }
    const int G = TILE / h->Nmax;
    const int ngroups = (h->S + G - 1) / G;
    const int grid = ngroups < h->num_sms ? ngroups : h->num_sms;
    { ProfScope ps__(h, NPE_K_ROLLOUT); k_rollout<<<grid, NT, sizeof(RollSmem), h->stream>>>(A);
    h->launches++; }
    { cudaError_t e__ = cudaGetLastError(); if (e__ != cudaSuccess) return fail(h, NPE_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e__), __FILE__, __LINE__); }
    return NPE_OK;
}

// Rollout as a host loop of kernels over a window tensor in HBM (large scene counts; the tcgen05 variant).
static int rollout_by_kernels(npe_handle* h, int t0, int steps, int use_gt, int teacher, bool want_traj, bool want_metrics) {
    if (h->batch_form || h->S == 0) return fail(h, NPE_ERR_STATE, "rollout: upload scenes first");
    if (steps <= 0 || t0 < 0 || t0 + 2 > h->T) return fail(h, NPE_ERR_INVALID, "rollout: bad t0/steps");
    if (use_gt && t0 + 2 + steps > h->T)
        return fail(h, NPE_ERR_INVALID, "rollout: %d steps need %d frames of ground truth, scenes have %d", steps, t0 + 2 + steps, h->T);
    const int SN = h->S * h->Nmax;
    // foci = every non-obstacle object, window at t0 = 0 of the roll tensor
    std::vector<int> fs, fo, fmap((size_t)SN, -1);
    for (int s = 0; s < h->S; ++s)
        for (int o : h->scene_foci[s]) { fmap[(size_t)s * h->Nmax + o] = (int)fs.size(); fs.push_back(s); fo.push_back(o); }
    const int F = (int)fs.size();
    if (F == 0) return fail(h, NPE_ERR_INVALID, "rollout: no predictable objects");
    std::vector<int> zeros((size_t)F, 0);
    CK(h->roll.ensure((size_t)SN * 2 * D));
    CK(h->roll_scene.ensure(F)); CK(h->roll_obj.ensure(F)); CK(h->roll_t0.ensure(F)); CK(h->foc_of_obj.ensure(SN));
    CK(cudaMemcpyAsync(h->roll_scene.p, fs.data(), (size_t)F * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->roll_obj.p, fo.data(), (size_t)F * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->roll_t0.p, zeros.data(), (size_t)F * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->foc_of_obj.p, fmap.data(), (size_t)SN * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));   // host vectors go out of scope
    int rc = ensure_batch_buffers(h, F);
    if (rc) return rc;
    RollPostArgs A;
    A.roll = h->roll.p; A.states = h->states.p; A.n_obj = h->n_obj.p; A.foc_of_obj = h->foc_of_obj.p; A.pred = h->pred.p;
    A.S = h->S; A.Nmax = h->Nmax; A.T = h->T; A.t0 = t0; A.steps = steps; A.use_gt = use_gt; A.teacher = teacher;
    A.ball_zero = h->cfg.ball_zero_mode; A.traj = nullptr; A.metrics = nullptr; A.step = 0;
    if (want_traj) {
        CK(h->traj.ensure((size_t)SN * steps * D));
        CK(cudaMemsetAsync(h->traj.p, 0, (size_t)SN * steps * D * sizeof(float), h->stream));
        A.traj = h->traj.p;
    }
    if (want_metrics) {
        CK(h->metrics.ensure((size_t)steps * 7));
        CK(cudaMemsetAsync(h->metrics.p, 0, (size_t)steps * 7 * sizeof(double), h->stream));
        A.metrics = h->metrics.p;
    }
    A.slot_metrics = nullptr;
    if (want_metrics && h->want_slots) {
        CK(h->slot_metrics.ensure((size_t)steps * h->Nmax * 7));
        CK(cudaMemsetAsync(h->slot_metrics.p, 0, (size_t)steps * h->Nmax * 7 * sizeof(double), h->stream));
        A.slot_metrics = h->slot_metrics.p;
    }
    k_rollout_init<<<(SN * 2 * D + 255) / 256, 256, 0, h->stream>>>(h->roll.p, h->states.p, SN, h->T, t0);
    CKL();
    FocusSrc src;
    src.states = h->roll.p; src.n_obj = h->n_obj.p; src.Nmax = h->Nmax; src.T = 2;
    src.foc_scene = h->roll_scene.p; src.foc_obj = h->roll_obj.p; src.foc_t0 = h->roll_t0.p; src.F = F;
    // every object of every scene is a focus, in order (ball scenes): the throughput builder derives (scene, object) from the
    // focus index instead of fetching three index arrays and the object count -- two dependent round trips less per block
    bool dense = F > FPB && F == SN;
    for (int s = 0; dense && s < h->S; ++s)
        for (int o = 0; o < h->Nmax; ++o) if (h->scene_foci[s][o] != o) { dense = false; break; }
    if (dense) { src.foc_scene = nullptr; src.foc_obj = nullptr; src.foc_t0 = nullptr; }
    const float thr = nbr_threshold_sq(h->cfg.nbrhdsize * 60.f);
    const int nblk = (F + FPB - 1) / FPB;
    const int tc = h->precision;
    // Scenes in which every object is a focus, no ground truth, no trajectory / metric output (npe_rollout_resident on ball
    // scenes): the 3xTF32 decoder kernel does the post-process and the window shift itself, no separate pass over the window
    const bool fused_post = tc == NPE_PRECISION_TF32X3_TC && !use_gt && !want_traj && !want_metrics && !h->want_slots &&
                            (long long)F == [&] { long long n = 0; for (int s = 0; s < h->S; ++s) n += (long long)h->scene_foci[s].size(); return n; }() &&
                            [&] { for (int s = 0; s < h->S; ++s) if ((int)h->scene_foci[s].size() != h->Nmax) return false; return true; }() &&
                            !getenv("NPE_NO_FUSED_POST");
    // fused post-process + dense focus list over scenes of 2^k <= 32 objects: the decoder kernel also keeps the new positions as
    // a contiguous array for the next step's pair builder (k_build_pairs_scan, by_shuffle)
    const bool compact = fused_post && dense && (h->Nmax & (h->Nmax - 1)) == 0 && h->Nmax <= 32;
    if (compact) CK(h->roll_pos.ensure((size_t)F));
    for (int step = 0; step < steps; ++step) {
        {
            ProfScope ps__(h, NPE_K_BUILD_PAIRS);
            const PairOut o = {h->keep.p, h->nbr.p, h->cnt.p, h->block_sums.p, h->foc_row.p, nullptr};
            if ((rc = launch_build_pairs(h, src, thr, o, false))) return rc;
        }
        if ((rc = launch_forward_kernels(h, h->roll.p, F, nullptr, nullptr, tc, nullptr, false, nullptr, fused_post ? h->roll.p : nullptr,
                                         compact ? h->roll_pos.p : nullptr))) return rc;
        if (compact) src.pos_compact = h->roll_pos.p;                 // the next step's builder reads what this decoder kernel wrote
        A.step = step;
        if (!fused_post) {
            ProfScope ps__(h, NPE_K_ROLLOUT);
            k_rollout_post<<<(SN + POST_OBJ - 1) / POST_OBJ, 256, 0, h->stream>>>(A);
            h->launches++;
        }
    }
    CK(cudaGetLastError());
    h->F = 0; h->pairs_valid = false; h->fwd_valid = false;   // the batch buffers were reused
    h->roll_F = F;
    return NPE_OK;
}

static int rollout_dispatch(npe_handle* h, int t0, int steps, int use_gt, int teacher, bool want_traj, bool want_metrics) {
    if (h->precision != NPE_PRECISION_FP32 || h->rollout_mode == NPE_ROLLOUT_BY_KERNELS)
        return rollout_by_kernels(h, t0, steps, use_gt, teacher, want_traj, want_metrics);
    return rollout_impl(h, t0, steps, use_gt, teacher, want_traj, want_metrics);
}

int npe_rollout(npe_handle* h, int32_t t0, int32_t steps, int32_t use_gt, int32_t teacher, float* traj_out, double* metrics_out) {
    NEED(h);
    int rc = rollout_dispatch(h, t0, steps, use_gt, teacher, traj_out != nullptr, metrics_out != nullptr);
    if (rc) return rc;
    if (traj_out) CK(cudaMemcpyAsync(traj_out, h->traj.p, (size_t)h->S * h->Nmax * steps * D * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    if (metrics_out) CK(cudaMemcpyAsync(metrics_out, h->metrics.p, (size_t)steps * 7 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NPE_OK;
}

int npe_rollout_slots(npe_handle* h, int32_t t0, int32_t steps, int32_t use_gt, int32_t teacher, float* traj_out, double* slot_metrics_out) {
    NEED(h);
    if (!slot_metrics_out) return fail(h, NPE_ERR_INVALID, "rollout_slots: slot_metrics_out is NULL");
    h->want_slots = true;
    int rc = rollout_dispatch(h, t0, steps, use_gt, teacher, traj_out != nullptr, true);
    h->want_slots = false;
    if (rc) return rc;
    if (traj_out) CK(cudaMemcpyAsync(traj_out, h->traj.p, (size_t)h->S * h->Nmax * steps * D * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaMemcpyAsync(slot_metrics_out, h->slot_metrics.p, (size_t)steps * h->Nmax * 7 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NPE_OK;
}

int npe_rollout_resident(npe_handle* h, int32_t t0, int32_t steps, int64_t* object_steps_out) {
    NEED(h);
    int rc = rollout_dispatch(h, t0, steps, 0, 0, false, false);
    if (rc) return rc;
    if (object_steps_out) *object_steps_out = h->scene_pred_objects * (int64_t)steps;
    return NPE_OK;
}

int npe_set_precision(npe_handle* h, int32_t mode) {
    NEED(h);
    if (mode != NPE_PRECISION_FP32 && mode != NPE_PRECISION_TF32_TC && mode != NPE_PRECISION_TF32X3_TC) return fail(h, NPE_ERR_INVALID, "set_precision: unknown mode %d", mode);
    h->precision = mode;
    h->fwd_valid = false;
    return NPE_OK;
}

int npe_set_rollout_mode(npe_handle* h, int32_t mode) {
    NEED(h);
    if (mode != NPE_ROLLOUT_PERSISTENT && mode != NPE_ROLLOUT_BY_KERNELS) return fail(h, NPE_ERR_INVALID, "set_rollout_mode: unknown mode %d", mode);
    h->rollout_mode = mode;
    return NPE_OK;
}

int npe_comm_unique_id(void* id128_out) {
    std::string err;
    if (!id128_out) return NPE_ERR_INVALID;
    if (!g_nccl.load(err)) { g_create_error = err; return NPE_ERR_NCCL; }
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != 0) { g_create_error = "ncclGetUniqueId failed"; return NPE_ERR_NCCL; }
    memcpy(id128_out, &id, sizeof id);
    return NPE_OK;
}

int npe_comm_init(npe_handle* h, int32_t rank, int32_t nranks, const void* id128) {
    NEED(h);
    if (!id128 || nranks < 1 || rank < 0 || rank >= nranks) return fail(h, NPE_ERR_INVALID, "comm_init: bad arguments");
    std::string err;
    if (!g_nccl.load(err)) return fail(h, NPE_ERR_NCCL, "%s", err.c_str());
    ncclUniqueId id;
    memcpy(&id, id128, sizeof id);
    ncclResult_t r = g_nccl.CommInitRank(&h->comm, nranks, id, rank);
    if (r != 0) return fail(h, NPE_ERR_NCCL, "ncclCommInitRank: %s", g_nccl.GetErrorString ? g_nccl.GetErrorString(r) : "?");
    h->rank = rank; h->nranks = nranks;
    return NPE_OK;
}

int npe_peer_export(npe_handle* h, void* handle64_out) {
    NEED(h);
    if (!handle64_out) return fail(h, NPE_ERR_INVALID, "peer_export: null output");
    static_assert(sizeof(cudaIpcMemHandle_t) == NPE_PEER_HANDLE_BYTES, "IPC handle size");
    if (!h->xregion) {
        CK(cudaMalloc((void**)&h->xregion, DP_REGION_BYTES));
        CK(cudaMemset(h->xregion, 0, DP_REGION_BYTES));
    }
    cudaIpcMemHandle_t hd;
    CK(cudaIpcGetMemHandle(&hd, h->xregion));
    memcpy(handle64_out, &hd, sizeof hd);
    return NPE_OK;
}

int npe_peer_attach(npe_handle* h, int32_t rank, int32_t nranks, const void* all_handles) {
    NEED(h);
    if (!all_handles || nranks < 1 || nranks > MAX_PEERS || rank < 0 || rank >= nranks)
        return fail(h, NPE_ERR_INVALID, "peer_attach: bad arguments (at most %d ranks)", MAX_PEERS);
    if (!h->xregion) return fail(h, NPE_ERR_STATE, "peer_attach: call npe_peer_export first");
    for (int r = 0; r < nranks; ++r) {
        if (r == rank) { h->peer_ptr[r] = h->xregion; continue; }
        cudaIpcMemHandle_t hd, own;
        memcpy(&hd, (const char*)all_handles + (size_t)r * NPE_PEER_HANDLE_BYTES, sizeof hd);
        // a handle that names this rank's own region (single-process tests of the missing-peer path): no IPC open
        CK(cudaIpcGetMemHandle(&own, h->xregion));
        if (memcmp(&hd, &own, sizeof hd) == 0) { h->peer_ptr[r] = h->xregion; continue; }
        CK(cudaIpcOpenMemHandle(&h->peer_ptr[r], hd, cudaIpcMemLazyEnablePeerAccess));
        h->peer_opened[r] = true;
    }
    h->rank = rank; h->nranks = nranks;
    h->peer_ready = true;
    h->peer_step = 0;
    return NPE_OK;
}

int npe_allreduce_grads(npe_handle* h) {
    NEED(h);
    if (!h->comm) return fail(h, NPE_ERR_STATE, "allreduce_grads: no communicator (npe_comm_init)");
    return run_allreduce(h);
}

#define PRIO_TABLE(fn)                                                                                              \
    if (table < 0 || table >= npe_handle::PRIO_TABLES) return fail(h, NPE_ERR_INVALID, fn ": table must be in [0, %d)", npe_handle::PRIO_TABLES); \
    auto& W = h->prio_w[table]; const int NB = h->prio_nb[table];

int npe_priority_init(npe_handle* h, int32_t table, int32_t num_batches, float alpha) {
    NEED(h);
    if (table < 0 || table >= npe_handle::PRIO_TABLES) return fail(h, NPE_ERR_INVALID, "priority_init: table must be in [0, %d)", npe_handle::PRIO_TABLES);
    if (num_batches <= 0 || alpha < 0.f || alpha > 1.f) return fail(h, NPE_ERR_INVALID, "priority_init: bad arguments");
    CK(h->prio_w[table].ensure(num_batches)); CK(h->prio_cum.ensure(num_batches)); CK(h->prio_out2.ensure(2));
    CK(cudaMemsetAsync(h->prio_w[table].p, 0, (size_t)num_batches * sizeof(float), h->stream));
    h->prio_nb[table] = num_batches; h->prio_alpha[table] = alpha;
    return NPE_OK;
}

int npe_priority_update(npe_handle* h, int32_t table, const int32_t* ids, const float* weights, int32_t n) {
    NEED(h);
    PRIO_TABLE("priority_update")
    if (!NB) return fail(h, NPE_ERR_STATE, "priority_update: call npe_priority_init first");
    if (!ids || n <= 0) return fail(h, NPE_ERR_INVALID, "priority_update: bad arguments");
    CK(h->prio_ids.ensure(n));
    CK(cudaMemcpyAsync(h->prio_ids.p, ids, (size_t)n * sizeof(int), cudaMemcpyHostToDevice, h->stream));
    const float* vals; int stride;
    if (weights) {
        CK(h->prio_vals.ensure(n));
        CK(cudaMemcpyAsync(h->prio_vals.p, weights, (size_t)n * sizeof(float), cudaMemcpyHostToDevice, h->stream));
        vals = h->prio_vals.p; stride = 1;
    } else {   // weights == NULL: the losses of the last npe_train_steps call, step i -> ids[i] (no host round trip of the losses)
        if (h->loss_hist.cap < (size_t)n * 4) return fail(h, NPE_ERR_STATE, "priority_update: no per-step losses of %d steps on the device", n);
        vals = h->loss_hist.p; stride = 4;
    }
    k_priority_update<<<(n + 255) / 256, 256, 0, h->stream>>>(W.p, NB, h->prio_ids.p, vals, stride, n);
    CKL();
    CK(cudaStreamSynchronize(h->stream));   // the host arrays may go out of scope
    return NPE_OK;
}

int npe_priority_sample(npe_handle* h, int32_t table, float pow_, uint64_t seed, int32_t n, int32_t* ids_out) {
    NEED(h);
    PRIO_TABLE("priority_sample")
    if (!NB) return fail(h, NPE_ERR_STATE, "priority_sample: call npe_priority_init first");
    if (!ids_out || n <= 0) return fail(h, NPE_ERR_INVALID, "priority_sample: bad arguments");
    CK(h->prio_ids.ensure((size_t)n + 1));
    CK(cudaMemsetAsync(h->prio_ids.p + n, 0, sizeof(int), h->stream));
    CK(h->prio_cum.ensure(NB));
    k_priority_sample<<<1, 1024, 0, h->stream>>>(W.p, NB, pow_ > 0.f ? pow_ : 1.f, h->prio_alpha[table], (unsigned long long)seed, n,
                                                 h->prio_cum.p, h->prio_ids.p, h->prio_ids.p + n);
    CKL();
    std::vector<int> tmp((size_t)n + 1);
    CK(cudaMemcpyAsync(tmp.data(), h->prio_ids.p, tmp.size() * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (tmp[(size_t)n]) return fail(h, NPE_ERR_STATE, "priority_sample: the weight table is not full (a batch has weight <= 0); "
                                                      "the reference asserts here (priority_sampler.lua:24-33)");
    memcpy(ids_out, tmp.data(), (size_t)n * sizeof(int));
    return NPE_OK;
}

int npe_priority_hardest(npe_handle* h, int32_t table, float* weight_out, int32_t* id_out) {
    NEED(h);
    PRIO_TABLE("priority_hardest")
    if (!NB) return fail(h, NPE_ERR_STATE, "priority_hardest: call npe_priority_init first");
    k_priority_hardest<<<1, 1024, 0, h->stream>>>(W.p, NB, h->prio_out2.p);
    CKL();
    float o[2];
    CK(cudaMemcpyAsync(o, h->prio_out2.p, sizeof o, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (weight_out) *weight_out = o[0];
    if (id_out) *id_out = (int32_t)o[1];
    return NPE_OK;
}

int npe_sync(npe_handle* h) {
    NEED(h);
    CK(cudaStreamSynchronize(h->stream));
    if (h->chain_prof_dev) {
        long long q[96];
        cudaMemcpy(q, h->chain_prof_dev, sizeof q, cudaMemcpyDeviceToHost);
        for (int w = 0; w < 2; ++w) {
            const long long* b = q + 64 + 12 * w;
            if (b[0]) fprintf(stderr, "[pair builder, %s CTA, ns after its start] block 1: bits %lld, count stored %lld | block 2: start %lld, bits %lld, count stored %lld | "
                                      "arrives at the barrier %lld, passes it %lld | lists written %lld / %lld | barrier released by CTA %lld at %lld (first CTA's clock)\n",
                              w ? "last" : "first", b[2] - b[0], b[3] - b[0], b[6] - b[0], b[7] - b[0], b[8] - b[0], b[11] - b[0], b[4] - b[0], b[5] - b[0],
                              b[10] - b[0], q[64 + 31], q[64 + 30] - q[64]);
        }
        fprintf(stderr, "[pair forward chain, CTA 0 thread 0] gather %lld, wait for MMA %lld, epilogue %lld, total %lld cycles\n", q[0], q[1], q[2], q[3]);
        for (int part = 0; part < 2; ++part) {
            const long long* w = q + 32 * part;
            fprintf(stderr, "[%s weight gradients, CTA 0] loader group 0: issue loads %lld, wait for the stage %lld, data + split + stores %lld, units %lld, loop %lld | "
                            "mma: wait %lld, issue %lld, loop %lld | epilogue %lld (of which waiting for the last MMAs %lld), kernel body %lld cycles\n",
                    part ? "decoder" : "pair", w[11], w[10], w[12], w[13], w[14], w[15], w[16], w[17], w[18], w[20], w[19]);
        }
    }
    if (chain_prof(h) && (h->graphs_captured || h->graphs_set))
        fprintf(stderr, "[npe_train_steps rank %d] step graphs: %lld captured + instantiated, %lld re-parameterised in place, %lld groups replayed as plain launches; host time per re-parameterised group: %.1f us for its 8 steps + %.1f us for the launch; slot nodes %zu %zu %zu %zu\n",
                h->rank, h->graphs_captured, h->graphs_set, h->graphs_replayed, h->graph_fill_us / (double)std::max(1ll, h->graphs_set), h->graph_launch_us / (double)std::max(1ll, h->graphs_set), h->step_nodes[0].size(), h->step_nodes[1].size(), h->step_nodes[2].size(), h->step_nodes[3].size());
    if (h->dp_prof_dev) {
        // the last two k_dp_step launches (by stamp parity) and, when npe_step_trace is armed, the last k_step_fused: absolute
        // globaltimer values (ns, last 9 digits) so that the two kernels of one step can be put on one time line
        static long long q[8 * DP_NBLK];
        cudaMemcpy(q, h->dp_prof_dev, sizeof q, cudaMemcpyDeviceToHost);
        for (int half = 0; half < 2; ++half) {
            long long lo[4], hi[4];
            for (int k = 0; k < 4; ++k) { lo[k] = 1ll << 62; hi[k] = 0; }
            for (int b = 0; b < DP_NBLK - 1; ++b)
                for (int k = 0; k < 4; ++k) { const long long v = q[half * 4 * DP_NBLK + 4 * b + k]; if (v) { lo[k] = std::min(lo[k], v); hi[k] = std::max(hi[k], v); } }
            fprintf(stderr, "[k_dp_step rank %d, stamp parity %d, absolute ns mod 1e9, first .. last CTA] wait passed %lld..%lld | local sum done %lld..%lld | "
                            "peers' values in %lld..%lld | optimiser done %lld..%lld\n", h->rank, half, lo[0] % 1000000000, hi[0] % 1000000000,
                    lo[1] % 1000000000, hi[1] % 1000000000, lo[2] % 1000000000, hi[2] % 1000000000, lo[3] % 1000000000, hi[3] % 1000000000);
        }
        if (h->step_trace) {
            static unsigned long long tr[64 * 8];
            cudaMemcpy(tr, h->step_trace, sizeof tr, cudaMemcpyDeviceToHost);
            unsigned long long lo[6], hi[6];
            for (int k = 0; k < 6; ++k) { lo[k] = ~0ull; hi[k] = 0; }
            for (int b = 0; b < 64; ++b)
                for (int k = 0; k < 6; ++k) { const unsigned long long v = tr[b * 8 + k]; if (v) { lo[k] = std::min(lo[k], v); hi[k] = std::max(hi[k], v); } }
            fprintf(stderr, "[k_step_fused rank %d, last launch, absolute ns mod 1e9, first .. last CTA] static loads done %llu..%llu | wait passed %llu..%llu | "
                            "hand-over 1 %llu..%llu | hand-over 2 %llu..%llu | gradients done %llu..%llu | partials stored %llu..%llu\n", h->rank,
                    lo[0] % 1000000000, hi[0] % 1000000000, lo[1] % 1000000000, hi[1] % 1000000000, lo[2] % 1000000000, hi[2] % 1000000000,
                    lo[3] % 1000000000, hi[3] % 1000000000, lo[4] % 1000000000, hi[4] % 1000000000, lo[5] % 1000000000, hi[5] % 1000000000);
        }
    }
    if (h->peer_ready) return check_peer_timeout(h);
    return NPE_OK;
}

int npe_set_option(npe_handle* h, int32_t option, int32_t value) {
    NEED(h);
    switch (option) {
        case NPE_OPT_TRAIN_FWD_TC: h->train_fwd_tc = value != 0; break;
        case NPE_OPT_TRAIN_STASH: h->train_stash = value != 0; break;
        case NPE_OPT_BWD_TC: h->bwd_tc = value != 0; break;
        case NPE_OPT_PEER_TIMEOUT_MS:
            if (value < 1) return fail(h, NPE_ERR_INVALID, "set_option: peer timeout must be >= 1 ms");
            h->peer_timeout_ms = value; break;
        case NPE_OPT_SMALL_BATCH_MAX: h->small_batch_max = value; break;
        case NPE_OPT_DEC_STACKED: if (value < 0 || value > 2) return fail(h, NPE_ERR_INVALID, "set_option: dec_stacked is 0, 1 or 2"); h->dec_stacked = value; break;
        default: return fail(h, NPE_ERR_INVALID, "set_option: unknown option %d", option);
    }
    h->fwd_valid = false; h->hact_valid = false; h->dact_valid = false;
    return NPE_OK;
}

void* npe_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
}
void npe_host_free(void* p) { if (p) cudaFreeHost(p); }

int npe_step_trace(npe_handle* h, uint64_t* stamps_out, int32_t max_ctas) {
    NEED(h);
    if (!h->step_trace) {
        CK(cudaMalloc((void**)&h->step_trace, 512 * 8 * sizeof(unsigned long long)));
        CK(cudaMemset(h->step_trace, 0, 512 * 8 * sizeof(unsigned long long)));
    }
    if (stamps_out && max_ctas > 0) {
        CK(cudaStreamSynchronize(h->stream));
        CK(cudaMemcpy(stamps_out, h->step_trace, (size_t)(max_ctas < 512 ? max_ctas : 512) * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    }
    return NPE_OK;
}

int npe_profile_select(npe_handle* h, int32_t kernel_id) {
    NEED(h);
    if (kernel_id < -1 || kernel_id >= NPE_K_COUNT) return fail(h, NPE_ERR_INVALID, "profile_select: bad kernel id");
    CK(cudaStreamSynchronize(h->stream));
    h->prof_kernel = kernel_id;
    h->prof_used = 0;
    h->prof_seen = 0;
    return NPE_OK;
}
int npe_profile_period(npe_handle* h, int32_t period) {
    NEED(h);
    if (period < 1) return fail(h, NPE_ERR_INVALID, "profile_period: must be >= 1");
    h->prof_period = period;
    return NPE_OK;
}
int npe_profile_read(npe_handle* h, int32_t* launches_out, double* total_ms_out) {
    NEED(h);
    CK(cudaStreamSynchronize(h->stream));
    double tot = 0.0;
    for (size_t i = 0; i + 1 < h->prof_used; i += 2) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, h->prof_events[i], h->prof_events[i + 1]));
        tot += ms;
    }
    if (launches_out) *launches_out = (int32_t)(h->prof_used / 2);
    if (total_ms_out) *total_ms_out = tot;
    return NPE_OK;
}

int npe_event_record(npe_handle* h, int32_t slot) {
    NEED(h);
    if (slot < 0 || slot >= 16) return fail(h, NPE_ERR_INVALID, "event slot out of range");
    CK(cudaEventRecord(h->events[slot], h->stream));
    return NPE_OK;
}
int npe_event_elapsed_ms(npe_handle* h, int32_t a, int32_t b, float* ms_out) {
    NEED(h);
    if (a < 0 || a >= 16 || b < 0 || b >= 16 || !ms_out) return fail(h, NPE_ERR_INVALID, "event slot out of range");
    CK(cudaEventSynchronize(h->events[b]));
    CK(cudaEventElapsedTime(ms_out, h->events[a], h->events[b]));
    return NPE_OK;
}

int64_t npe_launch_count(const npe_handle* h) { return h ? h->launches : 0; }

float npe_nbr_threshold_sq(float thr_px) { return nbr_threshold_sq(thr_px); }

int npe_debug_read(npe_handle* h, int32_t which, float* out, int64_t nfloats) {
    NEED(h);
    const float* src = nullptr; size_t cap = 0;
    switch (which) {
        case NPE_DBG_HACT: src = h->hact.p; cap = h->hact.cap; break;
        case NPE_DBG_DZACT: src = h->dzact.p; cap = h->dzact.cap; break;
        case NPE_DBG_DACT: src = h->dact.p; cap = h->dact.cap; break;
        case NPE_DBG_DDZ: src = h->ddz.p; cap = h->ddz.cap; break;
        case NPE_DBG_DS: src = h->ds.p; cap = h->ds.cap; break;
        case NPE_DBG_HP: src = h->Hp.p; cap = h->Hp.cap; break;
        case NPE_DBG_STATES: src = h->states.p; cap = h->states.cap; break;
        case NPE_DBG_ROLL: src = h->roll.p; cap = h->roll.cap; break;
        default: return fail(h, NPE_ERR_INVALID, "debug_read: unknown buffer %d", which);
    }
    if (!out || nfloats <= 0 || !src || (size_t)nfloats > cap) return fail(h, NPE_ERR_INVALID, "debug_read: buffer %d holds %zu floats", which, cap);
    CK(cudaMemcpyAsync(out, src, (size_t)nfloats * sizeof(float), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return NPE_OK;
}

int npe_batch_stats(npe_handle* h, int64_t* foci_out, int64_t* active_pairs_out) {
    NEED(h);
    int rc = build_pairs(h);
    if (rc) return rc;
    u64 tot = 0;
    CK(cudaMemcpyAsync(&tot, h->active_total, sizeof tot, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    int info[2] = {0, 0};
    CK(cudaMemcpy(info, h->info_dev, sizeof info, cudaMemcpyDeviceToHost));
    if (info[1]) return fail(h, NPE_ERR_NOMEM, "active-pair list overflow: %llu pairs > capacity %d", (unsigned long long)tot, h->pair_cap);
    if (foci_out) *foci_out = h->F;
    if (active_pairs_out) *active_pairs_out = (int64_t)tot;
    return NPE_OK;
}

}  // extern "C"
