// sm_100a kernels of the NPE hot path (FP32 FFMA variant).
//
//   k_build_pairs     K1  12-NN context trim + neighbourhood mask per focus (bit-exact integer/bit outputs)
//   k_forward         K2+K3  fused: focus encoder, pair MLP on ACTIVE pairs, per-focus segmented sum, decoder,
//                     sigmoid split, per-example losses
//   k_dec_backward    K4a fused decoder recompute + input-grad chain + weight grads (register-persistent)
//   k_pair_backward   K4b fused pair-MLP recompute + input-grad chain + weight grads (register-persistent)
//   k_reduce_grads    fixed-order cross-CTA reduction of the per-CTA gradient partials (deterministic)
//   k_adam / k_rmsprop  K5 optimiser steps on the flat vector
//   k_rollout         K6  persistent autoregressive rollout, scene state resident in shared memory
//   k_window_targets  K7  relative targets of the selected windows
//
// Reference semantics: src/lua/npe.lua:120-336 (unpack_batch, select_neighbors, apply_mask, fp, fp_batch, bp),
// modules.lua:10-26,46-88, data_process.lua:51-96, eval.lua:122-234,294-384, npe.lua:386-458, infer.lua:9-74.
#pragma once
#include <cstdint>
#include "npe_layout.h"
#include "npe_tile.cuh"

namespace npe {

typedef unsigned long long u64;

// Where the foci of the current batch live: padded scene tensor + focus index arrays.
struct FocusSrc {
    const float* states;   // (S, Nmax, T, 22)
    const int* n_obj;      // (S)
    int Nmax, T;
    const int* foc_scene;  // (F); nullptr (k_build_pairs_scan only): focus f = object f % Nmax of scene f / Nmax at t0 = 0, all scenes full
    const int* foc_obj;    // (F)
    const int* foc_t0;     // (F)
    int F;
    const float2* pos_compact = nullptr;   // dense lists only (rollout): positions of the last frame per focus, contiguous; or nullptr
};

__device__ __forceinline__ size_t row_off(const FocusSrc& s, int scene, int obj, int t) {
    return (((size_t)scene * s.Nmax + obj) * s.T + t) * D;
}

// ----------------------------------------------------------------------------------------------------------------
// weights: the flat (reference layout) vector is mirrored in a PACKED image = the exact shared-memory layout
// [pair image (SW_PAIR_TOTAL) | decoder image (SW_DEC_TOTAL)], maintained by k_pack_weights and written through by
// the optimiser kernels.  Kernels stage it with one TMA bulk copy per image (cp.async.bulk + mbarrier).
// ----------------------------------------------------------------------------------------------------------------
constexpr int WPACK_TOTAL = SW_PAIR_TOTAL + SW_DEC_TOTAL;

__host__ __device__ __forceinline__ int pack_index(int i) {
    if (i < OFF_ENC_C) { const int n = i / IN, k = i - n * IN; return SW_ENC_T + n * LD + k; }
    if (i < OFF_PAIR) { const int r = i - OFF_ENC_C, n = r / IN, k = r - n * IN; return SW_ENC_C + n * LD + k; }
    if (i < OFF_DEC1_W) {
        const int r0 = i - OFF_PAIR, l = r0 / 2500, r = r0 - l * 2500, n = r / H, k = r - n * H;
        return SW_PAIR + l * SW_PAIR_SZ + n * LD + k;
    }
    const int base = SW_PAIR_TOTAL;
    if (i < OFF_DEC1_B) {
        const int r = i - OFF_DEC1_W, n = r / DEC_IN, k = r - n * DEC_IN;
        return base + SW_DEC1 + n * LD1 + (k < H ? k : k + 2);
    }
    if (i < OFF_DEC2_W) return base + SW_DEC1 + (i - OFF_DEC1_B) * LD1 + ONE_COL;
    if (i < OFF_DEC5_W) {
        const int j = i - OFF_DEC2_W, li = j / 2550, r = j - li * 2550;
        if (r < 2500) { const int n = r / H, k = r - n * H; return base + SW_DEC2 + li * SW_PAIR_SZ + n * LD + k; }
        return base + SW_DEC2 + li * SW_PAIR_SZ + (r - 2500) * LD + ONE_COL;
    }
    if (i < OFF_DEC5_B) { const int r = i - OFF_DEC5_W, n = r / H, k = r - n * H; return base + SW_DEC5 + n * LD + k; }
    return base + SW_DEC5 + (i - OFF_DEC5_B) * LD + ONE_COL;
}

// Rebuilds the whole packed image: zeros, the unit entries that regenerate the constant-1 column, then the scatter.
__global__ void k_pack_zero(float* __restrict__ wpack) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= WPACK_TOTAL) return;
    float v = 0.f;
    const int d = i - SW_PAIR_TOTAL;
    if (d == SW_DEC1 + ONE_COL * LD1 + ONE_COL) v = 1.f;
    for (int li = 0; li < 3; ++li)
        if (d == SW_DEC2 + li * SW_PAIR_SZ + ONE_COL * LD + ONE_COL) v = 1.f;
    wpack[i] = v;
}
__global__ void k_pack_scatter(const float* __restrict__ params, float* __restrict__ wpack) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < NPARAM) wpack[pack_index(i)] = params[i];
}

// Programmatic dependent launch (PDL): the forward/backward kernels of one step are launched with the
// programmatic-stream-serialization attribute, so a kernel's prologue (mbarrier init, TMA weight staging, shared-memory
// clears -- nothing that depends on the previous kernel) overlaps the tail of its predecessor.  pdl_wait() must be
// executed by EVERY CTA before it touches predecessor outputs or exits (grid completion is what successors wait on).
// Without the launch attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

struct WeightStage {
    unsigned long long bar;
};
// Thread 0 arms the barrier and issues the bulk copies; everyone waits with stage_wait() before the first use.
__device__ __forceinline__ void stage_init(WeightStage& st) {
    if (threadIdx.x == 0) mbar_init(&st.bar, 1);
}
// after stage_init + a CTA barrier
__device__ __forceinline__ void stage_copy(WeightStage& st, float* sWp, float* sWd, const float* __restrict__ wpack) {
    if (threadIdx.x == 0) {
        unsigned bytes = 0;
        if (sWp) bytes += SW_PAIR_TOTAL * 4;
        if (sWd) bytes += SW_DEC_TOTAL * 4;
        mbar_expect_tx(&st.bar, bytes);
        if (sWp) bulk_g2s(sWp, wpack, SW_PAIR_TOTAL * 4, &st.bar);
        if (sWd) bulk_g2s(sWd, wpack + SW_PAIR_TOTAL, SW_DEC_TOTAL * 4, &st.bar);
    }
}
__device__ __forceinline__ void stage_issue(WeightStage& st, float* sWp, float* sWd, const float* __restrict__ wpack) {
    if (threadIdx.x == 0) mbar_init(&st.bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned bytes = 0;
        if (sWp) bytes += SW_PAIR_TOTAL * 4;
        if (sWd) bytes += SW_DEC_TOTAL * 4;
        mbar_expect_tx(&st.bar, bytes);
        if (sWp) bulk_g2s(sWp, wpack, SW_PAIR_TOTAL * 4, &st.bar);
        if (sWd) bulk_g2s(sWd, wpack + SW_PAIR_TOTAL, SW_DEC_TOTAL * 4, &st.bar);
    }
}
__device__ __forceinline__ void stage_wait(WeightStage& st) { mbar_wait(&st.bar, 0); }

// ----------------------------------------------------------------------------------------------------------------
// K1: pair table.  One warp per focus.  Context index c in [0, N-1) enumerates the other objects in original order
// (object j = c + (c >= focus)).  keep = contexts surviving the k-nearest trim (data_process.lua:51-70; ties: smaller
// distance, then lower index); nbr = kept contexts inside the neighbourhood (npe.lua:170-206).
// Distance arithmetic follows compute_euc_dist (data_utils.lua:276-285) with one rounding per op: no FMA contraction.
// ----------------------------------------------------------------------------------------------------------------
// The neighbourhood test of the reference is  fl(800 * fl(sqrt(s))) <= thr  on the squared distance s.  Both roundings
// are monotone, so the set of s that pass is  s <= thr_s  for ONE float thr_s (the largest s that passes; the host finds
// it by bisection over the float bit patterns, npe_api.cu nbr_threshold_sq): the same bits without the IEEE square root,
// which is only taken where contexts have to be RANKED by distance (distinct s can round to equal distances, and ties
// go to the lower index).
__device__ __forceinline__ float pair_sqdist(float x, float y, float fx, float fy) {
    const float dx = __fsub_rn(x, fx), dy = __fsub_rn(y, fy);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}
// A set of OBJECTS of a scene (bit o = object o) -> the same set as CONTEXTS of focus n (context c = object c + (c >= n)).
__device__ __forceinline__ u64 objects_to_contexts(u64 m, int n) {
    return (m & ((1ull << n) - 1ull)) | (((m >> n) >> 1) << n);
}
// One focus per warp, in OBJECT space: lane l holds the squared distances focus -> object l (s[0]) and -> object l + 32
// (s[1]).  Object order is context order, so ranks and the lower-index tie rule carry over unchanged; the two sets are
// renumbered to context indices at the end.
__device__ __forceinline__ void warp_pairs_core(const float (&s)[2], int N, int n, int kmax, float thr_s, u64& keep, u64& nbr,
                                                bool want_keep) {
    const int lane = threadIdx.x & 31;
    const int Cn = N - 1;
    const u64 V = (N >= 64 ? ~0ull : ((1ull << N) - 1ull)) & ~(1ull << n);     // the other objects of the scene
    bool in[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) in[h] = ((V >> (lane + 32 * h)) & 1ull) && s[h] <= thr_s;
    const u64 W = (u64)__ballot_sync(0xffffffffu, in[0]) | ((u64)__ballot_sync(0xffffffffu, in[1]) << 32);
    const int k = (kmax > 0 && kmax < Cn) ? kmax : Cn;
    if (k >= Cn) { keep = objects_to_contexts(V, n); nbr = objects_to_contexts(W, n); return; }
    // Trim to the k nearest (ties: lower index first).  The neighbourhood W is a PREFIX of that order, so ranks never
    // have to be taken over all Cn contexts (O(Cn^2) shuffles per focus):
    //   |W| >= k : the k nearest all lie in W; rank the members of W among W             (|W| shuffles)
    //   |W| <  k : every member of W is kept and nbr = W; the kept set also holds the k - |W| nearest OUTSIDE W, which
    //              only the API view of the pair table shows (want_keep): rank the outsiders among the outsiders
    const int w = __popcll(W);
    const bool inside = w >= k;
    if (!inside && !want_keep) { keep = objects_to_contexts(W, n); nbr = keep; return; }
    float d[2];
    bool valid[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        valid[h] = (V >> (lane + 32 * h)) & 1ull;
        d[h] = valid[h] ? __fsqrt_rn(s[h]) : __int_as_float(0x7f800000);
    }
    u64 bits = inside ? W : (V & ~W);
    int rank[2] = {0, 0};
    while (bits) {
        const int op = __ffsll((long long)bits) - 1;
        bits &= bits - 1;
        const float dp = __shfl_sync(0xffffffffu, op < 32 ? d[0] : d[1], op & 31);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int o = lane + 32 * h;
            rank[h] += (dp < d[h]) || (dp == d[h] && op < o);
        }
    }
    bool kp[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) kp[h] = inside ? (in[h] && rank[h] < k) : (in[h] || (valid[h] && rank[h] < k - w));
    const u64 K = (u64)__ballot_sync(0xffffffffu, kp[0]) | ((u64)__ballot_sync(0xffffffffu, kp[1]) << 32);
    keep = objects_to_contexts(K, n);
    nbr = inside ? keep : objects_to_contexts(W, n);
}
// object j at pos + j * stride (global or shared memory)
__device__ __forceinline__ void warp_pairs(const float* pos, int stride, int N, int n, int kmax, float thr_s,
                                           u64& keep, u64& nbr, bool want_keep = true) {
    const int lane = threadIdx.x & 31;
    const float fx = pos[(size_t)n * stride], fy = pos[(size_t)n * stride + 1];
    float s[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int o = lane + 32 * h;
        s[h] = __int_as_float(0x7f800000);
        if (o < N) s[h] = pair_sqdist(pos[(size_t)o * stride], pos[(size_t)o * stride + 1], fx, fy);
    }
    warp_pairs_core(s, N, n, kmax, thr_s, keep, nbr, want_keep);
}
// Sub-warp variant of warp_pairs for small scenes: a group of LANES (8 or 16) lanes handles one focus with at most
// LANES contexts, so a warp builds 32 / LANES pair tables at once (8-ball scenes: 4 foci per warp instead of 1 with 7
// of 32 lanes busy).  Same arithmetic, same tie rule, bit-identical results.
template <int LANES>
__device__ __forceinline__ float group_sqdist(const float* pos, int stride, int Cn, int n) {
    const int gl = threadIdx.x & (LANES - 1);
    float s = __int_as_float(0x7f800000);
    if (gl < Cn) {
        const int j = gl + (gl >= n);
        s = pair_sqdist(pos[(size_t)j * stride], pos[(size_t)j * stride + 1], pos[(size_t)n * stride], pos[(size_t)n * stride + 1]);
    }
    return s;
}
template <int LANES>
__device__ __forceinline__ void group_core(float s, int Cn, int kmax, float thr_s, u64& keep, u64& nbr) {
    const int lane = threadIdx.x & 31, gl = lane & (LANES - 1), gbase = lane & ~(LANES - 1);
    const bool valid = gl < Cn;
    const int k = (kmax > 0 && kmax < Cn) ? kmax : Cn;
    int rank = 0;
    if (__any_sync(0xffffffffu, k < Cn)) {                // no group of this warp trims: no ranks, no square root
        const float d = valid ? __fsqrt_rn(s) : __int_as_float(0x7f800000);
#pragma unroll
        for (int cp = 0; cp < LANES; ++cp) {
            const float dp = __shfl_sync(0xffffffffu, d, gbase + cp);
            if (cp < Cn) rank += (dp < d) || (dp == d && cp < gl);
        }
    }
    const bool kp = valid && (k >= Cn || rank < k);
    const unsigned bk = __ballot_sync(0xffffffffu, kp);
    const unsigned bm = __ballot_sync(0xffffffffu, kp && s <= thr_s);
    const unsigned gm = LANES == 32 ? 0xffffffffu : ((1u << LANES) - 1u);
    keep = (u64)((bk >> gbase) & gm);
    nbr = (u64)((bm >> gbase) & gm);
}
template <int LANES>
__device__ __forceinline__ void group_pairs(const float* pos, int stride, int N, int n, int kmax, float thr_s, bool active,
                                            u64& keep, u64& nbr) {
    const int Cn = active ? N - 1 : 0;
    group_core<LANES>(group_sqdist<LANES>(pos, stride, Cn, n), Cn, kmax, thr_s, keep, nbr);
}

// K1 (+K7): pair bits, focus row offset, the relative target y = frame(t0+2) with columns 0..5 minus frame(t0+1)
// (relative_pair, data_process.lua:87-96), and the first level of the active-pair scan (exclusive prefix of the counts
// inside the CTA + the CTA total).  64 foci per CTA.  LANES = 32: one warp per focus (up to 63 contexts), two foci per
// warp in turn, 1024 threads; LANES = 16 / 8: 2 / 4 foci per warp at once (scenes with at most 17 / 9 objects).
constexpr int FPB = 64;   // foci per CTA of the pair builder
constexpr int NT_PAIRS = 1024;
template <int LANES>
__device__ __forceinline__ void build_pairs_body(const FocusSrc& src, int kmax, float thr_s, u64* __restrict__ keep_out,
                                                 u64* __restrict__ nbr_out, int* __restrict__ local_prefix,
                                                 int* __restrict__ block_sums, long long* __restrict__ foc_row,
                                                 float* __restrict__ y, u64* __restrict__ active_total,
                                                 int* __restrict__ pair_start, int* __restrict__ pair_foc,
                                                 long long* __restrict__ pair_crow, int cap, int* __restrict__ info,
                                                 const int blk, const bool single, const bool want_keep) {
    __shared__ int s_cnt[FPB];
    __shared__ int s_pre[FPB];
    __shared__ u64 s_nbr[FPB];
    // Position cache: the foci of a CTA come in runs that share (scene, t0) -- all objects of a window are consecutive
    // foci -- and every focus of a run needs the positions of ALL objects of that scene at frame t0 + 1.  Object rows are
    // T * 22 floats apart, so each warp-wide position load touches 32 different lines; loading a run's positions ONCE
    // into shared memory (instead of once per focus) removes the pair builder's dominant cost on large scenes
    // (64-ball scenes: 64 foci per run).  Same values, same arithmetic: bit-identical tables.
    __shared__ float2 s_pos[FPB][MAXOBJ];      // [run][object]  (32 KB)
    __shared__ int s_run[FPB];                 // run index of every focus of this CTA
    __shared__ int s_head[FPB];                // focus index that opens run r
    __shared__ int s_nruns;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int GPW = 32 / LANES;                       // foci per warp per pass
    const int nwarps = blockDim.x >> 5;
    const int gl = lane & (LANES - 1), grp = lane / LANES;
    if (warp < 2) {                                       // warps 0, 1: foci 0..31, 32..63
        const int fl = threadIdx.x, f = blk * FPB + fl;
        const bool act = f < src.F;
        int sc = -1, t0 = -1;
        if (act) { sc = src.foc_scene[f]; t0 = src.foc_t0[f]; }
        const int psc = __shfl_up_sync(0xffffffffu, sc, 1), pt0 = __shfl_up_sync(0xffffffffu, t0, 1);
        bool head = act && (lane == 0 || sc != psc || t0 != pt0);
        if (act && lane == 0 && warp == 1) {              // run may continue across the warp boundary
            const int f1 = f - 1;
            head = src.foc_scene[f1] != sc || src.foc_t0[f1] != t0;
        }
        const unsigned hb = __ballot_sync(0xffffffffu, head);
        if (lane == 0) s_cnt[warp] = __popc(hb);          // s_cnt is rewritten below, after the barrier
        __syncwarp();
        const int before = __popc(hb & ((2u << lane) - 1u)) - 1;   // runs opened up to and including this focus, minus one
        s_pre[fl] = before;                               // per-warp run index (warp 1 is shifted below)
        s_nbr[fl] = head ? 1ull : 0ull;
    }
    __syncthreads();
    if (threadIdx.x < FPB) {
        const int fl = threadIdx.x;
        const int shift = fl >= 32 ? s_cnt[0] : 0;
        const int r = s_pre[fl] + shift;
        s_run[fl] = r;
        if (s_nbr[fl]) s_head[r] = fl;
        if (fl == 0) s_nruns = s_cnt[0] + s_cnt[1];
    }
    __syncthreads();
    for (int r = warp; r < s_nruns; r += nwarps) {        // one warp loads the positions of one run
        const int f = blk * FPB + s_head[r];
        const int sc = src.foc_scene[f], t0 = src.foc_t0[f], N = src.n_obj[sc];
        const float* base = src.states + row_off(src, sc, 0, t0 + 1);
        for (int j = lane; j < N; j += 32) s_pos[r][j] = *reinterpret_cast<const float2*>(base + (size_t)j * src.T * D);
    }
    __syncthreads();
    for (int base = 0; base < FPB; base += nwarps * GPW) {
        const int fl = base + warp * GPW + grp;
        const int f = blk * FPB + fl;
        const bool active = fl < FPB && f < src.F;
        int c = 0;
        u64 keep = 0, nbr = 0;
        size_t ro = 0;
        const float* pos = reinterpret_cast<const float*>(s_pos[active ? s_run[fl] : 0]);
        if (LANES == 32) {
            if (active) {
                const int s = src.foc_scene[f], n = src.foc_obj[f], t0 = src.foc_t0[f];
                warp_pairs(pos, 2, src.n_obj[s], n, kmax, thr_s, keep, nbr, want_keep);
                ro = row_off(src, s, n, t0);
            }
        } else {
            int s = 0, n = 0, t0 = 0, N = 1;
            if (active) { s = src.foc_scene[f]; n = src.foc_obj[f]; t0 = src.foc_t0[f]; N = src.n_obj[s]; ro = row_off(src, s, n, t0); }
            group_pairs<LANES>(pos, 2, N, n, kmax, thr_s, active, keep, nbr);
        }
        if (active) {
            if (y) {
                const float* p = src.states + ro;
                for (int col = gl; col < D; col += LANES) {
                    float v = p[2 * D + col];
                    if (col < 6) v = __fsub_rn(v, p[D + col]);
                    y[(size_t)f * D + col] = v;
                }
            }
            c = __popcll(nbr);
            if (gl == 0) { keep_out[f] = keep; nbr_out[f] = nbr; foc_row[f] = (long long)ro; s_nbr[fl] = nbr; }
        }
        if (gl == 0 && fl < FPB) s_cnt[fl] = c;
    }
    __syncthreads();
    if (warp == 0) {
        const int c0 = s_cnt[lane], c1 = s_cnt[lane + 32];
        int i0 = c0, i1 = c1;
        for (int off = 1; off < 32; off <<= 1) {
            const int v0 = __shfl_up_sync(0xffffffffu, i0, off), v1 = __shfl_up_sync(0xffffffffu, i1, off);
            if (lane >= off) { i0 += v0; i1 += v1; }
        }
        const int tot0 = __shfl_sync(0xffffffffu, i0, 31), tot1 = __shfl_sync(0xffffffffu, i1, 31);
        const int f = blk * FPB + lane;
        if (f < src.F) local_prefix[f] = i0 - c0;
        if (f + 32 < src.F) local_prefix[f + 32] = tot0 + i1 - c1;
        s_pre[lane] = i0 - c0; s_pre[lane + 32] = tot0 + i1 - c1;
        if (lane == 0) {
            block_sums[blk] = tot0 + tot1;
            if (single) *active_total = (u64)(tot0 + tot1); else atomicAdd(active_total, (u64)(tot0 + tot1));
        }
    }
    if (!single) return;
    // single-CTA batch (F <= 64, e.g. the reference's batch of 50): build the dense pair list here and save a launch
    __syncthreads();
    if (threadIdx.x < FPB && (int)threadIdx.x < src.F) {
        const int f = threadIdx.x;
        int base = s_pre[f];
        pair_start[f] = min(base, cap);
        u64 bits = s_nbr[f];
        const int fo = src.foc_obj[f];
        const long long sbase = (long long)row_off(src, src.foc_scene[f], 0, src.foc_t0[f]);
        const int objstride = src.T * D;
        while (bits) {
            const int c = __ffsll((long long)bits) - 1;
            bits &= bits - 1;
            if (base < cap) { pair_foc[base] = f; pair_crow[base] = sbase + (long long)(c + (c >= fo)) * objstride; }
            ++base;
        }
        if (f == src.F - 1) { pair_start[src.F] = min(base, cap); info[0] = min(base, cap); info[1] = base > cap; }
    }
}

template <int LANES>
__global__ void __launch_bounds__(NT_PAIRS) k_build_pairs(FocusSrc src, int kmax, float thr_s, u64* __restrict__ keep_out,
                                                          u64* __restrict__ nbr_out, int* __restrict__ local_prefix,
                                                          int* __restrict__ block_sums, long long* __restrict__ foc_row,
                                                          float* __restrict__ y, u64* __restrict__ active_total,
                                                          int* __restrict__ pair_start, int* __restrict__ pair_foc,
                                                          long long* __restrict__ pair_crow, int cap, int* __restrict__ info, int want_keep) {
    build_pairs_body<LANES>(src, kmax, thr_s, keep_out, nbr_out, local_prefix, block_sums, foc_row, y, active_total, pair_start,
                            pair_foc, pair_crow, cap, info, (int)blockIdx.x, gridDim.x == 1, want_keep != 0);
}

// The batcher for a whole SCHEDULE of small batches (train(): main.lua:299-304 over a device-resident schedule): CTA s
// builds the complete pair table of step s (at most 64 foci) into slice s of the schedule buffers, so the per-step
// launch of the pair builder disappears from the training loop.  fmax = slice stride in foci, cap = in pairs.
struct SchedOut {
    u64* keep; u64* nbr; int* cnt; int* block_sums; long long* foc_row; float* y; u64* total;
    int* pair_start; int* pair_foc; long long* pair_crow; int* info;
};
template <int LANES>
__global__ void __launch_bounds__(NT_PAIRS) k_build_pairs_sched(FocusSrc src, const int* __restrict__ step_first,
                                                                const int* __restrict__ step_F, int fmax, int cap, int kmax,
                                                                float thr_s, SchedOut o) {
    const int s = blockIdx.x;
    const int first = step_first[s];
    src.foc_scene += first; src.foc_obj += first; src.foc_t0 += first; src.F = step_F[s];
    const size_t fo = (size_t)s * fmax, po = (size_t)s * cap;
    build_pairs_body<LANES>(src, kmax, thr_s, o.keep + fo, o.nbr + fo, o.cnt + fo, o.block_sums + s, o.foc_row + fo, o.y + fo * D,
                            o.total + s, o.pair_start + (size_t)s * (fmax + 1), o.pair_foc + po, o.pair_crow + po, cap,
                            o.info + 2 * s, 0, true, true);
}

// K1 for more than one CTA of foci -- the throughput regime (S-64: 65 536 foci, configs[4] rollout: 524 288 per step).
// One launch builds the pair bits AND the dense pair list: for pair p of focus f (contexts in ascending index order)
// pair_foc[p] = f, pair_crow[p] = float offset of the context's window; pairs beyond `cap` are dropped and flagged.
// History, from ncu: the first generation (a warp per focus, 1024-thread CTAs, second kernel for the list) ran its phases
// one after the other with most warps parked at a barrier, 3.5 waves of 8 us; a warp per focus costs ~130 warp
// instructions per focus whatever the helper (shuffles, ballots, 64-bit mask arithmetic for 2 pair tests per lane), and a
// 32-wide look-back walked the predecessors of the last of 1024 blocks in 30 serial L2 round trips.  Now:
//   * a THREAD per focus: it walks the objects of its scene (positions: broadcast / L1-resident loads, the loop is
//     uniform for the foci of a window) and keeps the neighbourhood as a 64-bit object set; ~10 instructions per pair
//     test for 32 foci at once.  The k-nearest trim only runs where more than k objects are inside the neighbourhood (or
//     for the API view of the kept set), thread-locally.
//   * blocks of 256 foci; the relative targets are 14 loads per thread issued before the pair loop
//   * two phases around ONE grid barrier (the launch is cooperative): every block writes its pair count, then every block
//     sums the counts of its predecessors (a coalesced read of at most a few thousand words) and writes its part of the list
// Same arithmetic, order and tie rule as the first generation (bit-identical tables).
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
constexpr int NT_SCAN = 256;               // threads = foci per block
struct PairScan {
    unsigned* count;      // [blocks]: active pairs of every block of 256 foci (phase 1 -> phase 2)
    unsigned* bar;        // grid barrier: [0] top counter, [1] release word (= epoch of the last launch), [2 ..] group counters; zero / stale between launches
    unsigned epoch;       // differs from the previous launch on the same barrier words
    int nblk;             // blocks of NT_SCAN foci
    long long* prof;      // tools only (NPE_CHAIN_PROF): phase time stamps of CTA 0 and of the last CTA
};
struct ScanArgs {          // one parameter block: the kernel is launched with cudaLaunchCooperativeKernel
    FocusSrc src; int kmax; float thr_s; int want_keep;
    u64* keep; u64* nbr; long long* foc_row; float* y; u64* active_total; int* pair_start; int* pair_foc; long long* pair_crow;
    int cap; int* info; PairScan sc;
};
// The `need` nearest members of the object set `cand` (ties: lower index), thread-local: O(|cand|^2) distance evaluations
// from L1.  Distances are the reference's: sqrt of the squared distance, one rounding per operation.
__device__ __noinline__ u64 nearest_of(const float* pos, int stride, float fx, float fy, u64 cand, int need) {
    u64 out = 0;
    for (u64 a = cand; a; a &= a - 1) {
        const int o = __ffsll((long long)a) - 1;
        const float d = __fsqrt_rn(pair_sqdist(pos[(size_t)o * stride], pos[(size_t)o * stride + 1], fx, fy));
        int rank = 0;
        for (u64 b = cand; b; b &= b - 1) {
            const int p = __ffsll((long long)b) - 1;
            const float dp = __fsqrt_rn(pair_sqdist(pos[(size_t)p * stride], pos[(size_t)p * stride + 1], fx, fy));
            rank += (dp < d) || (dp == d && p < o);
        }
        if (rank < need) out |= 1ull << o;
    }
    return out;
}
template <bool HAS_Y>     // training batches also get their relative targets (28 more registers: 4 CTAs per SM instead of 8)
__global__ void __launch_bounds__(NT_SCAN, HAS_Y ? 4 : 8) k_build_pairs_scan(const ScanArgs A) {
    const FocusSrc& src = A.src;
    const int kmax = A.kmax, want_keep = A.want_keep, cap = A.cap;
    const float thr_s = A.thr_s;
    u64* __restrict__ keep_out = A.keep; u64* __restrict__ nbr_out = A.nbr;
    long long* __restrict__ foc_row = A.foc_row; float* __restrict__ y = A.y;
    u64* __restrict__ active_total = A.active_total; int* __restrict__ pair_start = A.pair_start;
    int* __restrict__ pair_foc = A.pair_foc; long long* __restrict__ pair_crow = A.pair_crow; int* __restrict__ info = A.info;
    const PairScan& sc = A.sc;
    __shared__ int s_wtot[NT_SCAN / 32];
    __shared__ unsigned s_part[NT_SCAN / 32];
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int stride = src.T * D;
    const bool stamp = sc.prof && t == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1);
    long long* pf = sc.prof + (blockIdx.x == 0 ? 0 : 12);
    if (stamp) pf[0] = (long long)global_ns();
    // ================= phase 1: pair bits, tables, targets and the pair COUNT of every block of this CTA =================
    // Static assignment: CTA b takes blocks b, b + G, b + 2G, ...; the launch is COOPERATIVE (all G CTAs resident), which is
    // what makes the grid barrier below legal.  History: a chained scan with look-back (status words polled by the
    // successors) spent most of the kernel waiting -- 300 k threads polling 128 cache lines delay the very stores they wait
    // for, whichever way the polling was organised (32 / 256 / 1024 predecessors per round, CTA-wide or one warp); handing the
    // blocks out through an atomic ticket cost 15 us for 2 x 1184 same-address atomics.
    int it = 0;
    for (int blk = blockIdx.x; blk < sc.nblk; blk += gridDim.x, ++it) {
        if (stamp && it < 2) pf[1 + 5 * it] = (long long)global_ns();
        // ---- this thread's focus ----
        const int f = blk * NT_SCAN + t;
        const bool active = f < src.F;
        int scn = 0, n = 0, t0 = 0, N = 1;
        long long ro = 0;
        if (active) {
            if (src.foc_scene) { scn = src.foc_scene[f]; n = src.foc_obj[f]; t0 = src.foc_t0[f]; N = src.n_obj[scn]; }
            else { scn = f / src.Nmax; n = f - scn * src.Nmax; t0 = 0; N = src.Nmax; }   // dense: every object of every (full) scene, in order
            ro = (long long)row_off(src, scn, n, t0);
            foc_row[f] = ro;
        }
        // relative target: frame(t0 + 2), columns 0..5 minus frame(t0 + 1) (relative_pair, data_process.lua:87-96).  All 14
        // loads of the row are issued here and consumed after the pair test (a loop over the columns waited for one
        // HBM round trip per iteration: 40 % of the kernel's stall samples)
        float2 ya[HAS_Y ? 11 : 1], yb[HAS_Y ? 3 : 1];
        if (HAS_Y && active) {
            const float* p = src.states + ro;
#pragma unroll
            for (int k2 = 0; k2 < 11; ++k2) ya[k2] = ld2(p + 2 * D + 2 * k2);
#pragma unroll
            for (int k2 = 0; k2 < 3; ++k2) yb[k2] = ld2(p + D + 2 * k2);
        }
        // ---- neighbourhood as an object set: bit o = object o of the scene is inside ----
        const float* pos = src.states + row_off(src, scn, 0, t0 + 1);
        unsigned m_lo = 0, m_hi = 0;
        float fx = 0.f, fy = 0.f;
        // dense focus lists over scenes of 2^k <= 32 objects (the configs[4] rollout): the objects of a scene are adjacent lanes,
        // so a thread loads ITS position only and takes the others by shuffle -- one HBM round trip instead of three
        const bool by_shuffle = !src.foc_scene && (src.Nmax & (src.Nmax - 1)) == 0 && src.Nmax <= 32;
        if (by_shuffle) {
            // the fused rollout keeps the positions of the last frame as a contiguous float2 array (written by the decoder kernel):
            // a coalesced 8-byte load instead of one 32-byte sector per object out of the 176-byte window rows
            if (active) {
                const float2 fp = src.pos_compact ? src.pos_compact[f] : *reinterpret_cast<const float2*>(pos + (size_t)n * stride);
                fx = fp.x; fy = fp.y;
            }
            const int base = lane & ~(src.Nmax - 1);
            for (int o = 0; o < src.Nmax; ++o) {
                const float qx = __shfl_sync(0xffffffffu, fx, base + o), qy = __shfl_sync(0xffffffffu, fy, base + o);
                m_lo |= (unsigned)(pair_sqdist(qx, qy, fx, fy) <= thr_s) << o;
            }
        } else if (active) {
            const float2 fp = *reinterpret_cast<const float2*>(pos + (size_t)n * stride);
            fx = fp.x; fy = fp.y;
            const int n_lo = N < 32 ? N : 32;
#pragma unroll 4
            for (int o = 0; o < n_lo; ++o) {
                const float2 q = *reinterpret_cast<const float2*>(pos + (size_t)o * stride);
                m_lo |= (unsigned)(pair_sqdist(q.x, q.y, fx, fy) <= thr_s) << o;
            }
#pragma unroll 4
            for (int o = 32; o < N; ++o) {
                const float2 q = *reinterpret_cast<const float2*>(pos + (size_t)o * stride);
                m_hi |= (unsigned)(pair_sqdist(q.x, q.y, fx, fy) <= thr_s) << (o - 32);
            }
        }
        const int Cn = N - 1;
        const u64 V = active ? ((N >= 64 ? ~0ull : ((1ull << N) - 1ull)) & ~(1ull << n)) : 0ull;   // the other objects of the scene
        const u64 W = (((u64)m_hi << 32) | m_lo) & V;
        const int k = (kmax > 0 && kmax < Cn) ? kmax : Cn;
        u64 K = V, Nb = W;                                 // kept set, neighbours (object space)
        if (k < Cn) {
            // Trim to the k nearest (ties: lower index).  The neighbourhood is a PREFIX of that order, so:
            //   |W| >= k : the k nearest all lie in W                       -> rank W among itself (only if |W| > k)
            //   |W| <  k : all of W is kept and nbr = W; the kept set also holds the k - |W| nearest outsiders, which
            //              only the API view of the pair table shows (want_keep)
            const int w = __popcll(W);
            if (w > k) { K = nearest_of(pos, stride, fx, fy, W, k); Nb = K; }
            else if (w == k || !want_keep) K = W;
            else K = W | nearest_of(pos, stride, fx, fy, V & ~W, k - w);
        }
        const u64 keep = objects_to_contexts(K, n), nbr = objects_to_contexts(Nb, n);
        const int c = __popcll(nbr);
        if (active) { keep_out[f] = keep; nbr_out[f] = nbr; }
        if (HAS_Y && active) {
            float* yo = y + (size_t)f * D;
#pragma unroll
            for (int k2 = 0; k2 < 11; ++k2) {
                float2 v = ya[k2];
                if (k2 < 3) { v.x = __fsub_rn(v.x, yb[k2].x); v.y = __fsub_rn(v.y, yb[k2].y); }
                *reinterpret_cast<float2*>(yo + 2 * k2) = v;
            }
        }
        if (stamp && it < 2) pf[2 + 5 * it] = (long long)global_ns();
        // ---- the block's pair count ----
        int tot = c;
        for (int off = 16; off > 0; off >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, off);
        if (lane == 0) s_wtot[warp] = tot;
        __syncthreads();
        if (t == 0) {
            int agg = 0;
#pragma unroll
            for (int w8 = 0; w8 < NT_SCAN / 32; ++w8) agg += s_wtot[w8];
            sc.count[blk] = (unsigned)agg;
        }
        __syncthreads();                                  // s_wtot is rewritten by the next block
        if (stamp && it < 2) pf[3 + 5 * it] = (long long)global_ns();
    }
    // ================= grid barrier: every block count of the launch is in memory =================
    // Hierarchical arrival (groups of 32 CTAs on their own counters, the last of a group arrives at the top counter): 1184
    // same-address atomics cost ~7 us on this part, 32 + 37 of them well under one.  The last arrivers reset the counters; one
    // thread per CTA polls the release word (= launch epoch), with a back-off.
    if (t == 0) {
        __threadfence();
        if (stamp) pf[11] = (long long)global_ns();
        const int G = (int)gridDim.x, g = (int)blockIdx.x >> 5, gsize = min(32, G - 32 * g), ngroups = (G + 31) >> 5;
        volatile unsigned* rel = sc.bar + 1;
        if (atomicAdd(sc.bar + 2 + g, 1u) == (unsigned)gsize - 1u) {
            sc.bar[2 + g] = 0u;
            if (atomicAdd(sc.bar, 1u) == (unsigned)ngroups - 1u) {
                sc.bar[0] = 0u;
                __threadfence();
                *rel = sc.epoch;
                if (sc.prof) { sc.prof[30] = (long long)global_ns(); sc.prof[31] = blockIdx.x; }
            }
        }
        int spins = 0;
        while (*rel != sc.epoch) {
            if (++spins > (1 << 20)) __trap();            // cannot happen: cooperative launch, every CTA is resident
            __nanosleep(200);
        }
        __threadfence();
    }
    __syncthreads();
    if (stamp) pf[4] = (long long)global_ns();
    // ================= phase 2: exclusive prefix of the block counts, then every focus writes its part of the pair list =========
    it = 0;
    for (int blk = blockIdx.x; blk < sc.nblk; blk += gridDim.x, ++it) {
        unsigned part = 0;
        for (int j = t; j < blk; j += NT_SCAN) part += __ldcg(sc.count + j);
        for (int off = 16; off > 0; off >>= 1) part += __shfl_xor_sync(0xffffffffu, part, off);
        if (lane == 0) s_part[warp] = part;
        // this thread's focus again: its neighbour bits come back from the table it wrote in phase 1
        const int f = blk * NT_SCAN + t;
        const bool active = f < src.F;
        int scn = 0, n = 0, t0 = 0;
        u64 nbr = 0;
        if (active) {
            if (src.foc_scene) { scn = src.foc_scene[f]; n = src.foc_obj[f]; t0 = src.foc_t0[f]; }
            else { scn = f / src.Nmax; n = f - scn * src.Nmax; }
            nbr = nbr_out[f];
        }
        const int c = __popcll(nbr);
        int incl = c;
        for (int off = 1; off < 32; off <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += v;
        }
        if (lane == 31) s_wtot[warp] = incl;
        __syncthreads();
        unsigned excl = 0;
        int wbase = 0;
#pragma unroll
        for (int w8 = 0; w8 < NT_SCAN / 32; ++w8) { excl += s_part[w8]; if (w8 < warp) wbase += s_wtot[w8]; }
        if (active) {
            int base = (int)excl + wbase + incl - c;
            pair_start[f] = min(base, cap);
            const long long sbase = (long long)row_off(src, scn, 0, t0);
            u64 bits = nbr;
            while (bits) {
                const int cc = __ffsll((long long)bits) - 1;
                bits &= bits - 1;
                if (base < cap) {
                    pair_foc[base] = f;
                    pair_crow[base] = sbase + (long long)(cc + (cc >= n)) * stride;
                }
                ++base;
            }
            if (f == src.F - 1) {
                pair_start[src.F] = min(base, cap);
                info[0] = min(base, cap);      // P
                info[1] = base > cap;          // overflow flag
                *active_total = (u64)base;
            }
        }
        if (stamp && it < 2) pf[5 + 5 * it] = (long long)global_ns();
        __syncthreads();                                  // s_part / s_wtot are rewritten by the next block
    }
}

// API view of the pair table: ctx_idx (F,K) ascending kept context indices (-1 padded), mask (F,K), cnt (F).
__global__ void k_expand_pairs(const u64* __restrict__ keep, const u64* __restrict__ nbr, int F, int K,
                               int* __restrict__ ctx_idx, unsigned char* __restrict__ mask, int* __restrict__ cnt) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    u64 kb = keep[f];
    const u64 nb = nbr[f];
    int slot = 0;
    while (kb) {
        const int c = __ffsll((long long)kb) - 1;
        kb &= kb - 1;
        if (slot < K) {
            if (ctx_idx) ctx_idx[(size_t)f * K + slot] = c;
            if (mask) mask[(size_t)f * K + slot] = (unsigned char)((nb >> c) & 1ull);
        }
        ++slot;
    }
    if (cnt) cnt[f] = slot;
    for (; slot < K; ++slot) {
        if (ctx_idx) ctx_idx[(size_t)f * K + slot] = -1;
        if (mask) mask[(size_t)f * K + slot] = 0;
    }
}

// f1 (the step before the path): raw matter-js state (12 floats: px, py, vx, vy, a, av, mass, oid, size, g, f, p) -> the
// 22-float normalised one-hot state, on the device.  Follows data_process:normalize (data_process.lua:119-139) and
// properties2onehotall (:184-207): positions / 800, velocities / 60, angles wrapped (a > pi loses 2 pi) then / pi, each
// operation rounded once; categorical columns one-hot (mass {1, 5, 25, 1e30}, oid {1, 2, 3}, size {0.5, 1, 2}, flags {0, 1}).
// A value outside its category set raises *bad.  One thread per object-frame.
__global__ void k_raw_to_state(const float* __restrict__ raw, float* __restrict__ st, long long n, int* __restrict__ bad) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* r = raw + i * 12;
    float o[D];
#pragma unroll
    for (int k = 0; k < D; ++k) o[k] = 0.f;
    o[0] = __fdiv_rn(r[0], 800.f); o[1] = __fdiv_rn(r[1], 800.f);
    o[2] = __fdiv_rn(r[2], 60.f); o[3] = __fdiv_rn(r[3], 60.f);
    const float PI_F = 3.14159274101257324f, M2PI = -6.28318548202514648f;
#pragma unroll
    for (int k = 4; k < 6; ++k) {
        const float a = r[k];
        o[k] = __fdiv_rn(__fadd_rn(a, __fmul_rn(M2PI, a > PI_F ? 1.f : 0.f)), PI_F);
    }
    bool ok = true;
    const float m = r[6];
    const int mi = m == 1.f ? 0 : m == 5.f ? 1 : m == 25.f ? 2 : m == 1e30f ? 3 : -1;
    ok &= mi >= 0; if (mi >= 0) o[6 + mi] = 1.f;
    const float oid = r[7];
    const int oi = oid == 1.f ? 0 : oid == 2.f ? 1 : oid == 3.f ? 2 : -1;
    ok &= oi >= 0; if (oi >= 0) o[10 + oi] = 1.f;
    const float sz = r[8];
    const int si = sz == 0.5f ? 0 : sz == 1.f ? 1 : sz == 2.f ? 2 : -1;
    ok &= si >= 0; if (si >= 0) o[13 + si] = 1.f;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const float v = r[9 + q];
        const int bi = v == 0.f ? 0 : v == 1.f ? 1 : -1;
        ok &= bi >= 0; if (bi >= 0) o[16 + 2 * q + bi] = 1.f;
    }
    if (!ok) *bad = 1;
    float* w = st + i * D;
#pragma unroll
    for (int k = 0; k < D; k += 2) *reinterpret_cast<float2*>(w + k) = make_float2(o[k], o[k + 1]);
}

// Gathered focus rows (F,44) for parity checks.
__global__ void k_gather_this(FocusSrc src, float* __restrict__ this_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = idx / IN, c = idx - f * IN;
    if (f >= src.F) return;
    this_out[(size_t)f * IN + c] = src.states[row_off(src, src.foc_scene[f], src.foc_obj[f], src.foc_t0[f]) + c];
}

// ----------------------------------------------------------------------------------------------------------------
// shared helpers
// ----------------------------------------------------------------------------------------------------------------
// Gather rows of 44 floats (two frames) into dst[r][LDD]; rows >= nrows are zero-filled.  Row offsets are multiples
// of 22 floats (8-byte aligned): float2 loads.
template <int ROWS, int LDD>
__device__ __forceinline__ void gather_rows(float* dst, const float* __restrict__ base, const long long* offs, int nrows) {
    for (int idx = threadIdx.x; idx < ROWS * (IN / 2); idx += NT) {
        const int r = idx / (IN / 2), c2 = idx - r * (IN / 2);
        float2 v = make_float2(0.f, 0.f);
        if (r < nrows) v = *reinterpret_cast<const float2*>(base + offs[r] + 2 * c2);
        *reinterpret_cast<float2*>(dst + r * LDD + 2 * c2) = v;
    }
}

__device__ __forceinline__ float relu(float v) { return fmaxf(v, 0.f); }

// u1..u4 and raw output o of the decoder on a 64-row focus tile; z = sZ [64][104].  u[i] receives u_{i+1}.
template <int ROWS>
__device__ __forceinline__ void decoder_forward(const float* sWd, const float* sZ, float* const* u, float* o, int nrows) {
    gemm_nt_r<ROWS, 7, LD1, LD1>(sZ, sWd + SW_DEC1, 24, nrows, EpiNT{u[0], LD, 52, 1});
    __syncthreads();
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {
        gemm_nt_r<ROWS, 7, LD, LD>(u[i], sWd + SW_DEC2 + i * SW_PAIR_SZ, 13, nrows, EpiNT{u[i + 1], LD, 52, 1});
        __syncthreads();
    }
    gemm_nt_r<ROWS, 3, LD, LD>(u[3], sWd + SW_DEC5, 13, nrows, EpiNT{o, LD, 24, 0});
    __syncthreads();
}

// Load a focus tile: z[r] = [s x 50 (zero here) | 1 | 0 | this_past(44) | 0 x 8]; rows >= nf are all zero.
template <int ROWS>
__device__ __forceinline__ void load_focus_tile(float* sZ, const float* __restrict__ states, const long long* foff, int nf) {
    for (int idx = threadIdx.x; idx < ROWS * (ZCOL_F / 2); idx += NT) {
        const int r = idx / (ZCOL_F / 2), c2 = idx - r * (ZCOL_F / 2);
        float2 v = make_float2(0.f, 0.f);
        if (c2 == ONE_COL / 2 && r < nf) v.x = 1.f;
        *reinterpret_cast<float2*>(sZ + r * LD1 + 2 * c2) = v;
    }
    for (int idx = threadIdx.x; idx < ROWS * (LD1 - ZCOL_F) / 2; idx += NT) {
        const int r = idx / ((LD1 - ZCOL_F) / 2), c2 = idx - r * ((LD1 - ZCOL_F) / 2);
        float2 v = make_float2(0.f, 0.f);
        if (r < nf && c2 < IN / 2) v = *reinterpret_cast<const float2*>(states + foff[r] + 2 * c2);
        *reinterpret_cast<float2*>(sZ + r * LD1 + ZCOL_F + 2 * c2) = v;
    }
}

// ----------------------------------------------------------------------------------------------------------------
// K2a: pair-MLP forward on DENSE tiles of 64 active pairs (independent of focus tiles): both encoder halves per pair
// (like the reference, which re-encodes the focus for every pair), 5 pairwise layers, h5 rows to Hp (P,52).
// ----------------------------------------------------------------------------------------------------------------
template <int ROWS>
struct PairMeta {
    long long frow[ROWS];
    long long crow[ROWS];
    int pf[ROWS];
};

template <int ROWS>
struct PairFwdSmem {
    float Wp[SW_PAIR_TOTAL];
    float Xf[ROWS * LD];
    float Xc[ROWS * LD];
    float P0[ROWS * LD];
    float P1[ROWS * LD];
    PairMeta<ROWS> M;
    WeightStage W;
};

// h0 = [relu(Wt f) | relu(Wc c)] into h0buf (cols 50..55 zeroed), then h_l = relu(L_l h_{l-1}) into bufs[l].
// Activation stashes between forward and backward, LAYER-MAJOR: [slot][row][HSLOT floats] (52 used; 224-byte rows keep
// every row on 32-byte sector boundaries and make a tile of one layer a contiguous block for TMA bulk copies).
// Pair stash: slots h0..h5, rows = active pairs (stash_rows = row capacity); decoder stash: slots {pair sum | 1 | 0,
// u1..u4}, rows = foci (written by the 3xTF32 decoder forward for the tcgen05 backward).
constexpr int HSLOT = 56;

// Copies rows [0, nrows) x 52 columns of a shared tile to the stash slot `l` of pairs [lo, lo + nrows).
__device__ __forceinline__ void stash_store(float* __restrict__ hact, long long stash_rows, int lo, int l, const float* tile, int nrows) {
    for (int idx = threadIdx.x; idx < nrows * 13; idx += NT) {
        const int r = idx / 13, c4 = idx - r * 13;
        st4(hact + ((size_t)l * stash_rows + lo + r) * HSLOT + 4 * c4, ld4(tile + r * LD + 4 * c4));
    }
}

// hact != nullptr (training): every h_l is also written to the global activation stash so that the backward kernel
// loads it instead of recomputing the chain -- with the LDS-bound FP32 tiles a pass costs ~3.5 k cycles per 64 rows,
// the 2.5 KB/pair round trip (mostly L2) about a quarter of the seven passes it replaces.
template <int ROWS>
__device__ __forceinline__ void pair_chain_dense(const float* sWp, const float* sXf, const float* sXc, int nrows,
                                                 float* const* bufs, float* __restrict__ hact = nullptr, int lo = 0,
                                                 long long stash_rows = 0) {
    float* h0 = bufs[0];
    for (int idx = threadIdx.x; idx < ROWS * 6; idx += NT) h0[(idx / 6) * LD + H + (idx % 6)] = 0.f;
    gemm_nt_r<ROWS, 4, LD, LD>(sXf, sWp + SW_ENC_T, 11, nrows, EpiNT{h0, LD, H2, 1});
    gemm_nt_r<ROWS, 4, LD, LD>(sXc, sWp + SW_ENC_C, 11, nrows, EpiNT{h0 + H2, LD, H2, 1});
    __syncthreads();
    if (hact) stash_store(hact, stash_rows, lo, 0, h0, nrows);
#pragma unroll 1
    for (int l = 0; l < NL; ++l) {
        gemm_nt_r<ROWS, 7, LD, LD>(bufs[l], sWp + SW_PAIR + l * SW_PAIR_SZ, 13, nrows, EpiNT{bufs[l + 1], LD, 52, 1});
        __syncthreads();
        if (hact) stash_store(hact, stash_rows, lo, l + 1, bufs[l + 1], nrows);
    }
}

template <int ROWS>
__device__ __forceinline__ void load_pair_meta(PairMeta<ROWS>& M, const int* __restrict__ pair_foc,
                                               const long long* __restrict__ pair_crow,
                                               const long long* __restrict__ foc_row, int lo, int nrows) {
    if (threadIdx.x < nrows) {
        const int pf = pair_foc[lo + threadIdx.x];
        M.pf[threadIdx.x] = pf;
        M.crow[threadIdx.x] = pair_crow[lo + threadIdx.x];
        M.frow[threadIdx.x] = foc_row[pf];
    }
}

template <int ROWS>
__global__ void __launch_bounds__(NT, 1)
k_pair_forward(const float* __restrict__ states, const int* __restrict__ pair_foc, const long long* __restrict__ pair_crow,
               const long long* __restrict__ foc_row, const int* __restrict__ info, const float* __restrict__ wpack,
               float* __restrict__ Hp, float* __restrict__ hact, long long stash_rows) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PairFwdSmem<ROWS>& S = *reinterpret_cast<PairFwdSmem<ROWS>*>(smem_raw);
    // Programmatic dependent launch.  This kernel's predecessor is either the pair builder (tables complete before
    // any CTA of this grid starts) or the previous step's reduce + optimiser kernel, which releases its dependents at
    // once: everything that does not depend on the NEW weights -- pair list, gathered state rows -- is loaded while
    // that kernel is still running.  The weight image is staged after griddepcontrol.wait, and only then are this
    // kernel's own dependents released (they stage weights in their prologues).
    stage_init(S.W);
    const int P = info[0];
    const int lo0 = blockIdx.x * ROWS;
    if (lo0 < P) {
        load_pair_meta(S.M, pair_foc, pair_crow, foc_row, lo0, min(ROWS, P - lo0));
        __syncthreads();
        gather_rows<ROWS, LD>(S.Xf, states, S.M.frow, min(ROWS, P - lo0));
        gather_rows<ROWS, LD>(S.Xc, states, S.M.crow, min(ROWS, P - lo0));
    }
    __syncthreads();
    pdl_wait();
    pdl_launch_dependents();
    stage_copy(S.W, S.Wp, nullptr, wpack);
    if (lo0 >= P) { stage_wait(S.W); return; }   // never exit with a bulk copy in flight
    bool staged = false;
    float* const bufs[NL + 1] = {S.P0, S.P1, S.P0, S.P1, S.P0, S.P1};
    for (int lo = lo0; lo < P; lo += gridDim.x * ROWS) {
        const int nrows = min(ROWS, P - lo);
        if (lo != lo0) {
            __syncthreads();
            load_pair_meta(S.M, pair_foc, pair_crow, foc_row, lo, nrows);
            __syncthreads();
            gather_rows<ROWS, LD>(S.Xf, states, S.M.frow, nrows);
            gather_rows<ROWS, LD>(S.Xc, states, S.M.crow, nrows);
            __syncthreads();
        }
        if (!staged) { stage_wait(S.W); staged = true; }
        pair_chain_dense<ROWS>(S.Wp, S.Xf, S.Xc, nrows, bufs, hact, lo, stash_rows);
        for (int idx = threadIdx.x; idx < nrows * 13; idx += NT) {
            const int r = idx / 13, c4 = idx - r * 13;
            st4(Hp + (size_t)(lo + r) * 52 + 4 * c4, ld4(S.P1 + r * LD + 4 * c4));
        }
    }
}

// ----------------------------------------------------------------------------------------------------------------
// K2b + K3: per-focus segmented sum of the pair outputs (fixed order -> deterministic), decoder, sigmoid split,
// per-example losses.
// ----------------------------------------------------------------------------------------------------------------
template <int ROWS>
struct DecFwdSmem {
    float Wd[SW_DEC_TOTAL];
    float Z[ROWS * LD1];
    float P0[ROWS * LD];
    float P1[ROWS * LD];
    float O[ROWS * LD];
    long long foff[ROWS];
    int pstart[ROWS + 1];
    WeightStage W;
};

__device__ __forceinline__ void focus_sums(float* sZ, const int* pstart, const float* __restrict__ Hp, int nf) {
    for (int idx = threadIdx.x; idx < nf * 13; idx += NT) {
        const int f = idx / 13, c4 = idx - f * 13;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int p = pstart[f]; p < pstart[f + 1]; ++p) {
            const float4 v = ld4(Hp + (size_t)p * 52 + 4 * c4);
            a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        if (c4 == 12) { a.z = 1.f; a.w = 0.f; }   // columns 50, 51: constant-1 column, pad
        st4(sZ + f * LD1 + 4 * c4, a);
    }
}

template <int ROWS>
__global__ void __launch_bounds__(NT, 1)
k_dec_forward(const float* __restrict__ states, const long long* __restrict__ foc_row, int F,
              const int* __restrict__ pair_start, const float* __restrict__ Hp, const float* __restrict__ wpack,
              const float* __restrict__ y, float* __restrict__ pred, float* __restrict__ loss_ex) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    DecFwdSmem<ROWS>& S = *reinterpret_cast<DecFwdSmem<ROWS>*>(smem_raw);
    pdl_launch_dependents();
    stage_issue(S.W, nullptr, S.Wd, wpack);
    pdl_wait();
    bool staged = false;
    float* const ubufs[4] = {S.P0, S.P1, S.P0, S.P1};
    const int ntiles = (F + ROWS - 1) / ROWS;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int f0 = tile * ROWS;
        const int nf = min(ROWS, F - f0);
        __syncthreads();
        if (threadIdx.x < nf) S.foff[threadIdx.x] = foc_row[f0 + threadIdx.x];
        if (threadIdx.x <= nf) S.pstart[threadIdx.x] = pair_start[f0 + threadIdx.x];
        __syncthreads();
        load_focus_tile<ROWS>(S.Z, states, S.foff, nf);
        __syncthreads();
        focus_sums(S.Z, S.pstart, Hp, nf);
        __syncthreads();
        if (!staged) { stage_wait(S.W); staged = true; }
        decoder_forward<ROWS>(S.Wd, S.Z, ubufs, S.O, nf);
        for (int idx = threadIdx.x; idx < nf * OUT; idx += NT) {
            const int r = idx / OUT, n = idx - r * OUT;
            float v = S.O[r * LD + n];
            if (n >= 6) v = 1.f / (1.f + expf(-v));   // Sigmoid on the property columns (modules.lua:80-86)
            pred[(size_t)(f0 + r) * OUT + n] = v;
        }
        if (loss_ex && y && threadIdx.x < nf) {
            const int r = threadIdx.x;
            const float* yy = y + (size_t)(f0 + r) * OUT;
            const float e2 = S.O[r * LD + 2] - yy[2], e3 = S.O[r * LD + 3] - yy[3], e5 = S.O[r * LD + 5] - yy[5];
            const float lv = __fadd_rn(__fmul_rn(e2, e2), __fmul_rn(e3, e3)), lw = __fmul_rn(e5, e5);
            loss_ex[(size_t)(f0 + r) * 3 + 0] = __fdiv_rn(__fadd_rn(lv, lw), 3.f);   // npe.lua:271-272
            loss_ex[(size_t)(f0 + r) * 3 + 1] = __fdiv_rn(lv, 2.f);                  // npe.lua:273
            loss_ex[(size_t)(f0 + r) * 3 + 2] = lw;                                  // npe.lua:274
        }
    }
}

// Batch losses from the per-example ones (npe.lua:238-246): {sum/(3B), vel/(2B), angvel/B}; double accumulation like
// THNN's accreal.
__global__ void k_loss_reduce(const float* __restrict__ loss_ex, int F, float* __restrict__ loss3, double* __restrict__ sums_out) {
    __shared__ double sv[256], sw[256];
    double av = 0.0, aw = 0.0;
    for (int f = threadIdx.x; f < F; f += 256) {
        av += 2.0 * (double)loss_ex[(size_t)f * 3 + 1];
        aw += (double)loss_ex[(size_t)f * 3 + 2];
    }
    sv[threadIdx.x] = av; sw[threadIdx.x] = aw;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { sv[threadIdx.x] += sv[threadIdx.x + s]; sw[threadIdx.x] += sw[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double B = (double)F;
        loss3[0] = (float)((sv[0] + sw[0]) / (3.0 * B));
        loss3[1] = (float)(sv[0] / (2.0 * B));
        loss3[2] = (float)(sw[0] / B);
        if (sums_out) { sums_out[0] = sv[0]; sums_out[1] = sw[0]; sums_out[2] = B; }
    }
}

// Register tile -> 16 consecutive floats of the tile-major partial.
__device__ __forceinline__ void tile_store(float* __restrict__ dst, const float (&acc)[16]) {
#pragma unroll
    for (int q = 0; q < 4; ++q) st4(dst + 4 * q, make_float4(acc[4 * q], acc[4 * q + 1], acc[4 * q + 2], acc[4 * q + 3]));
}

// Decoder weight-gradient register tiles -> this CTA's partial (tile-major, npe_layout.h).  Thread t owns tiles t and
// t + NT of layer 1 (13 x 24 tiles of 4 x 4), tile t of layers 2..4 (13 x 13) and tile t - 169 of layer 5 (6 x 13).
__device__ __forceinline__ void dec_partials_store(float* __restrict__ out, const float (&a1a)[16], const float (&a1b)[16],
                                                   const float (&a2)[16], const float (&a3)[16], const float (&a4)[16],
                                                   const float (&a5)[16]) {
    const int t = threadIdx.x;
    tile_store(out + PT_DEC1A + t * 16, a1a);
    if (t + NT < 13 * 24) tile_store(out + PT_DEC1B + t * 16, a1b);
    if (t < 169) {
        tile_store(out + PT_DEC2 + t * 16, a2);
        tile_store(out + PT_DEC2 + PT_LAYER + t * 16, a3);
        tile_store(out + PT_DEC2 + 2 * PT_LAYER + t * 16, a4);
    }
    if (t >= 169 && t < 169 + 78) tile_store(out + PT_DEC5 + (t - 169) * 16, a5);
}

// ----------------------------------------------------------------------------------------------------------------
// K4a: decoder backward.  d_pred = 2 (p - y) vlambda / (2B) on cols 2,3; 2 (p - y) lambda / B on col 5; zero
// elsewhere (npe.lua:301-320).  Weight-gradient tiles live in registers for the whole kernel.
// ----------------------------------------------------------------------------------------------------------------
template <int ROWS>
struct DecBwdSmem {
    float Wd[SW_DEC_TOTAL];
    float Z[ROWS * LD1];
    float U[4][ROWS * LD];
    float O[ROWS * LD];
    float Da[ROWS * LD];
    float Db[ROWS * LD];
    long long foff[ROWS];
    int pstart[ROWS + 1];
    WeightStage W;
};

template <int ROWS>
__global__ void __launch_bounds__(NT, 1)
k_dec_backward(const float* __restrict__ states, const long long* __restrict__ foc_row, int F,
               const int* __restrict__ pair_start, const float* __restrict__ Hp, const float* __restrict__ wpack,
               const float* __restrict__ y, float* __restrict__ ds_out, float* __restrict__ partial,
               float scale_v, float scale_w, float* __restrict__ pred, float* __restrict__ loss_ex) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    DecBwdSmem<ROWS>& S = *reinterpret_cast<DecBwdSmem<ROWS>*>(smem_raw);
    pdl_launch_dependents();
    for (int idx = threadIdx.x; idx < ROWS * LD; idx += NT) S.Db[idx] = 0.f;   // pad columns 52..55 are read by the 16-row mapping
    // the predecessor (pair forward) releases this grid only after the previous optimiser step is complete, so the
    // weight image and the static inputs of the first tile (focus rows, pair ranges) load while it runs
    stage_issue(S.W, nullptr, S.Wd, wpack);
    const int t = threadIdx.x;
    const int ntiles = (F + ROWS - 1) / ROWS;
    auto load_tile_inputs = [&](int tile) {
        const int f0 = tile * ROWS;
        const int nf = min(ROWS, F - f0);
        __syncthreads();
        if (t < nf) S.foff[t] = foc_row[f0 + t];
        if (t <= nf) S.pstart[t] = pair_start[f0 + t];
        __syncthreads();
        load_focus_tile<ROWS>(S.Z, states, S.foff, nf);
        for (int idx = t; idx < ROWS * LD; idx += NT) S.Da[idx] = 0.f;
        __syncthreads();
    };
    if ((int)blockIdx.x < ntiles) load_tile_inputs(blockIdx.x);
    pdl_wait();
    bool staged = false;
    float* const ubufs[4] = {S.U[0], S.U[1], S.U[2], S.U[3]};
    float a1a[16], a1b[16], a2[16], a3[16], a4[16], a5[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a1a[i] = a1b[i] = a2[i] = a3[i] = a4[i] = a5[i] = 0.f; }
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int f0 = tile * ROWS;
        const int nf = min(ROWS, F - f0);
        if (tile != (int)blockIdx.x) load_tile_inputs(tile);
        focus_sums(S.Z, S.pstart, Hp, nf);
        __syncthreads();
        if (!staged) { stage_wait(S.W); staged = true; }
        decoder_forward<ROWS>(S.Wd, S.Z, ubufs, S.O, nf);
        if (pred) {   // fused training step: this kernel IS the decoder forward (model:fp outputs) as well
            for (int idx = t; idx < nf * OUT; idx += NT) {
                const int r = idx / OUT, n = idx - r * OUT;
                float v = S.O[r * LD + n];
                if (n >= 6) v = 1.f / (1.f + expf(-v));
                pred[(size_t)(f0 + r) * OUT + n] = v;
            }
            if (loss_ex && t < nf) {
                const float* yy = y + (size_t)(f0 + t) * OUT;
                const float e2 = S.O[t * LD + 2] - yy[2], e3 = S.O[t * LD + 3] - yy[3], e5 = S.O[t * LD + 5] - yy[5];
                const float lv = __fadd_rn(__fmul_rn(e2, e2), __fmul_rn(e3, e3)), lw = __fmul_rn(e5, e5);
                loss_ex[(size_t)(f0 + t) * 3 + 0] = __fdiv_rn(__fadd_rn(lv, lw), 3.f);
                loss_ex[(size_t)(f0 + t) * 3 + 1] = __fdiv_rn(lv, 2.f);
                loss_ex[(size_t)(f0 + t) * 3 + 2] = lw;
            }
        }
        if (t < nf) {
            const float* yy = y + (size_t)(f0 + t) * OUT;
            S.Da[t * LD + 2] = (S.O[t * LD + 2] - yy[2]) * scale_v;
            S.Da[t * LD + 3] = (S.O[t * LD + 3] - yy[3]) * scale_v;
            S.Da[t * LD + 5] = (S.O[t * LD + 5] - yy[5]) * scale_w;
        }
        __syncthreads();
        // layer 5 (22 x 50): dz = Da (24 cols), input u4
        if (t >= 169 && t < 169 + 78) wgrad<13, LD, LD, true>(a5, S.Da, S.U[3], t - 169, nf);
        gemm_nn_r<ROWS, 13, LD, LD>(S.Da, S.Wd + SW_DEC5, 6, nf, EpiNN{S.Db, LD, S.U[3], ROWS});
        __syncthreads();
        // layer 4: dz = Db, input u3
        if (t < 169) wgrad<13, LD, LD, true>(a4, S.Db, S.U[2], t, nf);
        gemm_nn_r<ROWS, 13, LD, LD>(S.Db, S.Wd + SW_DEC2 + 2 * SW_PAIR_SZ, 13, nf, EpiNN{S.Da, LD, S.U[2], ROWS});
        __syncthreads();
        // layer 3: dz = Da, input u2
        if (t < 169) wgrad<13, LD, LD, true>(a3, S.Da, S.U[1], t, nf);
        gemm_nn_r<ROWS, 13, LD, LD>(S.Da, S.Wd + SW_DEC2 + SW_PAIR_SZ, 13, nf, EpiNN{S.Db, LD, S.U[1], ROWS});
        __syncthreads();
        // layer 2: dz = Db, input u1
        if (t < 169) wgrad<13, LD, LD, true>(a2, S.Db, S.U[0], t, nf);
        gemm_nn_r<ROWS, 13, LD, LD>(S.Db, S.Wd + SW_DEC2, 13, nf, EpiNN{S.Da, LD, S.U[0], ROWS});
        __syncthreads();
        // layer 1 (50 x 94 (+bias)): dz = Da, input z (24 k-chunks); input grad only for the pair-sum columns
        wgrad<24, LD, LD1, true>(a1a, S.Da, S.Z, t, nf);
        if (t + NT < 13 * 24) wgrad<24, LD, LD1, true>(a1b, S.Da, S.Z, t + NT, nf);
        gemm_nn_r<ROWS, 13, LD, LD1>(S.Da, S.Wd + SW_DEC1, 13, nf, EpiNN{ds_out + (size_t)f0 * 52, 52, nullptr, nf});
    }
    dec_partials_store(partial + (size_t)blockIdx.x * PT_TOTAL, a1a, a1b, a2, a3, a4, a5);
}

// Pair-MLP weight-gradient register tiles -> this CTA's partial (tile-major).  Thread t < 169 owns tile t of every
// pairwise layer; threads 169..245 own one tile of each encoder (acc[0] = context encoder, acc[1] = focus encoder).
__device__ __forceinline__ void pair_partials_store(float* __restrict__ out, const float (&acc)[NL][16]) {
    const int t = threadIdx.x;
    if (t < 169) {
#pragma unroll
        for (int l = 0; l < NL; ++l) tile_store(out + PT_PAIR + l * PT_LAYER + t * 16, acc[l]);
    } else if (t < 169 + 77) {
        tile_store(out + PT_ENC_C + (t - 169) * 16, acc[0]);
        tile_store(out + PT_ENC_T + (t - 169) * 16, acc[1]);
    }
}

// ----------------------------------------------------------------------------------------------------------------
// K4b: pair-MLP backward on dense 64-pair tiles.  The forward chain is recomputed in the tile (no activation round
// trip through HBM); dz5 = ds[focus(p)] * (h5 > 0) (CAddTable backward = copy; apply_mask is the identity on the
// active list, npe.lua:326-328); weight-gradient tiles stay in registers for the whole kernel.
// ----------------------------------------------------------------------------------------------------------------
template <int ROWS>
struct PairBwdSmem {
    float Wp[SW_PAIR_TOTAL];
    float Xf[ROWS * LD];
    float Xc[ROWS * LD];
    float Hh[NL + 1][ROWS * LD];
    float Da[ROWS * LD];
    float Db[ROWS * LD];
    PairMeta<ROWS> M;
    WeightStage W;
};

template <int ROWS>
__global__ void __launch_bounds__(NT, 1)
k_pair_backward(const float* __restrict__ states, const int* __restrict__ pair_foc, const long long* __restrict__ pair_crow,
                const long long* __restrict__ foc_row, const int* __restrict__ info, const float* __restrict__ wpack,
                const float* __restrict__ ds_in, float* __restrict__ partial, const float* __restrict__ hact,
                long long stash_rows) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PairBwdSmem<ROWS>& S = *reinterpret_cast<PairBwdSmem<ROWS>*>(smem_raw);
    pdl_launch_dependents();
    for (int idx = threadIdx.x; idx < ROWS * LD; idx += NT) { S.Da[idx] = 0.f; S.Db[idx] = 0.f; }   // pad columns stay zero
    stage_issue(S.W, S.Wp, nullptr, wpack);
    const int P = info[0];
    const int t = threadIdx.x;
    const int lo0 = blockIdx.x * ROWS;
    if (lo0 < P) {   // static inputs of the first tile load while the decoder kernel runs
        load_pair_meta(S.M, pair_foc, pair_crow, foc_row, lo0, min(ROWS, P - lo0));
        __syncthreads();
        gather_rows<ROWS, LD>(S.Xf, states, S.M.frow, min(ROWS, P - lo0));
        gather_rows<ROWS, LD>(S.Xc, states, S.M.crow, min(ROWS, P - lo0));
        __syncthreads();
    }
    pdl_wait();
    float acc[NL][16];
#pragma unroll
    for (int l = 0; l < NL; ++l)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[l][i] = 0.f;
    const bool enc_thread = (t >= 169 && t < 169 + 77);
    float* const bufs[NL + 1] = {S.Hh[0], S.Hh[1], S.Hh[2], S.Hh[3], S.Hh[4], S.Hh[5]};
    if ((int)blockIdx.x * ROWS >= P) stage_wait(S.W);   // never exit with a bulk copy in flight
    if ((int)blockIdx.x * ROWS < P) {
        bool staged = false;
        for (int lo = lo0; lo < P; lo += gridDim.x * ROWS) {
            const int nrows = min(ROWS, P - lo);
            if (lo != lo0) {
                __syncthreads();
                load_pair_meta(S.M, pair_foc, pair_crow, foc_row, lo, nrows);
                __syncthreads();
                gather_rows<ROWS, LD>(S.Xf, states, S.M.frow, nrows);
                gather_rows<ROWS, LD>(S.Xc, states, S.M.crow, nrows);
                __syncthreads();
            }
            if (!staged) { stage_wait(S.W); staged = true; }
            if (hact) {
                // activations of this tile from the forward kernel's stash (rows >= nrows and pad columns zero)
                for (int idx = t; idx < (NL + 1) * ROWS * (LD / 4); idx += NT) {
                    const int l = idx / (ROWS * (LD / 4)), rem = idx - l * (ROWS * (LD / 4));
                    const int r = rem / (LD / 4), c4 = rem - r * (LD / 4);
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < nrows && c4 < 13) v = ld4(hact + ((size_t)l * stash_rows + lo + r) * HSLOT + 4 * c4);
                    st4(S.Hh[l] + r * LD + 4 * c4, v);
                }
                __syncthreads();
            } else {
                pair_chain_dense<ROWS>(S.Wp, S.Xf, S.Xc, nrows, bufs);
            }
            // dz5 = ds[focus] * (h5 > 0); rows >= nrows zero
            for (int idx = t; idx < ROWS * (LD / 4); idx += NT) {
                const int r = idx / (LD / 4), c4 = idx - r * (LD / 4);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < nrows && c4 < 13) {
                    v = ld4(ds_in + (size_t)S.M.pf[r] * 52 + 4 * c4);
                    const float4 hv = ld4(S.Hh[NL] + r * LD + 4 * c4);
                    v.x = hv.x > 0.f ? v.x : 0.f; v.y = hv.y > 0.f ? v.y : 0.f;
                    v.z = hv.z > 0.f ? v.z : 0.f; v.w = hv.w > 0.f ? v.w : 0.f;
                }
                st4(S.Da + r * LD + 4 * c4, v);
            }
            __syncthreads();
            float* da = S.Da;
            float* db = S.Db;
#pragma unroll
            for (int l = NL - 1; l >= 0; --l) {
                const float* hin = S.Hh[l];
                if (t < 169) wgrad<13, LD, LD, true>(acc[l], da, hin, t, nrows);
                gemm_nn_r<ROWS, 13, LD, LD>(da, S.Wp + SW_PAIR + l * SW_PAIR_SZ, 13, nrows, EpiNN{db, LD, hin, ROWS});
                __syncthreads();
                float* tmp = da; da = db; db = tmp;
            }
            // da = dz0: cols 0..24 -> focus encoder (input Xf), cols 25..49 -> context encoder (input Xc)
            if (enc_thread) {
                wgrad<11, LD, LD, false>(acc[0], da, S.Xc, t - 169, nrows, H2);
                wgrad<11, LD, LD, true>(acc[1], da, S.Xf, t - 169, nrows, 0);
            }
        }
    }
    pair_partials_store(partial + (size_t)blockIdx.x * PT_TOTAL, acc);
}

// grad[i] = sum over CTA partials in CTA order (deterministic); pair part and decoder part have their own grids.
// OPT = 1 / 2 fuses the Adam / RMSprop update (and the write-through to the packed weight image) into the same
// kernel: used whenever no all-reduce sits between the reduction and the optimiser (single GPU).
struct OptArgs {
    float* x; float* m; float* v; float* wpack;
    float b1, omb1, b2, omb2, eps, step;   // adam: step = lr * sqrt(1 - b2^t) / (1 - b1^t); rmsprop: b1 = alpha, step = lr
    // hi / lo TF32 weight images of the tcgen05 kernels, written through like wpack when the step uses them (else nullptr):
    // three re-packing launches per training step (4 us each, latency-bound) disappear
    float* wtc3_hi; float* wtc3_lo; float* wtt3_hi; float* wtt3_lo; float* wenc_hi; float* wenc_lo;
};
__host__ __device__ __forceinline__ int tc_index(int i);      // npe_tc.cuh: forward / K-major image
__host__ __device__ __forceinline__ int tt_index(int i);      // npe_tcbwd.cuh: transposed image of the input-gradient chains (-1: not part of it)
__host__ __device__ __forceinline__ int enc64_index(int i);   // npe_chain.cuh: 64-row encoder images
__device__ __forceinline__ void opt_write_images(int i, float xn, const OptArgs& o) {
    if (!o.wtc3_hi) return;
    const float hi = __uint_as_float(__float_as_uint(xn) & 0xFFFFE000u), lo = xn - hi;
    const int a = tc_index(i);
    o.wtc3_hi[a] = hi; o.wtc3_lo[a] = lo;
    const int b = tt_index(i);
    if (b >= 0) { o.wtt3_hi[b] = hi; o.wtt3_lo[b] = lo; }
    if (i < OFF_PAIR) { const int e = enc64_index(i); o.wenc_hi[e] = hi; o.wenc_lo[e] = lo; }
}
// Last block (index gridDim.x - 1, launched as an extra block) optionally reduces the per-example losses and publishes
// {loss, vel, angvel, overflow flag} to `loss4_out` (device or host-mapped memory): saves a launch and a copy per step.
struct LossArgs {
    const float* loss_ex; int F; float* loss4_out; const int* info;
};
__device__ __forceinline__ void block_loss_reduce(const LossArgs& L) {
    __shared__ double sv[256], sw[256];
    double av = 0.0, aw = 0.0;
    for (int f = threadIdx.x; f < L.F; f += 256) {
        av += 2.0 * (double)__ldcg(L.loss_ex + (size_t)f * 3 + 1);
        aw += (double)__ldcg(L.loss_ex + (size_t)f * 3 + 2);
    }
    sv[threadIdx.x] = av; sw[threadIdx.x] = aw;
    __syncthreads();
    for (int s = 128; s > 0; s >>= 1) {
        if (threadIdx.x < s) { sv[threadIdx.x] += sv[threadIdx.x + s]; sw[threadIdx.x] += sw[threadIdx.x + s]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const double B = (double)L.F;
        L.loss4_out[0] = (float)((sv[0] + sw[0]) / (3.0 * B));
        L.loss4_out[1] = (float)(sv[0] / (2.0 * B));
        L.loss4_out[2] = (float)(sw[0] / B);
        L.loss4_out[3] = L.info ? (float)L.info[1] : 0.f;
        __threadfence_system();
    }
}

// Adam (OPT = 1) / RMSprop (OPT = 2) update of parameter i with gradient a, one rounding per operation (upstream optim
// semantics, main.lua:135-144,301-302), plus the write-through to the packed weight image.  OPT = 0: nothing.
// The optimiser state of parameter i; it does not depend on the gradient kernels, so the reduce kernels fetch it
// BEFORE griddepcontrol.wait (through L2: a parameter may be updated by a different SM every step).
struct OptState { float m, v, x; };
template <int OPT>
__device__ __forceinline__ OptState opt_load(int i, const OptArgs& o) {
    OptState s{0.f, 0.f, 0.f};
    if (OPT != 0) { s.m = __ldcg(o.m + i); s.x = __ldcg(o.x + i); }
    if (OPT == 1) s.v = __ldcg(o.v + i);
    return s;
}
template <int OPT>
__device__ __forceinline__ void opt_update(int i, float a, const OptArgs& o, const OptState& s) {
    if (OPT == 1) {
        const float mi = __fadd_rn(__fmul_rn(o.b1, s.m), __fmul_rn(o.omb1, a));
        const float vi = __fadd_rn(__fmul_rn(o.b2, s.v), __fmul_rn(o.omb2, __fmul_rn(a, a)));
        o.m[i] = mi; o.v[i] = vi;
        const float xn = __fsub_rn(s.x, __fdiv_rn(__fmul_rn(o.step, mi), __fadd_rn(__fsqrt_rn(vi), o.eps)));
        o.x[i] = xn;
        o.wpack[pack_index(i)] = xn;
        opt_write_images(i, xn, o);
    } else if (OPT == 2) {
        const float mi = __fadd_rn(__fmul_rn(o.b1, s.m), __fmul_rn(o.omb1, __fmul_rn(a, a)));
        o.m[i] = mi;
        const float xn = __fsub_rn(s.x, __fdiv_rn(__fmul_rn(o.step, a), __fadd_rn(__fsqrt_rn(mi), o.eps)));
        o.x[i] = xn;
        o.wpack[pack_index(i)] = xn;
        opt_write_images(i, xn, o);
    }
}

// Flat (reference) parameter index -> offset inside a CTA's tile-major partial: element (n, k) of a layer lives in tile
// (n / 4, k / 4) of the thread mapping of wgrad, at 4 (n % 4) + (k % 4).  Decoder biases are column 50 (the constant-1
// activation column); decoder layer 1 skips the pad column 51 (z = [s(50) | 1 | 0 | this_past(44)]).
__device__ __forceinline__ int partial_index(int i) {
    int base, n, k, kt;                                   // kt = tiles per row of the layer
    if (i < OFF_PAIR) {                                   // encoders, 25 x 44
        const int r = i < OFF_ENC_C ? i : i - OFF_ENC_C;
        base = i < OFF_ENC_C ? PT_ENC_T : PT_ENC_C; n = r / IN; k = r - n * IN; kt = 11;
    } else if (i < OFF_DEC1_W) {                          // pairwise layers, 50 x 50
        const int r0 = i - OFF_PAIR, l = r0 / 2500, r = r0 - l * 2500;
        base = PT_PAIR + l * PT_LAYER; n = r / H; k = r - n * H; kt = 13;
    } else if (i < OFF_DEC2_W) {                          // decoder layer 1, 50 x 94 + bias
        if (i < OFF_DEC1_B) { const int r = i - OFF_DEC1_W; n = r / DEC_IN; k = r - n * DEC_IN; if (k >= H) k += 2; }
        else { n = i - OFF_DEC1_B; k = ONE_COL; }
        const int tile = (n >> 2) * 24 + (k >> 2);
        return (tile < NT ? PT_DEC1A + tile * 16 : PT_DEC1B + (tile - NT) * 16) + 4 * (n & 3) + (k & 3);
    } else if (i < OFF_DEC5_W) {                          // decoder layers 2..4, 50 x 50 + bias
        const int r0 = i - OFF_DEC2_W, l = r0 / 2550, r = r0 - l * 2550;
        base = PT_DEC2 + l * PT_LAYER; kt = 13;
        if (r < 2500) { n = r / H; k = r - n * H; } else { n = r - 2500; k = ONE_COL; }
    } else {                                              // decoder layer 5, 22 x 50 + bias
        const int r = i - OFF_DEC5_W;
        base = PT_DEC5; kt = 13;
        if (r < OUT * H) { n = r / H; k = r - n * H; } else { n = r - OUT * H; k = ONE_COL; }
    }
    return base + ((n >> 2) * kt + (k >> 2)) * 16 + 4 * (n & 3) + (k & 3);
}

// Sum of partial[b][i] over b < np in a fixed order (deterministic), sixteen loads in flight at a time: the additions keep
// their order (padding with +0 does not change a sum), only the L2 latencies overlap.  RS = 1: one thread per parameter,
// CTA order.  RS = 4 (many partials: 148 per part in the throughput regime, i.e. ten dependent rounds of L2 latency for one
// thread): four adjacent lanes sum a contiguous quarter of the partials each, then (q0 + q1) + (q2 + q3).
template <int RS>
__device__ __forceinline__ float ordered_partial_sum(const float* __restrict__ partial, int np, int i, int sub, bool live) {
    float a = 0.f;
    if (live) {
        const float* src = partial + partial_index(i);
        const int chunk = (np + RS - 1) / RS, b_lo = sub * chunk, b_hi = min(np, b_lo + chunk);
        for (int b0 = b_lo; b0 < b_hi; b0 += 16) {
            float v[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) v[j] = (b0 + j < b_hi) ? __ldcg(src + (size_t)(b0 + j) * PT_TOTAL) : 0.f;
#pragma unroll
            for (int j = 0; j < 16; ++j) a += v[j];
        }
    }
    if (RS == 4) {
        a += __shfl_xor_sync(0xffffffffu, a, 1);
        a += __shfl_xor_sync(0xffffffffu, a, 2);
    }
    return a;
}

template <int OPT, int RS>
__global__ void k_reduce_grads(const float* __restrict__ partial, int nparts_pair, int nparts_dec, float* __restrict__ grad, OptArgs o,
                               LossArgs L) {
    // dependents are released at once: the pair forward kernel of the next step loads its (static) pair list and state
    // rows while this kernel runs and touches the weight image only after its own griddepcontrol.wait
    pdl_launch_dependents();
    const int gid = blockIdx.x * blockDim.x + threadIdx.x, i = gid / RS, sub = gid % RS;
    const bool loss_block = L.loss4_out && blockIdx.x == gridDim.x - 1;
    const bool live = !loss_block && i < NPARAM, owner = live && sub == 0;
    OptState st{0.f, 0.f, 0.f};
    if (owner) st = opt_load<OPT>(i, o);
    pdl_wait();
    if (loss_block) { block_loss_reduce(L); return; }
    const int np = (i < NPARAM_PAIR) ? nparts_pair : nparts_dec;
    const float a = ordered_partial_sum<RS>(partial, np, i, sub, live);
    if (!owner) return;
    grad[i] = a;
    opt_update<OPT>(i, a, o, st);
}

// ----------------------------------------------------------------------------------------------------------------
// Data-parallel step over NVLink peer memory: local reduction of the CTA partials, gradient exchange, optimiser and
// weight-image write-through in ONE kernel.  Every rank owns an exchange region (CUDA IPC) of 8-byte cells
//     cell[2 buffers][MAX_PEERS source ranks][DP_NPAD]  =  {gradient bits, step stamp}
// Thread i of rank q reduces parameter i over the CTA partials and PUSHES {value, stamp} into cell [buf][q][i] of every
// rank's region with ONE aligned 8-byte store per peer (posted NVLink writes; value and flag travel together, so no
// fence, no separate flag and no read round trip — the idea of NCCL's LL protocol).  It then polls the cells [buf][r][i]
// of its OWN region until every rank's stamp has arrived, sums the values in rank order (the same sum on every rank:
// replicas stay bit-identical) and applies Adam / RMSprop.  Parameters are independent: no barrier of any kind.
// Buffers alternate by step parity; a rank can only push step s+2 into a buffer after it has consumed every peer's
// step s+1 cell, which that peer pushed after finishing step s, so no "done" handshake is needed.  Every thread pushes
// before it waits, so ranks cannot deadlock each other.
// ----------------------------------------------------------------------------------------------------------------
constexpr int MAX_PEERS = 8;
constexpr int DP_NBLK = (NPARAM + 255) / 256;       // 111 CTAs
constexpr int DP_NPAD = DP_NBLK * 256;
constexpr size_t DP_REGION_BYTES = 2ull * MAX_PEERS * DP_NPAD * sizeof(uint2);
struct DpArgs {
    uint2* region[MAX_PEERS];             // exchange regions of all ranks (own region at index `rank`)
    int nranks, rank;
    int* abort;                           // info[1]: set to 3 when a peer's cell does not arrive within timeout_ns
    unsigned long long timeout_ns;
    long long* prof;                      // tools only (NPE_CHAIN_PROF): 4 globaltimer stamps per CTA
    int debug_mode;                       // tools only (NPE_DP_DEBUG): 1 = push but do not wait (WRONG sums), 2 = neither push nor wait
};
__device__ __forceinline__ void st_cell(uint2* p, float v, unsigned stamp) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(stamp) : "memory");
}
__device__ __forceinline__ uint2 ld_cell(const uint2* p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}

// Pushes this rank's value of parameter i into every peer's region and returns the rank-order sum over all ranks.
__device__ __forceinline__ float peer_exchange(int i, float a, const DpArgs& P, int buf, unsigned stamp) {
    const size_t mine = ((size_t)buf * MAX_PEERS + P.rank) * DP_NPAD + i;
    if (P.debug_mode == 2) return a;
    for (int r = 1; r < P.nranks; ++r) {                 // remote ranks
        const int q = P.rank + r < P.nranks ? P.rank + r : P.rank + r - P.nranks;
        st_cell(P.region[q] + mine, a, stamp);
    }
    if (P.debug_mode == 1) return a;
    const uint2* in = P.region[P.rank] + (size_t)buf * MAX_PEERS * DP_NPAD + i;
    // All peers' cells are read TOGETHER (one L2 round trip for the common case that they have all arrived; polling them
    // one after the other costs nranks - 1 dependent round trips), then only the missing ones are polled again.
    uint2 c[MAX_PEERS];
    unsigned missing = 0;
#pragma unroll
    for (int r = 0; r < MAX_PEERS; ++r) {
        c[r] = make_uint2(__float_as_uint(a), stamp);
        if (r < P.nranks && r != P.rank) c[r] = ld_cell(in + (size_t)r * DP_NPAD);
    }
#pragma unroll
    for (int r = 0; r < MAX_PEERS; ++r) missing |= (c[r].y != stamp ? 1u : 0u) << r;
    if (missing) {
        // Bounded wait (like wait_flags of the one-grid step): a peer that faulted, stopped stepping or issued a different
        // number of steps must surface as an error on this rank, not as a wedged GPU.  Ranks have to issue the same number
        // of npe_train_step calls.
        const unsigned long long t0 = global_ns();
        int spins = 0;
        while (missing) {
#pragma unroll
            for (int r = 0; r < MAX_PEERS; ++r)
                if (missing >> r & 1u) {
                    c[r] = ld_cell(in + (size_t)r * DP_NPAD);
                    if (c[r].y == stamp) missing &= ~(1u << r);
                }
            if (missing && ++spins > 64) {
                __nanosleep(64);
                if ((spins & 255) == 0 &&
                    (*(volatile int*)P.abort == 3 || global_ns() - t0 > P.timeout_ns)) { *P.abort = 3; break; }
            }
        }
    }
    float g = 0.f;
#pragma unroll
    for (int r = 0; r < MAX_PEERS; ++r)                  // rank order: the same sum on every rank
        if (r < P.nranks) g += __uint_as_float(c[r].x);
    return g;
}

template <int OPT, int RS>
__global__ void __launch_bounds__(256) k_dp_step(const float* __restrict__ partial, int nparts_pair, int nparts_dec, const __grid_constant__ DpArgs P, int buf,
                                                 unsigned stamp, float* __restrict__ grad, OptArgs o, LossArgs L) {
    pdl_launch_dependents();   // as in k_reduce_grads: the next kernel stages the weight image only after its wait
    const int gid = blockIdx.x * blockDim.x + threadIdx.x, i = gid / RS, sub = gid % RS;
    const bool loss_block = L.loss4_out && blockIdx.x == gridDim.x - 1;
    const bool live = !loss_block && i < NPARAM, owner = live && sub == 0;
    OptState st{0.f, 0.f, 0.f};
    if (owner) st = opt_load<OPT>(i, o);
    pdl_wait();
    if (loss_block) { block_loss_reduce(L); return; }
    const bool stamp_it = P.prof && threadIdx.x == 0;
    long long* const prof = P.prof + (stamp & 1u) * 4 * DP_NBLK;      // the last TWO launches stay readable
    if (stamp_it) prof[4 * blockIdx.x] = (long long)global_ns();
    const int np = (i < NPARAM_PAIR) ? nparts_pair : nparts_dec;
    float a = ordered_partial_sum<RS>(partial, np, i, sub, live);
    if (stamp_it) prof[4 * blockIdx.x + 1] = (long long)global_ns();
    if (!owner) return;
    a = peer_exchange(i, a, P, buf, stamp);
    if (stamp_it) prof[4 * blockIdx.x + 2] = (long long)global_ns();
    grad[i] = a;
    opt_update<OPT>(i, a, o, st);
    if (stamp_it) prof[4 * blockIdx.x + 3] = (long long)global_ns();
}

// ----------------------------------------------------------------------------------------------------------------
// K5: optimisers (upstream optim rock semantics; call sites main.lua:135-144,301-302)
// ----------------------------------------------------------------------------------------------------------------
__global__ void k_adam(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                       float* __restrict__ wpack, int n, float b1, float omb1, float b2, float omb2, float eps, float step) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float mi = __fadd_rn(__fmul_rn(b1, m[i]), __fmul_rn(omb1, gi));
    const float vi = __fadd_rn(__fmul_rn(b2, v[i]), __fmul_rn(omb2, __fmul_rn(gi, gi)));
    m[i] = mi; v[i] = vi;
    const float xn = __fsub_rn(x[i], __fdiv_rn(__fmul_rn(step, mi), __fadd_rn(__fsqrt_rn(vi), eps)));
    x[i] = xn;
    wpack[pack_index(i)] = xn;
}

__global__ void k_rmsprop(float* __restrict__ x, const float* __restrict__ g, float* __restrict__ m,
                          float* __restrict__ wpack, int n, float alpha, float oma, float eps, float lr) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float gi = g[i];
    const float mi = __fadd_rn(__fmul_rn(alpha, m[i]), __fmul_rn(oma, __fmul_rn(gi, gi)));
    m[i] = mi;
    const float xn = __fsub_rn(x[i], __fdiv_rn(__fmul_rn(lr, gi), __fadd_rn(__fsqrt_rn(mi), eps)));
    x[i] = xn;
    wpack[pack_index(i)] = xn;
}

// ----------------------------------------------------------------------------------------------------------------
// Device-side priority sampler (priority_sampler.lua:4-54; the `-rs`-off training path): a table of the last loss seen per
// batch lives in HBM; a draw is uniform with probability alpha and otherwise multinomial over weight^pow.  One CTA:
// sharpened weights -> inclusive scan (double) -> `n` draws by binary search, counter-based random numbers (same seed,
// same draws).  flag[0] = 1 when a non-uniform draw met a table that is not full (a weight <= 0; the reference asserts).
// ----------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {      // splitmix64 finaliser
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ double u01(unsigned long long seed, unsigned long long ctr) {
    return (double)(mix64(seed ^ mix64(ctr)) >> 11) * (1.0 / 9007199254740992.0);
}
__global__ void __launch_bounds__(1024) k_priority_sample(const float* __restrict__ w, int nb, float pow_, float alpha,
                                                          unsigned long long seed, int n, double* __restrict__ cum,
                                                          int* __restrict__ ids_out, int* __restrict__ flag) {
    __shared__ double s_part[1024];
    __shared__ int s_bad;
    const int t = threadIdx.x;
    if (t == 0) s_bad = 0;
    __syncthreads();
    const int per = (nb + 1023) / 1024, lo = t * per, hi = min(nb, lo + per);
    double acc = 0.0;
    bool bad = false;
    for (int i = lo; i < hi; ++i) {
        const float wi = w[i];
        bad |= !(wi > 0.f);
        acc += (double)powf(wi, pow_);
    }
    if (bad) s_bad = 1;
    s_part[t] = acc;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {          // inclusive scan of the 1024 chunk sums
        const double v = t >= off ? s_part[t - off] : 0.0;
        __syncthreads();
        s_part[t] += v;
        __syncthreads();
    }
    double run = t ? s_part[t - 1] : 0.0;
    for (int i = lo; i < hi; ++i) { run += (double)powf(w[i], pow_); cum[i] = run; }
    __syncthreads();
    const double total = s_part[1023];
    for (int d = t; d < n; d += 1024) {
        int id;
        if (u01(seed, 2ull * d) < (double)alpha) {
            id = min(nb - 1, (int)(u01(seed, 2ull * d + 1) * nb));
        } else {
            if (s_bad) *flag = 1;
            const double x = u01(seed, 2ull * d + 1) * total;
            int a = 0, b = nb - 1;                       // first index with cum > x
            while (a < b) { const int mid = (a + b) >> 1; if (cum[mid] > x) b = mid; else a = mid + 1; }
            id = a;
        }
        ids_out[d] = id + 1;                             // 1-based like the Lua table
    }
}
// weight[ids[i] - 1] = value i; values from a host-provided array or from the per-step losses of npe_train_steps (loss4 rows)
__global__ void k_priority_update(float* __restrict__ w, int nb, const int* __restrict__ ids, const float* __restrict__ vals, int stride, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int id = ids[i];
    if (id >= 1 && id <= nb) w[id - 1] = vals[(size_t)i * stride];
}
__global__ void __launch_bounds__(1024) k_priority_hardest(const float* __restrict__ w, int nb, float* __restrict__ out2) {
    __shared__ float sv[1024];
    __shared__ int si[1024];
    float best = -1e38f; int bi = 0;
    for (int i = threadIdx.x; i < nb; i += 1024) if (w[i] > best) { best = w[i]; bi = i; }
    sv[threadIdx.x] = best; si[threadIdx.x] = bi;
    __syncthreads();
    for (int s = 512; s > 0; s >>= 1) {
        if (threadIdx.x < s) {
            const float o = sv[threadIdx.x + s]; const int oi = si[threadIdx.x + s];
            if (o > sv[threadIdx.x] || (o == sv[threadIdx.x] && oi < si[threadIdx.x])) { sv[threadIdx.x] = o; si[threadIdx.x] = oi; }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) { out2[0] = sv[0]; out2[1] = (float)(si[0] + 1); }
}

__global__ void k_fill(float* p, int n, float v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace npe
