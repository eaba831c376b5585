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
    const int* foc_scene;  // (F)
    const int* foc_obj;    // (F)
    const int* foc_t0;     // (F)
    int F;
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

struct WeightStage {
    unsigned long long bar;
};
// Thread 0 arms the barrier and issues the bulk copies; everyone waits with stage_wait() before the first use.
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
__device__ __forceinline__ void warp_pairs(const float* pos, int stride, int N, int n, int kmax, float thr_px,
                                           u64& keep, u64& nbr) {
    const int lane = threadIdx.x & 31;
    const int Cn = N - 1;
    const float fx = pos[(size_t)n * stride], fy = pos[(size_t)n * stride + 1];
    float d[2];
    bool valid[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int c = lane + 32 * h;
        valid[h] = c < Cn;
        d[h] = __int_as_float(0x7f800000);
        if (valid[h]) {
            const int j = c + (c >= n);
            const float dx = __fsub_rn(pos[(size_t)j * stride], fx);
            const float dy = __fsub_rn(pos[(size_t)j * stride + 1], fy);
            d[h] = __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
        }
    }
    const int k = (kmax > 0 && kmax < Cn) ? kmax : Cn;
    bool kp[2] = {valid[0], valid[1]};
    if (k < Cn) {
        int rank[2] = {0, 0};
        for (int cp = 0; cp < Cn; ++cp) {
            const float dp = __shfl_sync(0xffffffffu, cp < 32 ? d[0] : d[1], cp & 31);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int c = lane + 32 * h;
                rank[h] += (dp < d[h]) || (dp == d[h] && cp < c);
            }
        }
        kp[0] = valid[0] && rank[0] < k;
        kp[1] = valid[1] && rank[1] < k;
    }
    const unsigned k0 = __ballot_sync(0xffffffffu, kp[0]), k1 = __ballot_sync(0xffffffffu, kp[1]);
    const unsigned m0 = __ballot_sync(0xffffffffu, kp[0] && (__fmul_rn(d[0], 800.f) <= thr_px));
    const unsigned m1 = __ballot_sync(0xffffffffu, kp[1] && (__fmul_rn(d[1], 800.f) <= thr_px));
    keep = (u64)k0 | ((u64)k1 << 32);
    nbr = (u64)m0 | ((u64)m1 << 32);
}

__global__ void __launch_bounds__(NT) k_build_pairs(FocusSrc src, int kmax, float thr_px, u64* __restrict__ keep_out,
                                                    u64* __restrict__ nbr_out, u64* __restrict__ active_total) {
    const int f = blockIdx.x * (NT / 32) + (threadIdx.x >> 5);
    if (f >= src.F) return;
    const int s = src.foc_scene[f], n = src.foc_obj[f], t0 = src.foc_t0[f];
    u64 keep, nbr;
    warp_pairs(src.states + row_off(src, s, 0, t0 + 1), src.T * D, src.n_obj[s], n, kmax, thr_px, keep, nbr);
    if ((threadIdx.x & 31) == 0) {
        keep_out[f] = keep;
        nbr_out[f] = nbr;
        atomicAdd(active_total, (u64)__popcll(nbr));
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

// K7: relative targets y = frame(t0+2) with columns 0..5 minus frame(t0+1) (relative_pair, data_process.lua:87-96),
// plus the gathered focus rows for parity checks.
__global__ void k_window_targets(FocusSrc src, float* __restrict__ y, float* __restrict__ this_out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = idx / IN, c = idx - f * IN;
    if (f >= src.F) return;
    const float* p = src.states + row_off(src, src.foc_scene[f], src.foc_obj[f], src.foc_t0[f]);
    if (this_out) this_out[(size_t)f * IN + c] = p[c];
    if (y && c < D) {
        float v = p[2 * D + c];
        if (c < 6) v = __fsub_rn(v, p[D + c]);
        y[(size_t)f * D + c] = v;
    }
}

// ----------------------------------------------------------------------------------------------------------------
// shared helpers for the fused kernels
// ----------------------------------------------------------------------------------------------------------------
struct TileMeta {           // lives in shared memory
    long long foff[TILE];   // float offset of each focus row (frame t0) in `states`
    long long sbase[TILE];  // float offset of object 0, frame t0, of the focus's scene
    u64 nbr[TILE];          // active-context bits
    int fobj[TILE];         // focus object index within its scene
    int start[TILE + 1];    // exclusive scan of active-pair counts
    long long px[TILE];     // per pair row: float offset of the context row
    int pf[TILE];           // per pair row: local focus index
};

// Fill per-row metadata of pair sub-tile [lo, lo + 64).
__device__ __forceinline__ void fill_pair_meta(TileMeta& M, int nf, int lo, int objstride) {
    const int f = threadIdx.x;
    if (f < nf) {
        u64 bits = M.nbr[f];
        int pos = M.start[f] - lo;
        while (bits) {
            const int c = __ffsll((long long)bits) - 1;
            bits &= bits - 1;
            if (pos >= 0 && pos < TILE) {
                const int j = c + (c >= M.fobj[f]);
                M.pf[pos] = f;
                M.px[pos] = M.sbase[f] + (long long)j * objstride;
            }
            ++pos;
        }
    }
}

// Gather rows of 44 floats (two frames) into dst[r][LDD]; rows >= nrows are zero-filled.  Row offsets are multiples
// of 22 floats (8-byte aligned): float2 loads.
template <int LDD>
__device__ __forceinline__ void gather_rows(float* dst, const float* __restrict__ base, const long long* offs, int nrows,
                                            int nrows_total) {
    for (int idx = threadIdx.x; idx < nrows_total * (IN / 2); idx += NT) {
        const int r = idx / (IN / 2), c2 = idx - r * (IN / 2);
        float2 v = make_float2(0.f, 0.f);
        if (r < nrows) v = *reinterpret_cast<const float2*>(base + offs[r] + 2 * c2);
        *reinterpret_cast<float2*>(dst + r * LDD + 2 * c2) = v;
    }
}

__device__ __forceinline__ float relu(float v) { return fmaxf(v, 0.f); }

// h0 .. h5 of one pair sub-tile.  X rows are already gathered.  bufs[l] receives h_l (l = 0..5); bufs may alias in
// ping-pong fashion for the forward pass (bufs[l] != bufs[l-1]).
__device__ __forceinline__ void pair_chain_forward(const float* sWp, const float* sX, const float* sEf, int ldef,
                                                   const int* pf, int nrows, float* const* bufs) {
    float* h0 = bufs[0];
    // focus half of the encoder output is shared by every pair of the focus; pad columns 50..55 are zero
    for (int idx = threadIdx.x; idx < TILE * 32; idx += NT) {
        const int r = idx >> 5, n = idx & 31;
        if (n < H2) h0[r * LD + n] = (r < nrows) ? sEf[pf[r] * ldef + n] : 0.f;
        else if (n >= 26) h0[r * LD + (n - 26) + H] = 0.f;   // n = 26..31 -> cols 50..55
    }
    gemm_nt<7, LD, LD, 2>(sX, sWp + SW_ENC_C, 11, EpiNT{h0 + H2, LD, H2, 1});
    __syncthreads();
#pragma unroll 1
    for (int l = 0; l < NL; ++l) {
        gemm_nt<13, LD, LD, 2>(bufs[l], sWp + SW_PAIR + l * SW_PAIR_SZ, 13, EpiNT{bufs[l + 1], LD, 52, 1});
        __syncthreads();
    }
}

// u1..u4 and raw output o of the decoder on a 64-row focus tile; z = sZ [64][104].  u[i] receives u_{i+1}.
__device__ __forceinline__ void decoder_forward(const float* sWd, const float* sZ, float* const* u, float* o) {
    gemm_nt<13, LD1, LD1, 2>(sZ, sWd + SW_DEC1, 24, EpiNT{u[0], LD, 52, 1});
    __syncthreads();
#pragma unroll 1
    for (int i = 0; i < 3; ++i) {
        gemm_nt<13, LD, LD, 2>(u[i], sWd + SW_DEC2 + i * SW_PAIR_SZ, 13, EpiNT{u[i + 1], LD, 52, 1});
        __syncthreads();
    }
    gemm_nt<6, LD, LD, 2>(u[3], sWd + SW_DEC5, 13, EpiNT{o, LD, 24, 0});
    __syncthreads();
}

// Load a focus tile: z[r] = [0 x 50 | 1 | 0 | this_past(44) | 0 x 8]; rows >= nf are all zero.
__device__ __forceinline__ void load_focus_tile(float* sZ, const float* __restrict__ states, const long long* foff, int nf) {
    for (int idx = threadIdx.x; idx < TILE * (ZCOL_F / 2); idx += NT) {
        const int r = idx / (ZCOL_F / 2), c2 = idx - r * (ZCOL_F / 2);
        float2 v = make_float2(0.f, 0.f);
        if (c2 == ONE_COL / 2 && r < nf) v.x = 1.f;
        *reinterpret_cast<float2*>(sZ + r * LD1 + 2 * c2) = v;
    }
    for (int idx = threadIdx.x; idx < TILE * (LD1 - ZCOL_F) / 2; idx += NT) {
        const int r = idx / ((LD1 - ZCOL_F) / 2), c2 = idx - r * ((LD1 - ZCOL_F) / 2);
        float2 v = make_float2(0.f, 0.f);
        if (r < nf && c2 < IN / 2) v = *reinterpret_cast<const float2*>(states + foff[r] + 2 * c2);
        *reinterpret_cast<float2*>(sZ + r * LD1 + ZCOL_F + 2 * c2) = v;
    }
}

__device__ __forceinline__ void scan_counts(TileMeta& M, int nf) {
    if (threadIdx.x == 0) {
        int acc = 0;
        for (int f = 0; f < TILE; ++f) {
            M.start[f] = acc;
            acc += (f < nf) ? __popcll(M.nbr[f]) : 0;
        }
        M.start[TILE] = acc;
    }
}

// s[f][0..49] += sum of h5 rows of focus f inside sub-tile [lo, lo+64): fixed order -> deterministic.
template <int LDS_>
__device__ __forceinline__ void segmented_add(float* sS, const float* h5, const TileMeta& M, int nfoc, int lo) {
    for (int idx = threadIdx.x; idx < nfoc * H; idx += NT) {
        const int f = idx / H, n = idx - f * H;
        const int r0 = max(M.start[f], lo) - lo, r1 = min(M.start[f + 1], lo + TILE) - lo;
        if (r1 > r0) {
            float acc = sS[f * LDS_ + n];
            for (int r = r0; r < r1; ++r) acc += h5[r * LD + n];
            sS[f * LDS_ + n] = acc;
        }
    }
}

// ----------------------------------------------------------------------------------------------------------------
// K2 + K3: fused forward
// ----------------------------------------------------------------------------------------------------------------
struct FwdSmem {
    float Wp[SW_PAIR_TOTAL];
    float Wd[SW_DEC_TOTAL];
    float Z[TILE * LD1];
    float Ef[TILE * 28];
    float X[TILE * LD];
    float P0[TILE * LD];
    float P1[TILE * LD];
    TileMeta M;
    WeightStage W;
};

__global__ void __launch_bounds__(NT, 1)
k_forward(FocusSrc src, const u64* __restrict__ nbr, const float* __restrict__ wpack, const float* __restrict__ y,
          float* __restrict__ pred, float* __restrict__ loss_ex, float* __restrict__ s_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FwdSmem& S = *reinterpret_cast<FwdSmem*>(smem_raw);
    stage_issue(S.W, S.Wp, S.Wd, wpack);
    bool staged = false;
    const int ntiles = (src.F + TILE - 1) / TILE;
    const int objstride = src.T * D;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int f0 = tile * TILE;
        const int nf = min(TILE, src.F - f0);
        __syncthreads();
        if (threadIdx.x < TILE) {
            const int r = threadIdx.x;
            if (r < nf) {
                const int s = src.foc_scene[f0 + r], o = src.foc_obj[f0 + r], t0 = src.foc_t0[f0 + r];
                S.M.foff[r] = (long long)row_off(src, s, o, t0);
                S.M.sbase[r] = (long long)row_off(src, s, 0, t0);
                S.M.fobj[r] = o;
                S.M.nbr[r] = nbr[f0 + r];
            } else {
                S.M.nbr[r] = 0;
            }
        }
        __syncthreads();
        load_focus_tile(S.Z, src.states, S.M.foff, nf);
        scan_counts(S.M, nf);
        __syncthreads();
        if (!staged) { stage_wait(S.W); staged = true; }
        gemm_nt<7, LD1, LD, 2>(S.Z + ZCOL_F, S.Wp + SW_ENC_T, 11, EpiNT{S.Ef, 28, 28, 1});
        __syncthreads();
        const int P = S.M.start[TILE];
        float* const bufs[NL + 1] = {S.P0, S.P1, S.P0, S.P1, S.P0, S.P1};
        for (int lo = 0; lo < P; lo += TILE) {
            const int nrows = min(TILE, P - lo);
            fill_pair_meta(S.M, nf, lo, objstride);
            __syncthreads();
            gather_rows<LD>(S.X, src.states, S.M.px, nrows, TILE);
            __syncthreads();
            pair_chain_forward(S.Wp, S.X, S.Ef, 28, S.M.pf, nrows, bufs);
            segmented_add<LD1>(S.Z, S.P1, S.M, nf, lo);
            __syncthreads();
        }
        if (s_out) {
            for (int idx = threadIdx.x; idx < nf * 52; idx += NT) {
                const int r = idx / 52, n = idx - r * 52;
                s_out[(size_t)(f0 + r) * 52 + n] = (n < H) ? S.Z[r * LD1 + n] : 0.f;
            }
        }
        // decoder: u1 -> P0, u2 -> P1, u3 -> P0, u4 -> P1, o -> X
        float* const ubufs[4] = {S.P0, S.P1, S.P0, S.P1};
        decoder_forward(S.Wd, S.Z, ubufs, S.X);
        for (int idx = threadIdx.x; idx < nf * OUT; idx += NT) {
            const int r = idx / OUT, n = idx - r * OUT;
            float v = S.X[r * LD + n];
            if (n >= 6) v = 1.f / (1.f + expf(-v));   // Sigmoid on the property columns (modules.lua:80-86)
            pred[(size_t)(f0 + r) * OUT + n] = v;
        }
        if (loss_ex && y && threadIdx.x < nf) {
            const int r = threadIdx.x;
            const float* yy = y + (size_t)(f0 + r) * OUT;
            const float e2 = S.X[r * LD + 2] - yy[2], e3 = S.X[r * LD + 3] - yy[3], e5 = S.X[r * LD + 5] - yy[5];
            const float lv = __fadd_rn(__fmul_rn(e2, e2), __fmul_rn(e3, e3)), lw = __fmul_rn(e5, e5);
            loss_ex[(size_t)(f0 + r) * 3 + 0] = __fdiv_rn(__fadd_rn(lv, lw), 3.f);   // npe.lua:271-272
            loss_ex[(size_t)(f0 + r) * 3 + 1] = __fdiv_rn(lv, 2.f);                  // npe.lua:273
            loss_ex[(size_t)(f0 + r) * 3 + 2] = lw;                                  // npe.lua:274
        }
    }
}

// Batch losses from the per-example ones (npe.lua:238-246): {sum/(3B), vel/(2B), angvel/B}; double accumulation like
// THNN's accreal.  Also leaves the three raw sums in sums_out (for the data-parallel all-reduce).
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

// ----------------------------------------------------------------------------------------------------------------
// K4a: decoder backward.  d_pred = 2 (p - y) vlambda / (2B) on cols 2,3; 2 (p - y) lambda / B on col 5; zero
// elsewhere (npe.lua:301-320).  Weight-gradient tiles live in registers for the whole kernel.
// ----------------------------------------------------------------------------------------------------------------
struct DecBwdSmem {
    float Wd[SW_DEC_TOTAL];
    float Z[TILE * LD1];
    float U[4][TILE * LD];
    float O[TILE * LD];
    float Da[TILE * LD];
    float Db[TILE * LD];
    long long foff[TILE];
    WeightStage W;
};

__device__ __forceinline__ float4 relu_mask4(float4 v, const float* h) {
    const float4 hv = ld4(h);
    return make_float4(hv.x > 0.f ? v.x : 0.f, hv.y > 0.f ? v.y : 0.f, hv.z > 0.f ? v.z : 0.f, hv.w > 0.f ? v.w : 0.f);
}

__global__ void __launch_bounds__(NT, 1)
k_dec_backward(FocusSrc src, const float* __restrict__ wpack, const float* __restrict__ y,
               const float* __restrict__ s_in, float* __restrict__ ds_out, float* __restrict__ partial,
               float scale_v, float scale_w) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    DecBwdSmem& S = *reinterpret_cast<DecBwdSmem*>(smem_raw);
    stage_issue(S.W, nullptr, S.Wd, wpack);
    bool staged = false;
    float* const ubufs[4] = {S.U[0], S.U[1], S.U[2], S.U[3]};
    float a1a[16], a1b[16], a2[16], a3[16], a4[16], a5[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { a1a[i] = a1b[i] = a2[i] = a3[i] = a4[i] = a5[i] = 0.f; }
    const int t = threadIdx.x;
    const int ntiles = (src.F + TILE - 1) / TILE;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int f0 = tile * TILE;
        const int nf = min(TILE, src.F - f0);
        __syncthreads();
        if (t < nf) S.foff[t] = (long long)row_off(src, src.foc_scene[f0 + t], src.foc_obj[f0 + t], src.foc_t0[f0 + t]);
        __syncthreads();
        load_focus_tile(S.Z, src.states, S.foff, nf);
        __syncthreads();
        for (int idx = t; idx < nf * H; idx += NT) {
            const int r = idx / H, n = idx - r * H;
            S.Z[r * LD1 + n] = s_in[(size_t)(f0 + r) * 52 + n];
        }
        for (int idx = t; idx < TILE * LD; idx += NT) S.Da[idx] = 0.f;
        __syncthreads();
        if (!staged) { stage_wait(S.W); staged = true; }
        decoder_forward(S.Wd, S.Z, ubufs, S.O);
        if (t < nf) {
            const float* yy = y + (size_t)(f0 + t) * OUT;
            S.Da[t * LD + 2] = (S.O[t * LD + 2] - yy[2]) * scale_v;
            S.Da[t * LD + 3] = (S.O[t * LD + 3] - yy[3]) * scale_v;
            S.Da[t * LD + 5] = (S.O[t * LD + 5] - yy[5]) * scale_w;
        }
        __syncthreads();
        // layer 5 (22 x 50): dz = Da (24 cols), input u4
        if (t >= 169 && t < 169 + 78) wgrad<13, LD, LD, true>(a5, S.Da, S.U[3], t - 169, nf);
        gemm_nn<13, LD, LD>(S.Da, S.Wd + SW_DEC5, 6, EpiNN{S.Db, LD, S.U[3], TILE});
        __syncthreads();
        // layer 4: dz = Db, input u3
        if (t < 169) wgrad<13, LD, LD, true>(a4, S.Db, S.U[2], t, nf);
        gemm_nn<13, LD, LD>(S.Db, S.Wd + SW_DEC2 + 2 * SW_PAIR_SZ, 13, EpiNN{S.Da, LD, S.U[2], TILE});
        __syncthreads();
        // layer 3: dz = Da, input u2
        if (t < 169) wgrad<13, LD, LD, true>(a3, S.Da, S.U[1], t, nf);
        gemm_nn<13, LD, LD>(S.Da, S.Wd + SW_DEC2 + SW_PAIR_SZ, 13, EpiNN{S.Db, LD, S.U[1], TILE});
        __syncthreads();
        // layer 2: dz = Db, input u1
        if (t < 169) wgrad<13, LD, LD, true>(a2, S.Db, S.U[0], t, nf);
        gemm_nn<13, LD, LD>(S.Db, S.Wd + SW_DEC2, 13, EpiNN{S.Da, LD, S.U[0], TILE});
        __syncthreads();
        // layer 1 (50 x 94 (+bias)): dz = Da, input z (24 k-chunks); input grad only for the pair-sum columns
        wgrad<24, LD, LD1, true>(a1a, S.Da, S.Z, t, nf);
        if (t + NT < 13 * 24) wgrad<24, LD, LD1, true>(a1b, S.Da, S.Z, t + NT, nf);
        gemm_nn<13, LD, LD1>(S.Da, S.Wd + SW_DEC1, 13, EpiNN{ds_out + (size_t)f0 * 52, 52, nullptr, nf});
    }
    // write the register tiles into this CTA's partial (reference flat layout)
    float* out = partial + (size_t)blockIdx.x * NPARAM;
    auto put1 = [&](const float (&acc)[16], int tileid) {
        const int a = tileid / 24, b = tileid - a * 24;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = 4 * a + i, zc = 4 * b + j;
                if (n < H) {
                    if (zc < H) out[OFF_DEC1_W + n * DEC_IN + zc] = acc[4 * i + j];
                    else if (zc == ONE_COL) out[OFF_DEC1_B + n] = acc[4 * i + j];
                    else if (zc >= ZCOL_F) out[OFF_DEC1_W + n * DEC_IN + zc - 2] = acc[4 * i + j];
                }
            }
    };
    put1(a1a, t);
    if (t + NT < 13 * 24) put1(a1b, t + NT);
    auto putm = [&](const float (&acc)[16], int layer) {   // layer = 2..4
        const int a = t / 13, b = t - a * 13;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = 4 * a + i, k = 4 * b + j;
                if (n < H) {
                    if (k < H) out[off_dec_w(layer) + n * H + k] = acc[4 * i + j];
                    else if (k == ONE_COL) out[off_dec_b(layer) + n] = acc[4 * i + j];
                }
            }
    };
    if (t < 169) { putm(a2, 2); putm(a3, 3); putm(a4, 4); }
    if (t >= 169 && t < 169 + 78) {
        const int e = t - 169, a = e / 13, b = e - a * 13;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = 4 * a + i, k = 4 * b + j;
                if (n < OUT) {
                    if (k < H) out[OFF_DEC5_W + n * H + k] = a5[4 * i + j];
                    else if (k == ONE_COL) out[OFF_DEC5_B + n] = a5[4 * i + j];
                }
            }
    }
}

// ----------------------------------------------------------------------------------------------------------------
// K4b: pair-MLP backward (encoders + 5 pairwise layers).  32 foci per tile, 64 pair rows per sub-tile; forward is
// recomputed in the tile (no activation round trip through HBM); masked pairs never enter (apply_mask on the
// CAddTable gradient, npe.lua:326-328, is the identity on the active list).
// ----------------------------------------------------------------------------------------------------------------
constexpr int FT_B = 32;
struct PairBwdSmem {
    float Wp[SW_PAIR_TOTAL];
    float X[TILE * LD];
    float Hh[NL + 1][TILE * LD];
    float Da[TILE * LD];
    float Db[TILE * LD];
    float Fr[FT_B * LD];     // focus rows (44 valid)
    float Ds[FT_B * LD];     // gradient into the pair sum
    float Ef[FT_B * 28];     // focus half of the encoder output
    float De[FT_B * 28];     // gradient into it (summed over the focus's pairs)
    TileMeta M;
    WeightStage W;
};

__global__ void __launch_bounds__(NT, 1)
k_pair_backward(FocusSrc src, const u64* __restrict__ nbr, const float* __restrict__ wpack,
                const float* __restrict__ ds_in, float* __restrict__ partial) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PairBwdSmem& S = *reinterpret_cast<PairBwdSmem*>(smem_raw);
    stage_issue(S.W, S.Wp, nullptr, wpack);
    bool staged = false;
    float* const bufs[NL + 1] = {S.Hh[0], S.Hh[1], S.Hh[2], S.Hh[3], S.Hh[4], S.Hh[5]};
    float acc[NL][16];
#pragma unroll
    for (int l = 0; l < NL; ++l)
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[l][i] = 0.f;
    const int t = threadIdx.x;
    const bool enc_thread = (t >= 169 && t < 169 + 77);
    const int ntiles = (src.F + FT_B - 1) / FT_B;
    const int objstride = src.T * D;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int f0 = tile * FT_B;
        const int nf = min(FT_B, src.F - f0);
        __syncthreads();
        if (t < TILE) {
            if (t < nf) {
                const int s = src.foc_scene[f0 + t], o = src.foc_obj[f0 + t], t0 = src.foc_t0[f0 + t];
                S.M.foff[t] = (long long)row_off(src, s, o, t0);
                S.M.sbase[t] = (long long)row_off(src, s, 0, t0);
                S.M.fobj[t] = o;
                S.M.nbr[t] = nbr[f0 + t];
            } else {
                S.M.nbr[t] = 0;
            }
        }
        __syncthreads();
        gather_rows<LD>(S.Fr, src.states, S.M.foff, nf, FT_B);
        for (int idx = t; idx < FT_B * LD; idx += NT) {
            const int r = idx / LD, n = idx - r * LD;
            S.Ds[idx] = (r < nf && n < H) ? ds_in[(size_t)(f0 + r) * 52 + n] : 0.f;
        }
        for (int idx = t; idx < FT_B * 28; idx += NT) S.De[idx] = 0.f;
        scan_counts(S.M, nf);
        __syncthreads();
        if (!staged) { stage_wait(S.W); staged = true; }
        gemm_nt<7, LD, LD, 1>(S.Fr, S.Wp + SW_ENC_T, 11, EpiNT{S.Ef, 28, 28, 1});
        __syncthreads();
        const int P = S.M.start[TILE];
        for (int lo = 0; lo < P; lo += TILE) {
            const int nrows = min(TILE, P - lo);
            fill_pair_meta(S.M, nf, lo, objstride);
            __syncthreads();
            gather_rows<LD>(S.X, src.states, S.M.px, nrows, TILE);
            __syncthreads();
            pair_chain_forward(S.Wp, S.X, S.Ef, 28, S.M.pf, nrows, bufs);
            // dz5 = ds[focus] * (h5 > 0); rows >= nrows zero
            for (int idx = t; idx < TILE * (LD / 4); idx += NT) {
                const int r = idx / (LD / 4), c4 = idx - r * (LD / 4);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < nrows && c4 < 13) v = relu_mask4(ld4(S.Ds + S.M.pf[r] * LD + 4 * c4), S.Hh[NL] + r * LD + 4 * c4);
                st4(S.Da + r * LD + 4 * c4, v);
            }
            __syncthreads();
            float* da = S.Da;
            float* db = S.Db;
#pragma unroll
            for (int l = NL - 1; l >= 0; --l) {
                const float* hin = S.Hh[l];
                if (t < 169) wgrad<13, LD, LD, true>(acc[l], da, hin, t, nrows);
                gemm_nn<13, LD, LD>(da, S.Wp + SW_PAIR + l * SW_PAIR_SZ, 13, EpiNN{db, LD, hin, TILE});
                __syncthreads();
                float* tmp = da; da = db; db = tmp;
            }
            // da = dz0: cols 0..24 focus-encoder part, 25..49 context-encoder part
            if (enc_thread) wgrad<11, LD, LD, false>(acc[0], da, S.X, t - 169, nrows, H2);
            for (int idx = t; idx < nf * H2; idx += NT) {
                const int f = idx / H2, n = idx - f * H2;
                const int r0 = max(S.M.start[f], lo) - lo, r1 = min(S.M.start[f + 1], lo + TILE) - lo;
                if (r1 > r0) {
                    float a = S.De[f * 28 + n];
                    for (int r = r0; r < r1; ++r) a += da[r * LD + n];
                    S.De[f * 28 + n] = a;
                }
            }
            __syncthreads();
        }
        if (enc_thread) wgrad<11, 28, LD, true>(acc[1], S.De, S.Fr, t - 169, nf);
    }
    float* out = partial + (size_t)blockIdx.x * NPARAM;
    if (t < 169) {
        const int a = t / 13, b = t - a * 13;
#pragma unroll
        for (int l = 0; l < NL; ++l)
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const int n = 4 * a + i, k = 4 * b + j;
                    if (n < H && k < H) out[OFF_PAIR + l * 2500 + n * H + k] = acc[l][4 * i + j];
                }
    } else if (enc_thread) {
        const int e = t - 169, a = e / 11, b = e - a * 11;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = 4 * a + i, k = 4 * b + j;
                if (n < H2) {
                    out[OFF_ENC_C + n * IN + k] = acc[0][4 * i + j];
                    out[OFF_ENC_T + n * IN + k] = acc[1][4 * i + j];
                }
            }
    }
}

// grad[i] = sum over CTA partials in CTA order (deterministic); pair part and decoder part have their own grids.
__global__ void k_reduce_grads(const float* __restrict__ partial, int nparts_pair, int nparts_dec, float* __restrict__ grad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NPARAM) return;
    const int np = (i < NPARAM_PAIR) ? nparts_pair : nparts_dec;
    float a = 0.f;
    for (int b = 0; b < np; ++b) a += partial[(size_t)b * NPARAM + i];
    grad[i] = a;
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

__global__ void k_fill(float* p, int n, float v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace npe
