// CTA-cooperative FP32 tile micro-kernels over shared memory (256 threads, 64-row tiles).
//
//   gemm_nt : C[r][n]  = act(sum_k A[r][k] * W[n][k])        (forward;    both operands k-contiguous)
//   gemm_nn : dA[r][k] = mask(sum_n dZ[r][n] * W[n][k])      (input grad; W rows contiguous in the OUTPUT index)
//   wgrad   : dW[n][k] += sum_r dZ[r][n] * Hm[r][k]          (weight grad; per-thread 4x4 register tile, persistent)
//
// Thread mapping of the two GEMMs (tid = threadIdx.x):  ks = tid & 1 (2-way split of the reduction index, adjacent
// lanes), cg = (tid >> 1) & 7 (output-column group), rg = tid >> 4 (row group, 0..15).  A thread owns the FOUR rows
// {rg + 16 j} and output columns {cg + 8 i}; the two ks lanes exchange partial sums with one shuffle per output and
// lane ks finalises rows j in {2 ks, 2 ks + 1}.
//
// Why 4 rows per thread: the shared-memory crossbar delivers 128 B/clk/SM to lanes and an LDS.128 costs 4 wavefronts
// per warp whether or not lanes broadcast (measured: 3.4 wavefronts per LDS.128, tools/microbench/gemm_pass), so the
// loaded bytes per FMA decide the speed: 64x52x52 costs 5408 * (1/RT + 1/CT) LDS cycles for an RT x CT register tile
// (2x13: 3120, measured 3881 cycles/pass; 4x7: 2124) against 1352 FMA cycles.
// Why __noinline__ + runtime epilogue: one fully unrolled GEMM per layer made 190-220 KB of SASS per kernel and the
// top stall was instruction fetch (ncu: no_instruction 3.3 per issue at 2 warps/scheduler).
#pragma once
#include "npe_layout.h"

namespace npe {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float2 ld2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void axpy4(float4& acc, float a, float4 w) {
    acc.x = fmaf(a, w.x, acc.x); acc.y = fmaf(a, w.y, acc.y); acc.z = fmaf(a, w.z, acc.z); acc.w = fmaf(a, w.w, acc.w);
}

// Epilogue of gemm_nt: out[row * ldo + n] = relu ? max(v, 0) : v   for n < nmax.
struct EpiNT {
    float* out;
    int ldo;
    int nmax;
    int relu;
};

// C[rg + 16 j][cg + 8 i] for j < 4, i < NI; reduction over kc chunks of 4.  Rows >= 16 * jmax are skipped.
template <int NI, int LDA, int LDW>
__device__ __noinline__ void gemm_nt(const float* __restrict__ A, const float* __restrict__ W, int kc, int nrows, EpiNT e) {
    const int tid = threadIdx.x;
    const int ks = tid & 1, cg = (tid >> 1) & 7, rg = tid >> 4;
    const int jmax = (nrows + 15) >> 4;
    float acc[4][NI];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < NI; ++i) acc[j][i] = 0.f;
    const float* a = A + rg * LDA + 4 * ks;
    const float* w = W + cg * LDW + 4 * ks;
#pragma unroll 1
    for (int c = ks; c < kc; c += 2) {
        float4 x[4], wv[NI];
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = (j < jmax) ? ld4(a + 16 * j * LDA) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int i = 0; i < NI; ++i) wv[i] = ld4(w + 8 * i * LDW);
        // component-major: consecutive FMAs hit different accumulators (dependency distance 4 NI)
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j][i] = fmaf(x[j].x, wv[i].x, acc[j][i]);
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j][i] = fmaf(x[j].y, wv[i].y, acc[j][i]);
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j][i] = fmaf(x[j].z, wv[i].z, acc[j][i]);
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j][i] = fmaf(x[j].w, wv[i].w, acc[j][i]);
        a += 8; w += 8;
    }
    __syncwarp();
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
        float* o = e.out + (rg + 16 * (2 * ks + jj)) * e.ldo + cg;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const float send = ks ? acc[jj][i] : acc[2 + jj][i];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
            float mine = (ks ? acc[2 + jj][i] : acc[jj][i]) + recv;
            if (e.relu) mine = fmaxf(mine, 0.f);
            if (cg + 8 * i < e.nmax) o[8 * i] = mine;
        }
    }
}

// Epilogue of gemm_nn: v (float4 at output chunk kc of `row`) is optionally masked by hmask[row][4 kc..] > 0 (ReLU
// derivative) and stored to out[row * ldo + 4 kc] for row < nrows.  `out` may be shared or global memory.
struct EpiNN {
    float* out;
    int ldo;
    const float* hmask;   // row stride LD, or nullptr
    int nrows;
};

// dA[rg + 16 j][4 kc .. 4 kc + 3] for kc = cg + 8 m < KCO; reduction over nc chunks of 4 n (ks lane takes
// n = 4 q + 2 ks + {0, 1}).
template <int KCO, int LDZ, int LDW>
__device__ __noinline__ void gemm_nn(const float* __restrict__ dZ, const float* __restrict__ W, int nc, int nrows, EpiNN e) {
    constexpr int MI = (KCO + 7) / 8;
    const int tid = threadIdx.x;
    const int ks = tid & 1, cg = (tid >> 1) & 7, rg = tid >> 4;
    const int jmax = (nrows + 15) >> 4;
    float4 acc[4][MI];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int m = 0; m < MI; ++m) acc[j][m] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* zp = dZ + rg * LDZ + 2 * ks;
    const float* wa = W + (2 * ks) * LDW + 4 * cg;
#pragma unroll 1
    for (int q = 0; q < nc; ++q) {
        float2 z[4];
        float4 w0[MI], w1[MI];
#pragma unroll
        for (int j = 0; j < 4; ++j) z[j] = (j < jmax) ? ld2(zp + 16 * j * LDZ) : make_float2(0.f, 0.f);
#pragma unroll
        for (int m = 0; m < MI; ++m) {
            if (cg + 8 * m < KCO) { w0[m] = ld4(wa + 32 * m); w1[m] = ld4(wa + LDW + 32 * m); }
            else { w0[m] = make_float4(0.f, 0.f, 0.f, 0.f); w1[m] = w0[m]; }
        }
#pragma unroll
        for (int m = 0; m < MI; ++m)
#pragma unroll
            for (int j = 0; j < 4; ++j) axpy4(acc[j][m], z[j].x, w0[m]);
#pragma unroll
        for (int m = 0; m < MI; ++m)
#pragma unroll
            for (int j = 0; j < 4; ++j) axpy4(acc[j][m], z[j].y, w1[m]);
        zp += 4; wa += 4 * LDW;
    }
    __syncwarp();
#pragma unroll
    for (int jj = 0; jj < 2; ++jj) {
        const int row = rg + 16 * (2 * ks + jj);
#pragma unroll
        for (int m = 0; m < MI; ++m) {
            const float4 s = ks ? acc[jj][m] : acc[2 + jj][m];
            float4 r;
            r.x = __shfl_xor_sync(0xffffffffu, s.x, 1); r.y = __shfl_xor_sync(0xffffffffu, s.y, 1);
            r.z = __shfl_xor_sync(0xffffffffu, s.z, 1); r.w = __shfl_xor_sync(0xffffffffu, s.w, 1);
            const float4 o = ks ? acc[2 + jj][m] : acc[jj][m];
            const int kc = cg + 8 * m;
            if (kc < KCO && row < e.nrows) {
                float4 v = make_float4(o.x + r.x, o.y + r.y, o.z + r.z, o.w + r.w);
                if (e.hmask) {
                    const float4 hv = ld4(e.hmask + row * LD + 4 * kc);
                    v.x = hv.x > 0.f ? v.x : 0.f; v.y = hv.y > 0.f ? v.y : 0.f;
                    v.z = hv.z > 0.f ? v.z : 0.f; v.w = hv.w > 0.f ? v.w : 0.f;
                }
                st4(e.out + (size_t)row * e.ldo + 4 * kc, v);
            }
        }
    }
}

// ---- 16-row variants (latency path: small batches are split over many CTAs) ------------------------------------
// rg = tid >> 6 (0..3) owns rows {rg + 4 j}; cg = (tid >> 3) & 7; the reduction index is split EIGHT ways over
// adjacent lanes (ks = tid & 7) so that a 16-row pass still occupies all 256 threads with the same loaded-bytes-per-FMA
// ratio as the 64-row mapping; the eight partial sums are combined with a 3-round butterfly that also scatters
// ownership (rows, then columns), ~25 shuffles per thread.
template <int NI, int LDA, int LDW>
__device__ __noinline__ void gemm_nt16(const float* __restrict__ A, const float* __restrict__ W, int kc, int nrows, EpiNT e) {
    const int tid = threadIdx.x;
    const int ks = tid & 7, cg = (tid >> 3) & 7, rg = tid >> 6;
    float acc[4][NI];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int i = 0; i < NI; ++i) acc[j][i] = 0.f;
    const float* a = A + rg * LDA + 4 * ks;
    const float* w = W + cg * LDW + 4 * ks;
#pragma unroll 1
    for (int c = ks; c < kc; c += 8) {
        float4 x[4], wv[NI];
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = ld4(a + 4 * j * LDA);
#pragma unroll
        for (int i = 0; i < NI; ++i) wv[i] = ld4(w + 8 * i * LDW);
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j][i] = fmaf(x[j].x, wv[i].x, acc[j][i]);
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j][i] = fmaf(x[j].y, wv[i].y, acc[j][i]);
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j][i] = fmaf(x[j].z, wv[i].z, acc[j][i]);
#pragma unroll
        for (int i = 0; i < NI; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[j][i] = fmaf(x[j].w, wv[i].w, acc[j][i]);
        a += 32; w += 32;
    }
    __syncwarp();
    // round 1 (lane ^ 4): keep rows {0,1} (bit clear) or {2,3} (bit set)
    const int b2 = (ks >> 2) & 1, b1 = (ks >> 1) & 1, b0 = ks & 1;
    float r1[2][NI];
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const float send = b2 ? acc[jj][i] : acc[2 + jj][i];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 4);
            r1[jj][i] = (b2 ? acc[2 + jj][i] : acc[jj][i]) + recv;
        }
    // round 2 (lane ^ 2): keep one of the two rows
    float r2[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        const float send = b1 ? r1[0][i] : r1[1][i];
        const float recv = __shfl_xor_sync(0xffffffffu, send, 2);
        r2[i] = (b1 ? r1[1][i] : r1[0][i]) + recv;
    }
    // round 3 (lane ^ 1): keep columns i with (i & 1) == b0
    const int row = rg + 4 * (2 * b2 + b1);
    float* o = e.out + row * e.ldo + cg;
#pragma unroll
    for (int i2 = 0; i2 < (NI + 1) / 2; ++i2) {
        const int ie = 2 * i2, io = 2 * i2 + 1;           // even / odd column slots of this pair
        const float even = r2[ie], odd = (io < NI) ? r2[io] : 0.f;
        const float send = b0 ? even : odd;
        const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
        float mine = (b0 ? odd : even) + recv;
        const int i = b0 ? io : ie;
        if (e.relu) mine = fmaxf(mine, 0.f);
        if (i < NI && row < 16 && cg + 8 * i < e.nmax) o[8 * i] = mine;
    }
    (void)nrows;
}

// dA[rg + 4 j][4 kc..] for kc = cg + 8 m < KCO (m < 2); lane ks takes n = 8 q + ks.
template <int KCO, int LDZ, int LDW>
__device__ __noinline__ void gemm_nn16(const float* __restrict__ dZ, const float* __restrict__ W, int nq, int nrows, EpiNN e) {
    constexpr int MI = (KCO + 7) / 8;
    static_assert(MI == 2, "gemm_nn16 assumes two output chunks per thread");
    const int tid = threadIdx.x;
    const int ks = tid & 7, cg = (tid >> 3) & 7, rg = tid >> 6;
    float4 acc[4][MI];
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int m = 0; m < MI; ++m) acc[j][m] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* zp = dZ + rg * LDZ + ks;
    const float* wa = W + ks * LDW + 4 * cg;
#pragma unroll 1
    for (int q = 0; q < nq; ++q) {
        float z[4];
        float4 w0[MI];
#pragma unroll
        for (int j = 0; j < 4; ++j) z[j] = zp[4 * j * LDZ];
#pragma unroll
        for (int m = 0; m < MI; ++m) w0[m] = (cg + 8 * m < KCO) ? ld4(wa + 32 * m) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int m = 0; m < MI; ++m)
#pragma unroll
            for (int j = 0; j < 4; ++j) axpy4(acc[j][m], z[j], w0[m]);
        zp += 8; wa += 8 * LDW;
    }
    __syncwarp();
    const int b2 = (ks >> 2) & 1, b1 = (ks >> 1) & 1, b0 = ks & 1;
    auto xchg = [](float4 send, int lane_mask) {
        float4 r;
        r.x = __shfl_xor_sync(0xffffffffu, send.x, lane_mask); r.y = __shfl_xor_sync(0xffffffffu, send.y, lane_mask);
        r.z = __shfl_xor_sync(0xffffffffu, send.z, lane_mask); r.w = __shfl_xor_sync(0xffffffffu, send.w, lane_mask);
        return r;
    };
    auto add4 = [](float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); };
    float4 r1[2][MI];
#pragma unroll
    for (int jj = 0; jj < 2; ++jj)
#pragma unroll
        for (int m = 0; m < MI; ++m)
            r1[jj][m] = add4(b2 ? acc[2 + jj][m] : acc[jj][m], xchg(b2 ? acc[jj][m] : acc[2 + jj][m], 4));
    float4 r2[MI];
#pragma unroll
    for (int m = 0; m < MI; ++m) r2[m] = add4(b1 ? r1[1][m] : r1[0][m], xchg(b1 ? r1[0][m] : r1[1][m], 2));
    float4 v = add4(b0 ? r2[1] : r2[0], xchg(b0 ? r2[0] : r2[1], 1));
    const int row = rg + 4 * (2 * b2 + b1), kc = cg + 8 * b0;
    if (kc < KCO && row < e.nrows) {
        if (e.hmask) {
            const float4 hv = ld4(e.hmask + row * LD + 4 * kc);
            v.x = hv.x > 0.f ? v.x : 0.f; v.y = hv.y > 0.f ? v.y : 0.f;
            v.z = hv.z > 0.f ? v.z : 0.f; v.w = hv.w > 0.f ? v.w : 0.f;
        }
        st4(e.out + (size_t)row * e.ldo + 4 * kc, v);
    }
    (void)nrows;
}

}  // namespace npe
#include "npe_mma16.cuh"
namespace npe {

// ROWS-dispatching wrappers used by the kernels (ROWS = 64: throughput mapping on the FFMA pipe, ROWS = 16: latency
// mapping on mma.sync with the 3xTF32 split, npe_mma16.cuh; gemm_nt16 / gemm_nn16 above are its FP32 predecessors, kept
// for tools/microbench).  kc / nc are literals at every call site, so the dispatch folds at compile time.
template <int ROWS, int NI, int LDA, int LDW>
__device__ __forceinline__ void gemm_nt_r(const float* A, const float* W, int kc, int nrows, EpiNT e) {
    if (ROWS == 64) gemm_nt<NI, LDA, LDW>(A, W, kc, nrows, e);
    else if (kc == 24) mma_nt16<NI, 24, LDA, LDW>(A, W, e);
    else if (kc == 13) mma_nt16<NI, 13, LDA, LDW>(A, W, e);
    else mma_nt16<NI, 11, LDA, LDW>(A, W, e);
}
// nc = reduction length in chunks of 4 n.
template <int ROWS, int KCO, int LDZ, int LDW>
__device__ __forceinline__ void gemm_nn_r(const float* dZ, const float* W, int nc, int nrows, EpiNN e) {
    if (ROWS == 64) gemm_nn<KCO, LDZ, LDW>(dZ, W, nc, nrows, e);
    else if (nc == 13) mma_nn16<KCO, 13, LDZ, LDW>(dZ, W, e);
    else mma_nn16<KCO, 6, LDZ, LDW>(dZ, W, e);
}

// acc[4 i + j] += sum_{r < nrows} dZ[r][zoff + 4 a + i] * Hm[r][4 b + j];  (a, b) = (tile / KCH, tile % KCH).
// ZALIGNED: zoff + 4 a is 16-byte aligned (one LDS.128), else four scalar loads.
template <int KCH, int LDZ, int LDH, bool ZALIGNED>
__device__ __forceinline__ void wgrad(float (&acc)[16], const float* __restrict__ dZ, const float* __restrict__ Hm,
                                      int tile, int nrows, int zoff = 0) {
    const int a = tile / KCH, b = tile - a * KCH;
    const float* zp = dZ + zoff + 4 * a;
    const float* hp = Hm + 4 * b;
#pragma unroll 4
    for (int r = 0; r < nrows; ++r) {
        float4 z;
        if (ZALIGNED) z = ld4(zp);
        else { z.x = zp[0]; z.y = zp[1]; z.z = zp[2]; z.w = zp[3]; }
        const float4 h = ld4(hp);
        zp += LDZ; hp += LDH;
        acc[0] = fmaf(z.x, h.x, acc[0]);   acc[1] = fmaf(z.x, h.y, acc[1]);   acc[2] = fmaf(z.x, h.z, acc[2]);   acc[3] = fmaf(z.x, h.w, acc[3]);
        acc[4] = fmaf(z.y, h.x, acc[4]);   acc[5] = fmaf(z.y, h.y, acc[5]);   acc[6] = fmaf(z.y, h.z, acc[6]);   acc[7] = fmaf(z.y, h.w, acc[7]);
        acc[8] = fmaf(z.z, h.x, acc[8]);   acc[9] = fmaf(z.z, h.y, acc[9]);   acc[10] = fmaf(z.z, h.z, acc[10]); acc[11] = fmaf(z.z, h.w, acc[11]);
        acc[12] = fmaf(z.w, h.x, acc[12]); acc[13] = fmaf(z.w, h.y, acc[13]); acc[14] = fmaf(z.w, h.z, acc[14]); acc[15] = fmaf(z.w, h.w, acc[15]);
    }
}

// wgrad over ALL rows of a 16-row tile, fully unrolled: rows beyond the tile's valid rows must hold zeros in dZ and
// finite values in Hm (the one-grid step guarantees both), so the result equals wgrad(..., nrows).  Being one basic
// block, the compiler interleaves these FFMAs with the HMMA chain of the input-gradient pass issued right after it.
template <int KCH, int LDZ, int LDH, bool ZALIGNED>
__device__ __forceinline__ void wgrad16(float (&acc)[16], const float* __restrict__ dZ, const float* __restrict__ Hm,
                                        int tile, int zoff = 0) {
    const int a = tile / KCH, b = tile - a * KCH;
    const float* zp = dZ + zoff + 4 * a;
    const float* hp = Hm + 4 * b;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        float4 z;
        if (ZALIGNED) z = ld4(zp + r * LDZ);
        else { z.x = zp[r * LDZ]; z.y = zp[r * LDZ + 1]; z.z = zp[r * LDZ + 2]; z.w = zp[r * LDZ + 3]; }
        const float4 h = ld4(hp + r * LDH);
        acc[0] = fmaf(z.x, h.x, acc[0]);   acc[1] = fmaf(z.x, h.y, acc[1]);   acc[2] = fmaf(z.x, h.z, acc[2]);   acc[3] = fmaf(z.x, h.w, acc[3]);
        acc[4] = fmaf(z.y, h.x, acc[4]);   acc[5] = fmaf(z.y, h.y, acc[5]);   acc[6] = fmaf(z.y, h.z, acc[6]);   acc[7] = fmaf(z.y, h.w, acc[7]);
        acc[8] = fmaf(z.z, h.x, acc[8]);   acc[9] = fmaf(z.z, h.y, acc[9]);   acc[10] = fmaf(z.z, h.z, acc[10]); acc[11] = fmaf(z.z, h.w, acc[11]);
        acc[12] = fmaf(z.w, h.x, acc[12]); acc[13] = fmaf(z.w, h.y, acc[13]); acc[14] = fmaf(z.w, h.z, acc[14]); acc[15] = fmaf(z.w, h.w, acc[15]);
    }
}

// ---- TMA bulk copy (global -> shared, 1-D) with mbarrier completion: stages the packed weight image ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

}  // namespace npe
