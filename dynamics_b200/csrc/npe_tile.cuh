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
