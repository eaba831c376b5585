// CTA-cooperative FP32 tile micro-kernels over shared memory (256 threads, 64- or 32-row tiles).
//
//   gemm_nt : C[r][n]  = sum_k A[r][k] * W[n][k]     (forward;   both operands k-contiguous)
//   gemm_nn : dA[r][k] = sum_n dZ[r][n] * W[n][k]    (input grad; W rows contiguous in the OUTPUT index)
//   wgrad   : dW[n][k] += sum_r dZ[r][n] * Hm[r][k]  (weight grad; per-thread 4x4 register tile, persistent)
//
// Thread mapping for the two GEMMs (tid = threadIdx.x):  ks = tid & 1 (split of the reduction index),
// cg = (tid >> 1) & 3 (output-column group), rg = tid >> 3 (row group, 0..31).  A thread owns rows {rg, rg + 32}
// (RPT = 2) or {rg} (RPT = 1) and output columns {cg + 4 i}; the two ks lanes are adjacent and exchange partial
// sums with one shuffle per output, after which lane ks finalises row rg + 32 * ks (RPT = 2) / both hold row rg.
#pragma once
#include "npe_layout.h"

namespace npe {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ float dot4(float4 a, float4 b, float c) {
    c = fmaf(a.x, b.x, c); c = fmaf(a.y, b.y, c); c = fmaf(a.z, b.z, c); c = fmaf(a.w, b.w, c);
    return c;
}
__device__ __forceinline__ void axpy4(float4& acc, float a, float4 w) {
    acc.x = fmaf(a, w.x, acc.x); acc.y = fmaf(a, w.y, acc.y); acc.z = fmaf(a, w.z, acc.z); acc.w = fmaf(a, w.w, acc.w);
}

// C[r][cg + 4 i] for i < NI; K = 4 * KC.  epi(row, n, value) is called once per owned output.
template <int KC, int NI, int LDA, int LDW, int RPT, typename Epi>
__device__ __forceinline__ void gemm_nt(const float* __restrict__ A, const float* __restrict__ W, Epi epi) {
    const int tid = threadIdx.x;
    const int ks = tid & 1, cg = (tid >> 1) & 3, rg = tid >> 3;
    float acc0[NI], acc1[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) { acc0[i] = 0.f; acc1[i] = 0.f; }
    const float* a0 = A + rg * LDA;
    const float* a1 = A + (rg + 32) * LDA;
    const float* w = W + cg * LDW;
#pragma unroll
    for (int cc = 0; cc < (KC + 1) / 2; ++cc) {
        const int c = 2 * cc + ks;
        if (c < KC) {
            const float4 x0 = ld4(a0 + 4 * c);
            float4 x1 = x0;
            if (RPT == 2) x1 = ld4(a1 + 4 * c);
#pragma unroll
            for (int i = 0; i < NI; ++i) {
                const float4 wv = ld4(w + 4 * i * LDW + 4 * c);
                acc0[i] = dot4(x0, wv, acc0[i]);
                if (RPT == 2) acc1[i] = dot4(x1, wv, acc1[i]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
        if (RPT == 2) {
            const float send = ks ? acc0[i] : acc1[i];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
            const float mine = (ks ? acc1[i] : acc0[i]) + recv;
            epi(rg + 32 * ks, cg + 4 * i, mine);
        } else {
            const float tot = acc0[i] + __shfl_xor_sync(0xffffffffu, acc0[i], 1);
            if ((i & 1) == ks) epi(rg, cg + 4 * i, tot);
        }
    }
}

// dA[r][4 kc .. 4 kc + 3] for kc = cg + 4 m < KCO; reduction over n < 4 * NC (ks lane takes n = 4 q + 2 ks + {0,1}).
// epi(row, kc, float4 value).
template <int NC, int KCO, int LDZ, int LDW, typename Epi>
__device__ __forceinline__ void gemm_nn(const float* __restrict__ dZ, const float* __restrict__ W, Epi epi) {
    constexpr int MI = (KCO + 3) / 4;
    const int tid = threadIdx.x;
    const int ks = tid & 1, cg = (tid >> 1) & 3, rg = tid >> 3;
    float4 acc0[MI], acc1[MI];
#pragma unroll
    for (int m = 0; m < MI; ++m) { acc0[m] = make_float4(0.f, 0.f, 0.f, 0.f); acc1[m] = acc0[m]; }
    const float* z0p = dZ + rg * LDZ;
    const float* z1p = dZ + (rg + 32) * LDZ;
#pragma unroll
    for (int q = 0; q < NC; ++q) {
        const float4 z0 = ld4(z0p + 4 * q), z1 = ld4(z1p + 4 * q);
        const float za0 = ks ? z0.z : z0.x, zb0 = ks ? z0.w : z0.y;
        const float za1 = ks ? z1.z : z1.x, zb1 = ks ? z1.w : z1.y;
        const float* wa = W + (4 * q + 2 * ks) * LDW + 4 * cg;
#pragma unroll
        for (int m = 0; m < MI; ++m) {
            if (cg + 4 * m < KCO) {
                const float4 w0 = ld4(wa + 16 * m), w1 = ld4(wa + LDW + 16 * m);
                axpy4(acc0[m], za0, w0); axpy4(acc0[m], zb0, w1);
                axpy4(acc1[m], za1, w0); axpy4(acc1[m], zb1, w1);
            }
        }
    }
#pragma unroll
    for (int m = 0; m < MI; ++m) {
        const float4 s = ks ? acc0[m] : acc1[m];
        float4 r;
        r.x = __shfl_xor_sync(0xffffffffu, s.x, 1); r.y = __shfl_xor_sync(0xffffffffu, s.y, 1);
        r.z = __shfl_xor_sync(0xffffffffu, s.z, 1); r.w = __shfl_xor_sync(0xffffffffu, s.w, 1);
        const float4 o = ks ? acc1[m] : acc0[m];
        if (cg + 4 * m < KCO)
            epi(rg + 32 * ks, cg + 4 * m, make_float4(o.x + r.x, o.y + r.y, o.z + r.z, o.w + r.w));
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
        if (ZALIGNED) z = ld4(zp + r * LDZ);
        else { z.x = zp[r * LDZ]; z.y = zp[r * LDZ + 1]; z.z = zp[r * LDZ + 2]; z.w = zp[r * LDZ + 3]; }
        const float4 h = ld4(hp + r * LDH);
        acc[0] = fmaf(z.x, h.x, acc[0]);   acc[1] = fmaf(z.x, h.y, acc[1]);   acc[2] = fmaf(z.x, h.z, acc[2]);   acc[3] = fmaf(z.x, h.w, acc[3]);
        acc[4] = fmaf(z.y, h.x, acc[4]);   acc[5] = fmaf(z.y, h.y, acc[5]);   acc[6] = fmaf(z.y, h.z, acc[6]);   acc[7] = fmaf(z.y, h.w, acc[7]);
        acc[8] = fmaf(z.z, h.x, acc[8]);   acc[9] = fmaf(z.z, h.y, acc[9]);   acc[10] = fmaf(z.z, h.z, acc[10]); acc[11] = fmaf(z.z, h.w, acc[11]);
        acc[12] = fmaf(z.w, h.x, acc[12]); acc[13] = fmaf(z.w, h.y, acc[13]); acc[14] = fmaf(z.w, h.z, acc[14]); acc[15] = fmaf(z.w, h.w, acc[15]);
    }
}

}  // namespace npe
