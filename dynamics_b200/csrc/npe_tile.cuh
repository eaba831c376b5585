// CTA-cooperative FP32 tile micro-kernels over shared memory (256 threads, 64- or 32-row tiles).
//
//   gemm_nt : C[r][n]  = act(sum_k A[r][k] * W[n][k])        (forward;    both operands k-contiguous)
//   gemm_nn : dA[r][k] = mask(sum_n dZ[r][n] * W[n][k])      (input grad; W rows contiguous in the OUTPUT index)
//   wgrad   : dW[n][k] += sum_r dZ[r][n] * Hm[r][k]          (weight grad; per-thread 4x4 register tile, persistent)
//
// Thread mapping for the two GEMMs (tid = threadIdx.x):  ks = tid & 1 (split of the reduction index),
// cg = (tid >> 1) & 3 (output-column group), rg = tid >> 3 (row group, 0..31).  A thread owns rows {rg, rg + 32}
// (RPT = 2) or {rg} (RPT = 1) and output columns {cg + 4 i}; the two ks lanes are adjacent and exchange partial
// sums with one shuffle per output, after which lane ks finalises row rg + 32 * ks (RPT = 2) / both hold row rg.
//
// Code size matters more than unrolling here: the first version inlined one fully unrolled GEMM per layer (190-220 KB
// of SASS per kernel) and ncu showed "no instruction" (instruction-cache miss) as the top stall with 2 warps per
// scheduler.  The GEMMs are therefore __noinline__, shared by all layers (runtime pointers + a runtime epilogue
// descriptor), and the reduction loop is unrolled only twice (~4 KB bodies, inside the 32 KB L1.5 I-cache).
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

// Epilogue of gemm_nt: out[row * ldo + n] = relu ? max(v, 0) : v   for n < nmax.
struct EpiNT {
    float* out;
    int ldo;
    int nmax;
    int relu;
};

// C[r][cg + 4 i] for i < NI; reduction over kc chunks of 4.
template <int NI, int LDA, int LDW, int RPT>
__device__ __noinline__ void gemm_nt(const float* __restrict__ A, const float* __restrict__ W, int kc, EpiNT e) {
    const int tid = threadIdx.x;
    const int ks = tid & 1, cg = (tid >> 1) & 3, rg = tid >> 3;
    float acc0[NI], acc1[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) { acc0[i] = 0.f; acc1[i] = 0.f; }
    const float* a0 = A + rg * LDA + 4 * ks;
    const float* a1 = a0 + 32 * LDA;
    const float* w = W + cg * LDW + 4 * ks;
#pragma unroll 1
    for (int c = ks; c < kc; c += 2) {
        const float4 x0 = ld4(a0);
        float4 x1 = x0;
        if (RPT == 2) x1 = ld4(a1);
        float4 wv[NI];
#pragma unroll
        for (int i = 0; i < NI; ++i) wv[i] = ld4(w + 4 * i * LDW);
        // component-major order: consecutive FMAs hit different accumulators (dependency distance 2 NI), which is
        // what hides the 4-cycle FMA latency with only two warps per scheduler
#pragma unroll
        for (int i = 0; i < NI; ++i) { acc0[i] = fmaf(x0.x, wv[i].x, acc0[i]); if (RPT == 2) acc1[i] = fmaf(x1.x, wv[i].x, acc1[i]); }
#pragma unroll
        for (int i = 0; i < NI; ++i) { acc0[i] = fmaf(x0.y, wv[i].y, acc0[i]); if (RPT == 2) acc1[i] = fmaf(x1.y, wv[i].y, acc1[i]); }
#pragma unroll
        for (int i = 0; i < NI; ++i) { acc0[i] = fmaf(x0.z, wv[i].z, acc0[i]); if (RPT == 2) acc1[i] = fmaf(x1.z, wv[i].z, acc1[i]); }
#pragma unroll
        for (int i = 0; i < NI; ++i) { acc0[i] = fmaf(x0.w, wv[i].w, acc0[i]); if (RPT == 2) acc1[i] = fmaf(x1.w, wv[i].w, acc1[i]); }
        a0 += 8; a1 += 8; w += 8;
    }
    __syncwarp();
    if (RPT == 2) {
        float* o = e.out + (rg + 32 * ks) * e.ldo + cg;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            const float send = ks ? acc0[i] : acc1[i];
            const float recv = __shfl_xor_sync(0xffffffffu, send, 1);
            float mine = (ks ? acc1[i] : acc0[i]) + recv;
            if (e.relu) mine = fmaxf(mine, 0.f);
            if (cg + 4 * i < e.nmax) o[4 * i] = mine;
        }
    } else {
        float* o = e.out + rg * e.ldo + cg;
#pragma unroll
        for (int i = 0; i < NI; ++i) {
            float tot = acc0[i] + __shfl_xor_sync(0xffffffffu, acc0[i], 1);
            if (e.relu) tot = fmaxf(tot, 0.f);
            if ((i & 1) == ks && cg + 4 * i < e.nmax) o[4 * i] = tot;
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

// dA[r][4 kc .. 4 kc + 3] for kc = cg + 4 m < KCO; reduction over nc chunks of 4 n (ks lane takes n = 4 q + 2 ks + {0,1}).
template <int KCO, int LDZ, int LDW>
__device__ __noinline__ void gemm_nn(const float* __restrict__ dZ, const float* __restrict__ W, int nc, EpiNN e) {
    constexpr int MI = (KCO + 3) / 4;
    const int tid = threadIdx.x;
    const int ks = tid & 1, cg = (tid >> 1) & 3, rg = tid >> 3;
    float4 acc0[MI], acc1[MI];
#pragma unroll
    for (int m = 0; m < MI; ++m) { acc0[m] = make_float4(0.f, 0.f, 0.f, 0.f); acc1[m] = acc0[m]; }
    const float* z0p = dZ + rg * LDZ + 2 * ks;
    const float* z1p = z0p + 32 * LDZ;
    const float* wa = W + (2 * ks) * LDW + 4 * cg;
#pragma unroll 1
    for (int q = 0; q < nc; ++q) {
        const float2 z0 = *reinterpret_cast<const float2*>(z0p), z1 = *reinterpret_cast<const float2*>(z1p);
        float4 w0[MI], w1[MI];
#pragma unroll
        for (int m = 0; m < MI; ++m) {
            if (cg + 4 * m < KCO) { w0[m] = ld4(wa + 16 * m); w1[m] = ld4(wa + LDW + 16 * m); }
            else { w0[m] = make_float4(0.f, 0.f, 0.f, 0.f); w1[m] = w0[m]; }
        }
#pragma unroll
        for (int m = 0; m < MI; ++m) { axpy4(acc0[m], z0.x, w0[m]); axpy4(acc1[m], z1.x, w0[m]); }
#pragma unroll
        for (int m = 0; m < MI; ++m) { axpy4(acc0[m], z0.y, w1[m]); axpy4(acc1[m], z1.y, w1[m]); }
        z0p += 4; z1p += 4; wa += 4 * LDW;
    }
    __syncwarp();
    const int row = rg + 32 * ks;
#pragma unroll
    for (int m = 0; m < MI; ++m) {
        const float4 s = ks ? acc0[m] : acc1[m];
        float4 r;
        r.x = __shfl_xor_sync(0xffffffffu, s.x, 1); r.y = __shfl_xor_sync(0xffffffffu, s.y, 1);
        r.z = __shfl_xor_sync(0xffffffffu, s.z, 1); r.w = __shfl_xor_sync(0xffffffffu, s.w, 1);
        const float4 o = ks ? acc1[m] : acc0[m];
        const int kc = cg + 4 * m;
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
