// 16-row tile micro-kernels on the warp-level tensor-core path (mma.sync.m16n8k8 TF32 with the 3xTF32 split).
//
// The latency mapping of the NPE kernels (batches of at most a few hundred foci, e.g. the reference's batch of 50) runs
// tiles of 16 rows; one tile pass is 16 x 52 x 52 MACs, far too little to feed 256 FFMA lanes: the FP32 mapping
// (gemm_nt16 / gemm_nn16 in npe_tile.cuh) measures ~1700 cycles per pass, dominated by shared-memory latency, an
// 8-way split of the reduction and a 25-shuffle butterfly.  A 16-row tile is exactly ONE m16 fragment, so a pass is 7
// independent n8 output tiles (one per warp) of 7 k8 steps: every lane loads 6 words per step, no shuffles, no
// reduction tree.  tcgen05 is the wrong tool at this size (M >= 64, operands through TMEM / descriptors, one issuing
// thread, ~560 cycles for the MMA + commit + TMEM round trip alone); it carries the 128-row throughput path instead.
//
// Precision: every operand x is split into hi = x & 0xFFFFE000 (the TF32 the tensor core keeps) and lo = x - hi (exact
// in fp32); a product is hi*hi + hi*lo + lo*hi with fp32 accumulation: error ~2^-22 relative per term, i.e. FP32
// grade (tests keep the 1e-4 parity bar against the oracle).
//
// Fragment layout (PTX m16n8k8, g = lane >> 2, t = lane & 3):
//   A (16x8, row):  a0 = (g, t)  a1 = (g+8, t)  a2 = (g, t+4)  a3 = (g+8, t+4)
//   B (8x8,  col):  b0 = (k = t, n = g)  b1 = (k = t+4, n = g)
//   C (16x8):       c0 = (g, 2t)  c1 = (g, 2t+1)  c2 = (g+8, 2t)  c3 = (g+8, 2t+1)
#pragma once
// included by npe_tile.cuh (after EpiNT / EpiNN / ld2)
#include <cstdint>

namespace npe {

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xFFFFE000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}
// acc[0] += Ahi*Bhi ; acc[1] += Alo*Bhi ; acc[2] += Ahi*Blo   (three accumulators: independent dependency chains)
__device__ __forceinline__ void mma_tf32x3(float (&acc)[3][4], const float (&a)[4], const float (&b)[2]) {
    uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
#pragma unroll
    for (int i = 0; i < 2; ++i) split_tf32(b[i], bh[i], bl[i]);
    mma_tf32(acc[0], ah, bh);
    mma_tf32(acc[1], al, bh);
    mma_tf32(acc[2], ah, bl);
}
// all three products into ONE accumulator (weight gradients: the independent chains are the warp's other tiles)
__device__ __forceinline__ void mma_tf32x3_1(float (&c)[4], const float (&a)[4], const float (&b)[2]) {
    uint32_t ah[4], al[4], bh[2], bl[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) split_tf32(a[i], ah[i], al[i]);
#pragma unroll
    for (int i = 0; i < 2; ++i) split_tf32(b[i], bh[i], bl[i]);
    mma_tf32(c, al, bh);
    mma_tf32(c, ah, bl);
    mma_tf32(c, ah, bh);
}

// C[r][n] = act(sum_{k < 4 KC} A[r][k] * W[n][k]) for r < 16, n < 8 NI (stored for n < e.nmax); warp w owns n-tile w.
// Same contract as gemm_nt16.  Columns k >= 4 KC of A are never read as values (they may be uninitialised).
// Even and odd k-steps accumulate separately: six independent chains of at most four dependent MMAs.
template <int NI, int KC, int LDA, int LDW>
__device__ __forceinline__ void mma_nt16(const float* __restrict__ A, const float* __restrict__ W, EpiNT e) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    if (warp >= NI) return;
    constexpr int K = 4 * KC, STEPS = (K + 7) / 8;
    float acc[2][3][4];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[p][q][i] = 0.f;
    const float* a0 = A + g * LDA + t;
    const float* a1 = A + (g + 8) * LDA + t;
    const float* wp = W + (8 * warp + g) * LDW + t;
#pragma unroll
    for (int s = 0; s < STEPS; ++s) {
        const int k0 = 8 * s;
        const bool hi4 = k0 + 4 < K;                       // K is a multiple of 4: the upper half-step may be past the end
        const float a[4] = {a0[k0], a1[k0], hi4 ? a0[k0 + 4] : 0.f, hi4 ? a1[k0 + 4] : 0.f};
        const float b[2] = {wp[k0], hi4 ? wp[k0 + 4] : 0.f};
        mma_tf32x3(acc[s & 1], a, b);
    }
    float v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[i] = (acc[0][0][i] + acc[1][0][i]) + ((acc[0][1][i] + acc[1][1][i]) + (acc[0][2][i] + acc[1][2][i]));
        if (e.relu) v[i] = fmaxf(v[i], 0.f);
    }
    const int n = 8 * warp + 2 * t;
    float* o0 = e.out + g * e.ldo + n;
    float* o1 = e.out + (g + 8) * e.ldo + n;
    if (n < e.nmax) { o0[0] = v[0]; o1[0] = v[2]; }
    if (n + 1 < e.nmax) { o0[1] = v[1]; o1[1] = v[3]; }
}

// dA[r][k] = mask(sum_{n < 4 NC} dZ[r][n] * W[n][k]) for r < 16, k < 4 KCO; warp w owns output columns 8 w .. 8 w + 7.
// Same contract as gemm_nn16 (hmask: ReLU derivative taken from hmask[r][k] > 0; rows >= e.nrows are not stored).
template <int KCO, int NC, int LDZ, int LDW>
__device__ __forceinline__ void mma_nn16(const float* __restrict__ dZ, const float* __restrict__ W, EpiNN e) {
    constexpr int NT8 = (4 * KCO + 7) / 8;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
    if (warp >= NT8) return;
    constexpr int N = 4 * NC, STEPS = (N + 7) / 8;
    float acc[2][3][4];
#pragma unroll
    for (int p = 0; p < 2; ++p)
#pragma unroll
        for (int q = 0; q < 3; ++q)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[p][q][i] = 0.f;
    const float* z0 = dZ + g * LDZ + t;
    const float* z1 = dZ + (g + 8) * LDZ + t;
    const float* wp = W + t * LDW + 8 * warp + g;          // B(k = n index, col = output k): W[n0 + t][8 w + g]
#pragma unroll
    for (int s = 0; s < STEPS; ++s) {
        const int n0 = 8 * s;
        const bool hi4 = n0 + 4 < N;
        const float a[4] = {z0[n0], z1[n0], hi4 ? z0[n0 + 4] : 0.f, hi4 ? z1[n0 + 4] : 0.f};
        const float b[2] = {wp[n0 * LDW], hi4 ? wp[(n0 + 4) * LDW] : 0.f};
        mma_tf32x3(acc[s & 1], a, b);
    }
    const int k = 8 * warp + 2 * t;
    if (k >= 4 * KCO) return;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int row = g + 8 * h;
        if (row >= e.nrows) continue;
        float v[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int i = 2 * h + j;
            v[j] = (acc[0][0][i] + acc[1][0][i]) + ((acc[0][1][i] + acc[1][1][i]) + (acc[0][2][i] + acc[1][2][i]));
        }
        if (e.hmask) {
            const float2 hv = ld2(e.hmask + row * LD + k);
            v[0] = hv.x > 0.f ? v[0] : 0.f; v[1] = hv.y > 0.f ? v[1] : 0.f;
        }
        *reinterpret_cast<float2*>(e.out + (size_t)row * e.ldo + k) = make_float2(v[0], v[1]);
    }
}

// One 16 x 8 tile of a weight gradient: acc (C fragment) += sum_{r < nrows} dZ[r][n0 + m] * Hm[r][k0 + j] for m < 16,
// j < 8, reduction over the 16 rows of the tile (two k8 steps).  `c` persists across the tiles of the kernel.
// Columns n0 + m >= ncols of dZ are read as zero (the last m16 block of a layer may overhang the row).
template <int LDZ, int LDH>
__device__ __forceinline__ void mma_wgrad16(float (&c)[4], const float* __restrict__ dZ, const float* __restrict__ Hm,
                                            int n0, int k0, int nrows, int ncols) {
    const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int r0 = 0; r0 < 16; r0 += 8) {
        const bool v0 = r0 + t < nrows, v1 = r0 + t + 4 < nrows;
        const bool m0 = n0 + g < ncols, m1 = n0 + g + 8 < ncols;
        const float* zr0 = dZ + (r0 + t) * LDZ + n0 + g;
        const float* zr1 = dZ + (r0 + t + 4) * LDZ + n0 + g;
        // A(m, k = r) = dZ[r][n0 + m]
        const float a[4] = {v0 && m0 ? zr0[0] : 0.f, v0 && m1 ? zr0[8] : 0.f, v1 && m0 ? zr1[0] : 0.f, v1 && m1 ? zr1[8] : 0.f};
        // B(k = r, n = j) = Hm[r][k0 + j]
        const float b[2] = {v0 ? Hm[(r0 + t) * LDH + k0 + g] : 0.f, v1 ? Hm[(r0 + t + 4) * LDH + k0 + g] : 0.f};
        mma_tf32x3_1(c, a, b);
    }
}

}  // namespace npe
