// tcgen05 (5th-generation tensor core) building blocks for the TF32 variant of the NPE forward path.
//
//   D[128 x N] (fp32, TMEM) (+)= A[128 x 8] (tf32, smem) * B[N x 8]^T (tf32, smem)      one tcgen05.mma, kind::tf32
//
// Operand layout (both K-major, no swizzle; validated stand-alone in tools/microbench/umma_test.cu, error 2.9e-6 vs
// an fp64 product of tf32-truncated inputs): element (row, k) of a tile with ROWS rows lives at byte
//     (k / 4) * ROWS * 16 + row * 16 + (k % 4) * 4
// i.e. 8-row x 16-byte core matrices are contiguous 128-byte blocks; descriptor SBO = 128 B (next 8 rows), LBO =
// ROWS * 16 B (next 16-byte k-chunk), version 1, SWIZZLE_NONE.  The accumulator of row r is TMEM lane r, column n.
// A thread that owns row r therefore writes its next-layer activations with one STS.128 per k-chunk (a warp writes 512
// contiguous bytes) and reads the accumulator with tcgen05.ld.32x32b from its own lane.
#pragma once
#include <cstdint>
#include "npe_tile.cuh"

namespace npe {

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (sm_100)
    return d;
}

// instruction descriptor: D = F32, A = B = TF32, both K-major, M = 128, N = n
__device__ __forceinline__ uint32_t umma_idesc_tf32(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// K = 8 * ksteps product of an A tile (rows_a rows per chunk) and a B tile (rows_b rows per chunk) into TMEM columns
// [tmem_d, tmem_d + n).  Issued by ONE thread.
__device__ __forceinline__ void umma_gemm_tf32(uint32_t tmem_d, const float* sA, int rows_a, const float* sB, int rows_b,
                                               int n, int ksteps, bool accumulate_first) {
    const uint32_t idesc = umma_idesc_tf32(n);
    const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
    for (int s = 0; s < ksteps; ++s) {
        const uint64_t da = umma_desc(a0 + 2 * s * rows_a * 16, rows_a * 16, 128);
        const uint64_t db = umma_desc(b0 + 2 * s * rows_b * 16, rows_b * 16, 128);
        umma_tf32(tmem_d, da, db, idesc, (s > 0 || accumulate_first) ? 1u : 0u);
    }
}

// A operand in TENSOR MEMORY (lane = row, column = k; validated in tools/microbench/umma_test.cu, kernel k2).
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}

// 3xTF32: D (+)= Ahi*Bhi + Ahi*Blo + Alo*Bhi with A (hi at tmem_ahi, lo at tmem_alo) in tensor memory and the two
// images of B in shared memory.  K = 8 * ksteps.  FP32-grade (1.5e-6 relative in the stand-alone test).  One thread.
// The issue loop is on the critical path of every layer (one thread, ~20 dependent instructions per MMA when the 64-bit
// descriptor is rebuilt each time): the descriptor's low word (address field | LBO << 16) advances by a constant per
// K-step, the high word (SBO | version) is fixed.  `accumulate_first`: add to what the accumulator already holds.
__device__ __forceinline__ void umma_gemm_tf32x3_ts(uint32_t tmem_d, uint32_t tmem_ahi, uint32_t tmem_alo, const float* sBhi,
                                                    const float* sBlo, int rows_b, int n, int ksteps, bool accumulate_first = false) {
    const uint32_t idesc = umma_idesc_tf32(n);
    const uint32_t lbo = ((uint32_t)rows_b * 16u >> 4) << 16, step = 2u * (uint32_t)rows_b * 16u >> 4;
    const uint32_t bh = ((smem_u32(sBhi) >> 4) & 0x3FFFu) | lbo, bl = ((smem_u32(sBlo) >> 4) & 0x3FFFu) | lbo;
    const uint64_t hiw = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
    uint32_t acc = accumulate_first ? 1u : 0u;
#pragma unroll
    for (int pass = 0; pass < 3; ++pass) {
        const uint32_t a0 = pass == 2 ? tmem_alo : tmem_ahi;
        const uint32_t b0 = pass == 1 ? bl : bh;
#pragma unroll 4
        for (int s = 0; s < ksteps; ++s) {
            umma_tf32_ts(tmem_d, a0 + 8 * s, hiw | (uint64_t)(b0 + s * step), idesc, acc);
            acc = 1u;
        }
    }
}

// 3xTF32 with the hi and lo images of B STACKED along N in one shared-memory image ([k-chunk][hi rows (R) | lo rows (R)][4]):
//     D[0, R) (+)= Ahi Bhi + Alo Bhi        D[R, 2R) (+)= Ahi Blo             (the epilogue adds the two halves)
// i.e. 2 MMAs per K-step instead of 3.  A tcgen05.mma kind::tf32 costs ~58 cycles whatever its shape up to N = 112 and
// 65 at N = 128 (tools/microbench/umma_issue.cu), so the wide one is almost free.
__device__ __forceinline__ void umma_gemm_tf32x3_ts_stacked(uint32_t tmem_d, uint32_t tmem_ahi, uint32_t tmem_alo, const float* sW, int R,
                                                            int ksteps, bool accumulate_first = false) {
    const uint32_t idesc2 = umma_idesc_tf32(2 * R), idesc1 = umma_idesc_tf32(R);
    const uint32_t lbo = ((uint32_t)(2 * R) * 16u >> 4) << 16, step = 2u * (uint32_t)(2 * R) * 16u >> 4;
    const uint32_t b = ((smem_u32(sW) >> 4) & 0x3FFFu) | lbo;
    const uint64_t hiw = (uint64_t)((128u >> 4) | (1u << 14)) << 32;
    uint32_t acc = accumulate_first ? 1u : 0u;
#pragma unroll 4
    for (int s = 0; s < ksteps; ++s) {
        umma_tf32_ts(tmem_d, tmem_ahi + 8 * s, hiw | (uint64_t)(b + s * step), idesc2, acc);
        acc = 1u;
    }
#pragma unroll 4
    for (int s = 0; s < ksteps; ++s) umma_tf32_ts(tmem_d, tmem_alo + 8 * s, hiw | (uint64_t)(b + s * step), idesc1, 1u);
}

// This thread's row: 8 consecutive 32-bit columns starting at `col` of tensor memory.
__device__ __forceinline__ void tmem_st8(uint32_t tmem_base, int col, const float* v) {
    const uint32_t taddr = tmem_base + ((uint32_t)((threadIdx.x >> 5) & 3) * 32u << 16) + (uint32_t)col;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
                 "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
                 "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// hi = x with the low 13 mantissa bits cleared (what the tensor core keeps), lo = x - hi (exact in fp32)
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// Writes NCH chunks of 8 values of this thread's row, split into hi / lo, at tensor-memory columns col_hi / col_lo.
template <int NCH>
__device__ __forceinline__ void tmem_store_split(uint32_t tmem_base, int col_hi, int col_lo, const float* v) {
#pragma unroll
    for (int c = 0; c < NCH; ++c) {
        float h[8], l[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { h[i] = tf32_hi(v[8 * c + i]); l[i] = v[8 * c + i] - h[i]; }
        tmem_st8(tmem_base, col_hi + 8 * c, h);
        tmem_st8(tmem_base, col_lo + 8 * c, l);
    }
}

__device__ __forceinline__ void umma_commit(unsigned long long* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// One warp allocates `cols` TMEM columns (power of two >= 32) and publishes the base address through shared memory.
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* slot) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "n"(COLS));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(COLS));
}

// Accumulator row of this thread (TMEM lane = 32 * (warp % 4) + lane): 16 consecutive columns starting at `col`.
__device__ __forceinline__ void tmem_ld16(uint32_t tmem_base, int col, float* out) {
    uint32_t r[16];
    const uint32_t taddr = tmem_base + ((uint32_t)((threadIdx.x >> 5) & 3) * 32u << 16) + (uint32_t)col;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) out[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// round-to-nearest TF32 (the tensor core itself truncates the low 13 mantissa bits)
__device__ __forceinline__ float to_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__device__ __forceinline__ float4 to_tf32(float4 v) { return make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w)); }

}  // namespace npe
