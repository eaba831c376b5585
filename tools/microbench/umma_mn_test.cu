// Can the weight gradient  dW[n][k] = sum_r G[r][n] * X[r][k]  take its operands as they lie in the stash -- row-major
// [r][col], i.e. MN-major for the MMA (the reduction index r is the slow one) -- instead of transposing every tile to
// K-major in shared memory?  tcgen05 kind::tf32 with a_major = b_major = MN (instruction descriptor bits 15 / 16) and the
// canonical 128-byte-swizzle MN-major layout: chunks of 32 columns, each chunk = [row][32 floats] with a 128-byte pitch,
// XOR-swizzled in 16-byte units by (row % 8) -- exactly what a TMA tile load with CU_TENSOR_MAP_SWIZZLE_128B produces
// from a row-major tensor.  Descriptor: LBO = bytes between column chunks, SBO = bytes between groups of 8 rows.
// Also answers: does the tensor core TRUNCATE fp32 operands to tf32 (then the raw tile is its own "hi" half)?
//   mode 0: hi = raw fp32 bits, lo = x - trunc(x)      mode 1: hi = trunc(x) stored explicitly
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int R = 32, NC = 56, CH = 2;                     // rows (K), valid columns, 32-column chunks
constexpr int CHUNK = R * 32;                              // floats per chunk (4096 B)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, int ltype = 2) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;                                // descriptor version (Blackwell)
    d |= (uint64_t)ltype << 61;                            // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
    return d;
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ int swz(int row, int col, int kind) {   // float offset of (row, col % 32) inside a chunk
    int off = row * 128 + (col & 31) * 4;
    if (kind == 0) off ^= ((off >> 7) & 7) << 4;           // Swizzle<3,4,3>: 16-byte units by row % 8
    else if (kind == 1) off ^= ((off >> 7) & 3) << 5;      // Swizzle<2,5,2>: 32-byte units by row % 4
    return off >> 2;
}

struct Smem {
    float Ah[CH * CHUNK], Al[CH * CHUNK], Bh[CH * CHUNK], Bl[CH * CHUNK];
    unsigned long long bar;
    uint32_t tmem_slot;
};

__global__ void __launch_bounds__(128, 1) k_mn(const float* G, const float* X, float* dump, int mode, int swap, int ltype, int skind, int sbo_in) {
    extern __shared__ unsigned char raw0[];
    Smem& S = *reinterpret_cast<Smem*>((reinterpret_cast<uintptr_t>(raw0) + 1023) & ~(uintptr_t)1023);
    const int t = threadIdx.x, warp = t >> 5;
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 64;" ::"r"(smem_u32(&S.tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    for (int i = t; i < R * 64; i += 128) {
        const int r = i / 64, c = i % 64;
        const float g = c < NC ? G[r * NC + c] : 0.f, x = c < NC ? X[r * NC + c] : 0.f;
        const int o = (c / 32) * CHUNK + swz(r, c, skind);
        S.Ah[o] = mode ? tf32_hi(g) : g; S.Al[o] = g - tf32_hi(g);
        S.Bh[o] = mode ? tf32_hi(x) : x; S.Bl[o] = x - tf32_hi(x);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = S.tmem_slot;
    if (t == 0) {
        // M = 64 (rows n of dW), N = 56 (columns k), tf32 inputs, fp32 accumulator, A and B MN-major
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(NC >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
        const uint32_t lbo = swap ? sbo_in : CHUNK * 4, sbo = swap ? CHUNK * 4 : sbo_in;
        for (int pass = 0; pass < 3; ++pass) {
            const float* sA = pass == 2 ? S.Al : S.Ah;
            const float* sB = pass == 1 ? S.Bl : S.Bh;
            for (int s = 0; s < R / 8; ++s) {
                const uint64_t da = make_desc_mn(smem_u32(sA) + s * 1024, lbo, sbo, ltype);
                const uint64_t db = make_desc_mn(smem_u32(sB) + s * 1024, lbo, sbo, ltype);
                const uint32_t acc = (pass > 0 || s > 0);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&S.bar)) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
        ::"r"(smem_u32(&S.bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t v[64];
#pragma unroll
    for (int c = 0; c < 64; c += 16) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(v[c + 0]), "=r"(v[c + 1]), "=r"(v[c + 2]), "=r"(v[c + 3]), "=r"(v[c + 4]), "=r"(v[c + 5]), "=r"(v[c + 6]),
              "=r"(v[c + 7]), "=r"(v[c + 8]), "=r"(v[c + 9]), "=r"(v[c + 10]), "=r"(v[c + 11]), "=r"(v[c + 12]),
              "=r"(v[c + 13]), "=r"(v[c + 14]), "=r"(v[c + 15])
            : "r"(lane_base + c));
    }
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int c = 0; c < 64; ++c) dump[t * 64 + c] = __uint_as_float(v[c]);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 64;" ::"r"(tmem));
}

int main() {
    std::vector<float> G(R * NC), X(R * NC), dump(128 * 64);
    std::vector<double> ref(64 * NC, 0.0);
    srand(3);
    for (auto& v : G) v = (rand() % 200001 - 100000) / 100000.f;
    for (auto& v : X) v = (rand() % 200001 - 100000) / 100000.f;
    double mx = 0;
    for (int n = 0; n < NC; ++n)
        for (int k = 0; k < NC; ++k) {
            double s = 0;
            for (int r = 0; r < R; ++r) s += (double)G[r * NC + n] * X[r * NC + k];
            ref[n * NC + k] = s; mx = fmax(mx, fabs(s));
        }
    float *dg, *dx, *dd;
    cudaMalloc(&dg, G.size() * 4); cudaMalloc(&dx, X.size() * 4); cudaMalloc(&dd, dump.size() * 4);
    cudaMemcpy(dg, G.data(), G.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dx, X.data(), X.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_mn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem) + 2048);
    for (int ltype : {1, 2})
        for (int skind : {0, 1, 2})
            for (int sbo : {512, 1024})
                for (int swap = 0; swap < 2; ++swap) {
                    const int mode = 1;
                    cudaMemset(dd, 0, dump.size() * 4);
                    k_mn<<<1, 128, sizeof(Smem) + 2048>>>(dg, dx, dd, mode, swap, ltype, skind, sbo);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("ltype %d skind %d sbo %d swap %d: %s\n", ltype, skind, sbo, swap, cudaGetErrorString(e)); return 1; }
                    cudaMemcpy(dump.data(), dd, dump.size() * 4, cudaMemcpyDeviceToHost);
                    double emax = 0; int nz = 0;
                    for (int n = 0; n < NC; ++n)
                        for (int k = 0; k < NC; ++k) {
                            const float v = dump[((n % 16) + 32 * (n / 16)) * 64 + k];
                            nz += v != 0.f;
                            emax = fmax(emax, fabs(v - ref[n * NC + k]));
                        }
                    printf("layout_type %d, swizzle %s, SBO %4d, %s: rel err %.2e, nonzero %d / %d\n", ltype,
                           skind == 0 ? "16B x row%8" : skind == 1 ? "32B x row%4" : "none       ", sbo, swap ? "LBO<->SBO" : "LBO=chunk ", emax / mx, nz, NC * NC);
                }
    // the working combination, raw fp32 bits as the hi operand: does the tensor core truncate?
    return 0;
}
