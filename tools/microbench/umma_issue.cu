// How much does ONE tcgen05.mma kind::tf32 (K = 8) cost as a function of its shape?  The weight-gradient kernel issues
// M = 64, N = 56 / 96 MMAs and measures ~65 cycles each on the issuing thread, twice the max(M,128) * N / 256 floor: is that
// a per-instruction cost (then stacking two jobs into one M = 128, N = 112 MMA halves the tensor time) or does it scale?
// One CTA, one thread issues `reps` x 8 K-steps into one accumulator, commits, waits.  Operands are zeros (values are
// irrelevant for timing).  Prints cycles per MMA: issue loop alone, and until the commit's mbarrier completes.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
constexpr int KSTEPS = 8;
struct Smem {
    float A[KSTEPS * 2 * 128 * 4];      // [k / 4][row (M)][k % 4]
    float B[KSTEPS * 2 * 256 * 4];
    unsigned long long bar;
    uint32_t tmem_slot;
};

__global__ void __launch_bounds__(128, 1) k_issue(long long* out, int M, int N, int reps, int naccs) {
    extern __shared__ __align__(1024) unsigned char raw[];
    Smem& S = *reinterpret_cast<Smem*>(raw);
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < (int)(sizeof(S.A) / 4); i += 128) S.A[i] = 0.f;
    for (int i = t; i < (int)(sizeof(S.B) / 4); i += 128) S.B[i] = 0.f;
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S.tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = S.tmem_slot;
    long long t0 = 0, t1 = 0;
    if (t == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t d = tmem + (uint32_t)((r % naccs) * N);
#pragma unroll
            for (int s = 0; s < KSTEPS; ++s) {
                const uint64_t da = make_desc(smem_u32(S.A) + 2 * s * M * 16, M * 16, 128);
                const uint64_t db = make_desc(smem_u32(S.B) + 2 * s * N * 16, N * 16, 128);
                const uint32_t acc = (r >= naccs || s > 0);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
        }
        t1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&S.bar)) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
        ::"r"(smem_u32(&S.bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const long long t2 = clock64();
    if (t == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

// Same, but TWO threads (lane 0 of warps 0 and 1) issue concurrently into different accumulators: is the ~58-cycle cost a
// limit of the issuing thread or of the SM?
__global__ void __launch_bounds__(128, 1) k_issue2(long long* out, int M, int N, int reps, int nthreads) {
    extern __shared__ __align__(1024) unsigned char raw[];
    Smem& S = *reinterpret_cast<Smem*>(raw);
    __shared__ unsigned long long bar2[2];
    const int t = threadIdx.x, warp = t >> 5;
    for (int i = t; i < (int)(sizeof(S.A) / 4); i += 128) S.A[i] = 0.f;
    for (int i = t; i < (int)(sizeof(S.B) / 4); i += 128) S.B[i] = 0.f;
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar2[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&S.tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = S.tmem_slot;
    const long long t0 = clock64();
    if ((t & 31) == 0 && warp < nthreads) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint32_t d = tmem + (uint32_t)(warp * 256);
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int s = 0; s < KSTEPS; ++s) {
                const uint64_t da = make_desc(smem_u32(S.A) + 2 * s * M * 16, M * 16, 128);
                const uint64_t db = make_desc(smem_u32(S.B) + 2 * s * N * 16, N * 16, 128);
                const uint32_t acc = (r > 0 || s > 0);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar2[warp])) : "memory");
    }
    for (int w = 0; w < nthreads; ++w)
        asm volatile(
            "{\n\t.reg .pred p;\n\tWAIT_%=:\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
            ::"r"(smem_u32(&bar2[w])) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const long long t2 = clock64();
    if (t == 0) out[0] = t2 - t0;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}

int main() {
    long long* dc; cudaMalloc(&dc, 16);
    cudaFuncSetAttribute(k_issue, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem) + 1024);
    const int shapes[][2] = {{64, 56}, {64, 64}, {64, 96}, {64, 112}, {64, 128}, {64, 256}, {128, 64}, {128, 96}, {128, 112},
                             {128, 128}, {128, 160}, {128, 256}};
    for (auto& sh : shapes)
        for (int naccs : {1, 2}) {
            if (naccs * sh[1] > 512) continue;
            const int reps = 24;
            long long c[2] = {0, 0};
            for (int rep = 0; rep < 3; ++rep) {
                k_issue<<<1, 128, sizeof(Smem) + 1024>>>(dc, sh[0], sh[1], reps, naccs);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("M=%d N=%d: %s\n", sh[0], sh[1], cudaGetErrorString(e)); return 1; }
                cudaMemcpy(c, dc, 16, cudaMemcpyDeviceToHost);
            }
            const double n = reps * KSTEPS;
            printf("M=%3d N=%3d accumulators=%d: issue %.1f cycles/MMA, issue+drain %.1f cycles/MMA (floor max(M,128)*N/256 = %.1f)\n",
                   sh[0], sh[1], naccs, c[0] / n, c[1] / n, 128.0 * sh[1] / 256.0);
        }
    cudaFuncSetAttribute(k_issue2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem) + 1024);
    for (int n : {56, 64, 112})
        for (int nth : {1, 2}) {
            long long c = 0;
            for (int rep = 0; rep < 3; ++rep) {
                k_issue2<<<1, 128, sizeof(Smem) + 1024>>>(dc, 128, n == 56 ? 64 : n, 24, nth);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("k_issue2: %s\n", cudaGetErrorString(e)); return 1; }
                cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost);
            }
            printf("M=128 N=%3d, %d issuing thread(s), 24 x 8 MMAs each: %.1f cycles per MMA of ONE thread's stream (wall %lld)\n",
                   n == 56 ? 64 : n, nth, c / (24.0 * KSTEPS), c);
        }
    return 0;
}
