// Prototype for the tensor-core weight gradient of the throughput kernels (round-2 item):
//   dW[n][k] = sum_{r < 128} dZ[r][n] * H[r][k]           one 128-row tile, n < 64, k < 64
// as tcgen05.mma kind::tf32 with M = 64 (n), N = 64 (k), K = 128 (r), 3xTF32 (hi*hi + hi*lo + lo*hi), accumulator in TMEM.
// The reduction index of a weight gradient is the ROW index, so both operands must be K-major in r: every thread owns
// one row r of dZ and H (as the epilogue threads of the chain kernels do) and writes its 64 values TRANSPOSED into the
// canonical no-swizzle layout   element (row = n, kk = r)  at float offset (r / 4) * 64 * 4 + n * 4 + (r % 4).
// The test prints the error against an fp64 product, the cycle cost of transposition + 48 MMAs, and WHERE the 64 x 64
// accumulator of an M = 64 MMA lands in tensor memory (lane of row n), which the epilogue of a real kernel needs.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

constexpr int R = 128, NN = 64, KK = 64;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

struct Smem {
    float Ah[NN * R], Al[NN * R], Bh[KK * R], Bl[KK * R];   // 4 x 32 KB
    unsigned long long bar;
    uint32_t tmem_slot;
};

// mshift: instruction-descriptor M field = (Mval >> 4) << 24
__global__ void __launch_bounds__(128, 1) k_wgrad(const float* dZ, const float* H, float* dump, long long* cycles, int m_val) {
    extern __shared__ __align__(1024) unsigned char raw[];
    Smem& S = *reinterpret_cast<Smem*>(raw);
    const int t = threadIdx.x, warp = t >> 5;
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&S.bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(&S.tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = S.tmem_slot;
    // the rows are in registers when the timing starts (in the chain kernels they come out of the epilogue)
    float zr[NN], hr[KK];
#pragma unroll
    for (int n = 0; n < NN; ++n) { zr[n] = dZ[t * NN + n]; hr[n] = H[t * KK + n]; }
    __syncthreads();
    long long t0 = clock64();
    // transposed, split stores: thread t = row r; the column order is rotated by r / 4 so that the 8 threads of a warp that
    // share r % 4 hit different banks
    {
        const int r = t;
        const int base = (r >> 2) * (NN * 4) + (r & 3);
#pragma unroll
        for (int j = 0; j < NN; ++j) {
            const int n = (j + (r >> 2)) & (NN - 1);
            float z = 0.f, h = 0.f;
#pragma unroll
            for (int q = 0; q < NN; ++q) { if (q == j) { z = zr[q]; h = hr[q]; } }     // static register index; rotation by address only
            (void)n;
            const float zh = tf32_hi(z), hh = tf32_hi(h);
            S.Ah[base + j * 4] = zh; S.Al[base + j * 4] = z - zh;
            S.Bh[base + j * 4] = hh; S.Bl[base + j * 4] = h - hh;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    long long t1 = clock64();
    if (t == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KK >> 3) << 17) | ((uint32_t)(m_val >> 4) << 24);
        for (int pass = 0; pass < 3; ++pass) {
            const float* sA = pass == 2 ? S.Al : S.Ah;
            const float* sB = pass == 1 ? S.Bl : S.Bh;
            for (int s = 0; s < R / 8; ++s) {
                // one K = 8 step = two 16-byte k-chunks; chunk stride (LBO) = rows * 16 B, 8-row stride (SBO) = 128 B
                const uint64_t da = make_desc(smem_u32(sA) + 2 * s * NN * 16, NN * 16, 128);
                const uint64_t db = make_desc(smem_u32(sB) + 2 * s * KK * 16, KK * 16, 128);
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
    long long t2 = clock64();
    // the same 48 MMAs twice into two INDEPENDENT accumulators (two layers of a tile in flight): issue-bound or chain-bound?
    if (t == 0) {
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(KK >> 3) << 17) | ((uint32_t)(m_val >> 4) << 24);
        for (int pass = 0; pass < 3; ++pass) {
            const float* sA = pass == 2 ? S.Al : S.Ah;
            const float* sB = pass == 1 ? S.Bl : S.Bh;
            for (int s = 0; s < R / 8; ++s) {
                const uint64_t da = make_desc(smem_u32(sA) + 2 * s * NN * 16, NN * 16, 128);
                const uint64_t db = make_desc(smem_u32(sB) + 2 * s * KK * 16, KK * 16, 128);
                const uint32_t acc = (pass > 0 || s > 0);
                for (int which = 0; which < 2; ++which)
                    asm volatile(
                        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                        ::"r"(tmem + 64 + 64 * which), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&S.bar)) : "memory");
    }
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT2_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], 1;\n\t@p bra DONE2_%=;\n\tbra WAIT2_%=;\n\tDONE2_%=:\n\t}"
        ::"r"(smem_u32(&S.bar)) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    long long t3 = clock64();
    // dump all 128 lanes x 64 columns: the host finds where row n of the accumulator lives
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
    for (int c = 0; c < 64; ++c) dump[t * 64 + c] = __uint_as_float(v[c]);
    if (t == 0) { cycles[0] = t1 - t0; cycles[1] = t2 - t1; cycles[2] = t3 - t2; }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

int main() {
    std::vector<float> dZ(R * NN), H(R * KK), ref(NN * KK), dump(128 * 64);
    srand(3);
    for (auto& v : dZ) v = (rand() % 2001 - 1000) / 1000.f;
    for (auto& v : H) v = (rand() % 2001 - 1000) / 1000.f;
    double mx = 0;
    for (int n = 0; n < NN; ++n)
        for (int k = 0; k < KK; ++k) {
            double s = 0;
            for (int r = 0; r < R; ++r) s += (double)dZ[r * NN + n] * H[r * KK + k];
            ref[n * KK + k] = (float)s; mx = fmax(mx, fabs(s));
        }
    float *dz, *dh, *dd; long long* dc;
    cudaMalloc(&dz, dZ.size() * 4); cudaMalloc(&dh, H.size() * 4); cudaMalloc(&dd, dump.size() * 4); cudaMalloc(&dc, 24);
    cudaMemcpy(dz, dZ.data(), dZ.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dh, H.data(), H.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem) + 1024);
    for (int m_val : {64, 128}) {
        cudaMemset(dd, 0, dump.size() * 4);
        k_wgrad<<<1, 128, sizeof(Smem) + 1024>>>(dz, dh, dd, dc, m_val);
        cudaError_t e = cudaDeviceSynchronize();
        printf("M = %d: %s\n", m_val, cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        cudaMemcpy(dump.data(), dd, dump.size() * 4, cudaMemcpyDeviceToHost);
        long long cyc[3]; cudaMemcpy(cyc, dc, 24, cudaMemcpyDeviceToHost);
        // locate every accumulator row: the TMEM lane whose 64 columns match ref[n][:]
        int found = 0; double emax = 0; int lane_of[NN];
        for (int n = 0; n < NN; ++n) {
            lane_of[n] = -1;
            for (int lane = 0; lane < 128; ++lane) {
                double e2 = 0;
                for (int k = 0; k < KK; ++k) e2 = fmax(e2, fabs(dump[lane * 64 + k] - ref[n * KK + k]));
                if (e2 < 1e-3 * mx) { lane_of[n] = lane; emax = fmax(emax, e2); ++found; break; }
            }
        }
        printf("  rows located %d / %d, max err vs fp64 %.3e (rel %.2e), cycles: transpose+split (rows in registers) %lld, 48 MMA + commit %lld, 2 x 48 MMA into two accumulators %lld\n",
               found, NN, emax, emax / mx, cyc[0], cyc[1], cyc[2]);
        printf("  lane of row n: n=0 -> %d, 1 -> %d, 15 -> %d, 16 -> %d, 17 -> %d, 31 -> %d, 32 -> %d, 47 -> %d, 48 -> %d, 63 -> %d\n",
               lane_of[0], lane_of[1], lane_of[15], lane_of[16], lane_of[17], lane_of[31], lane_of[32], lane_of[47], lane_of[48], lane_of[63]);
    }
    return 0;
}
