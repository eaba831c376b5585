// Weight gradients of the throughput regime on tcgen05, operands taken AS THEY LIE in the stashes (reference
// src/lua/npe.lua:290-336: the accGradParameters half of model:bp; modules.lua:10-26,46-88 for the layer shapes).
//
//     dW_l[n][k] = sum_r G_l[r][n] * X_l[r][k]           G = stashed dz of the layer, X = stashed input of the layer
//
// The reduction index is the ROW index r, and the stashes are row-major [r][56]: for the MMA both operands are MN-major.
// The first tensor-core version (k_wgrad_tc, npe_tcbwd.cuh) transposed every tile to K-major in shared memory (TMA raw
// ring -> 8 transposer warps -> operand ring) and issued one M = 64, N = 56 MMA chain per layer.  Two measurements
// (tools/microbench/umma_issue.cu, umma_mn_test.cu) replaced it:
//   * a tcgen05.mma kind::tf32 costs ~58 cycles on the issuing thread whatever its shape up to N = 112 (then the
//     max(M,128) N / 256 floor): an M = 64, N = 56 MMA wastes 3/4 of the instruction.  So TWO layers share one MMA:
//     A = [G_a | G_b] (M = 128), B = [X_a | X_b] (N = 128 / 160); the diagonal blocks of the 128 x N accumulator are the
//     two weight gradients, the off-diagonal blocks are never read.  Half the instructions.
//   * kind::tf32 takes MN-major operands in exactly one shared-memory layout, SWIZZLE_128B_BASE32B: chunks of 32 columns,
//     each chunk [row][32 floats] with a 128-byte pitch, 32-byte units XOR-swizzled by (row % 4); descriptor layout type 1,
//     LBO = chunk stride, SBO = 512 (4 rows).  A row-major tile goes there with NO transposition: a thread loads 16 bytes
//     of a row and stores them (hi / lo split on the way) 16 bytes wide, conflict-free.
// So: 16 loader warps in 2 groups (loads of a unit -> registers, wait for the unit's operand stage, split + swizzled
// stores, arrive), a warp that prefetches the stash blocks of coming units into L2, one MMA warp (12 MMAs per unit of 32 rows and two layers), accumulators of all layers in
// tensor memory for the whole kernel, tile-major partials for the existing reduce / optimiser / exchange kernels.
// Gathered operands (state windows of the encoders and of decoder layer 1) are fetched by the same loader threads.
#pragma once
#include "npe_tcbwd.cuh"

namespace npe {

constexpr int WM_ROWS = 32;                    // rows per unit: K = 32 = 4 MMA k-steps
constexpr int WM_CHUNK = WM_ROWS * 32;         // floats of one 32-column chunk (4096 B)
constexpr int WM_GROUPS = 2;                   // loader groups (256 threads each; more would cap the kernel at 72 registers)
constexpr int WM_STAGES = 3;                   // operand stages
constexpr int WM_GT = 256;                     // threads per loader group
constexpr int WM_NT = WM_GROUPS * WM_GT + 64;  // + the MMA warp + the L2 prefetch warp
constexpr int WM_AHEAD = 8;                    // units the prefetch warp runs ahead of the MMA warp (8 x 28 KB per SM)
constexpr int WM_ACH = 4, WM_BCH = 5;          // chunks of A (M = 128) and of B (N <= 160)

struct WPair {
    WJob a, b;            // b.G == nullptr: a single layer (rows 64..127 of A are zeros)
    int has_b;
    int ca, cb;           // 32-column chunks of X_a / X_b in B: 2 (stash rows) or 3 (gathered, 96 columns)
    int tmem_col;         // accumulator columns [tmem_col, tmem_col + 32 (ca + cb))
};
struct WmArgs {
    WPair pair[3];
    int npairs;
    const int* nrows_dev;  // number of rows R (pairs: info[0]) or nullptr
    int nrows;
    float* partial;        // [gridDim.x][PT_TOTAL]
    long long* prof;       // tools only (NPE_CHAIN_PROF)
};
struct WmStage { float Ahi[WM_ACH * WM_CHUNK], Alo[WM_ACH * WM_CHUNK], Bhi[WM_BCH * WM_CHUNK], Blo[WM_BCH * WM_CHUNK]; };   // 72 KB
struct WmSmem {
    WmStage st[WM_STAGES];
    unsigned long long full[WM_STAGES], empty[WM_STAGES], done;
    uint32_t tmem;
    int consumed;          // units the MMA warp has issued (paces the prefetch warp)
};
static_assert(sizeof(WmSmem) <= 232448 - 1024, "weight-gradient stages exceed shared memory");

// Streaming loads that do not allocate an L1 line: with ~216 KB of the SM's 256 KB carved out as shared memory the L1 has
// a few dozen lines left, and every allocating load in flight holds one -- the loaders stalled in their load ISSUE.
__device__ __forceinline__ float4 ldg4_stream(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float2 ldg2_stream(const float* p) {
    float2 v;
    asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
}
// float offset of (row, column) of an operand: chunk = column / 32, 32-byte units swizzled by row % 4
__device__ __forceinline__ int wm_offset(int row, int col) {
    const int cc = col & 31;
    return (col >> 5) * WM_CHUNK + row * 32 + ((((cc >> 3) ^ (row & 3))) << 3) + (cc & 7);
}
__device__ __forceinline__ void wm_put4(float* hi, float* lo, int off, float4 v) {
    float4 h, l;
    h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    st4(hi + off, h); st4(lo + off, l);
}
__device__ __forceinline__ void wm_put2(float* hi, float* lo, int off, float2 v) {
    float2 h, l;
    h.x = tf32_hi(v.x); h.y = tf32_hi(v.y);
    l.x = v.x - h.x; l.y = v.y - h.y;
    *reinterpret_cast<float2*>(hi + off) = h; *reinterpret_cast<float2*>(lo + off) = l;
}
// Per-thread constants of the plain (stash block) path: item k of a thread is (row = tg / 16 + 16 k, 16-byte piece tg % 16) of
// a 32-row x 64-column block.  Consecutive lanes read consecutive addresses (rows are 224 bytes apart: a block is contiguous).
struct WmLane {
    int row, col;          // first item: row, first column of the piece
    int goff;              // float offset of the first item inside a stash block; the second is 16 rows further
    int soff;              // float offset inside an operand (chunk of column `col`); the second item is 16 rows = 512 floats further
    __device__ __forceinline__ void init(int tg) {
        row = tg >> 4; col = 4 * (tg & 15);
        goff = row * HSLOT + col;
        soff = (col >> 5) * WM_CHUNK + row * 32 + ((((col & 31) >> 3) ^ (row & 3)) << 3) + (col & 7);
    }
};
// blk = first row of the block (or nullptr: zeros); rows >= rows_left and columns >= valid read as zero
__device__ __forceinline__ void wm_load_plain(float4 (&v)[2], const float* blk, int valid, int rows_left, const WmLane& L) {
    const bool colok = blk && L.col < valid;
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    v[0] = (colok && L.row < rows_left) ? ldg4_stream(blk + L.goff) : z4;
    v[1] = (colok && L.row + 16 < rows_left) ? ldg4_stream(blk + L.goff + 16 * HSLOT) : z4;
}
// col0: first column of the block inside the operand (a multiple of 32)
__device__ __forceinline__ void wm_store_plain(float* hi, float* lo, int col0, const float4 (&v)[2], const WmLane& L) {
    const int o = (col0 >> 5) * WM_CHUNK + L.soff;
    wm_put4(hi, lo, o, v[0]);
    wm_put4(hi, lo, o + 16 * 32, v[1]);
}
// row offsets of this thread's gathered row (zero for rows >= R): fetched ONE UNIT AHEAD, so that the window loads of a
// gathered unit do not wait for a dependent round trip
__device__ __forceinline__ void wm_gather_offsets(long long (&o)[2], const WJob& J, long long row0, long long R, int tg) {
    const long long r = row0 + (tg >> 3);
    o[0] = 0; o[1] = 0;
    if (r < R) {
        o[0] = J.gs[0].offs[r];
        if (J.ngs == 2) o[1] = J.gs[1].offs[r];
    }
}
__device__ __forceinline__ void wm_load_gather(float2 (&w)[6], const WJob& J, const long long (&o)[2], long long row0, long long R, int tg) {
    const int row = tg >> 3, sub = tg & 7;
    const long long r = row0 + row;
    const bool live = r < R;
    const float2 z2 = make_float2(0.f, 0.f);
    if (J.ngs == 2) {
        const float* pa = J.gs[0].base + o[0];
        const float* pb = J.gs[1].base + o[1];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int col = 2 * (sub + 8 * k);
            w[k] = !live ? z2 : col < IN ? ldg2_stream(pa + col) : col < 48 ? z2 : col < 48 + IN ? ldg2_stream(pb + col - 48) : z2;
        }
    } else {
        const float* pz = J.X0 + (live ? r * HSLOT : 0);
        const float* pw = J.gs[0].base + o[0];
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const int col = 2 * (sub + 8 * k);
            w[k] = !live ? z2 : col < 52 ? ldg2_stream(pz + col) : ldg2_stream(pw + col - 52);
        }
    }
}
__device__ __forceinline__ void wm_store_gather(float* hi, float* lo, int col0, const float2 (&w)[6], int tg) {
    const int row = tg >> 3, sub = tg & 7;
#pragma unroll
    for (int k = 0; k < 6; ++k) wm_put2(hi, lo, wm_offset(row, col0 + 2 * (sub + 8 * k)), w[k]);
}

// HBM -> L2 only: no registers, no shared memory, so the bytes in flight are not bounded by the operand stages
__device__ __forceinline__ void prefetch_l2_bulk(const void* p, unsigned bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t umma_idesc_tf32_mn(int m, int n) {   // tf32 x tf32 -> f32, A and B MN-major
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// One unit of the loader: pair J of a launch of kind KIND (0: pair part, pairs (L5,L4) (L3,L2) (L1,encoders); 1: decoder part,
// pairs (D5,D4) (D3,D2) (D1)).  J and KIND are compile-time so that every pair gets straight-line code: the generic version
// spent ~500 instructions per unit on parameter indexing, 64-bit iterator arithmetic and branches over the operand kinds.
template <int KIND, int J>
__device__ __forceinline__ void wm_unit(const WmArgs& A, WmSmem& S, int u, long long row0, long long R, const WmLane& L, int tg,
                                        long long (&noff)[2], long long& pw0, long long& pw1, long long& pw2) {
    constexpr bool GATHER = J == 2;                        // pair 2 holds the gathered operand: X_b (encoders) / X_a (decoder layer 1)
    constexpr bool HAS_B = !(KIND == 1 && J == 2);
    const WPair& P = A.pair[J];
    const int s = u % WM_STAGES;
    WmStage& st = S.st[s];
    const long long ca0 = clock64();
    const int rows_left = (int)min((long long)WM_ROWS, R - row0);
    const long long boff = row0 * HSLOT;
    float4 ga[2], gb[2], xp[2];
    float2 gw[6];
    wm_load_plain(ga, P.a.G + boff, P.a.g_cols, rows_left, L);
    if (HAS_B) wm_load_plain(gb, P.b.G + boff, P.b.g_cols, rows_left, L);
    if (!GATHER) {
        wm_load_plain(xp, P.a.X0 + boff, P.a.x0_valid, rows_left, L);
    } else if (KIND == 0) {
        wm_load_plain(xp, P.a.X0 + boff, P.a.x0_valid, rows_left, L);
        const long long off[2] = {noff[0], noff[1]};
        wm_load_gather(gw, P.b, off, row0, R, tg);
        wm_gather_offsets(noff, P.b, row0 + 2 * WM_ROWS, R, tg);      // this group's next gathered unit: two quarter tiles on
    } else {
        const long long off[2] = {noff[0], noff[1]};
        wm_load_gather(gw, P.a, off, row0, R, tg);
        wm_gather_offsets(noff, P.a, row0 + 2 * WM_ROWS, R, tg);
    }
    float4 xq[2];
    if (!GATHER) wm_load_plain(xq, P.b.X0 + boff, P.b.x0_valid, rows_left, L);
    const long long c0 = clock64();
    if (u >= WM_STAGES) mbar_wait(&S.empty[s], (unsigned)(((u / WM_STAGES) - 1) & 1));
    const long long c1 = clock64();
    pw0 += c1 - c0; pw1 += c0 - ca0;
    wm_store_plain(st.Ahi, st.Alo, 0, ga, L);
    if (HAS_B) wm_store_plain(st.Ahi, st.Alo, 64, gb, L);
    else { const float4 z[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)}; wm_store_plain(st.Ahi, st.Alo, 64, z, L); }
    if (!GATHER) {
        wm_store_plain(st.Bhi, st.Blo, 0, xp, L);
        wm_store_plain(st.Bhi, st.Blo, 64, xq, L);
    } else if (KIND == 0) {
        wm_store_plain(st.Bhi, st.Blo, 0, xp, L);
        wm_store_gather(st.Bhi, st.Blo, 64, gw, tg);
    } else {
        wm_store_gather(st.Bhi, st.Blo, 0, gw, tg);
    }
    proxy_fence_async();
    mbar_arrive(&S.full[s]);
    pw2 += clock64() - c1;
}

template <int KIND>
__global__ void __launch_bounds__(WM_NT, 1) k_wgrad_mn(const __grid_constant__ WmArgs A) {
    extern __shared__ unsigned char wm_raw[];
    WmSmem& S = *reinterpret_cast<WmSmem*>(wm_raw + ((1024u - (smem_u32(wm_raw) & 1023u)) & 1023u));   // 1024-byte aligned stages
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    pdl_launch_dependents();   // latency regime only (launch_pdl): the next kernel's prologue may overlap this kernel's tail
    if (t == 0) {
        for (int i = 0; i < WM_STAGES; ++i) { mbar_init(&S.full[i], WM_GT); mbar_init(&S.empty[i], 1); }
        mbar_init(&S.done, 1);
        S.consumed = 0;
    }
    if (warp == WM_GROUPS * (WM_GT / 32)) tmem_alloc<512>(&S.tmem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();                // everything below reads the predecessors' stashes
    const long long R = A.nrows_dev ? (long long)*A.nrows_dev : (long long)A.nrows;
    const uint32_t tmem = S.tmem;
    WUnit it; it.init(R, 3);                               // 3 pairs per quarter tile; unit u = 3 * (local quarter tile) + pair
    const bool has_work = it.valid();
    const int nq = (int)(it.qt_end - it.qt);
    const bool prof = A.prof && blockIdx.x == 0 && lane == 0;
    long long pw0 = 0, pw1 = 0, pw2 = 0;
    const long long pt0 = clock64();
    if (warp < WM_GROUPS * (WM_GT / 32)) {
        // ---- loaders: group g takes the units u = g, g + 2, ...; unit u goes to stage u % 3.  Pair 2 (the gathered, slow one)
        //      falls to group (quarter tile) % 2: the groups take it in turn ----
        const int g = warp / (WM_GT / 32), tg = t - g * WM_GT;
        WmLane L; L.init(tg);
        long long noff[2] = {0, 0};
        if (g < nq) wm_gather_offsets(noff, KIND == 0 ? A.pair[2].b : A.pair[2].a, (it.qt + g) * WM_ROWS, R, tg);   // first gathered unit of this group
        for (int ql = 0; ql < nq; ++ql) {
            const long long row0 = (it.qt + ql) * WM_ROWS;
            const int u = 3 * ql;
            if (((u + 0) & 1) == g) wm_unit<KIND, 0>(A, S, u + 0, row0, R, L, tg, noff, pw0, pw1, pw2);
            if (((u + 1) & 1) == g) wm_unit<KIND, 1>(A, S, u + 1, row0, R, L, tg, noff, pw0, pw1, pw2);
            if (((u + 2) & 1) == g) wm_unit<KIND, 2>(A, S, u + 2, row0, R, L, tg, noff, pw0, pw1, pw2);
        }
        if (prof && warp == 0) { A.prof[10] = pw0; A.prof[11] = pw1; A.prof[12] = pw2; A.prof[13] = (3 * nq + 1 - g) / 2; A.prof[14] = clock64() - pt0; }
    } else if (warp == WM_GROUPS * (WM_GT / 32) + 1) {
        // ---- L2 prefetch: the stash blocks of the units WM_AHEAD ahead of the MMA warp (one bulk prefetch per block).  The
        //      loaders keep one unit per group in flight (registers are the limit); with the blocks already in L2 their loads
        //      return in a fraction of the HBM latency
        if (lane < 4) {
            int u = 0;
            for (; it.valid(); it.next(), ++u) {
                while (u >= *(volatile int*)&S.consumed + WM_AHEAD) __nanosleep(256);
                const WPair& P = A.pair[it.j];
                const float* src = lane == 0 ? P.a.G : lane == 1 ? (P.has_b ? P.b.G : nullptr) : lane == 2 ? P.a.X0 : (P.has_b ? P.b.X0 : nullptr);
                if (src) prefetch_l2_bulk(src + it.row0() * HSLOT, WM_ROWS * HSLOT * 4);
            }
        }
    } else if (lane == 0) {
        // ---- MMA issue: per unit 4 k-steps x (hi hi + hi lo + lo hi), M = 128, N = 32 (ca + cb) ----
        const uint32_t hiw = (512u >> 4) | (1u << 14) | (1u << 29);     // SBO = 512 B, descriptor version 1, SWIZZLE_128B_BASE32B
        const uint32_t lbo = (uint32_t)(WM_CHUNK * 4 >> 4) << 16;       // LBO = chunk stride
        uint32_t idesc[3], d_tm[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) { idesc[j] = umma_idesc_tf32_mn(128, 32 * (A.pair[j].ca + A.pair[j].cb)); d_tm[j] = tmem + A.pair[j].tmem_col; }
        const int nu = 3 * nq;
        int s = 0;
        for (int u = 0; u < nu; u += 3) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                const WmStage& st = S.st[s];
                const uint32_t a_lo[2] = {((smem_u32(st.Ahi) >> 4) & 0x3FFFu) | lbo, ((smem_u32(st.Alo) >> 4) & 0x3FFFu) | lbo};
                const uint32_t b_lo[2] = {((smem_u32(st.Bhi) >> 4) & 0x3FFFu) | lbo, ((smem_u32(st.Blo) >> 4) & 0x3FFFu) | lbo};
                const long long c0 = clock64();
                mbar_wait(&S.full[s], (unsigned)(((u + j) / WM_STAGES) & 1));
                const long long c1 = clock64();
                pw0 += c1 - c0;
                tc_fence_after();
                uint32_t acc = u > 0;
#pragma unroll
                for (int pass = 0; pass < 3; ++pass) {
                    const uint32_t a0 = a_lo[pass == 2], b0 = b_lo[pass == 1];
#pragma unroll
                    for (int ks = 0; ks < WM_ROWS / 8; ++ks) {           // a k-step = 8 rows = 1024 bytes of every chunk
                        const uint64_t da = ((uint64_t)hiw << 32) | (a0 + ks * (1024u >> 4));
                        const uint64_t db = ((uint64_t)hiw << 32) | (b0 + ks * (1024u >> 4));
                        umma_tf32(d_tm[j], da, db, idesc[j], acc);
                        acc = 1u;
                    }
                }
                umma_commit(&S.empty[s]);
                *(volatile int*)&S.consumed = u + j + 1;
                pw1 += clock64() - c1;
                s = s == WM_STAGES - 1 ? 0 : s + 1;
            }
        }
        umma_commit(&S.done);
        if (prof) { A.prof[15] = pw0; A.prof[16] = pw1; A.prof[17] = clock64() - pt0; }
    }
    // ---- epilogue: warp w < 12 takes lane quadrant w % 4 of pair w / 4; an M = 128 accumulator keeps row n in lane n ----
    const long long pe0 = clock64();
    if (warp < 12) {
        const int q = warp & 3, p = warp >> 2;
        const WPair& P = A.pair[p];
        mbar_wait(&S.done, 0);
        tc_fence_after();
        if (prof && warp == 0) A.prof[20] = clock64() - pe0;
        if (q < 2 || P.has_b) {
            const WJob& J = q < 2 ? P.a : P.b;
            const int col0 = P.tmem_col + (q < 2 ? 0 : 32 * P.ca);
            const int n = 32 * (q & 1) + lane;
            float* out = A.partial + (size_t)blockIdx.x * PT_TOTAL;
            float d[96];
            if (has_work) {
                tmem_ld16(tmem + col0, 0, d); tmem_ld16(tmem + col0, 16, d + 16);
                tmem_ld16(tmem + col0, 32, d + 32); tmem_ld16(tmem + col0, 48, d + 48);
                if (J.nb > 64) { tmem_ld16(tmem + col0, 64, d + 64); tmem_ld16(tmem + col0, 80, d + 80); }
                tmem_ld_wait();
            }
            wgrad_store_rows(out, J, n, d, !has_work);
        }
    }
    if (prof && warp == 0) { A.prof[18] = clock64() - pe0; A.prof[19] = clock64() - pt0; }
    tc_fence_before();
    __syncthreads();
    if (warp == WM_GROUPS * (WM_GT / 32)) tmem_dealloc<512>(tmem);
}

}  // namespace npe
