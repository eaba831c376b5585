// tcgen05 backward of the NPE for the throughput regime (model:bp, reference src/lua/npe.lua:290-336).
//
// The backward of a large batch is three dense contractions per layer on tiles of 128 rows (active pairs or foci):
//
//   input gradients   dH[r][k]  = sum_n dZ[r][n] * W[n][k]        k_dec_dgrad_tc3 / k_pair_dgrad_tc3
//   ReLU derivative   dZ'[r][k] = dH[r][k] * (h[r][k] > 0)        (epilogue of the same kernels; h from the forward stash)
//   weight gradients  dW[n][k]  = sum_r dZ[r][n] * X[r][k]        k_wgrad_tc
//
// All three run as 3xTF32 (hi*hi + hi*lo + lo*hi, FP32 accumulation in tensor memory): FP32-grade results, same
// building blocks as the 3xTF32 forward (npe_tc.cuh, npe_umma.cuh).
//
//  * dgrad chains: like the forward chain, the A operand (dZ, split into hi / lo) lives in TENSOR MEMORY and the B
//    operand is a shared-memory weight image -- here the TRANSPOSED image (row = input index k, K = output index n),
//    hi and lo, resident for the whole kernel.  Every thread owns one row: it reads the accumulator from TMEM, applies
//    the ReLU mask with its row of the forward stash, writes dZ of that layer to the gradient stash (consumed by the
//    weight-gradient kernel) and writes the hi / lo split back to TMEM as the next layer's A operand.
//  * wgrad: the reduction index is the ROW, so both operands must be K-major in r.  Producer warps read G = dZ and X
//    (stash rows, or gathered state rows) and store them transposed into the canonical no-swizzle K-major layout
//        element (n, r)  ->  ((r / 4) * ROWS + n) * 4 + (r % 4)   floats
//    with a lane mapping (4 rows x 8 columns per warp instruction) that makes every store instruction write 32
//    CONSECUTIVE floats (conflict-free by construction).  One elected thread of a dedicated warp issues
//    tcgen05.mma.kind::tf32 with M = 64 (n), N = 56 / 96 (k), K = 8 per instruction over a double-buffered stage ring
//    (full / empty mbarriers; tcgen05.commit releases a stage).  The accumulators of ALL layers of a CTA stay in tensor
//    memory across its tiles (5 x 64 + 96 columns); one epilogue at the end writes them to the CTA's gradient partial in
//    the tile-major layout that k_reduce_grads / k_dp_step already consume (fixed order -> deterministic).
#pragma once
#include "npe_tc.cuh"

namespace npe {

// ---- transposed weight images for the input-gradient chains: element (k_in, n_out) at ((n >> 2) * 64 + k) * 4 + (n & 3) ----
constexpr int TT_PAIR = 0;                                   // 5 x [14 K-chunks (n) x 64 rows (k)]
constexpr int TT_PAIR_TOTAL = NL * TC_HID_SZ;                // 17920 floats = 71680 B
constexpr int TT_DEC5 = 0;                                   // [6 K-chunks (n < 24) x 64 rows (k)]
constexpr int TT_DEC5_SZ = 6 * 64 * 4;                       // 1536
constexpr int TT_DEC2 = TT_DEC5 + TT_DEC5_SZ;                // dec layers 2..4
constexpr int TT_DEC1 = TT_DEC2 + 3 * TC_HID_SZ;             // rows k < 50 only: the gradient into the pair sum
constexpr int TT_DEC_TOTAL = TT_DEC1 + TC_HID_SZ;            // 15872 floats = 63488 B
constexpr int TT_TOTAL = TT_PAIR_TOTAL + TT_DEC_TOTAL;

__host__ __device__ __forceinline__ int tt_off(int base, int k, int n) { return base + ((n >> 2) * 64 + k) * 4 + (n & 3); }

// flat parameter index -> offset in the transposed image, -1 for parameters no input-gradient chain reads (encoders:
// no gradient into raw inputs; biases; decoder layer 1 columns of this_past)
__host__ __device__ __forceinline__ int tt_index(int i) {
    if (i < OFF_PAIR) return -1;
    if (i < OFF_DEC1_W) {
        const int r0 = i - OFF_PAIR, l = r0 / 2500, r = r0 - l * 2500, n = r / H, k = r - n * H;
        return tt_off(TT_PAIR + l * TC_HID_SZ, k, n);
    }
    const int base = TT_PAIR_TOTAL;
    if (i < OFF_DEC1_B) {
        const int r = i - OFF_DEC1_W, n = r / DEC_IN, k = r - n * DEC_IN;
        return k < H ? tt_off(base + TT_DEC1, k, n) : -1;
    }
    if (i < OFF_DEC2_W) return -1;
    if (i < OFF_DEC5_W) {
        const int j = i - OFF_DEC2_W, li = j / 2550, r = j - li * 2550;
        if (r >= 2500) return -1;
        const int n = r / H, k = r - n * H;
        return tt_off(base + TT_DEC2 + li * TC_HID_SZ, k, n);
    }
    if (i < OFF_DEC5_B) { const int r = i - OFF_DEC5_W, n = r / H, k = r - n * H; return tt_off(base + TT_DEC5, k, n); }
    return -1;
}

__global__ void k_pack_tt3_scatter(const float* __restrict__ params, float* __restrict__ w_hi, float* __restrict__ w_lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NPARAM) return;
    const int o = tt_index(i);
    if (o < 0) return;
    const float x = params[i], h = tf32_hi(x);
    w_hi[o] = h;
    w_lo[o] = x - h;
}

// ---- stashes (row-major, one slot of HSLOT = 56 floats per layer, 52 used; npe_kernels.cuh) ----
// pair:     hact  [pair ][6 slots]  h0..h5 (forward), dzact [pair ][6 slots]  dz0..dz5 (pair dgrad)
// decoder:  dact  [focus][5 slots]  {z[0..51] = pair sum | 1 | 0, u1..u4} (forward), ddz [focus][5 slots] dz1..dz5 (dec dgrad)
// ReLU bits (one 64-bit word per row and layer, layer-major -> coalesced): hmask [6][cap], umask [4][F]

__device__ __forceinline__ void row_load52(float* d, const float* __restrict__ src) {
#pragma unroll
    for (int c = 0; c < 13; ++c) {
        const float4 v = ld4(src + 4 * c);
        d[4 * c] = v.x; d[4 * c + 1] = v.y; d[4 * c + 2] = v.z; d[4 * c + 3] = v.w;
    }
}
__device__ __forceinline__ void row_store52(float* __restrict__ dst, const float* d) {
#pragma unroll
    for (int c = 0; c < 13; ++c) st4(dst + 4 * c, make_float4(d[4 * c], d[4 * c + 1], d[4 * c + 2], d[4 * c + 3]));
}
// accumulator row (64 columns) -> d
__device__ __forceinline__ void tc_load64(uint32_t tm, float* d) {
    tmem_ld16(tm, 0, d); tmem_ld16(tm, 16, d + 16); tmem_ld16(tm, 32, d + 32); tmem_ld16(tm, 48, d + 48);
    tmem_ld_wait();
}
// d[k] = d[k] * (h[k] > 0) for k < 50 with the forward's ReLU bits, zero beyond (column 50 is the constant-1 column of
// the decoder: no gradient)
__device__ __forceinline__ void relu_mask50(float* d, unsigned long long bits) {
    const unsigned lo = (unsigned)bits, hi = (unsigned)(bits >> 32);
#pragma unroll
    for (int k = 0; k < 56; ++k) d[k] = (k < H && (((k < 32 ? lo : hi) >> (k & 31)) & 1u)) ? d[k] : 0.f;
}

// ----------------------------------------------------------------------------------------------------------------
// weight gradients: dW_j[n][k] = sum_r G_j[r][n] * X_j[r][k] for a list of jobs that share the row index r.
//
// Pipeline per CTA (persistent over 128-row tiles; unit = 32 rows of one job):
//   TMA warp        one thread issues cp.async.bulk for the unit's G block and stash X block (32 rows x 224 B, contiguous in the
//                   layer-major stashes) into a 6-deep raw ring in shared memory (raw_full / raw_empty mbarriers)
//   8 transposer    read the raw blocks (conflict-free: pitch 56), split hi / lo and store them TRANSPOSED into the K-major
//   warps           operand images of a 2-deep operand ring; lane mapping 4 rows x 8 columns per instruction, so every store
//                   instruction writes 32 consecutive floats.  Gathered X segments (state windows addressed through a row
//                   offset table) are fetched with LDG one unit ahead.
//   MMA warp        12 x tcgen05.mma.kind::tf32 (M = 64, N = 56 / 96, K = 8; hi*hi + hi*lo + lo*hi) per unit into the job's
//                   accumulator in tensor memory; tcgen05.commit frees the operand stage
// The accumulators of all jobs stay in TMEM across the CTA's tiles; one epilogue writes them to the CTA's partial.
// ----------------------------------------------------------------------------------------------------------------
struct WGather {
    const float* base;            // value(R, c) = base[offs[R] + c]
    const long long* offs;
    int cols_valid;               // columns c >= cols_valid read as zero
    int cols_store;               // columns c < cols_store are written to the B image at column col0 + c
    int col0;
};
struct WJob {
    const float* G; int g_cols;                       // gradient rows: stash slot [rows][HSLOT]; columns >= g_cols are zero
    const float* X0; int x0_valid, x0_store;          // input rows from a stash slot (or nullptr): B image columns [0, x0_store)
    WGather gs[2]; int ngs;                           // gathered input segments
    int nb;                                           // rows of the B image (k): 56 or 96
    int tmem_col;                                     // accumulator columns [tmem_col, tmem_col + nb)
    int kind;                                         // epilogue: 0 uniform layer, 1 decoder layer 1 (tiles split A / B), 2 encoders
    int pt_base, kt, n_valid;                         // kind 0: partial base, tiles per row, valid accumulator rows
};
constexpr int WG_MAXJOBS = 6;
struct WArgs {
    WJob job[WG_MAXJOBS];
    int njobs;
    const int* nrows_dev;         // number of rows R (pairs: info[0]) or nullptr
    int nrows;                    // used when nrows_dev == nullptr
    float* partial;               // [gridDim.x][PT_TOTAL]
    long long* prof;              // tools only (NPE_CHAIN_PROF): cycle counters of CTA 0, or nullptr
};

constexpr int WG_ROWS = 32;       // rows per unit: K = 32 = 4 MMA k-steps
constexpr int WG_NRAW = 4;        // raw ring depth
constexpr int WG_NOP = 3;         // operand ring depth: transposers run up to 2 units ahead of the MMA warp
constexpr int WG_NT = 352;        // 8 transposer warps + MMA warp + TMA warp + gather warp
constexpr int WG_BLOCK = WG_ROWS * HSLOT;            // floats of one stash block (7168 B)
constexpr int WG_XG_PITCH = 98;   // floats: pitch of gathered X rows in the raw stage (8-byte stores of 32 lanes: conflict-free)
struct WRaw { float G[WG_BLOCK]; float X[WG_ROWS * WG_XG_PITCH]; };   // 7168 + 12544 B
struct WOp {
    float Ahi[64 * WG_ROWS], Alo[64 * WG_ROWS];       // G^T: [r / 4][n (64)][r % 4]
    float Bhi[96 * WG_ROWS], Blo[96 * WG_ROWS];       // X^T: [r / 4][k (nb)][r % 4]
};
struct WgradSmem {
    WOp op[WG_NOP];                                   // 3 x 40960 B
    WRaw raw[WG_NRAW];                                // 4 x 19712 B
    unsigned long long raw_full[WG_NRAW], raw_empty[WG_NRAW], op_full[WG_NOP], op_empty[WG_NOP], done;
    uint32_t tmem, pad;
};

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_only(unsigned long long* bar, unsigned bytes) {   // no arrival
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t umma_idesc_tf32_m64(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
}

// Unit iterator shared by the three roles.  The R rows are cut into 32-row quarter tiles; CTA b owns a contiguous, balanced
// range of them and runs every job on each (unit = quarter tile x job).
struct WUnit {
    long long qt, qt_end;
    int j, njobs;
    __device__ __forceinline__ void init(long long R, int njobs_) {
        const long long nq = (R + WG_ROWS - 1) / WG_ROWS;
        qt = nq * blockIdx.x / gridDim.x; qt_end = nq * (blockIdx.x + 1) / gridDim.x;
        j = 0; njobs = njobs_;
    }
    __device__ __forceinline__ bool valid() const { return qt < qt_end; }
    __device__ __forceinline__ long long row0() const { return qt * WG_ROWS; }
    __device__ __forceinline__ void next() { if (++j == njobs) { j = 0; ++qt; } }
};

__device__ __forceinline__ void wgrad_put(float* hi_img, float* lo_img, int off, const float (&v)[4]) {
    float4 h, l;
    h.x = tf32_hi(v[0]); h.y = tf32_hi(v[1]); h.z = tf32_hi(v[2]); h.w = tf32_hi(v[3]);
    l.x = v[0] - h.x; l.y = v[1] - h.y; l.z = v[2] - h.z; l.w = v[3] - h.w;
    st4(hi_img + off, h); st4(lo_img + off, l);
}
// Transposer work items of a unit: item i = (column n = i % ncols, row group rg = i / ncols), rows 4 rg .. 4 rg + 3.  A thread
// reads the four rows of its column (consecutive lanes = consecutive columns: conflict-free) and writes ONE 16-byte store
// per image: element (n, 4 rg .. 4 rg + 3) is contiguous in the K-major operand layout.
template <int NITEMS>
__device__ __forceinline__ void wgrad_transpose_block(float* hi_img, float* lo_img, int img_rows, int img_col0, const float* src,
                                                      int pitch, int ncols, int cols_valid, int rows_left) {
    const int t = threadIdx.x;
#pragma unroll
    for (int k = 0; k < NITEMS; ++k) {
        const int i = t + 256 * k, rg = i / ncols, n = i - rg * ncols;
        if (rg < 8) {
            float v[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float x = src[(4 * rg + r) * pitch + n];
                v[r] = (4 * rg + r < rows_left && n < cols_valid) ? x : 0.f;
            }
            wgrad_put(hi_img, lo_img, (rg * img_rows + img_col0 + n) * 4, v);
        }
    }
}
// raw blocks -> hi / lo operand images of one operand stage
__device__ __forceinline__ void wgrad_transpose(WOp& op, const WRaw& raw, const WJob& J, long long row0, long long R) {
    const int rows_left = (int)min((long long)WG_ROWS, R - row0);       // rows beyond R hold whatever the copy found: zero
    wgrad_transpose_block<2>(op.Ahi, op.Alo, 64, 0, raw.G, HSLOT, 56, J.g_cols, rows_left);
    if (J.ngs) wgrad_transpose_block<3>(op.Bhi, op.Blo, 96, 0, raw.X, WG_XG_PITCH, 96, 96, rows_left);   // assembled by the producer warp
    else wgrad_transpose_block<2>(op.Bhi, op.Blo, J.nb, 0, raw.X, HSLOT, 56, J.x0_valid, rows_left);
}

// Producer warp, gathered jobs: lane = row.  Assembles the 96-column X row in the raw stage (pitch 98 floats):
//   encoders         [focus window (44) | 0 x 4 | context window (44) | 0 x 4]      (row offsets gs[0].offs / gs[1].offs)
//   decoder layer 1  [stash row z[0..51] | focus window (44)]                        (stash X0, row offsets gs[0].offs)
__device__ __forceinline__ void wgrad_assemble(float* xs, const WJob& J, long long row0, long long R) {
    const int lane = threadIdx.x & 31;
    const long long row = row0 + lane;
    const bool ok = row < R;
    float* o = xs + lane * WG_XG_PITCH;
    float2 a[26], b[24];
    const float2 z2 = make_float2(0.f, 0.f);
    if (J.ngs == 2) {
        const float* pa = J.gs[0].base + (ok ? J.gs[0].offs[row] : 0);
        const float* pb = J.gs[1].base + (ok ? J.gs[1].offs[row] : 0);
#pragma unroll
        for (int c = 0; c < 24; ++c) { a[c] = (ok && c < 22) ? ld2(pa + 2 * c) : z2; b[c] = (ok && c < 22) ? ld2(pb + 2 * c) : z2; }
    } else {
        const float* pa = J.X0 + (ok ? row * HSLOT : 0);
        const float* pb = J.gs[0].base + (ok ? J.gs[0].offs[row] : 0);
#pragma unroll
        for (int c = 0; c < 13; ++c) {
            const float4 v = ok ? ld4(pa + 4 * c) : make_float4(0.f, 0.f, 0.f, 0.f);
            a[2 * c] = make_float2(v.x, v.y); a[2 * c + 1] = make_float2(v.z, v.w);
        }
#pragma unroll
        for (int c = 0; c < 22; ++c) b[c] = ok ? ld2(pb + 2 * c) : z2;
    }
    if (J.ngs == 2) {
#pragma unroll
        for (int c = 0; c < 24; ++c) { *reinterpret_cast<float2*>(o + 2 * c) = a[c]; *reinterpret_cast<float2*>(o + 48 + 2 * c) = b[c]; }
    } else {
#pragma unroll
        for (int c = 0; c < 26; ++c) *reinterpret_cast<float2*>(o + 2 * c) = a[c];
#pragma unroll
        for (int c = 0; c < 22; ++c) *reinterpret_cast<float2*>(o + 52 + 2 * c) = b[c];
    }
}

// Accumulator row n (d[0 .. nb)) -> the CTA's partial in the tile-major layout of npe_layout.h (what partial_index addresses).
// All register indices are static (no local memory).
__device__ __forceinline__ void wgrad_store_rows(float* __restrict__ out, const WJob& J, int n, const float* d, bool zero) {
    const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (J.kind == 2) {   // encoders: G = dz0; rows 0..24 x cols 0..43 -> focus encoder, rows 25..49 x cols 48..91 -> context
        if (n >= H) return;
        if (n < H2) {
            float* o = out + PT_ENC_T + (n >> 2) * 11 * 16 + 4 * (n & 3);
#pragma unroll
            for (int kc = 0; kc < 11; ++kc) st4(o + kc * 16, zero ? z4 : make_float4(d[4 * kc], d[4 * kc + 1], d[4 * kc + 2], d[4 * kc + 3]));
        } else {
            const int nn = n - H2;
            float* o = out + PT_ENC_C + (nn >> 2) * 11 * 16 + 4 * (nn & 3);
#pragma unroll
            for (int kc = 0; kc < 11; ++kc)
                st4(o + kc * 16, zero ? z4 : make_float4(d[48 + 4 * kc], d[49 + 4 * kc], d[50 + 4 * kc], d[51 + 4 * kc]));
        }
        return;
    }
    if (n >= J.n_valid) return;
    const int nk = min(J.nb >> 2, J.kt);
#pragma unroll
    for (int kc = 0; kc < 24; ++kc) {
        if (kc < nk) {
            float* o;
            if (J.kind == 1) {
                const int tile = (n >> 2) * 24 + kc;
                o = out + (tile < NT ? PT_DEC1A + tile * 16 : PT_DEC1B + (tile - NT) * 16) + 4 * (n & 3);
            } else {
                o = out + J.pt_base + ((n >> 2) * J.kt + kc) * 16 + 4 * (n & 3);
            }
            st4(o, zero ? z4 : make_float4(d[4 * kc], d[4 * kc + 1], d[4 * kc + 2], d[4 * kc + 3]));
        }
    }
}

__global__ void __launch_bounds__(WG_NT, 1) k_wgrad_tc(WArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WgradSmem& S = *reinterpret_cast<WgradSmem*>(smem_raw);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    pdl_launch_dependents();   // the next kernel's prologue may overlap this kernel's tail
    if (t == 0) {
        for (int i = 0; i < WG_NRAW; ++i) { mbar_init(&S.raw_full[i], 2); mbar_init(&S.raw_empty[i], 256); }
        for (int i = 0; i < WG_NOP; ++i) { mbar_init(&S.op_full[i], 256); mbar_init(&S.op_empty[i], 1); }
        mbar_init(&S.done, 1);
    }
    if (warp == 8) tmem_alloc<512>(&S.tmem);
    // zero the A images once: rows 56..63 are never written and must read as zero
    for (int i = t; i < WG_NOP * 2 * 64 * WG_ROWS / 4; i += WG_NT) {
        const int o = i / (2 * 64 * WG_ROWS / 4), r = i - o * (2 * 64 * WG_ROWS / 4);
        reinterpret_cast<float4*>(S.op[o].Ahi)[r] = make_float4(0.f, 0.f, 0.f, 0.f);      // Ahi and Alo are adjacent
    }
    proxy_fence_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();                // everything below reads the predecessors' stashes
    const long long R = A.nrows_dev ? (long long)*A.nrows_dev : (long long)A.nrows;
    const uint32_t tmem = S.tmem;
    WUnit it; it.init(R, A.njobs);
    const bool has_work = it.valid();
    const bool prof = A.prof && blockIdx.x == 0 && lane == 0;
    long long pw0 = 0, pw1 = 0, pw2 = 0, pt0 = clock64();
    if (warp == 9) {   // TMA warp: bulk copies of the stash blocks of every unit
        if (lane == 0) {
            for (int u = 0; it.valid(); it.next(), ++u) {
                const int s = u % WG_NRAW;
                const WJob& J = A.job[it.j];
                const long long c0 = clock64();
                if (u >= WG_NRAW) mbar_wait(&S.raw_empty[s], (unsigned)(((u / WG_NRAW) - 1) & 1));
                pw0 += clock64() - c0;
                mbar_expect_tx(&S.raw_full[s], (J.ngs ? 1u : 2u) * WG_BLOCK * 4);      // arrival 1 of 2
                bulk_g2s(S.raw[s].G, J.G + it.row0() * HSLOT, WG_BLOCK * 4, &S.raw_full[s]);
                if (!J.ngs) bulk_g2s(S.raw[s].X, J.X0 + it.row0() * HSLOT, WG_BLOCK * 4, &S.raw_full[s]);
            }
            if (prof) { A.prof[8] = pw0; A.prof[9] = clock64() - pt0; }
        }
    } else if (warp == 10) {   // gather warp (lane = row): assembles the X rows of gathered jobs; arrival 2 of 2 of every unit
        for (int u = 0; it.valid(); it.next(), ++u) {
            const int s = u % WG_NRAW;
            const WJob& J = A.job[it.j];
            // the stage must be free before EITHER arrival of its next use: an early arrival would complete the previous phase
            if (u >= WG_NRAW) mbar_wait(&S.raw_empty[s], (unsigned)(((u / WG_NRAW) - 1) & 1));
            if (J.ngs) {
                wgrad_assemble(S.raw[s].X, J, it.row0(), R);
                __syncwarp();
            }
            if (lane == 0) mbar_arrive(&S.raw_full[s]);
        }
    } else if (warp < 8) {
        int u = 0;
        for (; it.valid(); it.next(), ++u) {
            const int s = u % WG_NRAW, o = u % WG_NOP;
            const long long c0 = clock64();
            mbar_wait(&S.raw_full[s], (unsigned)((u / WG_NRAW) & 1));
            const long long c1 = clock64();
            if (u >= WG_NOP) mbar_wait(&S.op_empty[o], (unsigned)(((u / WG_NOP) - 1) & 1));
            const long long c2 = clock64();
            wgrad_transpose(S.op[o], S.raw[s], A.job[it.j], it.row0(), R);
            proxy_fence_async();
            mbar_arrive(&S.op_full[o]);
            mbar_arrive(&S.raw_empty[s]);
            pw0 += c1 - c0; pw1 += c2 - c1; pw2 += clock64() - c2;
        }
        if (prof && warp == 0) { A.prof[10] = pw0; A.prof[11] = pw1; A.prof[12] = pw2; A.prof[13] = u; A.prof[14] = clock64() - pt0; }
    } else if (lane == 0) {   // warp 8: MMA issue
        unsigned started = 0;
        for (int u = 0; it.valid(); it.next(), ++u) {
            const WJob& J = A.job[it.j];
            const int nb = J.nb, o = u % WG_NOP;
            const uint32_t idesc = umma_idesc_tf32_m64(nb);
            const uint32_t d_tm = tmem + J.tmem_col;
            // descriptors: low word = address field | LBO << 16, high word = SBO | version; a K-step advances the address field
            const uint32_t hiw = (128u >> 4) | (1u << 14);
            const uint32_t a_lo[2] = {((smem_u32(S.op[o].Ahi) >> 4) & 0x3FFFu) | ((64u * 16u >> 4) << 16),
                                      ((smem_u32(S.op[o].Alo) >> 4) & 0x3FFFu) | ((64u * 16u >> 4) << 16)};
            const uint32_t b_lo[2] = {((smem_u32(S.op[o].Bhi) >> 4) & 0x3FFFu) | (((uint32_t)nb * 16u >> 4) << 16),
                                      ((smem_u32(S.op[o].Blo) >> 4) & 0x3FFFu) | (((uint32_t)nb * 16u >> 4) << 16)};
            const uint32_t a_step = 2u * 64u * 16u >> 4, b_step = 2u * (uint32_t)nb * 16u >> 4;
            const long long c0 = clock64();
            mbar_wait(&S.op_full[o], (unsigned)((u / WG_NOP) & 1));
            pw0 += clock64() - c0;
            const long long c1 = clock64();
            tc_fence_after();
            uint32_t acc = (started >> it.j) & 1u;
#pragma unroll
            for (int pass = 0; pass < 3; ++pass) {
                const uint32_t a0 = a_lo[pass == 2], b0 = b_lo[pass == 1];
#pragma unroll
                for (int ks = 0; ks < WG_ROWS / 8; ++ks) {
                    const uint64_t da = ((uint64_t)hiw << 32) | (a0 + ks * a_step);
                    const uint64_t db = ((uint64_t)hiw << 32) | (b0 + ks * b_step);
                    umma_tf32(d_tm, da, db, idesc, acc);
                    acc = 1u;
                }
            }
            started |= 1u << it.j;
            umma_commit(&S.op_empty[o]);
            pw1 += clock64() - c1;
        }
        umma_commit(&S.done);
        if (prof) { A.prof[15] = pw0; A.prof[16] = pw1; A.prof[17] = clock64() - pt0; }
    }
    const long long pe0 = clock64();
    // epilogue: warps 0..3 own the four TMEM lane quadrants; an M = 64 accumulator keeps row n in lane (n % 16) + 32 (n / 16)
    if (warp < 4) {
        mbar_wait(&S.done, 0);
        tc_fence_after();
        if (prof && warp == 0) A.prof[20] = clock64() - pe0;
        float* out = A.partial + (size_t)blockIdx.x * PT_TOTAL;
        const int n = 16 * warp + lane;
        for (int j = 0; j < A.njobs; ++j) {
            const WJob& J = A.job[j];
            float d[96];
            if (has_work) {
                tmem_ld16(tmem + J.tmem_col, 0, d); tmem_ld16(tmem + J.tmem_col, 16, d + 16);
                tmem_ld16(tmem + J.tmem_col, 32, d + 32); tmem_ld16(tmem + J.tmem_col, 48, d + 48);
                if (J.nb > 64) { tmem_ld16(tmem + J.tmem_col, 64, d + 64); tmem_ld16(tmem + J.tmem_col, 80, d + 80); }
                tmem_ld_wait();
            }
            if (lane < 16) wgrad_store_rows(out, J, n, d, !has_work);
        }
    }
    if (prof && warp == 0) { A.prof[18] = clock64() - pe0; A.prof[19] = clock64() - pt0; }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<512>(tmem);
}

}  // namespace npe
