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
// weight gradients: dW_j[n][k] = sum_r G_j[r][n] * X_j[r][k] for a list of jobs (layers) that share the row index r.
// The job tables below describe where G and X of every layer live and where the result goes; the kernel is in
// npe_wgrad.cuh (operands stay row-major / MN-major, two layers per MMA, accumulators in tensor memory for the whole
// kernel, one epilogue into the CTA's tile-major partial).
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

constexpr int WG_ROWS = 32;       // rows per unit of the weight-gradient kernel: K = 32 = 4 MMA k-steps

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Unit iterator shared by the roles of the weight-gradient kernel (npe_wgrad.cuh).  The R rows are cut into 32-row quarter
// tiles; CTA b owns a contiguous, balanced range of them and runs every job on each (unit = quarter tile x job).
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

}  // namespace npe
