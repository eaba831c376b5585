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

// ---- stashes (row-major, one row of 52 floats per layer) ----
// pair:     hact  [pair ][6][52]  h0..h5 (forward), dzact [pair ][6][52]  dz0..dz5 (pair dgrad)
// decoder:  dact  [focus][5][52]  {z[0..51] = pair sum | 1 | 0, u1..u4} (forward), ddz [focus][5][52] dz1..dz5 (dec dgrad)
constexpr int DACT_STRIDE = 5 * 52;

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
// d[k] = d[k] * (h[k] > 0) for k < 50, zero beyond (column 50 is the constant-1 column of the decoder: no gradient)
__device__ __forceinline__ void relu_mask50(float* d, const float* h) {
#pragma unroll
    for (int k = 0; k < 56; ++k) d[k] = (k < H && h[k] > 0.f) ? d[k] : 0.f;
}

// ----------------------------------------------------------------------------------------------------------------
// decoder input-gradient chain on tiles of 128 foci: d_pred (npe.lua:301-320) -> dz5 -> ... -> dz1 -> ds
// ----------------------------------------------------------------------------------------------------------------
struct DecDgradSmem {
    float Whi[TT_DEC_TOTAL];       // 63488 B
    float Wlo[TT_DEC_TOTAL];       // 63488 B
    Tc3Ctl ctl;
};

__global__ void __launch_bounds__(NT, 1)
k_dec_dgrad_tc3(int F, const float* __restrict__ pred, const float* __restrict__ y, const float* __restrict__ dact,
                const float* __restrict__ wtt3, float scale_v, float scale_w, float* __restrict__ ddz,
                float* __restrict__ ds_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    DecDgradSmem& S = *reinterpret_cast<DecDgradSmem*>(smem_raw);
    const int t = threadIdx.x, warp = t >> 5, g = t >> 7, r = t & 127;
    if (t == 0) { mbar_init(&S.ctl.wbar, 1); mbar_init(&S.ctl.mbar[0], 1); mbar_init(&S.ctl.mbar[1], 1); }
    if (warp == 0) tmem_alloc<TM_COLS>(&S.ctl.tmem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = S.ctl.tmem + g * TM_GROUP;
    if (t == 0) {
        mbar_expect_tx(&S.ctl.wbar, 2 * TT_DEC_TOTAL * 4);
        bulk_g2s(S.Whi, wtt3 + TT_PAIR_TOTAL, TT_DEC_TOTAL * 4, &S.ctl.wbar);
        bulk_g2s(S.Wlo, wtt3 + TT_TOTAL + TT_PAIR_TOTAL, TT_DEC_TOTAL * 4, &S.ctl.wbar);
    }
    bool staged = false;
    uint32_t phase = 0;
    const int ntiles = (F + 127) / 128;
    for (int tile = 2 * blockIdx.x + g; tile < ntiles; tile += 2 * gridDim.x) {
        const int f0 = tile * 128;
        const bool live = r < min(128, F - f0);
        const size_t f = (size_t)(f0 + r);
        {   // dz5 = d_pred: columns 2, 3 (velocity) and 5 (angular velocity) only
            float d5[24];
#pragma unroll
            for (int k = 0; k < 24; ++k) d5[k] = 0.f;
            if (live) {
                const float* pp = pred + f * OUT;
                const float* yy = y + f * OUT;
                d5[2] = (pp[2] - yy[2]) * scale_v;
                d5[3] = (pp[3] - yy[3]) * scale_v;
                d5[5] = (pp[5] - yy[5]) * scale_w;
                float* o = ddz + f * DACT_STRIDE + 4 * 52;
#pragma unroll
                for (int c = 0; c < 6; ++c) st4(o + 4 * c, make_float4(d5[4 * c], d5[4 * c + 1], d5[4 * c + 2], d5[4 * c + 3]));
            }
            tmem_store_split<3>(tm, TM_AHI, TM_ALO, d5);
            tmem_st_wait();
        }
        if (!staged) { mbar_wait(&S.ctl.wbar, 0); staged = true; }
        tc_group_sync(g);
        // layers 5 .. 2: dh_{l-1} = dz_l W_l, dz_{l-1} = dh_{l-1} * (u_{l-1} > 0); then ds = dz1 W1[:, :50]
#pragma unroll 1
        for (int l = NL; l >= 1; --l) {
            if (r == 0) {
                if (l == NL) umma_gemm_tf32x3_ts(tm + TM_D, tm + TM_AHI, tm + TM_ALO, S.Whi + TT_DEC5, S.Wlo + TT_DEC5, 64, 64, 3);
                else if (l >= 2) umma_gemm_tf32x3_ts(tm + TM_D, tm + TM_AHI, tm + TM_ALO, S.Whi + TT_DEC2 + (l - 2) * TC_HID_SZ,
                                                     S.Wlo + TT_DEC2 + (l - 2) * TC_HID_SZ, 64, 64, 7);
                else umma_gemm_tf32x3_ts(tm + TM_D, tm + TM_AHI, tm + TM_ALO, S.Whi + TT_DEC1, S.Wlo + TT_DEC1, 64, 64, 7);
                umma_commit(&S.ctl.mbar[g]);
            }
            float h[52];
            if (l >= 2) {   // u_{l-1} of this focus from the forward stash, fetched while the MMAs run
                if (live) row_load52(h, dact + f * DACT_STRIDE + (l - 1) * 52);
                else {
#pragma unroll
                    for (int k = 0; k < 52; ++k) h[k] = 0.f;
                }
            }
            mbar_wait(&S.ctl.mbar[g], phase);
            tc_fence_after();
            float d[64];
            tc_load64(tm, d);
            if (l >= 2) {
                relu_mask50(d, h);
                if (live) row_store52(ddz + f * DACT_STRIDE + (l - 2) * 52, d);
                tmem_store_split<7>(tm, TM_AHI, TM_ALO, d);
                tmem_st_wait();
            } else if (live) {
                d[50] = 0.f; d[51] = 0.f;
                row_store52(ds_out + f * 52, d);
            }
            phase ^= 1;
            tc_group_sync(g);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TM_COLS>(S.ctl.tmem);
}

// ----------------------------------------------------------------------------------------------------------------
// pair-MLP input-gradient chain on tiles of 128 active pairs: dz5 = ds[focus] * (h5 > 0) -> ... -> dz0
// (CAddTable backward = copy; apply_mask is the identity on the active list, npe.lua:326-328)
// ----------------------------------------------------------------------------------------------------------------
struct PairDgradSmem {
    float Whi[TT_PAIR_TOTAL];      // 71680 B
    float Wlo[TT_PAIR_TOTAL];      // 71680 B
    Tc3Ctl ctl;
};

__global__ void __launch_bounds__(NT, 1)
k_pair_dgrad_tc3(const int* __restrict__ pair_foc, const long long* __restrict__ foc_row, const int* __restrict__ info,
                 const float* __restrict__ hact, const float* __restrict__ ds_in, const float* __restrict__ wtt3,
                 float* __restrict__ dzact, long long* __restrict__ pair_frow) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PairDgradSmem& S = *reinterpret_cast<PairDgradSmem*>(smem_raw);
    const int P = info[0];
    if ((int)blockIdx.x * 256 >= P) return;
    const int t = threadIdx.x, warp = t >> 5, g = t >> 7, r = t & 127;
    if (t == 0) { mbar_init(&S.ctl.wbar, 1); mbar_init(&S.ctl.mbar[0], 1); mbar_init(&S.ctl.mbar[1], 1); }
    if (warp == 0) tmem_alloc<TM_COLS>(&S.ctl.tmem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = S.ctl.tmem + g * TM_GROUP;
    if (t == 0) {
        mbar_expect_tx(&S.ctl.wbar, 2 * TT_PAIR_TOTAL * 4);
        bulk_g2s(S.Whi, wtt3, TT_PAIR_TOTAL * 4, &S.ctl.wbar);
        bulk_g2s(S.Wlo, wtt3 + TT_TOTAL, TT_PAIR_TOTAL * 4, &S.ctl.wbar);
    }
    bool staged = false;
    uint32_t phase = 0;
    for (int lo = (2 * blockIdx.x + g) * 128; lo < P; lo += 2 * gridDim.x * 128) {
        const bool live = r < min(128, P - lo);
        const size_t p = (size_t)(lo + r);
        {
            float d[56], h[52];
#pragma unroll
            for (int k = 0; k < 56; ++k) d[k] = 0.f;
            if (live) {
                const int pf = pair_foc[p];
                pair_frow[p] = foc_row[pf];                 // gather offset of the focus window for the encoder wgrad
                row_load52(d, ds_in + (size_t)pf * 52);
                row_load52(h, hact + p * HACT_STRIDE + NL * 52);
                relu_mask50(d, h);
                row_store52(dzact + p * HACT_STRIDE + NL * 52, d);
            }
            tmem_store_split<7>(tm, TM_AHI, TM_ALO, d);
            tmem_st_wait();
        }
        if (!staged) { mbar_wait(&S.ctl.wbar, 0); staged = true; }
        tc_group_sync(g);
#pragma unroll 1
        for (int l = NL; l >= 1; --l) {
            if (r == 0) {
                umma_gemm_tf32x3_ts(tm + TM_D, tm + TM_AHI, tm + TM_ALO, S.Whi + TT_PAIR + (l - 1) * TC_HID_SZ,
                                    S.Wlo + TT_PAIR + (l - 1) * TC_HID_SZ, 64, 64, 7);
                umma_commit(&S.ctl.mbar[g]);
            }
            float h[52];
            if (live) row_load52(h, hact + p * HACT_STRIDE + (l - 1) * 52);
            else {
#pragma unroll
                for (int k = 0; k < 52; ++k) h[k] = 0.f;
            }
            mbar_wait(&S.ctl.mbar[g], phase);
            tc_fence_after();
            float d[64];
            tc_load64(tm, d);
            relu_mask50(d, h);
            if (live) row_store52(dzact + p * HACT_STRIDE + (l - 1) * 52, d);
            if (l > 1) {
                tmem_store_split<7>(tm, TM_AHI, TM_ALO, d);
                tmem_st_wait();
            }
            phase ^= 1;
            tc_group_sync(g);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<TM_COLS>(S.ctl.tmem);
}

// ----------------------------------------------------------------------------------------------------------------
// weight gradients: dW_j[n][k] = sum_r G_j[r][n] * X_j[r][k] for a list of jobs that share the row index r
// ----------------------------------------------------------------------------------------------------------------
struct WSeg {
    const float* base;            // value(R, c) = base[(offs ? offs[R] : R * stride) + c]
    const long long* offs;
    long long stride;
    int cols_valid;               // columns c >= cols_valid read as zero
    int cols_store;               // columns c < cols_store are written to the B image at column col0 + c
    int col0;
};
struct WJob {
    const float* G; long long g_stride; int g_cols;   // gradient rows (row-major); columns >= g_cols are zero
    WSeg x[2]; int nx;                                // input rows: one or two column segments
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
};

constexpr int WG_ROWS = 64;       // rows r per stage (half a 128-row tile): K = 64 per stage = 8 MMA k-steps
constexpr int WG_NT = 288;        // 8 producer warps + 1 MMA warp
struct WStage {
    float Ahi[64 * WG_ROWS], Alo[64 * WG_ROWS];       // G^T: [r / 4][n (64)][r % 4]
    float Bhi[96 * WG_ROWS], Blo[96 * WG_ROWS];       // X^T: [r / 4][k (nb)][r % 4]
};
struct WgradSmem {
    WStage st[2];                                     // 2 x 81920 B
    unsigned long long full[2], empty[2], done;
    uint32_t tmem, pad;
};

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t umma_idesc_tf32_m64(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(64 >> 4) << 24);
}

// One producer thread's share of a stage: rows {4 w + rsub, 4 (w + 8) + rsub}, columns {8 g + nsub}.
__device__ __forceinline__ void wgrad_fill_stage(WStage& st, const WJob& J, long long row0, long long R) {
    const int t = threadIdx.x, w = t >> 5, lane = t & 31, rsub = lane & 3, nsub = lane >> 2;
    float gv[2][7];
    float xv[2][14];
    long long Rr[2];
    bool ok[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        Rr[q] = row0 + 4 * (w + 8 * q) + rsub;
        ok[q] = Rr[q] < R;
    }
    // issue every global load of this thread first (independent), then split and store
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const float* gp = J.G + Rr[q] * J.g_stride;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            const int n = 8 * c + nsub;
            gv[q][c] = (ok[q] && n < J.g_cols) ? __ldg(gp + n) : 0.f;
        }
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        if (s < J.nx) {
            const WSeg& X = J.x[s];
            const int ngroups = (X.cols_store + 7) >> 3;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float* xp = X.base + (ok[q] ? (X.offs ? X.offs[Rr[q]] : Rr[q] * X.stride) : 0);
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    const int cc = 8 * c + nsub;
                    xv[q][7 * s + c] = (c < ngroups && ok[q] && cc < X.cols_valid) ? __ldg(xp + cc) : 0.f;
                }
            }
        }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int rg = w + 8 * q;
        float* ah = st.Ahi + rg * 256 + lane;
        float* al = st.Alo + rg * 256 + lane;
#pragma unroll
        for (int c = 0; c < 7; ++c) {
            const float v = gv[q][c], hi = tf32_hi(v);
            ah[32 * c] = hi; al[32 * c] = v - hi;
        }
    }
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        if (s < J.nx) {
            const WSeg& X = J.x[s];
            const int ngroups = (X.cols_store + 7) >> 3;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const int rg = w + 8 * q;
                float* bh = st.Bhi + rg * J.nb * 4 + X.col0 * 4 + lane;
                float* bl = st.Blo + rg * J.nb * 4 + X.col0 * 4 + lane;
#pragma unroll
                for (int c = 0; c < 7; ++c) {
                    if (c < ngroups && 8 * c + nsub < X.cols_store) {
                        const float v = xv[q][7 * s + c], hi = tf32_hi(v);
                        bh[32 * c] = hi; bl[32 * c] = v - hi;
                    }
                }
            }
        }
    }
}

// Accumulator row -> the CTA's partial in the tile-major layout of npe_layout.h (what partial_index addresses).
__device__ __forceinline__ void wgrad_store_rows(float* __restrict__ out, const WJob& J, int n, const float* d, bool zero) {
    const int nchunks = J.nb >> 2;
    if (J.kind == 2) {   // encoders: G = dz0; rows 0..24 x cols 0..43 -> focus encoder, rows 25..49 x cols 48..91 -> context
        if (n >= H) return;
        const int nn = n < H2 ? n : n - H2, cb = n < H2 ? 0 : 12;
        float* o = out + (n < H2 ? PT_ENC_T : PT_ENC_C) + (nn >> 2) * 11 * 16 + 4 * (nn & 3);
        for (int kc = 0; kc < 11; ++kc) {
            const float* v = d + 4 * (cb + kc);
            st4(o + kc * 16, zero ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(v[0], v[1], v[2], v[3]));
        }
        return;
    }
    if (n >= J.n_valid) return;
    for (int kc = 0; kc < nchunks && kc < J.kt; ++kc) {
        float* o;
        if (J.kind == 1) {
            const int tile = (n >> 2) * 24 + kc;
            o = out + (tile < NT ? PT_DEC1A + tile * 16 : PT_DEC1B + (tile - NT) * 16) + 4 * (n & 3);
        } else {
            o = out + J.pt_base + ((n >> 2) * J.kt + kc) * 16 + 4 * (n & 3);
        }
        const float* v = d + 4 * kc;
        st4(o, zero ? make_float4(0.f, 0.f, 0.f, 0.f) : make_float4(v[0], v[1], v[2], v[3]));
    }
}

__global__ void __launch_bounds__(WG_NT, 1) k_wgrad_tc(WArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    WgradSmem& S = *reinterpret_cast<WgradSmem*>(smem_raw);
    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const long long R = A.nrows_dev ? (long long)*A.nrows_dev : (long long)A.nrows;
    const int ntiles = (int)((R + 127) / 128);
    if (t == 0) {
        mbar_init(&S.full[0], 256); mbar_init(&S.full[1], 256);
        mbar_init(&S.empty[0], 1); mbar_init(&S.empty[1], 1); mbar_init(&S.done, 1);
    }
    if (warp == 8) tmem_alloc<512>(&S.tmem);
    // zero both stages once: image rows / columns that no job writes (A rows 56..63) must read as zero
    for (int i = t; i < (int)(sizeof(S.st) / 16); i += WG_NT) reinterpret_cast<float4*>(S.st)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    proxy_fence_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.tmem;
    const bool has_work = (int)blockIdx.x < ntiles;
    if (warp < 8) {
        int u = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int j = 0; j < A.njobs; ++j) {
                for (int half = 0; half < 2; ++half, ++u) {
                    const int s = u & 1;
                    if (u >= 2) mbar_wait(&S.empty[s], (unsigned)(((u >> 1) - 1) & 1));
                    wgrad_fill_stage(S.st[s], A.job[j], (long long)tile * 128 + half * WG_ROWS, R);
                    proxy_fence_async();
                    mbar_arrive(&S.full[s]);
                }
            }
        }
    } else if (lane == 0) {
        int u = 0;
        unsigned started = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int j = 0; j < A.njobs; ++j) {
                const WJob& J = A.job[j];
                const uint32_t idesc = umma_idesc_tf32_m64(J.nb);
                for (int half = 0; half < 2; ++half, ++u) {
                    const int s = u & 1;
                    mbar_wait(&S.full[s], (unsigned)((u >> 1) & 1));
                    tc_fence_after();
                    const uint32_t ah = smem_u32(S.st[s].Ahi), al = smem_u32(S.st[s].Alo);
                    const uint32_t bh = smem_u32(S.st[s].Bhi), bl = smem_u32(S.st[s].Blo);
                    for (int pass = 0; pass < 3; ++pass) {
                        const uint32_t a0 = pass == 2 ? al : ah, b0 = pass == 1 ? bl : bh;
                        for (int ks = 0; ks < WG_ROWS / 8; ++ks) {
                            const uint64_t da = umma_desc(a0 + 2 * ks * 64 * 16, 64 * 16, 128);
                            const uint64_t db = umma_desc(b0 + 2 * ks * J.nb * 16, J.nb * 16, 128);
                            const uint32_t acc = ((started >> j) & 1u) | (uint32_t)(pass | ks);
                            umma_tf32(tmem + J.tmem_col, da, db, idesc, acc ? 1u : 0u);
                        }
                    }
                    started |= 1u << j;
                    umma_commit(&S.empty[s]);
                }
            }
        }
        umma_commit(&S.done);
    }
    // epilogue: warps 0..3 own the four TMEM lane quadrants; an M = 64 accumulator keeps row n in lane (n % 16) + 32 (n / 16)
    if (warp < 4) {
        mbar_wait(&S.done, 0);
        tc_fence_after();
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
    tc_fence_before();
    __syncthreads();
    if (warp == 8) tmem_dealloc<512>(tmem);
}

}  // namespace npe
