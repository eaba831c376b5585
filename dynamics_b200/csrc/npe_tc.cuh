// TF32 tcgen05 variant of the NPE forward path (north star: "a tf32/bf16 tcgen05 variant, treating pair-tiles of 128 as
// the M dimension").  Same data flow as the FP32 kernels (dense active-pair list -> pair MLP -> per-focus sum ->
// decoder), but every Linear layer is ONE chain of tcgen05.mma (M = 128 rows, N = 32/64, K = 8 per instruction) with
// the accumulator in TMEM; the epilogue warps read it with tcgen05.ld, apply ReLU and write the next layer's A operand
// straight into the canonical K-major shared-memory layout.  Inputs are rounded to TF32 (10-bit mantissa), the
// accumulation is FP32: results agree with the FP32 path to ~1e-3 relative (tests/test_gpu_parity.py, TF32 tolerance
// of SURVEY.md 8d), neighbourhood masks and pair indices stay bit-exact (they never touch the tensor path).
#pragma once
#include "npe_kernels.cuh"
#include "npe_umma.cuh"

namespace npe {

// ---- tensor-core weight image: every matrix as [k-chunk][n][4 floats] (K-major canonical layout), TF32-rounded ----
constexpr int TC_ENC_ROWS = 32, TC_HID_ROWS = 64, TC_OUT_ROWS = 32;
constexpr int TC_ENC_KC = 12, TC_HID_KC = 14, TC_D1_KC = 24;
constexpr int TC_ENC_T = 0;
constexpr int TC_ENC_C = TC_ENC_T + TC_ENC_KC * TC_ENC_ROWS * 4;           // 1536
constexpr int TC_PAIR = TC_ENC_C + TC_ENC_KC * TC_ENC_ROWS * 4;            // 3072
constexpr int TC_HID_SZ = TC_HID_KC * TC_HID_ROWS * 4;                     // 3584
constexpr int TC_PAIR_TOTAL = TC_PAIR + NL * TC_HID_SZ;                    // 20992 floats = 83968 B
constexpr int TC_DEC1 = 0;
constexpr int TC_DEC2 = TC_DEC1 + TC_D1_KC * TC_HID_ROWS * 4;              // 6144
constexpr int TC_DEC5 = TC_DEC2 + 3 * TC_HID_SZ;                           // 16896
constexpr int TC_DEC_TOTAL = TC_DEC5 + TC_HID_KC * TC_OUT_ROWS * 4;        // 18688 floats = 74752 B
constexpr int TC_TOTAL = TC_PAIR_TOTAL + TC_DEC_TOTAL;

__host__ __device__ __forceinline__ int tc_off(int base, int rows, int n, int k) { return base + ((k >> 2) * rows + n) * 4 + (k & 3); }

__host__ __device__ __forceinline__ int tc_index(int i) {
    if (i < OFF_ENC_C) { const int n = i / IN, k = i - n * IN; return tc_off(TC_ENC_T, TC_ENC_ROWS, n, k); }
    if (i < OFF_PAIR) { const int r = i - OFF_ENC_C, n = r / IN, k = r - n * IN; return tc_off(TC_ENC_C, TC_ENC_ROWS, n, k); }
    if (i < OFF_DEC1_W) {
        const int r0 = i - OFF_PAIR, l = r0 / 2500, r = r0 - l * 2500, n = r / H, k = r - n * H;
        return tc_off(TC_PAIR + l * TC_HID_SZ, TC_HID_ROWS, n, k);
    }
    const int base = TC_PAIR_TOTAL;
    if (i < OFF_DEC1_B) {
        const int r = i - OFF_DEC1_W, n = r / DEC_IN, k = r - n * DEC_IN;
        return tc_off(base + TC_DEC1, TC_HID_ROWS, n, k < H ? k : k + 2);
    }
    if (i < OFF_DEC2_W) return tc_off(base + TC_DEC1, TC_HID_ROWS, i - OFF_DEC1_B, ONE_COL);
    if (i < OFF_DEC5_W) {
        const int j = i - OFF_DEC2_W, li = j / 2550, r = j - li * 2550;
        if (r < 2500) { const int n = r / H, k = r - n * H; return tc_off(base + TC_DEC2 + li * TC_HID_SZ, TC_HID_ROWS, n, k); }
        return tc_off(base + TC_DEC2 + li * TC_HID_SZ, TC_HID_ROWS, r - 2500, ONE_COL);
    }
    if (i < OFF_DEC5_B) { const int r = i - OFF_DEC5_W, n = r / H, k = r - n * H; return tc_off(base + TC_DEC5, TC_OUT_ROWS, n, k); }
    return tc_off(base + TC_DEC5, TC_OUT_ROWS, i - OFF_DEC5_B, ONE_COL);
}

__global__ void k_pack_tc_zero(float* __restrict__ w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= TC_TOTAL) return;
    float v = 0.f;
    if (i == tc_off(TC_PAIR_TOTAL + TC_DEC1, TC_HID_ROWS, ONE_COL, ONE_COL)) v = 1.f;
    for (int li = 0; li < 3; ++li)
        if (i == tc_off(TC_PAIR_TOTAL + TC_DEC2 + li * TC_HID_SZ, TC_HID_ROWS, ONE_COL, ONE_COL)) v = 1.f;
    w[i] = v;
}
__global__ void k_pack_tc_scatter(const float* __restrict__ params, float* __restrict__ w) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < NPARAM) w[tc_index(i)] = to_tf32(params[i]);
}

// ----------------------------------------------------------------------------------------------------------------
// shared pieces
// ----------------------------------------------------------------------------------------------------------------
struct TcCtl {
    unsigned long long wbar;   // weight staging (TMA bulk copy)
    unsigned long long mbar;   // MMA completion (tcgen05.commit)
    uint32_t tmem;
    uint32_t pad;
};

// Epilogue helper: thread `row` (< 128) writes act[0 .. 4 * chunks) into the A tile [chunk][128][4] as TF32.
template <int CHUNKS>
__device__ __forceinline__ void store_a_row(float* sA, int row, const float* act) {
#pragma unroll
    for (int c = 0; c < CHUNKS; ++c)
        st4(sA + (c * 128 + row) * 4, to_tf32(make_float4(act[4 * c], act[4 * c + 1], act[4 * c + 2], act[4 * c + 3])));
}

// All threads: make generic-proxy smem writes visible to the tensor core, order TMEM reads before the barrier.
__device__ __forceinline__ void tc_sync() {
    proxy_fence_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
}

// ----------------------------------------------------------------------------------------------------------------
// pair MLP on dense tiles of 128 active pairs
// ----------------------------------------------------------------------------------------------------------------
struct PairTcSmem {
    float W[TC_PAIR_TOTAL];        // 83968 B
    float Af[TC_ENC_KC * 128 * 4]; // 24576 B   focus windows   (K = 48: 44 + zero pad)
    float Ac[TC_ENC_KC * 128 * 4]; // 24576 B   context windows
    float Ah[TC_HID_KC * 128 * 4]; // 28672 B   hidden activations (K = 56)
    long long frow[128];
    long long crow[128];
    TcCtl ctl;
};

__global__ void __launch_bounds__(NT, 1)
k_pair_forward_tc(const float* __restrict__ states, const int* __restrict__ pair_foc, const long long* __restrict__ pair_crow,
                  const long long* __restrict__ foc_row, const int* __restrict__ info, const float* __restrict__ wtc,
                  float* __restrict__ Hp) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PairTcSmem& S = *reinterpret_cast<PairTcSmem*>(smem_raw);
    const int P = info[0];
    if ((int)blockIdx.x * 128 >= P) return;
    const int t = threadIdx.x, warp = t >> 5;
    if (t == 0) { mbar_init(&S.ctl.wbar, 1); mbar_init(&S.ctl.mbar, 1); }
    if (warp == 0) tmem_alloc<64>(&S.ctl.tmem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.ctl.tmem;
    if (t == 0) {
        mbar_expect_tx(&S.ctl.wbar, TC_PAIR_TOTAL * 4);
        bulk_g2s(S.W, wtc, TC_PAIR_TOTAL * 4, &S.ctl.wbar);
    }
    bool staged = false;
    uint32_t phase = 0;
    for (int lo = blockIdx.x * 128; lo < P; lo += gridDim.x * 128) {
        const int nrows = min(128, P - lo);
        __syncthreads();
        if (t < 128) {
            if (t < nrows) {
                const int pf = pair_foc[lo + t];
                S.crow[t] = pair_crow[lo + t];
                S.frow[t] = foc_row[pf];
            }
        }
        __syncthreads();
        // gather both windows of every pair row into the K-major operand layout (TF32-rounded); chunk 11 = zero pad
        for (int idx = t; idx < 2 * TC_ENC_KC * 128; idx += NT) {
            const int which = idx / (TC_ENC_KC * 128), rem = idx - which * (TC_ENC_KC * 128);
            const int c = rem >> 7, r = rem & 127;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nrows && c < 11) {
                const float* p = states + (which ? S.crow[r] : S.frow[r]) + 4 * c;
                const float2 a = ld2(p), b = ld2(p + 2);
                v = to_tf32(make_float4(a.x, a.y, b.x, b.y));
            }
            st4((which ? S.Ac : S.Af) + (c * 128 + r) * 4, v);
        }
        if (!staged) { mbar_wait(&S.ctl.wbar, 0); staged = true; }
        tc_sync();
        if (t == 128) {
            tc_fence_after();
            // h0 = [ Wt f (cols 0..31) | Wc c (cols 32..63) ]
            umma_gemm_tf32(tmem, S.Af, 128, S.W + TC_ENC_T, TC_ENC_ROWS, 32, TC_ENC_KC / 2, false);
            umma_gemm_tf32(tmem + 32, S.Ac, 128, S.W + TC_ENC_C, TC_ENC_ROWS, 32, TC_ENC_KC / 2, false);
            umma_commit(&S.ctl.mbar);
        }
        if (t < 128) {
            mbar_wait(&S.ctl.mbar, phase);
            tc_fence_after();
            float d[64], act[56];
            tmem_ld16(tmem, 0, d); tmem_ld16(tmem, 16, d + 16); tmem_ld16(tmem, 32, d + 32); tmem_ld16(tmem, 48, d + 48);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 56; ++k) act[k] = k < H2 ? relu(d[k]) : (k < H ? relu(d[32 + k - H2]) : 0.f);
            store_a_row<TC_HID_KC>(S.Ah, t, act);
        }
        phase ^= 1;
        tc_sync();
#pragma unroll 1
        for (int l = 0; l < NL; ++l) {
            if (t == 128) {
                tc_fence_after();
                umma_gemm_tf32(tmem, S.Ah, 128, S.W + TC_PAIR + l * TC_HID_SZ, TC_HID_ROWS, 64, TC_HID_KC / 2, false);
                umma_commit(&S.ctl.mbar);
            }
            if (t < 128) {
                mbar_wait(&S.ctl.mbar, phase);
                tc_fence_after();
                float d[64];
                tmem_ld16(tmem, 0, d); tmem_ld16(tmem, 16, d + 16); tmem_ld16(tmem, 32, d + 32); tmem_ld16(tmem, 48, d + 48);
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 56; ++k) d[k] = relu(d[k]);
                if (l < NL - 1) {
                    store_a_row<TC_HID_KC>(S.Ah, t, d);
                } else if (t < nrows) {
                    float* o = Hp + (size_t)(lo + t) * 52;
#pragma unroll
                    for (int c = 0; c < 13; ++c) st4(o + 4 * c, make_float4(d[4 * c], d[4 * c + 1], d[4 * c + 2], d[4 * c + 3]));
                }
            }
            phase ^= 1;
            tc_sync();
        }
    }
    if (warp == 0) tmem_dealloc<64>(tmem);
}

// ----------------------------------------------------------------------------------------------------------------
// decoder on tiles of 128 foci
// ----------------------------------------------------------------------------------------------------------------
struct DecTcSmem {
    float W[TC_DEC_TOTAL];          // 74752 B
    float Az[TC_D1_KC * 128 * 4];   // 49152 B   z = [s(50) | 1 | 0 | this_past(44)]
    float Ah[TC_HID_KC * 128 * 4];  // 28672 B
    long long foff[128];
    int pstart[129];
    TcCtl ctl;
};

__global__ void __launch_bounds__(NT, 1)
k_dec_forward_tc(const float* __restrict__ states, const long long* __restrict__ foc_row, int F,
                 const int* __restrict__ pair_start, const float* __restrict__ Hp, const float* __restrict__ wtc,
                 const float* __restrict__ y, float* __restrict__ pred, float* __restrict__ loss_ex) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    DecTcSmem& S = *reinterpret_cast<DecTcSmem*>(smem_raw);
    const int t = threadIdx.x, warp = t >> 5;
    if (t == 0) { mbar_init(&S.ctl.wbar, 1); mbar_init(&S.ctl.mbar, 1); }
    if (warp == 0) tmem_alloc<64>(&S.ctl.tmem);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = S.ctl.tmem;
    if (t == 0) {
        mbar_expect_tx(&S.ctl.wbar, TC_DEC_TOTAL * 4);
        bulk_g2s(S.W, wtc + TC_PAIR_TOTAL, TC_DEC_TOTAL * 4, &S.ctl.wbar);
    }
    bool staged = false;
    uint32_t phase = 0;
    const int ntiles = (F + 127) / 128;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int f0 = tile * 128;
        const int nf = min(128, F - f0);
        __syncthreads();
        if (t < nf) S.foff[t] = foc_row[f0 + t];
        if (t <= nf && t < 129) S.pstart[t] = pair_start[f0 + t];
        __syncthreads();
        // z tile: chunks 0..12 = pair sums (cols 50, 51 = 1, 0), chunks 13..23 = focus window
        for (int idx = t; idx < TC_D1_KC * 128; idx += NT) {
            const int c = idx >> 7, r = idx & 127;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (r < nf) {
                if (c < 13) {
                    for (int p = S.pstart[r]; p < S.pstart[r + 1]; ++p) {
                        const float4 h = ld4(Hp + (size_t)p * 52 + 4 * c);
                        v.x += h.x; v.y += h.y; v.z += h.z; v.w += h.w;
                    }
                    if (c == 12) { v.z = 1.f; v.w = 0.f; }
                } else {
                    const float* p = states + S.foff[r] + 4 * (c - 13);
                    const float2 a = ld2(p), b = ld2(p + 2);
                    v = make_float4(a.x, a.y, b.x, b.y);
                }
            }
            st4(S.Az + (c * 128 + r) * 4, to_tf32(v));
        }
        if (!staged) { mbar_wait(&S.ctl.wbar, 0); staged = true; }
        tc_sync();
#pragma unroll 1
        for (int l = 0; l < NL; ++l) {
            if (t == 128) {
                tc_fence_after();
                if (l == 0) umma_gemm_tf32(tmem, S.Az, 128, S.W + TC_DEC1, TC_HID_ROWS, 64, TC_D1_KC / 2, false);
                else if (l < NL - 1) umma_gemm_tf32(tmem, S.Ah, 128, S.W + TC_DEC2 + (l - 1) * TC_HID_SZ, TC_HID_ROWS, 64, TC_HID_KC / 2, false);
                else umma_gemm_tf32(tmem, S.Ah, 128, S.W + TC_DEC5, TC_OUT_ROWS, 32, TC_HID_KC / 2, false);
                umma_commit(&S.ctl.mbar);
            }
            if (t < 128) {
                mbar_wait(&S.ctl.mbar, phase);
                tc_fence_after();
                if (l < NL - 1) {
                    float d[64];
                    tmem_ld16(tmem, 0, d); tmem_ld16(tmem, 16, d + 16); tmem_ld16(tmem, 32, d + 32); tmem_ld16(tmem, 48, d + 48);
                    tmem_ld_wait();
#pragma unroll
                    for (int k = 0; k < 56; ++k) d[k] = relu(d[k]);
                    store_a_row<TC_HID_KC>(S.Ah, t, d);
                } else {
                    float o[32];
                    tmem_ld16(tmem, 0, o); tmem_ld16(tmem, 16, o + 16);
                    tmem_ld_wait();
                    if (t < nf) {
                        float* pr = pred + (size_t)(f0 + t) * OUT;
#pragma unroll
                        for (int n = 0; n < OUT; ++n) pr[n] = n < 6 ? o[n] : 1.f / (1.f + expf(-o[n]));
                        if (loss_ex && y) {
                            const float* yy = y + (size_t)(f0 + t) * OUT;
                            const float e2 = o[2] - yy[2], e3 = o[3] - yy[3], e5 = o[5] - yy[5];
                            const float lv = __fadd_rn(__fmul_rn(e2, e2), __fmul_rn(e3, e3)), lw = __fmul_rn(e5, e5);
                            loss_ex[(size_t)(f0 + t) * 3 + 0] = __fdiv_rn(__fadd_rn(lv, lw), 3.f);
                            loss_ex[(size_t)(f0 + t) * 3 + 1] = __fdiv_rn(lv, 2.f);
                            loss_ex[(size_t)(f0 + t) * 3 + 2] = lw;
                        }
                    }
                }
            }
            phase ^= 1;
            tc_sync();
        }
    }
    if (warp == 0) tmem_dealloc<64>(tmem);
}

// ================================================================================================================
// 3xTF32 variant: FP32-grade forward on the tensor cores.  The activations never touch shared memory: each epilogue
// thread owns one row, reads the accumulator from TMEM, applies ReLU, splits the result into hi / lo TF32 halves and
// writes them back to TENSOR MEMORY as the A operand of the next layer (tcgen05.st; A-from-TMEM MMA).  Shared memory
// holds only the hi and lo images of the weights.  The kernels live in npe_chain.cuh (warp-specialised chains); this
// file keeps the weight images and the shared helpers.
// TMEM columns per tile slot: [0,64) accumulator, [64,160) A hi, [160,256) A lo.
// ================================================================================================================
constexpr int TM_D = 0, TM_AHI = 64, TM_ALO = 160, TM_GROUP = 256, TM_COLS = 512;

__global__ void k_pack_tc3_scatter(const float* __restrict__ params, float* __restrict__ w_hi, float* __restrict__ w_lo) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NPARAM) return;
    const float x = params[i], h = tf32_hi(x);
    w_hi[tc_index(i)] = h;
    w_lo[tc_index(i)] = x - h;
}

// ReLU of the 56 leading accumulator columns of this thread's row.
__device__ __forceinline__ void tc_load_relu(uint32_t tm, float* d) {
    tmem_ld16(tm, 0, d); tmem_ld16(tm, 16, d + 16); tmem_ld16(tm, 32, d + 32); tmem_ld16(tm, 48, d + 48);
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 56; ++k) d[k] = relu(d[k]);
}

// ----------------------------------------------------------------------------------------------------------------
// Rollout by kernels (large scene counts): the window tensor roll (S, Nmax, 2, 22) lives in HBM/L2; every step runs the
// pair builder, the two forward kernels and this post-process kernel, which applies simulate_all_postprocess
// (eval.lua:155-182; update_position / update_angle npe.lua:386-458) per object, accumulates the step metrics and
// shifts the window in place (each thread owns one object: reads its own last frame, then overwrites).
// ----------------------------------------------------------------------------------------------------------------
struct RollPostArgs {
    float* roll;              // (S, Nmax, 2, 22) current window
    const float* states;      // (S, Nmax, T, 22) ground truth / initial frames
    const int* n_obj;
    const int* foc_of_obj;    // (S * Nmax) focus index of each object, -1 for obstacles / empty slots
    const float* pred;        // (F, 22) relative predictions
    int S, Nmax, T, t0, step, steps, use_gt, teacher, ball_zero;
    float* traj;              // (S, Nmax, steps, 22) or null
    double* metrics;          // (steps, 7) or null
    double* slot_metrics;     // (steps, Nmax, 7) or null
};

// One thread per (object, state column): 22 consecutive threads touch 22 consecutive floats (coalesced); the thread of
// column 0 also computes the object's step metrics.
constexpr int POST_OBJ = 11, POST_NT = POST_OBJ * D;   // 242 threads: a CTA owns whole objects (the shift is in place)
__global__ void __launch_bounds__(256) k_rollout_post(RollPostArgs A) {   // 256 threads, the last 14 idle
    __shared__ double red[7][8];
    const long long gid = (long long)blockIdx.x * POST_NT + threadIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const float PI_F = 3.14159274101257324f, TWO_PI_F = 6.28318548202514648f;
    double mt[7] = {0, 0, 0, 0, 0, 0, 0};
    const long long idx = gid / D;
    const int c = (int)(gid - idx * D);
    bool live = false;
    float q = 0.f, lc = 0.f, gtc = 0.f;
    float* w = nullptr;
    if (threadIdx.x < POST_NT && idx < (long long)A.S * A.Nmax) {
        const int s = (int)(idx / A.Nmax), ob = (int)(idx - (long long)s * A.Nmax);
        if (ob < A.n_obj[s]) {
            live = true;
            w = A.roll + (size_t)idx * 2 * D;
            const float* last = w + D;
            lc = last[c];
            const float* gt = A.use_gt ? A.states + (((size_t)idx * A.T) + A.t0 + 2 + A.step) * D : nullptr;
            if (gt) gtc = gt[c];
            const int f = A.foc_of_obj[idx];
            if (f < 0) {
                q = gt ? gt[c] : lc;
            } else {
                const float* o = A.pred + (size_t)f * OUT;
                if (c < 2) q = __fdiv_rn(__fadd_rn(__fmul_rn(lc, 800.f), __fmul_rn(last[c + 2], 60.f)), 800.f);
                else if (c == 4) {
                    float ang = __fadd_rn(__fmul_rn(lc, PI_F), __fmul_rn(last[5], PI_F));
                    const float gtpi = ang > PI_F ? 1.f : 0.f, lepi = ang <= -PI_F ? 1.f : 0.f;
                    ang = __fadd_rn(ang, __fmul_rn(-TWO_PI_F, gtpi));
                    ang = __fadd_rn(ang, __fmul_rn(TWO_PI_F, lepi));
                    q = __fdiv_rn(ang, PI_F);
                } else if (c < 6) q = __fadd_rn(o[c], lc);
                else q = lc;
                if ((c == 4 || c == 5) && A.ball_zero && last[10] == 1.f) q = 0.f;
                if (c == 0 && gt) {
                    const float y2 = __fsub_rn(gt[2], last[2]), y3 = __fsub_rn(gt[3], last[3]), y5 = __fsub_rn(gt[5], last[5]);
                    const float e2 = o[2] - y2, e3 = o[3] - y3, e5 = o[5] - y5;
                    const float lv = __fadd_rn(__fmul_rn(e2, e2), __fmul_rn(e3, e3)), lw = __fmul_rn(e5, e5);
                    mt[0] = __fdiv_rn(__fadd_rn(lv, lw), 3.f); mt[1] = __fdiv_rn(lv, 2.f); mt[2] = lw; mt[3] = 1.0;
                    const float pvx = __fmul_rn(__fadd_rn(o[2], last[2]), 60.f), pvy = __fmul_rn(__fadd_rn(o[3], last[3]), 60.f);
                    const float gvx = __fmul_rn(__fadd_rn(y2, last[2]), 60.f), gvy = __fmul_rn(__fadd_rn(y3, last[3]), 60.f);
                    const float pm = __fsqrt_rn(__fadd_rn(__fmul_rn(pvx, pvx), __fmul_rn(pvy, pvy)));
                    const float gm = __fsqrt_rn(__fadd_rn(__fmul_rn(gvx, gvx), __fmul_rn(gvy, gvy)));
                    if (gm != 0.f) {
                        mt[4] = __fdiv_rn(__fadd_rn(__fmul_rn(pvx, gvx), __fmul_rn(pvy, gvy)), __fmul_rn(pm, gm));
                        mt[5] = fabsf(__fsub_rn(1.f, __fdiv_rn(pm, gm)));
                        mt[6] = 1.0;
                    }
                    if (A.slot_metrics) {
                        double* m = A.slot_metrics + ((size_t)A.step * A.Nmax + ob) * 7;
#pragma unroll
                        for (int i = 0; i < 7; ++i)
                            if (mt[i] != 0.0) atomicAdd(m + i, mt[i]);
                    }
                }
            }
            if (A.traj) A.traj[(((size_t)idx * A.steps) + A.step) * D + c] = q;
        }
    }
    __syncthreads();   // all reads of the object's last frame (by its 22 threads, same CTA) precede the in-place shift
    if (live) {
        w[c] = lc;
        w[D + c] = (A.teacher && A.use_gt) ? gtc : q;
    }
    if (A.metrics && A.use_gt) {
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            double v = mt[i];
            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
            if (lane == 0) red[i][warp] = v;
        }
        __syncthreads();
        if (threadIdx.x < 7) {
            double v = 0.0;
            for (int w8 = 0; w8 < 8; ++w8) v += red[threadIdx.x][w8];
            if (v != 0.0) atomicAdd(A.metrics + (size_t)A.step * 7 + threadIdx.x, v);
        }
    }
}

// roll[s][n][0..1] = states[s][n][t0 .. t0+1]
__global__ void k_rollout_init(float* __restrict__ roll, const float* __restrict__ states, int S_Nmax, int T, int t0) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= S_Nmax * 2 * D) return;
    const int o = idx / (2 * D), c = idx - o * 2 * D;
    roll[idx] = states[((size_t)o * T + t0) * D + c];
}

}  // namespace npe
