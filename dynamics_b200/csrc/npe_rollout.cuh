// K6: persistent autoregressive rollout.  One CTA owns a group of scenes (G = 64 / Nmax scenes, <= 64 object slots);
// the two-frame window of every object stays in shared memory for all steps, every object is the focus once per step
// (Jacobi update: all foci read the same window, eval.lua:294-384), and only the predicted states leave the SM.
//
// Per step and focus (reference eval.lua:122-182, npe.lua:386-458):
//   contexts = other objects of the scene -> 12-NN trim -> neighbourhood mask -> NPE forward -> relative prediction
//   q[0:6] = pred[0:6] + last[0:6];  q[6:22] = last[6:22]
//   q[0:2] = ((last_pos * 800) + (last_vel * 60)) / 800          -- integrates the PREVIOUS velocity
//   ang = (last_a * pi) + (last_av * pi), wrapped to (-pi, pi], / pi
//   ball foci: a = av = 0;  obstacle foci are never predicted (ground truth / static copy)
// Metrics per step (eval.lua:337-356, infer.lua:9-74): sums of per-example losses, cosine and relative-magnitude
// errors of the predicted vs ground-truth velocity, with their valid counts.
#pragma once
#include "npe_kernels.cuh"

namespace npe {

struct RollMeta {
    u64 nbr[TILE];            // active-context bits per slot
    int start[TILE + 1];      // exclusive scan of active-pair counts
    int pf[TILE];             // per pair row: focus slot
    int px[TILE];             // per pair row: context slot
};

struct RollSmem {
    float Wp[SW_PAIR_TOTAL];
    float Wd[SW_DEC_TOTAL];
    float Z[TILE * LD1];      // cols 52..95 hold the window [frame0 | frame1] of every object slot
    float Ef[TILE * 28];
    float P0[TILE * LD];
    float P1[TILE * LD];      // also the gathered context rows X (dead once h0 is built)
    float O[TILE * LD];
    float New[TILE * 24];
    RollMeta M;
    long long foff[TILE];
    int kind[TILE];           // 0 = empty slot, 1 = predicted focus, 2 = obstacle
    double red[7][NT / 32];
    WeightStage W;
};

struct RollArgs {
    const float* states;      // (S, Nmax, T, 22)
    const int* n_obj;
    int S, Nmax, T;
    int t0, steps, use_gt, teacher, kmax, ball_zero;
    float thr_s;
    const float* wpack;       // packed weight image
    float* traj;              // (S, Nmax, steps, 22) or null
    double* metrics;          // (steps, 7) or null
    double* slot_metrics;     // (steps, Nmax, 7) or null: the same sums kept per object slot (eval.lua:348-358)
};

__global__ void __launch_bounds__(NT, 1) k_rollout(RollArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    RollSmem& S = *reinterpret_cast<RollSmem*>(smem_raw);
    stage_issue(S.W, S.Wp, S.Wd, A.wpack);
    bool staged = false;
    float* const ubufs[4] = {S.P0, S.P1, S.P0, S.P1};
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int G = TILE / A.Nmax;
    const int ngroups = (A.S + G - 1) / G;
    const float PI_F = 3.14159274101257324f;          // (float)math.pi
    const float TWO_PI_F = 6.28318548202514648f;      // (float)(2*math.pi)
    for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
        const int s0 = grp * G;
        const int ns = min(G, A.S - s0);
        const int nslots = ns * A.Nmax;
        __syncthreads();
        // load the initial window; slot = g * Nmax + o
        if (t < TILE) {
            const int g = t / A.Nmax, o = t - g * A.Nmax;
            int kind = 0;
            if (t < nslots && o < A.n_obj[s0 + g]) kind = 1;
            S.kind[t] = kind;
            S.foff[t] = (((long long)(s0 + min(g, ns - 1)) * A.Nmax + o) * A.T + A.t0) * D;
        }
        __syncthreads();
        for (int idx = t; idx < TILE * LD1; idx += NT) S.Z[idx] = 0.f;
        __syncthreads();
        for (int idx = t; idx < TILE * IN; idx += NT) {
            const int r = idx / IN, c = idx - r * IN;
            if (S.kind[r]) S.Z[r * LD1 + ZCOL_F + c] = A.states[S.foff[r] + c];
        }
        __syncthreads();
        if (t < TILE && S.kind[t] && S.Z[t * LD1 + ZCOL_F + D + 11] == 1.f) S.kind[t] = 2;   // oid one-hot: obstacle
        __syncthreads();
        for (int step = 0; step < A.steps; ++step) {
            // ---- pair table on the last frame (K1 in-kernel) ----
            for (int r = warp; r < TILE; r += NT / 32) {
                u64 keep = 0, nbr = 0;
                if (S.kind[r] == 1) {
                    const int g = r / A.Nmax, o = r - g * A.Nmax;
                    warp_pairs(S.Z + (g * A.Nmax) * LD1 + ZCOL_F + D, LD1, A.n_obj[s0 + g], o, A.kmax, A.thr_s, keep, nbr, false);
                }
                if (lane == 0) S.M.nbr[r] = nbr;
            }
            for (int idx = t; idx < TILE * 26; idx += NT) {
                const int r = idx / 26, c2 = idx - r * 26;
                float2 v = make_float2(0.f, 0.f);
                if (c2 == ONE_COL / 2 && S.kind[r] == 1) v.x = 1.f;
                *reinterpret_cast<float2*>(S.Z + r * LD1 + 2 * c2) = v;
            }
            __syncthreads();
            if (t == 0) {
                int acc = 0;
                for (int f = 0; f < TILE; ++f) { S.M.start[f] = acc; acc += __popcll(S.M.nbr[f]); }
                S.M.start[TILE] = acc;
            }
            if (!staged) { stage_wait(S.W); staged = true; }
            // focus half of the encoder once per focus
            gemm_nt<4, LD1, LD>(S.Z + ZCOL_F, S.Wp + SW_ENC_T, 11, TILE, EpiNT{S.Ef, 28, 28, 1});
            __syncthreads();
            const int P = S.M.start[TILE];
            for (int lo = 0; lo < P; lo += TILE) {
                const int nrows = min(TILE, P - lo);
                if (t < TILE) {
                    u64 bits = S.M.nbr[t];
                    int pos = S.M.start[t] - lo;
                    const int fo = t % A.Nmax, sb = (t / A.Nmax) * A.Nmax;
                    while (bits) {
                        const int c = __ffsll((long long)bits) - 1;
                        bits &= bits - 1;
                        if (pos >= 0 && pos < TILE) { S.M.pf[pos] = t; S.M.px[pos] = sb + c + (c >= fo); }
                        ++pos;
                    }
                }
                __syncthreads();
                // X (context rows) -> P1; h0 -> P0 = [Ef[focus] | relu(Wc x) | 0]
                for (int idx = t; idx < TILE * (IN / 4); idx += NT) {
                    const int r = idx / (IN / 4), c4 = idx - r * (IN / 4);
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (r < nrows) v = ld4(S.Z + S.M.px[r] * LD1 + ZCOL_F + 4 * c4);
                    st4(S.P1 + r * LD + 4 * c4, v);
                }
                for (int idx = t; idx < TILE * 32; idx += NT) {
                    const int r = idx >> 5, n = idx & 31;
                    if (n < H2) S.P0[r * LD + n] = (r < nrows) ? S.Ef[S.M.pf[r] * 28 + n] : 0.f;
                    else if (n >= 26) S.P0[r * LD + (n - 26) + H] = 0.f;
                }
                __syncthreads();
                gemm_nt<4, LD, LD>(S.P1, S.Wp + SW_ENC_C, 11, nrows, EpiNT{S.P0 + H2, LD, H2, 1});
                __syncthreads();
                float* hin = S.P0;
                float* hout = S.P1;
#pragma unroll 1
                for (int l = 0; l < NL; ++l) {
                    gemm_nt<7, LD, LD>(hin, S.Wp + SW_PAIR + l * SW_PAIR_SZ, 13, nrows, EpiNT{hout, LD, 52, 1});
                    __syncthreads();
                    float* tmp = hin; hin = hout; hout = tmp;
                }
                // h5 is in `hin` (= P1 after five swaps); fixed-order segmented add into the pair-sum columns
                for (int idx = t; idx < TILE * H; idx += NT) {
                    const int f = idx / H, n = idx - f * H;
                    const int r0 = max(S.M.start[f], lo) - lo, r1 = min(S.M.start[f + 1], lo + TILE) - lo;
                    if (r1 > r0) {
                        float a = S.Z[f * LD1 + n];
                        for (int r = r0; r < r1; ++r) a += hin[r * LD + n];
                        S.Z[f * LD1 + n] = a;
                    }
                }
                __syncthreads();
            }
            decoder_forward<TILE>(S.Wd, S.Z, ubufs, S.O, TILE);
            // ---- post-process, metrics ----
            double mt[7] = {0, 0, 0, 0, 0, 0, 0};
            if (t < TILE && S.kind[t]) {
                const float* last = S.Z + t * LD1 + ZCOL_F + D;
                const float* o = S.O + t * LD;
                const int g = t / A.Nmax, ob = t - g * A.Nmax;
                const float* gt = A.use_gt ? A.states + ((((size_t)(s0 + g) * A.Nmax + ob) * A.T) + A.t0 + 2 + step) * D : nullptr;
                float q[D];
                if (S.kind[t] == 2) {
#pragma unroll
                    for (int c = 0; c < D; ++c) q[c] = gt ? gt[c] : last[c];
                } else {
#pragma unroll
                    for (int c = 0; c < 6; ++c) q[c] = __fadd_rn(o[c], last[c]);
#pragma unroll
                    for (int c = 6; c < D; ++c) q[c] = last[c];
                    q[0] = __fdiv_rn(__fadd_rn(__fmul_rn(last[0], 800.f), __fmul_rn(last[2], 60.f)), 800.f);
                    q[1] = __fdiv_rn(__fadd_rn(__fmul_rn(last[1], 800.f), __fmul_rn(last[3], 60.f)), 800.f);
                    float ang = __fadd_rn(__fmul_rn(last[4], PI_F), __fmul_rn(last[5], PI_F));
                    const float gtpi = ang > PI_F ? 1.f : 0.f, lepi = ang <= -PI_F ? 1.f : 0.f;
                    ang = __fadd_rn(ang, __fmul_rn(-TWO_PI_F, gtpi));
                    ang = __fadd_rn(ang, __fmul_rn(TWO_PI_F, lepi));
                    q[4] = __fdiv_rn(ang, PI_F);
                    if (A.ball_zero && last[10] == 1.f) { q[4] = 0.f; q[5] = 0.f; }
                    if (gt) {
                        // fp_batch losses against gt made relative to the PREDICTED last frame (eval.lua:126-133)
                        const float y2 = __fsub_rn(gt[2], last[2]), y3 = __fsub_rn(gt[3], last[3]), y5 = __fsub_rn(gt[5], last[5]);
                        const float e2 = o[2] - y2, e3 = o[3] - y3, e5 = o[5] - y5;
                        const float lv = __fadd_rn(__fmul_rn(e2, e2), __fmul_rn(e3, e3)), lw = __fmul_rn(e5, e5);
                        mt[0] = __fdiv_rn(__fadd_rn(lv, lw), 3.f); mt[1] = __fdiv_rn(lv, 2.f); mt[2] = lw; mt[3] = 1.0;
                        // angle_magnitude (infer.lua:32-74): un-relative both, scale by vnc
                        const float pvx = __fmul_rn(__fadd_rn(o[2], last[2]), 60.f), pvy = __fmul_rn(__fadd_rn(o[3], last[3]), 60.f);
                        const float gvx = __fmul_rn(__fadd_rn(y2, last[2]), 60.f), gvy = __fmul_rn(__fadd_rn(y3, last[3]), 60.f);
                        const float pm = __fsqrt_rn(__fadd_rn(__fmul_rn(pvx, pvx), __fmul_rn(pvy, pvy)));
                        const float gm = __fsqrt_rn(__fadd_rn(__fmul_rn(gvx, gvx), __fmul_rn(gvy, gvy)));
                        if (gm != 0.f) {
                            const float num = __fadd_rn(__fmul_rn(pvx, gvx), __fmul_rn(pvy, gvy));
                            mt[4] = __fdiv_rn(num, __fmul_rn(pm, gm));
                            mt[5] = fabsf(__fsub_rn(1.f, __fdiv_rn(pm, gm)));
                            mt[6] = 1.0;
                        }
                    }
                }
                if (A.slot_metrics && mt[3] != 0.0) {
                    double* m = A.slot_metrics + ((size_t)step * A.Nmax + ob) * 7;
#pragma unroll
                    for (int i = 0; i < 7; ++i)
                        if (mt[i] != 0.0) atomicAdd(m + i, mt[i]);
                }
#pragma unroll
                for (int c = 0; c < D; ++c) S.New[t * 24 + c] = q[c];
                if (A.traj) {
                    float* dst = A.traj + ((((size_t)(s0 + g) * A.Nmax + ob) * A.steps) + step) * D;
#pragma unroll
                    for (int c = 0; c < D; ++c) dst[c] = q[c];
                }
            }
            if (A.metrics && A.use_gt) {
#pragma unroll
                for (int i = 0; i < 7; ++i) {
                    double v = mt[i];
                    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                    if (lane == 0) S.red[i][warp] = v;
                }
            }
            __syncthreads();
            if (A.metrics && A.use_gt && t < 7) {
                double v = 0.0;
                for (int w = 0; w < NT / 32; ++w) v += S.red[t][w];
                atomicAdd(A.metrics + (size_t)step * 7 + t, v);
            }
            // ---- shift the window: frame0 <- frame1, frame1 <- new (or ground truth when teacher forcing) ----
            for (int idx = t; idx < TILE * D; idx += NT) {
                const int r = idx / D, c = idx - r * D;
                if (S.kind[r]) {
                    float* w = S.Z + r * LD1 + ZCOL_F;
                    float nv = S.New[r * 24 + c];
                    if (A.teacher && A.use_gt) {
                        const int g = r / A.Nmax, ob = r - g * A.Nmax;
                        nv = A.states[((((size_t)(s0 + g) * A.Nmax + ob) * A.T) + A.t0 + 2 + step) * D + c];
                    }
                    const float f1 = w[D + c];
                    w[c] = f1;
                    w[D + c] = nv;
                }
            }
            __syncthreads();
        }
    }
}

}  // namespace npe
