// One kernel for the forward + backward of a SMALL batch (the reference's batch of 50): the latency path.
//
// With 16-row tiles a batch of 50 foci is 4 decoder tiles and ~13 pair tiles; as separate kernels (pair forward ->
// decoder forward/backward -> pair backward) every boundary costs a grid drain, a launch hand-off and a round trip of
// the pair activations through global memory.  Here every tile is a CTA of ONE grid and the two dependencies are
// point-to-point flags in global memory:
//
//   pair CTA p :  pair MLP forward on its 16 active pairs (activations stay in shared memory) -> Hp rows, flag P[p]
//                 ... waits for the decoder tiles that own its foci: flag D[d] ...
//                 dz5 = ds[focus] * relu', input-gradient chain + weight gradients -> partial[p] (pair part)
//   decoder CTA d: waits for the pair tiles that hold its foci's pairs: flag P[p]
//                 per-focus sums, decoder forward, prediction / loss, input-gradient chain -> ds, flag D[d]
//                 (the pair CTAs continue from here) ... weight gradients -> partial[d] (decoder part)
//
// Every CTA owns exactly one tile (the host only takes this path when pair tiles + decoder tiles fit the SM count, so
// all CTAs are resident and no CTA waits on work that has not been scheduled); a pair CTA publishes before it waits
// and decoder CTAs wait only on pair CTAs' first phase, so the wait graph has no cycle.  Waits are bounded: a wait that
// does not complete in ~0.1 s raises `abort` (reported by the host as an error) instead of hanging the device.
// Arithmetic, tile assignment and summation orders are those of k_pair_forward / k_dec_backward / k_pair_backward with
// ROWS = 16, so the results are bit-identical to the separate kernels.
#pragma once
#include "npe_kernels.cuh"

namespace npe {

constexpr int SR = 16;   // rows per tile on this path

struct StepArgs {
    const float* states; const int* pair_foc; const long long* pair_crow; const long long* foc_row; const int* pair_start;
    const int* info; const float* wpack; const float* y;
    float* Hp; float* ds; float* partial; float* pred; float* loss_ex;
    int F, gp;                      // foci; number of pair CTAs (decoder CTAs follow in the grid)
    float scale_v, scale_w;
    unsigned* flags;                // [0, gp): pair tile forward done; [gp, gp + gd): decoder tile ds done
    unsigned stamp;
    int* abort;
    unsigned long long* trace;      // optional (tools): 8 globaltimer stamps per CTA at the phase boundaries
};

struct StepDecSmem {
    float Wd[SW_DEC_TOTAL];
    float Z[SR * LD1];
    float U[4][SR * LD];
    float O[SR * LD];
    float Dz[5][SR * LD];           // dz1..dz5: kept so that the weight gradients can follow the input-gradient chain
    long long foff[SR];
    int pstart[SR + 1];
    WeightStage W;
};
union StepSmem {
    PairBwdSmem<SR> pair;
    StepDecSmem dec;
};

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// thread 0 waits until flags[lo..hi] carry `stamp`; bounded
__device__ __forceinline__ void wait_flags(const unsigned* flags, int lo, int hi, unsigned stamp, int* abort) {
    if (threadIdx.x == 0) {
        for (int i = lo; i <= hi; ++i) {
            int spins = 0;
            while (ld_acquire_gpu(flags + i) != stamp) {
                __nanosleep(32);
                if (++spins > (1 << 22)) { *abort = 2; break; }
            }
        }
    }
    __syncthreads();
}
__device__ __forceinline__ void publish_flag(unsigned* flag, unsigned stamp) {
    __syncthreads();                                   // the CTA's global writes precede the release (cumulativity)
    if (threadIdx.x == 0) { __threadfence(); st_release_gpu(flag, stamp); }
}
__device__ __forceinline__ void trace_mark(const StepArgs& A, int slot) {
    if (A.trace && threadIdx.x == 0) {
        unsigned long long tns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
        A.trace[(size_t)blockIdx.x * 8 + slot] = tns;
    }
}
__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
// The reduction of the partials + optimiser stays a separate kernel (k_reduce_grads / k_dp_step, 111 CTAs, one
// parameter per thread): folded into this grid behind an all-CTA barrier it was slower (17 CTAs x 256 threads walk
// 6-7 parameters each, i.e. 6-7 serial L2 latency chains: 39 us per launch against 30 + 4).
__global__ void __launch_bounds__(NT, 1) k_step_fused(StepArgs A) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    StepSmem& U_ = *reinterpret_cast<StepSmem*>(smem_raw);
    const int t = threadIdx.x;
    if ((int)blockIdx.x < A.gp) {
        // ------------------------------------------------ pair tile ------------------------------------------------
        PairBwdSmem<SR>& S = U_.pair;
        for (int idx = t; idx < SR * LD; idx += NT) { S.Da[idx] = 0.f; S.Db[idx] = 0.f; }
        stage_init(S.W);
        const int P = A.info[0];
        const int lo = blockIdx.x * SR;
        const int nrows = min(SR, P - lo);
        const bool active = lo < P;
        if (active) {   // static inputs: loaded while the previous kernel (reduce + optimiser) is still running
            load_pair_meta(S.M, A.pair_foc, A.pair_crow, A.foc_row, lo, nrows);
            __syncthreads();
            gather_rows<SR, LD>(S.Xf, A.states, S.M.frow, nrows);
            gather_rows<SR, LD>(S.Xc, A.states, S.M.crow, nrows);
        }
        __syncthreads();
        trace_mark(A, 0);
        pdl_wait();
        trace_mark(A, 1);
        pdl_launch_dependents();
        float acc[NL][16];
#pragma unroll
        for (int l = 0; l < NL; ++l)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[l][i] = 0.f;
        if (active) {
            stage_copy(S.W, S.Wp, nullptr, A.wpack);
            stage_wait(S.W);
            float* const bufs[NL + 1] = {S.Hh[0], S.Hh[1], S.Hh[2], S.Hh[3], S.Hh[4], S.Hh[5]};
            // rows >= nrows of the activation tiles: gather_rows zero-fills the inputs, so they are finite
            pair_chain_dense<SR>(S.Wp, S.Xf, S.Xc, nrows, bufs);
            for (int idx = t; idx < nrows * 13; idx += NT) {
                const int r = idx / 13, c4 = idx - r * 13;
                st4(A.Hp + (size_t)(lo + r) * 52 + 4 * c4, ld4(S.Hh[NL] + r * LD + 4 * c4));
            }
        }
        publish_flag(A.flags + blockIdx.x, A.stamp);       // inactive tiles publish too (nobody waits on them)
        trace_mark(A, 2);
        if (active) {
            // the decoder tiles that own this tile's foci (the pair list is sorted by focus)
            wait_flags(A.flags + A.gp, S.M.pf[0] / SR, S.M.pf[nrows - 1] / SR, A.stamp, A.abort);
            trace_mark(A, 3);
            for (int idx = t; idx < SR * (LD / 4); idx += NT) {
                const int r = idx / (LD / 4), c4 = idx - r * (LD / 4);
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (r < nrows && c4 < 13) {
                    v = ldcg4(A.ds + (size_t)S.M.pf[r] * 52 + 4 * c4);
                    const float4 hv = ld4(S.Hh[NL] + r * LD + 4 * c4);
                    v.x = hv.x > 0.f ? v.x : 0.f; v.y = hv.y > 0.f ? v.y : 0.f;
                    v.z = hv.z > 0.f ? v.z : 0.f; v.w = hv.w > 0.f ? v.w : 0.f;
                }
                st4(S.Da + r * LD + 4 * c4, v);
            }
            __syncthreads();
            float* da = S.Da;
            float* db = S.Db;
#pragma unroll
            for (int l = NL - 1; l >= 0; --l) {
                const float* hin = S.Hh[l];
                if (t < 169) wgrad16<13, LD, LD, true>(acc[l], da, hin, t);   // rows >= nrows: dz rows are zero
                gemm_nn_r<SR, 13, LD, LD>(da, S.Wp + SW_PAIR + l * SW_PAIR_SZ, 13, nrows, EpiNN{db, LD, hin, SR});
                __syncthreads();
                float* tmp = da; da = db; db = tmp;
            }
            if (t >= 169 && t < 169 + 77) {
                wgrad<11, LD, LD, false>(acc[0], da, S.Xc, t - 169, nrows, H2);
                wgrad<11, LD, LD, true>(acc[1], da, S.Xf, t - 169, nrows, 0);
            }
        }
        trace_mark(A, 4);
        pair_partials_store(A.partial + (size_t)blockIdx.x * PT_TOTAL, acc);
        trace_mark(A, 5);
    } else {
        // ---------------------------------------------------- decoder tile ----------------------------------------------------
        StepDecSmem& S = U_.dec;
        const int tile = blockIdx.x - A.gp;
        const int f0 = tile * SR;
        const int nf = min(SR, A.F - f0);
        for (int idx = t; idx < 5 * SR * LD; idx += NT) (&S.Dz[0][0])[idx] = 0.f;
        stage_init(S.W);
        if (t < nf) S.foff[t] = A.foc_row[f0 + t];
        if (t <= nf) S.pstart[t] = A.pair_start[f0 + t];
        __syncthreads();
        load_focus_tile<SR>(S.Z, A.states, S.foff, nf);
        __syncthreads();
        trace_mark(A, 0);
        pdl_wait();
        trace_mark(A, 1);
        pdl_launch_dependents();
        stage_copy(S.W, nullptr, S.Wd, A.wpack);
        const int p_lo = S.pstart[0], p_hi = S.pstart[nf];
        if (p_hi > p_lo) wait_flags(A.flags, p_lo / SR, (p_hi - 1) / SR, A.stamp, A.abort);
        trace_mark(A, 2);
        // per-focus sums of the pair outputs in list order (focus_sums, reading through L2: the rows were written by other SMs)
        for (int idx = t; idx < nf * 13; idx += NT) {
            const int f = idx / 13, c4 = idx - f * 13;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            const int p1 = S.pstart[f + 1];
            const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int p = S.pstart[f]; p < p1; p += 4) {       // four rows in flight; added in list order
                const float* src = A.Hp + (size_t)p * 52 + 4 * c4;
                const float4 v0 = ldcg4(src);
                const float4 v1 = p + 1 < p1 ? ldcg4(src + 52) : zero4;
                const float4 v2 = p + 2 < p1 ? ldcg4(src + 104) : zero4;
                const float4 v3 = p + 3 < p1 ? ldcg4(src + 156) : zero4;
                a.x += v0.x; a.y += v0.y; a.z += v0.z; a.w += v0.w;
                a.x += v1.x; a.y += v1.y; a.z += v1.z; a.w += v1.w;
                a.x += v2.x; a.y += v2.y; a.z += v2.z; a.w += v2.w;
                a.x += v3.x; a.y += v3.y; a.z += v3.z; a.w += v3.w;
            }
            if (c4 == 12) { a.z = 1.f; a.w = 0.f; }
            st4(S.Z + f * LD1 + 4 * c4, a);
        }
        __syncthreads();
        stage_wait(S.W);
        float* const ubufs[4] = {S.U[0], S.U[1], S.U[2], S.U[3]};
        decoder_forward<SR>(S.Wd, S.Z, ubufs, S.O, nf);
        float* dz5 = S.Dz[4];
        if (t < nf) {
            const float* yy = A.y + (size_t)(f0 + t) * OUT;
            const float e2 = S.O[t * LD + 2] - yy[2], e3 = S.O[t * LD + 3] - yy[3], e5 = S.O[t * LD + 5] - yy[5];
            dz5[t * LD + 2] = e2 * A.scale_v;
            dz5[t * LD + 3] = e3 * A.scale_v;
            dz5[t * LD + 5] = e5 * A.scale_w;
        }
        __syncthreads();
        // input-gradient chain first: ds is what the pair tiles are waiting for
        gemm_nn_r<SR, 13, LD, LD>(S.Dz[4], S.Wd + SW_DEC5, 6, nf, EpiNN{S.Dz[3], LD, S.U[3], SR});
        __syncthreads();
        gemm_nn_r<SR, 13, LD, LD>(S.Dz[3], S.Wd + SW_DEC2 + 2 * SW_PAIR_SZ, 13, nf, EpiNN{S.Dz[2], LD, S.U[2], SR});
        __syncthreads();
        gemm_nn_r<SR, 13, LD, LD>(S.Dz[2], S.Wd + SW_DEC2 + SW_PAIR_SZ, 13, nf, EpiNN{S.Dz[1], LD, S.U[1], SR});
        __syncthreads();
        gemm_nn_r<SR, 13, LD, LD>(S.Dz[1], S.Wd + SW_DEC2, 13, nf, EpiNN{S.Dz[0], LD, S.U[0], SR});
        __syncthreads();
        gemm_nn_r<SR, 13, LD, LD1>(S.Dz[0], S.Wd + SW_DEC1, 13, nf, EpiNN{A.ds + (size_t)f0 * 52, 52, nullptr, nf});
        publish_flag(A.flags + A.gp + tile, A.stamp);
        trace_mark(A, 3);
        // model:fp outputs (prediction with the sigmoid split, per-example losses): nobody in this grid waits for them
        for (int idx = t; idx < nf * OUT; idx += NT) {
            const int r = idx / OUT, n = idx - r * OUT;
            float v = S.O[r * LD + n];
            if (n >= 6) v = 1.f / (1.f + expf(-v));
            A.pred[(size_t)(f0 + r) * OUT + n] = v;
        }
        if (t < nf) {
            const float* yy = A.y + (size_t)(f0 + t) * OUT;
            const float e2 = S.O[t * LD + 2] - yy[2], e3 = S.O[t * LD + 3] - yy[3], e5 = S.O[t * LD + 5] - yy[5];
            const float lv = __fadd_rn(__fmul_rn(e2, e2), __fmul_rn(e3, e3)), lw = __fmul_rn(e5, e5);
            A.loss_ex[(size_t)(f0 + t) * 3 + 0] = __fdiv_rn(__fadd_rn(lv, lw), 3.f);
            A.loss_ex[(size_t)(f0 + t) * 3 + 1] = __fdiv_rn(lv, 2.f);
            A.loss_ex[(size_t)(f0 + t) * 3 + 2] = lw;
        }
        // weight gradients (off the critical path of the step: the pair tiles run their backward meanwhile)
        float a1a[16], a1b[16], a2[16], a3[16], a4[16], a5[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) { a1a[i] = a1b[i] = a2[i] = a3[i] = a4[i] = a5[i] = 0.f; }
        if (t >= 169 && t < 169 + 78) wgrad<13, LD, LD, true>(a5, S.Dz[4], S.U[3], t - 169, nf);
        if (t < 169) {
            wgrad<13, LD, LD, true>(a4, S.Dz[3], S.U[2], t, nf);
            wgrad<13, LD, LD, true>(a3, S.Dz[2], S.U[1], t, nf);
            wgrad<13, LD, LD, true>(a2, S.Dz[1], S.U[0], t, nf);
        }
        wgrad<24, LD, LD1, true>(a1a, S.Dz[0], S.Z, t, nf);
        if (t + NT < 13 * 24) wgrad<24, LD, LD1, true>(a1b, S.Dz[0], S.Z, t + NT, nf);
        trace_mark(A, 4);
        dec_partials_store(A.partial + (size_t)tile * PT_TOTAL, a1a, a1b, a2, a3, a4, a5);
        trace_mark(A, 5);
    }
}

}  // namespace npe
