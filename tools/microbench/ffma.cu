// FP32 issue-rate microbenchmark for sm_100a: scalar FFMA vs packed fma.rn.f32x2 (FFMA2), register-only.
// Answers: which instruction reaches the 128 FMA/clk/SM FP32 peak on B200?
#include <cuda_runtime.h>
#include <cstdio>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
    if (MODE == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
        }
    } else {
        unsigned long long p[8], aa, bb;
        asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a), "f"(a));
        asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b), "f"(b));
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(acc[2 * i]), "f"(acc[2 * i + 1]));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(aa), "l"(bb));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * i]), "=f"(acc[2 * i + 1]) : "l"(p[i]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// register-operand GEMM-like pattern: acc(16) += x[j] * w(16), all operands in ordinary registers
template <int MODE>
__global__ void __launch_bounds__(256) kg(float* out, const float* in, int iters) {
    float acc[16], w[16], x[4];
#pragma unroll
    for (int i = 0; i < 16; ++i) { acc[i] = 0.f; w[i] = in[threadIdx.x + 32 * i]; }
#pragma unroll
    for (int j = 0; j < 4; ++j) x[j] = in[threadIdx.x + 7 * j + 1];
    if (MODE == 0) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = fmaf(x[j], w[(i + j) & 15], acc[i]);
        }
    } else {
        unsigned long long p[8], wp[8], xp[4];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(acc[2 * i]), "f"(acc[2 * i + 1]));
            asm("mov.b64 %0, {%1, %2};" : "=l"(wp[i]) : "f"(w[2 * i]), "f"(w[2 * i + 1]));
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) asm("mov.b64 %0, {%1, %2};" : "=l"(xp[j]) : "f"(x[j]), "f"(x[j]));
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
#pragma unroll
                for (int i = 0; i < 8; ++i)
                    asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(xp[j]), "l"(wp[(i + j) & 7]));
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("mov.b64 {%0, %1}, %2;" : "=f"(acc[2 * i]), "=f"(acc[2 * i + 1]) : "l"(p[i]));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    int grid = sms * 8, iters = 20000;
    float* out;
    cudaMalloc(&out, grid * 256 * sizeof(float));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<grid, 256>>>(out, iters, 1.0001f, 0.5f);
            else k<1><<<grid, 256>>>(out, iters, 1.0001f, 0.5f);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            double fma = (double)grid * 256 * 16.0 * iters;
            printf("%s rep %d: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM at %d MHz nominal)\n", mode ? "FFMA2(f32x2)" : "FFMA scalar ",
                   rep, ms, 2 * fma / ms * 1e-9, fma / (ms * 1e-3) / sms / (prop.clockRate * 1e3), prop.clockRate / 1000);
        }
    }
    float* in;
    cudaMalloc(&in, 4096 * sizeof(float));
    cudaMemset(in, 0, 4096 * sizeof(float));
    for (int mode = 0; mode < 2; ++mode) {
        for (int occ = 0; occ < 2; ++occ) {
            int g2 = sms * (occ ? 8 : 2), it2 = 5000;
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) kg<0><<<g2, 256>>>(out, in, it2);
                else kg<1><<<g2, 256>>>(out, in, it2);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                double fma = (double)g2 * 256 * 64.0 * it2;
                printf("gemm-like %s ctas/sm=%d rep %d: %.3f ms  %.2f TFLOP/s  (%.1f FMA/clk/SM at nominal clock)\n",
                       mode ? "FFMA2" : "FFMA ", occ ? 8 : 2, rep, ms, 2 * fma / ms * 1e-9,
                       fma / (ms * 1e-3) / sms / (prop.clockRate * 1e3));
            }
        }
    }
    printf("cudaGetLastError: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
