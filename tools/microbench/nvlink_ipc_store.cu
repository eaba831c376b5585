// Does a grid that ends with posted stores into a peer's memory cost more at its boundary when that memory belongs to ANOTHER
// PROCESS (cudaIpcOpenMemHandle, what the data-parallel exchange uses) than with same-process peer access
// (nvlink_pingpong.cu)?  The child process owns GPU 1 and exports a buffer; the parent runs chains of 2000 dependent grids
// (CUDA graph, plain and programmatic edges) that store into local memory and into the imported buffer.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o nvlink_ipc_store nvlink_ipc_store.cu && ./nvlink_ipc_store
#include <cstdio>
#include <cstdlib>
#include <unistd.h>
#include <sys/wait.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__device__ __forceinline__ void st_cell(uint2* p, unsigned v, unsigned stamp) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v), "r"(stamp) : "memory");
}
__device__ __forceinline__ void idle(unsigned ns) {          // a training step's worth of time without link traffic
    if (!ns) return;
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t0));
    do { __nanosleep(500); asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); } while (t - t0 < ns);
}
__device__ float4* g_dirty;          // local scratch that every grid overwrites first (dirty lines in L2, like a training step leaves)
__device__ unsigned g_dirty_vec;     // float4 elements of it
__device__ __forceinline__ void dirty(unsigned r) {
    const unsigned n = g_dirty_vec, stride = gridDim.x * blockDim.x;
    for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) g_dirty[i] = make_float4(r, 1.f, 2.f, 3.f);
}
__global__ void k_store(uint2* dst, unsigned idle_ns) { idle(idle_ns); dirty(idle_ns); st_cell(dst + blockIdx.x * blockDim.x + threadIdx.x, 7u, 0u); }
__global__ void k_store_pdl(uint2* dst, unsigned idle_ns) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    idle(idle_ns);
    dirty(idle_ns);
    st_cell(dst + blockIdx.x * blockDim.x + threadIdx.x, 7u, 0u);
}

// the same chain launched grid by grid into the stream (no graph): what a training loop does
static float chain_stream(cudaStream_t st, uint2* dst, bool pdl, unsigned idle_ns) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(111); cfg.blockDim = dim3(256); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto go = [&](int n) {
        for (int r = 0; r < n; ++r) {
            if (pdl) CK(cudaLaunchKernelEx(&cfg, k_store_pdl, dst, idle_ns));
            else CK(cudaLaunchKernelEx(&cfg, k_store, dst, idle_ns));
        }
    };
    go(100);
    CK(cudaEventRecord(e0, st));
    go(1000);
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms;      // us per grid
}

static float chain(cudaStream_t st, uint2* dst, bool pdl, unsigned idle_ns) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(111); cfg.blockDim = dim3(256); cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    cudaGraph_t g; cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal));
    for (int r = 0; r < 500; ++r) {
        if (pdl) CK(cudaLaunchKernelEx(&cfg, k_store_pdl, dst, idle_ns));
        else CK(cudaLaunchKernelEx(&cfg, k_store, dst, idle_ns));
    }
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e0, st));
    for (int r = 0; r < 4; ++r) CK(cudaGraphLaunch(ge, st));
    CK(cudaEventRecord(e1, st));
    CK(cudaStreamSynchronize(st));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / 2.0f;      // us per grid
}

int main() {
    int up[2], down[2];
    if (pipe(up) || pipe(down)) return 1;
    const size_t bytes = 111 * 256 * sizeof(uint2);
    pid_t pid = fork();
    if (pid == 0) {          // child: GPU 1 owns the buffer
        int n = 0;
        CK(cudaGetDeviceCount(&n));
        if (n < 2) { cudaIpcMemHandle_t none = {}; (void)!write(up[1], &none, sizeof none); return 0; }
        CK(cudaSetDevice(1));
        uint2* buf;
        CK(cudaMalloc(&buf, bytes));
        CK(cudaMemset(buf, 0, bytes));
        CK(cudaDeviceSynchronize());
        cudaIpcMemHandle_t hnd;
        CK(cudaIpcGetMemHandle(&hnd, buf));
        (void)!write(up[1], &hnd, sizeof hnd);
        char c;
        (void)!read(down[0], &c, 1);      // parent is done
        return 0;
    }
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
    CK(cudaSetDevice(0));
    cudaIpcMemHandle_t hnd;
    if (read(up[0], &hnd, sizeof hnd) != (ssize_t)sizeof hnd) return 1;
    uint2 *remote, *local;
    CK(cudaIpcOpenMemHandle((void**)&remote, hnd, cudaIpcMemLazyEnablePeerAccess));
    CK(cudaMalloc(&local, bytes));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    float4* scratch;
    CK(cudaMalloc(&scratch, 64u << 20));
    CK(cudaMemcpyToSymbol(g_dirty, &scratch, sizeof scratch));
    {
        const unsigned nvec = 0;
        CK(cudaMemcpyToSymbol(g_dirty_vec, &nvec, sizeof nvec));
        for (int pdl = 0; pdl < 2; ++pdl) {
            const float a = chain_stream(st, local, pdl, 20000u), b = chain_stream(st, remote, pdl, 20000u);
            printf("STREAM launches (no graph), %s edges, 20000 ns idle: last store into local memory %.2f us per grid, into the peer process's memory %.2f (+%.2f)\n",
                   pdl ? "programmatic" : "plain       ", a, b, b - a);
        }
    }
    for (unsigned dirty_bytes : {0u, 1u << 20, 4u << 20, 16u << 20})
        for (unsigned idle_ns : {0u, 20000u})
            for (int pdl = 0; pdl < 2; ++pdl) {
                const unsigned nvec = dirty_bytes / 16;
                CK(cudaMemcpyToSymbol(g_dirty_vec, &nvec, sizeof nvec));
                const float a = chain(st, local, pdl, idle_ns), b = chain(st, remote, pdl, idle_ns);
                printf("%2u MB written locally per grid, %s edges, %5u ns idle: last store into local memory %.2f us per grid, into the peer process's memory %.2f (+%.2f)\n",
                       dirty_bytes >> 20, pdl ? "programmatic" : "plain       ", idle_ns, a, b, b - a);
            }
    (void)!write(down[1], "x", 1);
    waitpid(pid, nullptr, 0);
    return 0;
}
