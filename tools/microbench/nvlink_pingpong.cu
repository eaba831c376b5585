// One-way latency of a posted 8-byte {value, stamp} store into a peer GPU's memory as the data-parallel exchange uses it
// (npe_kernels.cuh: st_cell / ld_cell): GPU 0 and GPU 1 of one process bounce a stamp ROUNDS times; every wait is bounded.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o nvlink_pingpong nvlink_pingpong.cu && ./nvlink_pingpong
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__device__ __forceinline__ void st_cell(uint2* p, unsigned v, unsigned stamp) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(v), "r"(stamp) : "memory");
}
__device__ __forceinline__ uint2 ld_cell(const uint2* p) {
    uint2 v;
    asm volatile("ld.volatile.global.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long gns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

// side 0 sends stamp r and waits for the echo; side 1 waits for stamp r and echoes it.  Every thread bounces its OWN cell (as in
// the exchange: no synchronisation between threads); `width` lanes of every warp are active.
__global__ void k_pingpong(int side, uint2* mine, uint2* peer, int rounds, int width, unsigned long long* out) {
    const int lane = threadIdx.x & 31, cell = blockIdx.x * blockDim.x + threadIdx.x;
    if (lane >= width) return;
    unsigned long long t0 = 0;
    long long fails = 0;
    for (int r = 1; r <= rounds; ++r) {
        if (r == 11 && side == 0) t0 = gns();              // 10 warm-up rounds
        if (side == 0) st_cell(peer + cell, 7u, (unsigned)r);
        int spins = 0;
        while (ld_cell(mine + cell).y != (unsigned)r && ++spins < 2000000) {}
        if (spins >= 2000000) { ++fails; break; }
        if (side == 1) st_cell(peer + cell, 7u, (unsigned)r);
    }
    if (side == 0 && cell == 0) { out[0] = gns() - t0; out[1] = (unsigned long long)fails; }
}

// Same bounce with an idle gap before every round (side 0) and an optional early dummy store; side 0 accumulates only the
// time between its store and the echo.  The dummy goes to a cell nobody polls (cell + 1 of the last CTA's range is unused
// by construction: the scratch cell lives one region further, see main).
__global__ void k_idle_pingpong(int side, uint2* mine, uint2* peer, int rounds, int idle_ns, int lead_ns, unsigned long long* out) {
    const int cell = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long acc = 0;
    long long fails = 0;
    for (int r = 1; r <= rounds; ++r) {
        unsigned long long t0 = 0;
        if (side == 0) {
            if (idle_ns) {
                const unsigned long long s0 = gns();
                bool woke = false;
                while (gns() - s0 < (unsigned long long)idle_ns) {
                    if (lead_ns && !woke && gns() - s0 >= (unsigned long long)(idle_ns - lead_ns)) {
                        if (threadIdx.x == 0) st_cell(peer + 111 * 256 + blockIdx.x, 1u, 1u);     // wake-up store into a scratch cell
                        woke = true;
                    }
                    __nanosleep(200);
                }
            }
            t0 = gns();
            st_cell(peer + cell, 7u, (unsigned)r);
        }
        int spins = 0;
        while (ld_cell(mine + cell).y != (unsigned)r && ++spins < 4000000) {}
        if (spins >= 4000000) { ++fails; break; }
        if (side == 1) st_cell(peer + cell, 7u, (unsigned)r);
        if (side == 0 && r > 10) acc += gns() - t0;
    }
    if (side == 0 && cell == 0) { out[0] = acc; out[1] = (unsigned long long)fails; }
}

// What does a grid that ends with posted stores into PEER memory cost at its boundary?  1000 dependent launches of a grid
// whose only work is one 8-byte store per thread, into local or into peer memory.
__global__ void k_store_exit(uint2* dst, unsigned r) { st_cell(dst + blockIdx.x * blockDim.x + threadIdx.x, 7u, r); }

// the same with programmatic dependent launch: every grid releases its dependent at once and waits for its predecessor
// (griddepcontrol.wait: "all memory operations of the prerequisite grid performed") before it stores
__global__ void k_store_exit_pdl(uint2* dst, unsigned r) {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
    st_cell(dst + blockIdx.x * blockDim.x + threadIdx.x, 7u, r);
}

int main() {
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, 0, 1));
    if (!can) { printf("no peer access\n"); return 0; }
    uint2* cell[2];
    cudaStream_t st[2];
    unsigned long long* out;
    const int maxcells = 111 * 256;
    for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d));
        CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        CK(cudaMalloc(&cell[d], (maxcells + 256) * sizeof(uint2)));
        CK(cudaStreamCreate(&st[d]));
    }
    CK(cudaSetDevice(0));
    CK(cudaMallocManaged(&out, 16));
    const int rounds = 2010;
    const int shapes[][3] = {{1, 32, 1}, {1, 32, 32}, {1, 256, 32}, {16, 256, 32}, {111, 256, 32}, {111, 256, 8}, {111, 32, 32}};   // CTAs, threads, active lanes per warp
    for (auto& sh : shapes) {
        for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaMemset(cell[d], 0, maxcells * sizeof(uint2))); CK(cudaDeviceSynchronize()); }
        CK(cudaSetDevice(1));
        k_pingpong<<<sh[0], sh[1], 0, st[1]>>>(1, cell[1], cell[0], rounds, sh[2], out);
        CK(cudaSetDevice(0));
        k_pingpong<<<sh[0], sh[1], 0, st[0]>>>(0, cell[0], cell[1], rounds, sh[2], out);
        CK(cudaStreamSynchronize(st[0]));
        CK(cudaSetDevice(1));
        CK(cudaStreamSynchronize(st[1]));
        printf("%3d CTAs x %3d threads, %2d lanes per warp (%5d cells): round trip %.2f us, one way %.2f us (%llu timeouts)\n", sh[0], sh[1], sh[2],
               sh[0] * sh[1] / 32 * sh[2], out[0] / 1e3 / (rounds - 10), out[0] / 2e3 / (rounds - 10), out[1]);
    }
    {
        CK(cudaSetDevice(0));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int where = 0; where < 2; ++where) {
            // a graph of 500 dependent launches, so that the host's launch rate is not what is measured
            cudaGraph_t g; cudaGraphExec_t ge;
            CK(cudaStreamBeginCapture(st[0], cudaStreamCaptureModeGlobal));
            for (int r = 0; r < 500; ++r) k_store_exit<<<111, 256, 0, st[0]>>>(cell[where], 0u);
            CK(cudaStreamEndCapture(st[0], &g));
            CK(cudaGraphInstantiate(&ge, g, 0));
            CK(cudaGraphLaunch(ge, st[0]));
            CK(cudaEventRecord(e0, st[0]));
            for (int r = 0; r < 4; ++r) CK(cudaGraphLaunch(ge, st[0]));
            CK(cudaEventRecord(e1, st[0]));
            CK(cudaStreamSynchronize(st[0]));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("2000 dependent grids (graph) of one store per thread into %s memory: %.2f us per grid\n", where ? "PEER" : "local", ms / 2.0f);
        }
    }
    {
        CK(cudaSetDevice(0));
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int where = 0; where < 2; ++where) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(111); cfg.blockDim = dim3(256); cfg.stream = st[0];
            cudaLaunchAttribute at[1];
            at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; at[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            cudaGraph_t g; cudaGraphExec_t ge;
            CK(cudaStreamBeginCapture(st[0], cudaStreamCaptureModeGlobal));
            for (int r = 0; r < 500; ++r) CK(cudaLaunchKernelEx(&cfg, k_store_exit_pdl, cell[where], 0u));
            CK(cudaStreamEndCapture(st[0], &g));
            CK(cudaGraphInstantiate(&ge, g, 0));
            CK(cudaGraphLaunch(ge, st[0]));
            CK(cudaEventRecord(e0, st[0]));
            for (int r = 0; r < 4; ++r) CK(cudaGraphLaunch(ge, st[0]));
            CK(cudaEventRecord(e1, st[0]));
            CK(cudaStreamSynchronize(st[0]));
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("2000 programmatic-dependent grids (graph) of one store per thread into %s memory: %.2f us per grid\n", where ? "PEER" : "local", ms / 2.0f);
        }
    }
    // The exchange sends one burst per ~30-us training step.  Does an idle link answer as fast as a busy one?  Side 0 sleeps
    // `idle` ns before every round; optionally it sends a dummy store `lead` ns before the measured one.
    const int idles[][2] = {{0, 0}, {5000, 0}, {25000, 0}, {100000, 0}, {25000, 2000}, {25000, 5000}, {25000, 10000}};
    for (auto& id : idles) {
        for (int d = 0; d < 2; ++d) { CK(cudaSetDevice(d)); CK(cudaMemset(cell[d], 0, maxcells * sizeof(uint2))); CK(cudaDeviceSynchronize()); }
        CK(cudaSetDevice(1));
        k_idle_pingpong<<<111, 256, 0, st[1]>>>(1, cell[1], cell[0], 510, 0, 0, out);
        CK(cudaSetDevice(0));
        k_idle_pingpong<<<111, 256, 0, st[0]>>>(0, cell[0], cell[1], 510, id[0], id[1], out);
        CK(cudaStreamSynchronize(st[0]));
        CK(cudaSetDevice(1));
        CK(cudaStreamSynchronize(st[1]));
        printf("idle %6d ns between rounds, wake-up store %5d ns ahead: round trip %.2f us (%llu timeouts)\n", id[0], id[1], out[0] / 1e3 / 500, out[1]);
    }
    return 0;
}
