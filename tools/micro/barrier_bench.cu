// Micro-benchmark: cost of one grid-wide barrier on B200 for different grid shapes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o barrier_bench barrier_bench.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int& epoch)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch += gridDim.x;
        __threadfence();
        atomicAdd(bar, 1u);
        unsigned int v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory"); } while ((int)(v - epoch) < 0);
    }
    __syncthreads();
}
// all-to-all flags: every block publishes its epoch in its own 4-byte flag (no atomics, no same-address serialisation);
// warp 0 of every block polls all flags
__device__ __forceinline__ void flag_barrier(unsigned int* flags, unsigned int& epoch)
{
    __syncthreads();
    epoch++;
    if (threadIdx.x < 32) {
        if (threadIdx.x == 0) { asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(flags + blockIdx.x), "r"(epoch) : "memory"); }
        bool all;
        do {
            bool ok = true;
            for (unsigned int b = threadIdx.x; b < gridDim.x; b += 32) {
                unsigned int v;
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + b) : "memory");
                ok = ok && (int)(v - epoch) >= 0;
            }
            all = __all_sync(0xffffffffu, ok);
        } while (!all);
    }
    __syncthreads();
}
__global__ void k_flags(unsigned int* flags, int n, int* sink)
{
    unsigned int epoch = 0;
    for (int i = 0; i < n; i++) flag_barrier(flags, epoch);
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = epoch;
}
// two stages: groups of GRP blocks arrive on their own counter, the last arriver of a group arrives on the top counter,
// the last arriver there publishes the epoch in a flag everybody polls (polls do not hit the atomics' addresses)
#define GRP 16
__device__ __forceinline__ void tree_barrier(unsigned int* bars, unsigned int& epoch)
{
    __syncthreads();
    if (threadIdx.x == 0) {
        epoch++;
        const unsigned int g = blockIdx.x / GRP, ng = (gridDim.x + GRP - 1) / GRP;
        const unsigned int gsize = min((unsigned int)GRP, gridDim.x - g * GRP);
        __threadfence();
        if (atomicAdd(bars + 32 + g * 32, 1u) == epoch * gsize - 1) {
            if (atomicAdd(bars + 16, 1u) == epoch * ng - 1) { asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(bars), "r"(epoch) : "memory"); }
        }
        unsigned int v;
        do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bars) : "memory"); } while ((int)(v - epoch) < 0);
    }
    __syncthreads();
}
__global__ void k_tree(unsigned int* bars, int n, int* sink)
{
    unsigned int epoch = 0;
    for (int i = 0; i < n; i++) tree_barrier(bars, epoch);
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = epoch;
}
__global__ void k_custom(unsigned int* bar, int n, int* sink)
{
    unsigned int epoch = 0;
    for (int i = 0; i < n; i++) grid_barrier(bar, epoch);
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = epoch;
}
__global__ void k_cg(int n, int* sink)
{
    cg::grid_group g = cg::this_grid();
    for (int i = 0; i < n; i++) g.sync();
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = n;
}
// barrier + one dependent global load chain of depth d (models "read descriptor -> gather body")
__global__ void k_custom_chain(unsigned int* bar, int n, int* chain, int depth, int* sink)
{
    unsigned int epoch = 0; int x = threadIdx.x & 1;
    for (int i = 0; i < n; i++) {
        for (int d = 0; d < depth; d++) x = __ldcg(&chain[x]);
        grid_barrier(bar, epoch);
    }
    if (x == 12345) *sink = x;
}
int main()
{
    unsigned int* bar; int *sink, *chain;
    cudaMalloc(&bar, 1 << 16); cudaMalloc(&sink, 4); cudaMalloc(&chain, 1 << 20); cudaMemset(chain, 0, 1 << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int n = 2000;
    int shapes[][2] = {{148, 256}, {296, 256}, {444, 256}, {592, 256}, {148, 768}, {148, 1024}, {148, 128}};
    for (auto& s : shapes) {
        for (int variant = 0; variant < 6; variant++) {
            cudaMemset(bar, 0, 1 << 16);
            int depth = variant == 2 ? 1 : 3;
            void* a0[] = {&bar, &n, &sink}; void* a1[] = {&n, &sink}; void* a2[] = {&bar, &n, &chain, &depth, &sink};
            cudaEventRecord(e0);
            cudaError_t rc;
            if (variant == 0) rc = cudaLaunchCooperativeKernel((void*)k_custom, dim3(s[0]), dim3(s[1]), a0, 0, 0);
            else if (variant == 1) rc = cudaLaunchCooperativeKernel((void*)k_cg, dim3(s[0]), dim3(s[1]), a1, 0, 0);
            else if (variant == 4) rc = cudaLaunchCooperativeKernel((void*)k_flags, dim3(s[0]), dim3(s[1]), a0, 0, 0);
            else if (variant == 5) rc = cudaLaunchCooperativeKernel((void*)k_tree, dim3(s[0]), dim3(s[1]), a0, 0, 0);
            else rc = cudaLaunchCooperativeKernel((void*)k_custom_chain, dim3(s[0]), dim3(s[1]), a2, 0, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            const char* names[] = {"custom", "cg.sync", "custom+1 load", "custom+3 dependent loads", "flags all-to-all", "two-stage tree"};
            printf("grid %4d x %4d  %-26s %7.3f us/barrier  (%s)\n", s[0], s[1], names[variant], ms * 1e3 / n, cudaGetErrorString(rc));
        }
    }
    return 0;
}
