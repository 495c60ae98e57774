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
    cudaMalloc(&bar, 4); cudaMalloc(&sink, 4); cudaMalloc(&chain, 1 << 20); cudaMemset(chain, 0, 1 << 20);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int n = 2000;
    int shapes[][2] = {{148, 256}, {296, 256}, {444, 256}, {592, 256}, {148, 768}, {148, 1024}, {148, 128}};
    for (auto& s : shapes) {
        for (int variant = 0; variant < 4; variant++) {
            cudaMemset(bar, 0, 4);
            int depth = variant == 2 ? 1 : 3;
            void* a0[] = {&bar, &n, &sink}; void* a1[] = {&n, &sink}; void* a2[] = {&bar, &n, &chain, &depth, &sink};
            cudaEventRecord(e0);
            cudaError_t rc;
            if (variant == 0) rc = cudaLaunchCooperativeKernel((void*)k_custom, dim3(s[0]), dim3(s[1]), a0, 0, 0);
            else if (variant == 1) rc = cudaLaunchCooperativeKernel((void*)k_cg, dim3(s[0]), dim3(s[1]), a1, 0, 0);
            else rc = cudaLaunchCooperativeKernel((void*)k_custom_chain, dim3(s[0]), dim3(s[1]), a2, 0, 0);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
            const char* names[] = {"custom", "cg.sync", "custom+1 load", "custom+3 dependent loads"};
            printf("grid %4d x %4d  %-26s %7.3f us/barrier  (%s)\n", s[0], s[1], names[variant], ms * 1e3 / n, cudaGetErrorString(rc));
        }
    }
    return 0;
}
