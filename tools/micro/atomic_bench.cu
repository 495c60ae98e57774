// Same-address vs spread global atomics with a returned value (what push_successor / warp_reserve issue):
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o atomic_bench atomic_bench.cu && ./atomic_bench
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k_atom(int* ctr, int n_addr, int per_warp, int* sink)
{
    int lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int acc = 0;
    if (lane == 0)
        for (int i = 0; i < per_warp; i++) acc += atomicAdd(&ctr[((warp + i) % n_addr) * 32], 1);     // 128-byte spaced addresses
    if (acc == -1) sink[0] = acc;
}
int main()
{
    int *ctr, *sink; cudaMalloc(&ctr, 4096 * 128); cudaMalloc(&sink, 4); cudaMemset(ctr, 0, 4096 * 128);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 296, threads = 256, per_warp = 64;
    const int addrs[] = {1, 3, 8, 64, 1024};
    for (int a : addrs) {
        k_atom<<<blocks, threads>>>(ctr, a, per_warp, sink); cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k_atom<<<blocks, threads>>>(ctr, a, per_warp, sink);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)blocks * (threads / 32) * per_warp;
        printf("%5d addresses: %.0f atomics in %.3f ms = %.2f ns per atomic (chip-wide)\n", a, ops, ms, ms * 1e6 / ops);
    }
    return 0;
}
