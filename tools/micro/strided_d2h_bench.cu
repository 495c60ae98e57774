// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/strided_d2h_bench tools/micro/strided_d2h_bench.cu
// How should 12 bytes per 104-byte host record (the accumulated impulses of world->constraints) reach pinned host
// memory after the rest of the record has been copied early?  (a) cudaMemcpy2DAsync, width 12, dpitch 104;
// (b) a kernel that writes them straight into mapped host memory; (c) for scale: the contiguous 104-byte copy.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__global__ void k_scatter(const float* __restrict__ src, float* dst, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float a = src[3 * i], b = src[3 * i + 1], c = src[3 * i + 2];
    float* d = dst + (size_t)i * 26 + 20;
    d[0] = a; d[1] = b; d[2] = c;
}
int main()
{
    const int n = 6000000;
    float *dsrc, *dfull, *host, *hostdev;
    CK(cudaMalloc(&dsrc, sizeof(float) * 3 * (size_t)n));
    CK(cudaMalloc(&dfull, 104 * (size_t)n));
    CK(cudaHostAlloc(&host, 104 * (size_t)n, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer(&hostdev, host, 0));
    CK(cudaMemset(dsrc, 0, sizeof(float) * 3 * (size_t)n));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        CK(cudaMemcpyAsync(host, dfull, 104 * (size_t)n, cudaMemcpyDeviceToHost, 0));
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("contiguous 104 B x %d: %.3f ms (%.1f GB/s)\n", n, ms, 104.0 * n / ms / 1e6);
        cudaEventRecord(e0);
        CK(cudaMemcpy2DAsync((char*)host + 80, 104, dsrc, 12, 12, n, cudaMemcpyDeviceToHost, 0));
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("memcpy2D 12 B rows, dpitch 104: %.3f ms\n", ms);
        cudaEventRecord(e0);
        k_scatter<<<(n + 255) / 256, 256>>>(dsrc, hostdev, n);
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        CK(cudaGetLastError());
        printf("zero-copy scatter kernel: %.3f ms\n", ms);
        cudaEventRecord(e0);
        CK(cudaMemcpyAsync(host, dsrc, 12 * (size_t)n, cudaMemcpyDeviceToHost, 0));
        cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        printf("contiguous 12 B x %d: %.3f ms\n", n, ms);
    }
    return 0;
}
