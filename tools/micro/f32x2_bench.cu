// Packed FP32x2 on sm_100a: is fma.rn.f32x2 as fast as a scalar FFMA (issue rate and dependent latency)?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o f32x2_bench f32x2_bench.cu && ./f32x2_bench
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long pk(float a, float b) { return ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(a); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c)
{
    unsigned long long d;
    asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
template <int MODE> __global__ void k(float* out, int n, float seed)
{
    float x = seed + threadIdx.x * 1e-3f, y = seed * 0.5f, z = 1.0001f, w = 0.999f;
    unsigned long long p = pk(x, y), q = pk(z, w), m0 = pk(-0.0f, -0.0f), one = pk(1.0f, 1.0f);
    for (int i = 0; i < n; i++) {
        if (MODE == 0) {            // scalar: two independent mul+add chains (4 instructions per step without contraction)
            x = x * z; x = x + w; y = y * w; y = y + z;
        } else {                    // packed: the same two chains as 2 instructions per step
            p = fma2(p, q, m0); p = fma2(p, one, q);
        }
    }
    if (MODE == 0) out[blockIdx.x * blockDim.x + threadIdx.x] = x + y;
    else out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float((unsigned)p) + __uint_as_float((unsigned)(p >> 32));
}
int main()
{
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int n = 1 << 16;
    for (int blocks_per_sm : {1, 8}) {
        for (int mode = 0; mode < 2; mode++) {
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<148 * blocks_per_sm, 256>>>(out, n, 1.0f); else k<1><<<148 * blocks_per_sm, 256>>>(out, n, 1.0f);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
            }
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("%d blocks/SM, %s: %.3f ms for %d steps (2 mul + 2 add each) -> %.2f ns per step\n", blocks_per_sm,
                   mode ? "packed f32x2" : "scalar      ", ms, n, ms * 1e6 / n);
        }
    }
    return 0;
}
