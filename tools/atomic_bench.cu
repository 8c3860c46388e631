// Microbenchmark: throughput of same-address atomicAdd (one per warp) on sm_100a, as used for tile tickets and
// site allocation.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/atomic_bench.cu -o /tmp/atomic_bench
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned long long* ctr, int per_warp, int stride_words, unsigned long long* sink) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    unsigned long long acc = 0;
    for (int i = 0; i < per_warp; ++i) {
        if (lane == 0) acc += atomicAdd(ctr + (size_t)(stride_words ? (gw * stride_words) : 0), 1ull);
        __syncwarp();
    }
    if (acc == 0xFFFFFFFFFFFFFFFFull) *sink = acc;
}
int main() {
    unsigned long long* d; cudaMalloc(&d, 1 << 26); cudaMemset(d, 0, 1 << 26);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int stride : {0, 16}) for (int warps_per_cta : {1, 8}) for (int ctas : {148, 444, 1184}) {
        const int per_warp = 2000;
        k<<<ctas, warps_per_cta * 32>>>(d, 10, stride, d + 1);
        cudaEventRecord(a);
        k<<<ctas, warps_per_cta * 32>>>(d, per_warp, stride, d + 1);
        cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double n = (double)ctas * warps_per_cta * per_warp;
        printf("stride %2d ctas %4d warps/cta %d: %.2f ns per atomic (%.1f M atomics in %.3f ms)\n", stride, ctas, warps_per_cta, ms * 1e6 / n, n / 1e6, ms);
    }
    return 0;
}
