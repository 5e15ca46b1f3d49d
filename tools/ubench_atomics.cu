// Micro-benchmark: shared-memory atomics vs L2 (global) reductions for histogram accumulation on B200.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/ubench_atomics tools/ubench_atomics.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define WORDS 8448  // 8 features x 1056 words (C = 4)

template <int NS, int NG>
__global__ void __launch_bounds__(256) k(uint32_t* ghist, int iters, uint32_t* sink) {
    extern __shared__ uint32_t sh[];
    for (int i = threadIdx.x; i < WORDS; i += 256) sh[i] = 0;
    __syncthreads();
    uint32_t* gh = ghist + (size_t)blockIdx.x * WORDS;
    uint32_t x = threadIdx.x * 2654435761u + blockIdx.x * 40503u + 12345u;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < NS; ++j) {
            x = x * 1664525u + 1013904223u;
            atomicAdd(&sh[j * 1056 + (x >> 22)], 1u);
        }
#pragma unroll
        for (int j = 0; j < NG; ++j) {
            x = x * 1664525u + 1013904223u;
            atomicAdd(&gh[(NS + j) * 1056 + (x >> 22)], 1u);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) sink[blockIdx.x] = sh[x % WORDS];
}

template <int NS, int NG>
void run(uint32_t* gh, uint32_t* sink, int blocks) {
    int iters = 4000;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaFuncSetAttribute(k<NS, NG>, cudaFuncAttributeMaxDynamicSharedMemorySize, WORDS * 4);
    k<NS, NG><<<blocks, 256, WORDS * 4>>>(gh, 100, sink);
    cudaDeviceSynchronize();
    cudaEventRecord(a);
    k<NS, NG><<<blocks, 256, WORDS * 4>>>(gh, iters, sink);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    double at = (double)blocks * 256 * iters * (NS + NG);
    printf("smem %d + global %d per row: %.2f ms, %.1f G atomics/s, %.2f G rows/s (err %s)\n", NS, NG, ms, at / ms / 1e6,
           at / (NS + NG) / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int blocks = 148 * 4;
    uint32_t *gh, *sink;
    cudaMalloc(&gh, (size_t)blocks * WORDS * 4);
    cudaMemset(gh, 0, (size_t)blocks * WORDS * 4);
    cudaMalloc(&sink, blocks * 4);
    run<8, 0>(gh, sink, blocks);
    run<0, 8>(gh, sink, blocks);
    run<4, 4>(gh, sink, blocks);
    run<3, 5>(gh, sink, blocks);
    run<5, 3>(gh, sink, blocks);
    run<6, 2>(gh, sink, blocks);
    run<4, 0>(gh, sink, blocks);
    return 0;
}
