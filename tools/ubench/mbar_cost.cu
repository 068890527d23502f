// Micro-benchmark: what a SATISFIED mbarrier wait costs (the phase is already complete when the wait is issued), for the wait forms
// the kernels use; plus mbarrier.arrive and elect.sync.  build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../gpu-ai-inference-server_b200/csrc -o mbar_cost mbar_cost.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "umma_ptx.cuh"
using namespace b200::kernels;

__global__ void k(long long* out, int iters) {
    __shared__ uint64_t bar[2];
    if (threadIdx.x == 0) { MbarInit(&bar[0], 1); MbarInit(&bar[1], 1); FenceBarrierInit(); }
    __syncthreads();
    if (threadIdx.x == 0) MbarArrive(&bar[0]);   // phase 0 of bar[0] is complete from now on
    __syncthreads();
    long long t[8];
    if (threadIdx.x < 32) {
        long long a = clock64();
        for (int i = 0; i < iters; ++i) MbarWaitWarp(&bar[0], 0);
        t[0] = clock64() - a;
        a = clock64();
        for (int i = 0; i < iters; ++i) MbarWait(&bar[0], 0);
        t[1] = clock64() - a;
        a = clock64();
        uint32_t acc = 0;
        for (int i = 0; i < iters; ++i) acc += MbarTest(&bar[0], 0);
        t[2] = clock64() - a;
        a = clock64();
        for (int i = 0; i < iters; ++i) if (ElectOne()) acc += 1;
        t[3] = clock64() - a;
        a = clock64();
        for (int i = 0; i < iters; ++i) { if ((threadIdx.x & 31) == 0) MbarArrive(&bar[1]); __syncwarp(); }
        t[4] = clock64() - a;
        a = clock64();
        for (int i = 0; i < iters; ++i) { TcFenceAfter(); }
        t[5] = clock64() - a;
        a = clock64();
        for (int i = 0; i < iters; ++i) { FenceProxyAsync(); }
        t[6] = clock64() - a;
        if (threadIdx.x == 0) { for (int j = 0; j < 7; ++j) out[j] = t[j]; out[7] = acc; }
    }
}
int main() {
    long long* d; cudaMalloc(&d, 64);
    const int iters = 4096;
    k<<<1, 128>>>(d, iters); k<<<1, 128>>>(d, iters);
    cudaDeviceSynchronize();
    long long h[8]; cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    const char* names[7] = {"MbarWaitWarp (lane 0 try_wait + syncwarp), satisfied", "MbarWait (32 lanes try_wait), satisfied", "mbarrier.test_wait, satisfied",
                            "elect.sync", "lane-0 mbarrier.arrive + syncwarp", "tcgen05.fence::after_thread_sync", "fence.proxy.async.shared::cta"};
    for (int j = 0; j < 7; ++j) printf("%-60s %8.1f cycles\n", names[j], (double)h[j] / iters);
    return 0;
}
