// Micro-benchmark: HBM read bandwidth when only the first `used` bytes of every `pitch`-byte pixel are read
// (the access pattern of a DenseNet 1x1 conv reading a channel slice of the concat-in-place block buffer).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__global__ void rd(const uint4* __restrict__ p, size_t pixels, int pitch16, int used16, uint4* sink) {
    uint4 acc = make_uint4(0, 0, 0, 0);
    size_t total = pixels * used16;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t px = i / used16; int k = (int)(i - px * used16);
        uint4 v = __ldcs(p + px * pitch16 + k);
        acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
    }
    if (acc.x == 0x12345 && acc.y == 0x777) *sink = acc;
}
__global__ void wr(uint4* p, size_t pixels, int pitch16, int used16, int off16) {
    size_t total = pixels * used16;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t px = i / used16; int k = (int)(i - px * used16);
        p[px * pitch16 + off16 + k] = make_uint4((uint32_t)i, 1, 2, 3);
    }
}
__global__ void flush(uint4* p, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) p[i] = make_uint4(1, 2, 3, 4);
}
int main() {
    const size_t pixels = 802816;  // 256 x 56 x 56
    uint4 *buf, *fl, *sink;
    cudaMalloc(&buf, pixels * 1024); cudaMalloc(&fl, 512u << 20); cudaMalloc(&sink, 64);
    cudaMemset(buf, 1, pixels * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    struct C { int pitch, used; } cases[] = {{256, 256}, {256, 128}, {256, 64}, {256, 224}, {128, 128}, {512, 128}, {512, 160}, {512, 256}, {512, 512}, {1024, 256}, {1024, 640}, {1024, 1024}};
    for (auto c : cases) {
        float best = 1e9;
        for (int r = 0; r < 3; ++r) {
            flush<<<1184, 512>>>(fl, (512u << 20) / 16);
            cudaEventRecord(e0);
            rd<<<148 * 8, 512>>>(buf, pixels, c.pitch / 16, c.used / 16, sink);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
        }
        printf("read  pitch %4d used %4d : %7.1f us  useful %6.0f GB/s\n", c.pitch, c.used, best * 1e3, pixels * (double)c.used / best / 1e6);
    }
    struct W { int pitch, used, off; } wc[] = {{256, 128, 0}, {256, 32, 64}, {512, 32, 128}, {1024, 32, 256}, {128, 128, 0}, {256, 256, 0}};
    for (auto c : wc) {
        float best = 1e9;
        for (int r = 0; r < 3; ++r) {
            flush<<<1184, 512>>>(fl, (512u << 20) / 16);
            cudaEventRecord(e0);
            wr<<<148 * 8, 512>>>(buf, pixels, c.pitch / 16, c.used / 16, c.off / 16);
            flush<<<1184, 512>>>(fl, (256u << 20) / 16);   // force write-back of the dirty lines inside the timed region
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); best = ms < best ? ms : best;
        }
        printf("write pitch %4d used %4d : %7.1f us (incl. 256 MB flush write)\n", c.pitch, c.used, best * 1e3);
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
