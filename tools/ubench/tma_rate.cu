// Micro-benchmark: how fast ONE SM's TMA unit lands [rows][128 B] boxes (SWIZZLE_128B) as a function of the row pitch in global
// memory and of the footprint (L2-resident or streaming from HBM).  Every tcgen05 kernel of the engine feeds itself with such boxes
// (a 128-byte channel slice of each NHWC pixel), so this rate is a ceiling for all of them.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../gpu-ai-inference-server_b200/csrc -I ../../include -o tma_rate tma_rate.cu ../../gpu-ai-inference-server_b200/csrc/tmap.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "kernels.h"
#include "umma_ptx.cuh"

using namespace b200::kernels;

constexpr int kMaxStages = 12;

__global__ void __launch_bounds__(64, 1) tma_rate_kernel(const __grid_constant__ CUtensorMap tmap, int rows_total, int box_rows, int iters, int kStages,
                                                         long long* cycles_out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full[kMaxStages];
    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) MbarInit(&full[s], 1);
        FenceBarrierInit();
    }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int tiles = rows_total / box_rows;
        long long issue_cyc = 0;
        long long t0 = clock64();
        // producer and consumer in one warp: keep kStages boxes in flight, wait for the oldest, reissue
        for (int i = 0; i < iters + kStages; ++i) {
            const int s = i % kStages;
            if (i >= kStages) MbarWaitWarp(&full[s], ((i / kStages) - 1) & 1);
            if (i < iters) {
                const int tile = (int)(((unsigned)blockIdx.x * 7919u + (unsigned)i * gridDim.x) % (unsigned)tiles);
                if (ElectOne()) {
                    MbarArriveExpectTx(&full[s], (uint32_t)(box_rows * 128));
                    long long a = clock64();
                    TmaLoad2D(smem + s * (box_rows * 128), &tmap, &full[s], 0, tile * box_rows);
                    issue_cyc += clock64() - a;
                }
                __syncwarp();
            }
        }
        long long t1 = clock64();
        issue_cyc = __reduce_max_sync(0xffffffffu, (unsigned)issue_cyc);
        if (threadIdx.x == 0) { cycles_out[blockIdx.x] = t1 - t0; cycles_out[gridDim.x + blockIdx.x] = issue_cyc; }
    }
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long* d_cyc;
    cudaMalloc(&d_cyc, 2 * sms * sizeof(long long));
    const int smem = 1024 + 6 * 32768;
    cudaFuncSetAttribute(tma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    printf("TMA box [rows][128 B], SWIZZLE_128B, %d SMs\n", sms);
    printf("%8s %8s %8s %12s %12s %14s %12s\n", "pitch_B", "box_rows", "inflight", "footprint_MB", "B/clk/SM", "chip_TB/s@1.9", "cyc/box");
    for (size_t foot_mb : {32, 1024}) {
        for (int pitch : {256}) {
            for (int box_rows : {32, 128, 256})
            for (int kStages : {4, 6}) {
                if (kStages * box_rows * 128 > 6 * 32768) continue;
                const size_t rows = foot_mb * 1024 * 1024 / pitch;
                void* buf;
                if (cudaMalloc(&buf, rows * pitch) != cudaSuccess) { printf("alloc failed\n"); return 1; }
                cudaMemset(buf, 1, rows * pitch);
                TensorMap tm;
                const uint64_t dims[2] = {(uint64_t)pitch, (uint64_t)rows};
                const uint64_t strides[1] = {(uint64_t)pitch};
                const uint32_t box[2] = {128u, (uint32_t)box_rows};
                if (MakeTensorMap(&tm, buf, 1, 2, dims, strides, box, true) != 0) { printf("tensor map failed\n"); return 1; }
                const int iters = 4096 * 128 / box_rows / 4;
                for (int rep = 0; rep < 2; ++rep)
                    tma_rate_kernel<<<sms, 64, smem>>>(*reinterpret_cast<const CUtensorMap*>(&tm), (int)rows, box_rows, iters, kStages, d_cyc);
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
                long long h[512];
                cudaMemcpy(h, d_cyc, 2 * sms * sizeof(long long), cudaMemcpyDeviceToHost);
                double sum = 0, isum = 0;
                for (int i = 0; i < sms; ++i) { sum += (double)h[i]; isum += (double)h[sms + i]; }
                const double cyc = sum / sms;
                const double bpc = (double)iters * box_rows * 128 / cyc;
                printf("%8d %8d %8d %12zu %12.1f %14.2f %12.1f   issue %.0f cyc\n", pitch, box_rows, kStages, foot_mb, bpc, bpc * sms * 1.9e9 / 1e12, cyc / (double)iters, isum / sms / iters);
                cudaFree(buf);
            }
        }
    }
    return 0;
}
