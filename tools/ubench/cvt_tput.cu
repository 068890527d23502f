// Micro-benchmark: per-SM issue throughput of the conversion / packed-math instructions the fp8 path leans on.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cvt_tput cvt_tput.cu ; run: ./cvt_tput
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(uint32_t* out, int iters, uint32_t seed) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + threadIdx.x * 8 + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) {  // e4m3x2 -> f16x2
                uint32_t d; asm volatile("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(d) : "h"((unsigned short)a[i])); a[i] = d;
            } else if (OP == 1) {  // f16x2 -> e4m3x2
                unsigned short d; asm volatile("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(d) : "r"(a[i])); a[i] = d | (a[i] << 16);
            } else if (OP == 2) {  // f32,f32 -> e4m3x2 (relu)
                unsigned short d; asm volatile("cvt.rn.satfinite.relu.e4m3x2.f32 %0, %1, %2;" : "=h"(d) : "f"(__uint_as_float(a[i])), "f"(__uint_as_float(a[(i + 1) & 7]))); a[i] = d | 0x3f800000u;
            } else if (OP == 3) {  // hfma2.relu
                uint32_t d; asm volatile("fma.rn.relu.f16x2 %0, %1, %2, %1;" : "=r"(d) : "r"(a[i]), "r"(a[(i + 1) & 7])); a[i] = d;
            } else if (OP == 4) {  // f32,f32 -> bf16x2
                uint32_t d; asm volatile("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(__uint_as_float(a[i])), "f"(__uint_as_float(a[(i + 1) & 7]))); a[i] = d;
            } else if (OP == 5) {  // lop3 (baseline int)
                uint32_t d; asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a[i]), "r"(a[(i + 1) & 7]), "r"(seed)); a[i] = d;
            } else if (OP == 6) {  // ffma2
                unsigned long long x = ((unsigned long long)a[i] << 32) | a[(i + 1) & 7], d;
                asm volatile("fma.rn.f32x2 %0, %1, %1, %1;" : "=l"(d) : "l"(x)); a[i] = (uint32_t)d ^ (uint32_t)(d >> 32);
            } else if (OP == 7) {  // vmax4 unsigned (SIMD byte max, likely emulated)
                uint32_t d; asm volatile("vmax4.u32.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a[i]), "r"(a[(i + 1) & 7]), "r"(0)); a[i] = d;
            } else if (OP == 8) {  // bf16x2 fma relu
                uint32_t d; asm volatile("fma.rn.relu.bf16x2 %0, %1, %2, %1;" : "=r"(d) : "r"(a[i]), "r"(a[(i + 1) & 7])); a[i] = d;
            } else if (OP == 9) {  // prmt
                uint32_t d; asm volatile("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(d) : "r"(a[i]), "r"(a[(i + 1) & 7])); a[i] = d;
            } else if (OP == 10) {  // max.s16x2 (DPX-class)
                uint32_t d; asm volatile("max.s16x2 %0, %1, %2;" : "=r"(d) : "r"(a[i]), "r"(a[(i + 1) & 7])); a[i] = d;
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    if (s == 0x12345678u) out[threadIdx.x] = s;
}

template <int OP> void run(const char* name) {
    uint32_t* d; cudaMalloc(&d, 4096 * 4);
    const int iters = 4096;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<148, 1024>>>(d, 16, 1); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<OP><<<148, 1024>>>(d, iters, 1);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int mhz; cudaDeviceGetAttribute(&mhz, cudaDevAttrClockRate, 0);
    double cycles = ms * 1e-3 * mhz * 1e3;
    double warp_instr_per_sm = (double)iters * 8 * 32;  // 32 warps per SM
    printf("%-28s %8.3f ms  %6.2f warp-instr/clk/SM  (%5.1f lanes/clk/SM)  [%s]\n", name, ms, warp_instr_per_sm / cycles, 32 * warp_instr_per_sm / cycles, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d);
}

int main() {
    run<5>("lop3 (baseline)");
    run<0>("cvt f16x2 <- e4m3x2");
    run<1>("cvt e4m3x2 <- f16x2");
    run<2>("cvt e4m3x2.relu <- f32,f32");
    run<4>("cvt bf16x2.relu <- f32,f32");
    run<3>("fma.relu.f16x2");
    run<8>("fma.relu.bf16x2");
    run<6>("fma.f32x2");
    run<7>("vmax4.u32");
    run<9>("prmt");
    run<10>("max.s16x2");
    return 0;
}
