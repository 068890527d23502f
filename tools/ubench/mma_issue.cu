// Micro-benchmark: cost of ONE tcgen05.mma (cta_group::1, M = 128, K = 32 bytes) as a function of N, operand source and kind,
// issued back to back by one thread per SM while nothing else runs (the in-situ version of this experiment - the MMA warp of
// the streaming dense-layer kernel free-running - gave the same numbers).  What it shows on B200: a dispatch costs at least
// ~115-130 cycles regardless of N; only N = 256 reaches the nominal M*N/256 cycles.  DenseNet's convolutions have Cout = 128
// (1x1) and 32 (3x3; 96 with three taps stacked along N), so every MMA of the network runs at 40-55 % of the tensor pipe's rate
// before any other limit applies - the reason the >= 50 % tensor-roofline target of BASELINE.json is out of reach at M = 128.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I ../../gpu-ai-inference-server_b200/csrc -o mma_issue mma_issue.cu
// run:   ./mma_issue
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#include "umma_ptx.cuh"

using namespace b200::kernels;

// mode 0: SS (A and B from shared memory), 1: TS (A from tensor memory); kind 1: f8f6f4 (e4m3), 0: f16 (bf16); accs: accumulators cycled
__global__ void __launch_bounds__(128, 1) mma_issue_kernel(int n, int mode, int kind, int accs, int count, long long* cycles_out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_a = smem;                 // [128][128 B]
    uint8_t* s_b = smem + 16384;         // [256][128 B]
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { MbarInit(&bar, 1); FenceBarrierInit(); }
    if (warp == 1) TmemAlloc(&tmem_slot, 512);
    FenceProxyAsync();
    TcFenceBefore();
    __syncthreads();
    TcFenceAfter();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        const uint32_t idesc = MakeInstrDesc(kind == 1 ? 0 : 1, n);
        const uint64_t a_desc = MakeSmemDesc(SmemAddr(s_a)), b_desc = MakeSmemDesc(SmemAddr(s_b));
        long long t0 = 0, t1 = 0;
        if (ElectOne()) {
            t0 = clock64();
            for (int i = 0; i < count; ++i) {
                const uint32_t d = tmem + (uint32_t)(accs > 1 ? (i % accs) * 256 / accs * (accs == 2 ? 1 : 1) : 0);
                const int ks = i & 3;
                if (mode == 1) UmmaTS(d, tmem + 480 + ks * 8, b_desc + (uint64_t)(2 * ks), idesc, i >= accs ? 1u : 0u);
                else if (kind == 1) UmmaSS<1>(d, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc, i >= accs ? 1u : 0u);
                else UmmaSS<0>(d, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc, i >= accs ? 1u : 0u);
            }
            UmmaCommit(&bar);
        }
        __syncwarp();
        MbarWaitWarp(&bar, 0);
        t1 = clock64();
        t0 = __shfl_sync(0xffffffffu, t0, __ffs(__activemask()) - 1);  // whichever lane was elected holds t0; others hold 0
        long long tmax = t0;
        for (int o = 16; o; o >>= 1) { long long v = __shfl_xor_sync(0xffffffffu, tmax, o); tmax = v > tmax ? v : tmax; }
        if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - tmax;
    }
    TcFenceBefore();
    __syncthreads();
    if (warp == 1) { TcFenceAfter(); TmemDealloc(tmem, 512); }
}


// CTA pair (cta_group::2): M = 256 across the two CTAs of a cluster, each CTA holds N / 2 rows of B; the leader issues for both.
__global__ void __launch_bounds__(128, 1) mma_pair_kernel(int n, int count, long long* cycles_out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_a = smem;
    uint8_t* s_b = smem + 16384;
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = ClusterCtaRank();
    if (threadIdx.x == 0) { MbarInit(&bar, 1); FenceBarrierInit(); }
    if (warp == 1) TmemAlloc2(&tmem_slot, 512);
    FenceProxyAsync();
    TcFenceBefore();
    __syncthreads();
    ClusterSync();
    TcFenceAfter();
    const uint32_t tmem = tmem_slot;
    if (warp == 0) {
        const uint32_t idesc = MakeInstrDescM(0, n, 256);
        const uint64_t a_desc = MakeSmemDesc(SmemAddr(s_a)), b_desc = MakeSmemDesc(SmemAddr(s_b));
        long long t0 = clock64();
        if (rank == 0) {
            if (ElectOne()) {
                for (int i = 0; i < count; ++i) {
                    const int ks = i & 3;
                    UmmaSS2Fp8(tmem, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc, i ? 1u : 0u);
                }
                UmmaCommit2(&bar);
            }
            __syncwarp();
        }
        MbarWaitWarp(&bar, 0);
        long long t1 = clock64();
        if (threadIdx.x == 0) cycles_out[blockIdx.x] = t1 - t0;
    }
    TcFenceBefore();
    __syncthreads();
    ClusterSync();
    if (warp == 1) { TcFenceAfter(); TmemDealloc2(tmem, 512); }
}

int main() {
    int dev = 0, sms = 0, khz = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    long long* d_cyc;
    cudaMalloc(&d_cyc, sms * sizeof(long long));
    const int smem = 1024 + 16384 + 32768;
    cudaFuncSetAttribute(mma_issue_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int count = 2048;
    printf("tcgen05.mma cta_group::1 M=128 K=32B, %d back-to-back dispatches per SM on %d SMs (max clock %.0f MHz)\n", count, sms, khz / 1e3);
    printf("%-10s %-4s %5s %5s %12s %12s %10s\n", "kind", "A", "N", "accs", "cyc/mma", "nominal", "of_rate");
    struct Cfg { int n, mode, kind, accs; };
    const Cfg cfgs[] = {{16, 0, 1, 1}, {32, 0, 1, 1}, {64, 0, 1, 1}, {96, 0, 1, 1}, {128, 0, 1, 1}, {192, 0, 1, 1}, {256, 0, 1, 1},
                        {96, 0, 1, 2}, {128, 0, 1, 2}, {32, 1, 1, 1}, {128, 1, 1, 1}, {256, 1, 1, 1},
                        {32, 0, 0, 1}, {96, 0, 0, 1}, {128, 0, 0, 1}, {256, 0, 0, 1}};
    for (const Cfg& c : cfgs) {
        mma_issue_kernel<<<sms, 128, smem>>>(c.n, c.mode, c.kind, c.accs, count, d_cyc);  // warm-up
        mma_issue_kernel<<<sms, 128, smem>>>(c.n, c.mode, c.kind, c.accs, count, d_cyc);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        long long h[256];
        cudaMemcpy(h, d_cyc, sms * sizeof(long long), cudaMemcpyDeviceToHost);
        double sum = 0;
        for (int i = 0; i < sms; ++i) sum += (double)h[i];
        const double cyc = sum / sms / count, nominal = 128.0 * c.n / 256.0;
        printf("%-10s %-4s %5d %5d %12.1f %12.1f %9.0f%%\n", c.kind ? "f8f6f4" : "f16(bf16)", c.mode ? "tmem" : "smem", c.n, c.accs, cyc, nominal, 100.0 * nominal / cyc);
    }
    cudaFuncSetAttribute(mma_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int n : {32, 96, 128, 256}) {
        for (int rep = 0; rep < 2; ++rep) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3((unsigned)(sms & ~1));
            cfg.blockDim = dim3(128);
            cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            cudaLaunchKernelEx(&cfg, mma_pair_kernel, n, count, d_cyc);
        }
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("pair launch failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
        long long h[256];
        cudaMemcpy(h, d_cyc, (sms & ~1) * sizeof(long long), cudaMemcpyDeviceToHost);
        double sum = 0; int cnt = 0;
        for (int i = 0; i < (sms & ~1); i += 2) { sum += (double)h[i]; ++cnt; }
        const double cyc = sum / cnt / count, nominal = 128.0 * n / 256.0;
        printf("%-10s %-4s %5d %5s %12.1f %12.1f %9.0f%%   (cta_group::2, M = 256: per-SM nominal)\n", "f8f6f4", "pair", n, "1", cyc, nominal, 100.0 * nominal / cyc);
    }
    return 0;
}
