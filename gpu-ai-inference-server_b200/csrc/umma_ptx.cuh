// umma_ptx.cuh — PTX wrappers shared by the tcgen05 kernels (sm_100a only): mbarrier, TMA, TMEM
// allocation, tcgen05.mma/commit/ld, shared-memory matrix descriptors, element traits and the packed
// BatchNorm(+ReLU) prologue math.  Everything is internal-linkage (anonymous namespace) on purpose.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp8.h>

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <utility>

namespace b200 {
namespace kernels {
namespace {

constexpr int kThreads = 448;
constexpr int kTileM = 128;
constexpr int kRowBytes = 128;             // one K chunk of one row: 128 bytes (the swizzle span)
constexpr int kATileBytes = kTileM * kRowBytes;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t SmemAddr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void MbarInit(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(SmemAddr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void MbarArrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(SmemAddr(bar)) : "memory");
}
__device__ __forceinline__ void MbarArriveExpectTx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(SmemAddr(bar)), "r"(bytes) : "memory");
}
#ifndef B200_TRYWAIT_HINT_NS
#define B200_TRYWAIT_HINT_NS 0
#endif
__device__ __forceinline__ bool MbarTryWait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
#if B200_TRYWAIT_HINT_NS > 0
    // suspend-time hint: the hardware may keep the thread suspended up to this long before try_wait returns false
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(SmemAddr(bar)), "r"(parity), "r"((uint32_t)B200_TRYWAIT_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(SmemAddr(bar)), "r"(parity)
        : "memory");
#endif
    return ok != 0;
}
// Non-blocking probe (mbarrier.test_wait never suspends the thread).
__device__ __forceinline__ bool MbarTest(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(SmemAddr(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// A lost arrive must become an error, not a hung GPU: trap after a generous spin budget.
__device__ __noinline__ void MbarTimeout() {
    printf("conv_umma: mbarrier timeout (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
    __trap();
}
// (measured, tools/ubench/mbar_cost.cu: a try_wait on an ALREADY completed phase costs ~118 cycles, a test_wait 34 - so every wait
// probes with test_wait first; in a well-fed pipeline most waits are satisfied when they are issued)
__device__ __forceinline__ void MbarWait(uint64_t* bar, uint32_t parity) {
    if (MbarTest(bar, parity)) return;
    if (MbarTryWait(bar, parity)) return;
    uint32_t spins = 0;
    while (!MbarTryWait(bar, parity)) {
        if (++spins > (1u << 24)) MbarTimeout();
    }
}
#ifndef B200_POLL_SLEEP_NS
#define B200_POLL_SLEEP_NS 0
#endif
#ifndef B200_POLL_BACKOFF_AFTER
#define B200_POLL_BACKOFF_AFTER 0
#endif
// Warp-collective wait (every lane of a CONVERGED warp must call it): lane 0 polls, backing off with nanosleep, the
// other 31 lanes park at the warp barrier.  Measured on B200: try_wait returns after only ~60-80 cycles when the
// phase is still pending, so 18 warps x 32 lanes polling at full speed made 30-55 % of all executed instructions
// SYNCS polls and slowed the working warps down by up to 1.7x.
__device__ __forceinline__ void MbarWaitWarp(uint64_t* bar, uint32_t parity) {
    if ((threadIdx.x & 31) == 0 && !MbarTest(bar, parity)) {
        uint32_t spins = 0;
        while (!MbarTryWait(bar, parity)) {
#if B200_POLL_SLEEP_NS > 0
            if (spins >= B200_POLL_BACKOFF_AFTER) __nanosleep(B200_POLL_SLEEP_NS);
#endif
            if (++spins > (1u << 24)) MbarTimeout();
        }
    }
    __syncwarp();
}
// One lane of a fully converged warp; the compiler keeps the surrounding code on the uniform datapath
// (descriptors in uniform registers) instead of wrapping every tcgen05/TMA instruction in a waterfall loop.
__device__ __forceinline__ bool ElectOne() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "elect.sync _|P1, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// Programmatic dependent launch: `GridDepLaunch` lets the next kernel of the stream start its preamble (barrier
// init, TMEM allocation, weight TMA) while this grid is still running; `GridDepWait` blocks until the previous grid
// has completed and its memory is visible - every thread that touches activations calls it first.
__device__ __forceinline__ void GridDepWait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void GridDepLaunch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// Launch with the programmatic-stream-serialization attribute (the kernel must call GridDepWait()).
template <typename... KArgs, typename... Args>
inline cudaError_t LaunchPdl(void (*kern)(KArgs...), int grid, int block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

__device__ __forceinline__ void FenceBarrierInit() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void FenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void TcFenceBefore() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void TcFenceAfter() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void TmaLoad2D(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(SmemAddr(smem_dst)), "l"((uint64_t)map), "r"(SmemAddr(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void TmaLoad4D(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
        ::"r"(SmemAddr(smem_dst)), "l"((uint64_t)map), "r"(SmemAddr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void TmaLoad5D(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(SmemAddr(smem_dst)), "l"((uint64_t)map), "r"(SmemAddr(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
// TMA store of a shared-memory box (bulk async-group completion).
__device__ __forceinline__ void TmaStore2D(const CUtensorMap* map, const void* smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"((uint64_t)map), "r"(SmemAddr(smem_src)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void TmaStore5D(const CUtensorMap* map, const void* smem_src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"((uint64_t)map), "r"(SmemAddr(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void BulkCommit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void BulkWaitRead() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void BulkWait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void NamedBarSync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// not volatile, no memory clobber: the compiler may batch these (ordering comes from the mbarrier wait before / the
// proxy fence after, which are volatile with a memory clobber)
__device__ __forceinline__ uint4 LdsV4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ float4 LdsF4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void StsV4(uint32_t addr, uint4 v) {
    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// ---- packed epilogue math: two fp32 lanes per instruction, ReLU folded into the narrowing convert
__device__ __forceinline__ float2 Fma2(float2 a, float2 b, float2 c) {
    float2 d;
    asm("{\n\t.reg .b64 ra, rb, rc, rd;\n\t"
        "mov.b64 ra, {%2, %3};\n\tmov.b64 rb, {%4, %5};\n\tmov.b64 rc, {%6, %7};\n\t"
        "fma.rn.f32x2 rd, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rd;\n\t}"
        : "=f"(d.x), "=f"(d.y)
        : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
    return d;
}
// (lo, hi) -> two e4m3 in the low 16 bits, optional ReLU
template <bool RELU> __device__ __forceinline__ uint32_t CvtE4m3x2(float lo, float hi) {
    unsigned short d;
    if (RELU) asm("cvt.rn.satfinite.relu.e4m3x2.f32 %0, %1, %2;" : "=h"(d) : "f"(hi), "f"(lo));
    else asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(d) : "f"(hi), "f"(lo));
    return (uint32_t)d;
}
template <bool RELU> __device__ __forceinline__ uint32_t CvtBf16x2(float lo, float hi) {
    uint32_t d;
    if (RELU) asm("cvt.rn.relu.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    else asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
// y = [relu](acc * scale + bias) for 32 consecutive channels of one row -> packed OutT words (8 for e4m3, 16 for bf16).
// sc/bi: 32 floats each (registers or shared memory).
template <typename OutT, bool RELU>
__device__ __forceinline__ void EpiloguePack32(const uint32_t* acc, const float* sc, const float* bi, uint32_t* out) {
#pragma unroll
    for (int q = 0; q < 32; q += 4) {
        float2 a = Fma2(make_float2(__uint_as_float(acc[q]), __uint_as_float(acc[q + 1])), make_float2(sc[q], sc[q + 1]), make_float2(bi[q], bi[q + 1]));
        float2 b = Fma2(make_float2(__uint_as_float(acc[q + 2]), __uint_as_float(acc[q + 3])), make_float2(sc[q + 2], sc[q + 3]), make_float2(bi[q + 2], bi[q + 3]));
        if (sizeof(OutT) == 1) {
            out[q / 4] = CvtE4m3x2<RELU>(a.x, a.y) | (CvtE4m3x2<RELU>(b.x, b.y) << 16);
        } else {
            out[q / 2] = CvtBf16x2<RELU>(a.x, a.y);
            out[q / 2 + 1] = CvtBf16x2<RELU>(b.x, b.y);
        }
    }
}

// Same with scale/bias read from shared memory by 32-bit shared address (explicit LDS.128 broadcasts; going through a
// generic pointer made the compiler emit generic LD + 12 R2UR per call).
template <typename OutT, bool RELU>
__device__ __forceinline__ void EpiloguePack32Smem(const uint32_t* acc, uint32_t sc_addr, uint32_t bi_addr, uint32_t* out) {
#pragma unroll
    for (int q = 0; q < 32; q += 4) {
        const float4 s4 = LdsF4(sc_addr + q * 4), b4 = LdsF4(bi_addr + q * 4);
        float2 a = Fma2(make_float2(__uint_as_float(acc[q]), __uint_as_float(acc[q + 1])), make_float2(s4.x, s4.y), make_float2(b4.x, b4.y));
        float2 b = Fma2(make_float2(__uint_as_float(acc[q + 2]), __uint_as_float(acc[q + 3])), make_float2(s4.z, s4.w), make_float2(b4.z, b4.w));
        if (sizeof(OutT) == 1) {
            out[q / 4] = CvtE4m3x2<RELU>(a.x, a.y) | (CvtE4m3x2<RELU>(b.x, b.y) << 16);
        } else {
            out[q / 2] = CvtBf16x2<RELU>(a.x, a.y);
            out[q / 2 + 1] = CvtBf16x2<RELU>(b.x, b.y);
        }
    }
}

__device__ __forceinline__ void PrefetchTensorMap(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

__device__ __forceinline__ void TmemAlloc(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(SmemAddr(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void TmemDealloc(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T ; KIND 0: kind::f16 (bf16 operands), 1: kind::f8f6f4 (e4m3 operands)
template <int KIND>
__device__ __forceinline__ void UmmaSS(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if (KIND == 0) {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
            : "memory");
    }
}
// Arrives on `bar` once every previously issued tcgen05.mma of this thread has completed.
__device__ __forceinline__ void UmmaCommit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(SmemAddr(bar)) : "memory");
}

// ---- CTA pairs (cta_group::2): two CTAs of a cluster execute one tcgen05.mma of M = 256 together; each supplies its own 128 rows
// of A and HALF of B (N / 2 rows at the same shared-memory offset in both CTAs), so the B fetch per SM halves.
__device__ __forceinline__ uint32_t ClusterCtaRank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void ClusterSync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `local_smem_addr` (a shared::cta address) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t MapaShared(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
// RELAXED on purpose: a release at cluster scope first waits for every earlier global store of the thread to become visible
// (measured: 1.1 us per tile on an epilogue warp that stores its results with st.global).  What these arrives order is tensor-memory
// reads (tcgen05.wait::ld + tcgen05.fence before) and TMA-written shared memory (complete_tx), neither of which needs it.
__device__ __forceinline__ void MbarArriveCluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void TmemAlloc2(uint32_t* slot, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(SmemAddr(slot)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void TmemDealloc2(uint32_t addr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
// issued by the leader CTA (cluster rank 0) only
__device__ __forceinline__ void UmmaSS2Fp8(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrives on the barrier at this shared-memory offset in BOTH CTAs of the pair once every MMA issued before it has completed
__device__ __forceinline__ void UmmaCommit2(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(SmemAddr(bar)), "h"((unsigned short)3) : "memory");
}
__host__ __device__ constexpr uint32_t MakeInstrDescM(int fmt, int n, int m) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void TmemLoad32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void TmemLoadWait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void TmemLoad16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
// 16 consecutive 32-bit columns of this thread's TMEM lane (registers -> tensor memory)
__device__ __forceinline__ void TmemStore16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void TmemStoreWait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]^T (kind::f8f6f4): the A tile lives in tensor memory, row = lane, K packed 4 e4m3 per
// 32-bit column (a 32-byte K step = 8 columns)
__device__ __forceinline__ void UmmaTS(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// K-major, 128-byte swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start>>4 [0,14) | LBO>>4 [16,30) (unused for swizzled K-major, 1) | SBO>>4 [32,46) = 1024 B between
// 8-row groups | version=1 [46,48) | layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t MakeSmemDesc(uint32_t smem_byte_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_byte_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32=1 [4,6) | a_format [7,10) |
// b_format [10,13) | a/b K-major (0) | N>>3 [17,23) | M>>4 [24,29).
__host__ __device__ constexpr uint32_t MakeInstrDesc(int fmt, int n) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kTileM >> 4) << 24);
}

__device__ __forceinline__ uint4 LdgNc(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 LdgNc8(const void* p) {
    uint2 r;
    asm volatile("ld.global.nc.v2.u32 {%0, %1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

// ------------------------------------------------------------------ element traits
template <typename T> struct MmaElem;
template <> struct MmaElem<__nv_bfloat16> {
    static constexpr int kKind = 0, kFmt = 1;  // kind::f16, BF16
    static constexpr int kPerVec = 8;          // elements per 16 bytes
    static constexpr int kChunk = 64;          // elements per 128-byte K chunk
    static constexpr int kStepK = 16;          // elements per tcgen05.mma (32 bytes)
    __device__ static void Unpack(const uint4& v, float* f) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            f[2 * i] = __uint_as_float(w[i] << 16);
            f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
        }
    }
    __device__ static uint4 Pack(const float* f) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
};
template <> struct MmaElem<__nv_fp8_e4m3> {
    static constexpr int kKind = 1, kFmt = 0;  // kind::f8f6f4, E4M3
    static constexpr int kPerVec = 16;
    static constexpr int kChunk = 128;
    static constexpr int kStepK = 32;
    __device__ static void Unpack(const uint4& v, float* f) {
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                __half2_raw hr = __nv_cvt_fp8x2_to_halfraw2((__nv_fp8x2_storage_t)((w[i] >> (16 * h)) & 0xFFFFu), __NV_E4M3);
                float2 t = __half22float2(*reinterpret_cast<__half2*>(&hr));
                f[4 * i + 2 * h] = t.x;
                f[4 * i + 2 * h + 1] = t.y;
            }
        }
    }
    __device__ static uint4 Pack(const float* f) {
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(f[4 * i], f[4 * i + 1]), __NV_SATFINITE, __NV_E4M3);
            uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(f[4 * i + 2], f[4 * i + 3]), __NV_SATFINITE, __NV_E4M3);
            w[i] = lo | (hi << 16);
        }
        return make_uint4(w[0], w[1], w[2], w[3]);
    }
};

// ------------------------------------------------------------------ packed prologue math
// Folded BatchNorm (+ReLU) on packed pairs: one fma.rn[.relu] per two channels.
template <bool RELU> __device__ __forceinline__ uint32_t FmaBf16x2(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    if (RELU) asm("fma.rn.relu.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else asm("fma.rn.bf16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
template <bool RELU> __device__ __forceinline__ uint32_t FmaF16x2(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    if (RELU) asm("fma.rn.relu.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t E4m3x2ToF16x2(uint32_t v16) {
    uint32_t d;
    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(d) : "h"((unsigned short)v16));
    return d;
}
__device__ __forceinline__ uint32_t F16x2ToE4m3x2(uint32_t v) {
    unsigned short d;
    asm("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(d) : "r"(v));
    return (uint32_t)d;
}
// Packs two fp32 per-channel constants for the packed prologue of each MMA element type.
template <typename T> __device__ __forceinline__ uint32_t PackPair(float a, float b);
template <> __device__ __forceinline__ uint32_t PackPair<__nv_bfloat16>(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
template <> __device__ __forceinline__ uint32_t PackPair<__nv_fp8_e4m3>(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
// y = [relu](x * scale + shift) on one 16-byte piece; sc/sh hold one packed pair per two channels.
template <typename T, bool RELU> __device__ __forceinline__ uint4 ProloguePiece(uint4 v, const uint32_t* sc, const uint32_t* sh);
template <> __device__ __forceinline__ uint4 ProloguePiece<__nv_bfloat16, true>(uint4 v, const uint32_t* sc, const uint32_t* sh) {
    return make_uint4(FmaBf16x2<true>(v.x, sc[0], sh[0]), FmaBf16x2<true>(v.y, sc[1], sh[1]),
                      FmaBf16x2<true>(v.z, sc[2], sh[2]), FmaBf16x2<true>(v.w, sc[3], sh[3]));
}
template <> __device__ __forceinline__ uint4 ProloguePiece<__nv_bfloat16, false>(uint4 v, const uint32_t* sc, const uint32_t* sh) {
    return make_uint4(FmaBf16x2<false>(v.x, sc[0], sh[0]), FmaBf16x2<false>(v.y, sc[1], sh[1]),
                      FmaBf16x2<false>(v.z, sc[2], sh[2]), FmaBf16x2<false>(v.w, sc[3], sh[3]));
}
template <bool RELU> __device__ __forceinline__ uint32_t PrologueWordFp8(uint32_t w, const uint32_t* sc, const uint32_t* sh) {
    uint32_t lo = FmaF16x2<RELU>(E4m3x2ToF16x2(w & 0xFFFFu), sc[0], sh[0]);
    uint32_t hi = FmaF16x2<RELU>(E4m3x2ToF16x2(w >> 16), sc[1], sh[1]);
    return F16x2ToE4m3x2(lo) | (F16x2ToE4m3x2(hi) << 16);
}
template <> __device__ __forceinline__ uint4 ProloguePiece<__nv_fp8_e4m3, true>(uint4 v, const uint32_t* sc, const uint32_t* sh) {
    return make_uint4(PrologueWordFp8<true>(v.x, sc, sh), PrologueWordFp8<true>(v.y, sc + 2, sh + 2),
                      PrologueWordFp8<true>(v.z, sc + 4, sh + 4), PrologueWordFp8<true>(v.w, sc + 6, sh + 6));
}
template <> __device__ __forceinline__ uint4 ProloguePiece<__nv_fp8_e4m3, false>(uint4 v, const uint32_t* sc, const uint32_t* sh) {
    return make_uint4(PrologueWordFp8<false>(v.x, sc, sh), PrologueWordFp8<false>(v.y, sc + 2, sh + 2),
                      PrologueWordFp8<false>(v.z, sc + 4, sh + 4), PrologueWordFp8<false>(v.w, sc + 6, sh + 6));
}

// Transition layers: A row = SUM over the 2x2 input pixels of relu(x*scale+shift) (the 1/4 of the average is folded into
// the epilogue scale).  v[0..3] are the same 16-byte piece of the four pixels.
__device__ __forceinline__ uint32_t AddF16x2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t AddBf16x2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
template <typename T, bool RELU> __device__ __forceinline__ uint4 PoolPiece(const uint4* v, const uint32_t* sc, const uint32_t* sh);
template <bool RELU> __device__ __forceinline__ uint32_t PoolWordFp8(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, const uint32_t* sc, const uint32_t* sh) {
    uint32_t lo = AddF16x2(AddF16x2(FmaF16x2<RELU>(E4m3x2ToF16x2(w0 & 0xFFFFu), sc[0], sh[0]), FmaF16x2<RELU>(E4m3x2ToF16x2(w1 & 0xFFFFu), sc[0], sh[0])),
                           AddF16x2(FmaF16x2<RELU>(E4m3x2ToF16x2(w2 & 0xFFFFu), sc[0], sh[0]), FmaF16x2<RELU>(E4m3x2ToF16x2(w3 & 0xFFFFu), sc[0], sh[0])));
    uint32_t hi = AddF16x2(AddF16x2(FmaF16x2<RELU>(E4m3x2ToF16x2(w0 >> 16), sc[1], sh[1]), FmaF16x2<RELU>(E4m3x2ToF16x2(w1 >> 16), sc[1], sh[1])),
                           AddF16x2(FmaF16x2<RELU>(E4m3x2ToF16x2(w2 >> 16), sc[1], sh[1]), FmaF16x2<RELU>(E4m3x2ToF16x2(w3 >> 16), sc[1], sh[1])));
    return F16x2ToE4m3x2(lo) | (F16x2ToE4m3x2(hi) << 16);
}
template <> __device__ __forceinline__ uint4 PoolPiece<__nv_fp8_e4m3, true>(const uint4* v, const uint32_t* sc, const uint32_t* sh) {
    return make_uint4(PoolWordFp8<true>(v[0].x, v[1].x, v[2].x, v[3].x, sc, sh), PoolWordFp8<true>(v[0].y, v[1].y, v[2].y, v[3].y, sc + 2, sh + 2),
                      PoolWordFp8<true>(v[0].z, v[1].z, v[2].z, v[3].z, sc + 4, sh + 4), PoolWordFp8<true>(v[0].w, v[1].w, v[2].w, v[3].w, sc + 6, sh + 6));
}
template <> __device__ __forceinline__ uint4 PoolPiece<__nv_fp8_e4m3, false>(const uint4* v, const uint32_t* sc, const uint32_t* sh) {
    return make_uint4(PoolWordFp8<false>(v[0].x, v[1].x, v[2].x, v[3].x, sc, sh), PoolWordFp8<false>(v[0].y, v[1].y, v[2].y, v[3].y, sc + 2, sh + 2),
                      PoolWordFp8<false>(v[0].z, v[1].z, v[2].z, v[3].z, sc + 4, sh + 4), PoolWordFp8<false>(v[0].w, v[1].w, v[2].w, v[3].w, sc + 6, sh + 6));
}
template <bool RELU> __device__ __forceinline__ uint32_t PoolWordBf16(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t sc, uint32_t sh) {
    return AddBf16x2(AddBf16x2(FmaBf16x2<RELU>(w0, sc, sh), FmaBf16x2<RELU>(w1, sc, sh)), AddBf16x2(FmaBf16x2<RELU>(w2, sc, sh), FmaBf16x2<RELU>(w3, sc, sh)));
}
template <> __device__ __forceinline__ uint4 PoolPiece<__nv_bfloat16, true>(const uint4* v, const uint32_t* sc, const uint32_t* sh) {
    return make_uint4(PoolWordBf16<true>(v[0].x, v[1].x, v[2].x, v[3].x, sc[0], sh[0]), PoolWordBf16<true>(v[0].y, v[1].y, v[2].y, v[3].y, sc[1], sh[1]),
                      PoolWordBf16<true>(v[0].z, v[1].z, v[2].z, v[3].z, sc[2], sh[2]), PoolWordBf16<true>(v[0].w, v[1].w, v[2].w, v[3].w, sc[3], sh[3]));
}
template <> __device__ __forceinline__ uint4 PoolPiece<__nv_bfloat16, false>(const uint4* v, const uint32_t* sc, const uint32_t* sh) {
    return make_uint4(PoolWordBf16<false>(v[0].x, v[1].x, v[2].x, v[3].x, sc[0], sh[0]), PoolWordBf16<false>(v[0].y, v[1].y, v[2].y, v[3].y, sc[1], sh[1]),
                      PoolWordBf16<false>(v[0].z, v[1].z, v[2].z, v[3].z, sc[2], sh[2]), PoolWordBf16<false>(v[0].w, v[1].w, v[2].w, v[3].w, sc[3], sh[3]));
}

__device__ __forceinline__ void CpAsync16(uint32_t dst, const void* src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void CpAsync8(uint32_t dst, const void* src, bool valid) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(valid ? 8 : 0) : "memory");
}
__device__ __forceinline__ void CpAsyncCommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void CpAsyncWait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }


__device__ __forceinline__ uint64_t MakeSmemDescNoSwizzle(uint32_t smem_byte_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_byte_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    return d;  // layout_type 0 = SWIZZLE_NONE
}


}  // namespace
}  // namespace kernels
}  // namespace b200
