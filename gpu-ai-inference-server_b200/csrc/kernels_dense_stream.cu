// kernels_dense_stream.cu — ONE DenseNet dense layer, BN-ReLU-Conv1x1(->128)-BN-ReLU-Conv3x3(->32), of the LARGE-image
// blocks (56x56 and 28x28) as a single streaming kernel: the 128-channel bottleneck tensor never leaves the SM.
// Replaces, per layer, the six nodes (BatchNormalization, Relu, Conv, BatchNormalization, Relu, Conv) ONNX Runtime
// executes one by one inside `Ort::Session::Run` (reference inference_engine/src/model.cpp:1264-1270), and the engine's
// own conv1x1_tma + conv3x3_tma kernel pair, which round-trips 128 bytes per pixel through HBM between the two convs.
//
// A CTA walks DOWN the rows of its share of the images (a contiguous range of row groups, cut into strips at image
// boundaries) as a sliding window:
//   conv1 tile k   128 patch slots = RPT image rows x HV half rows x 32 slots (56 wide: 2 rows x 2 halves of 28 + halo,
//                  28 wide: 4 rows).  TMA lands the raw block-buffer rows [slot][128 B of channels]; the transform warps
//                  read them ONCE, apply this layer's folded BN1+ReLU in registers and write the result into TENSOR
//                  memory (tcgen05.st): the conv1 A operand never goes back to shared memory, the MMAs read it from
//                  TMEM (kind::f8f6f4, A in tensor memory) against the resident 1x1 weights.
//   epilogue 1     TMEM -> BN2+ReLU -> e4m3 -> 128-byte-swizzled patch tile in a RING of three tiles in shared memory;
//                  slots outside the image are written as zeros (= the 3x3 conv's padding).
//   conv2 tile     needs the patch rows of conv1 tiles k-1 and k: for each filter row one view of 128 consecutive slots
//                  (shifted by fr rows) times the three taps of that row stacked along N (N = 96, see kernels_conv3x3.cu);
//                  views that run past the end of the ring read a copy of ring tile 0's first rows kept behind the ring.
//   epilogue 2     adds the three column groups across neighbouring TMEM lanes (half rows are 32 slots = one warp's lanes,
//                  so the +-1 shuffles never cross a warp), scale/bias, e4m3, 16-byte stores into the layer's channel slice.
// The MMA issuer interleaves conv1 of tile k with conv2 of the tile pair (k-2, k-1): tcgen05.mma completes in issue
// order, which is also what makes the patch ring safe without an "empty" barrier.
// Per-channel constants (BN1 pairs, both epilogues' scale/bias) travel as a __grid_constant__ kernel parameter and are
// read through the constant cache: no shared-memory wavefronts are spent on broadcasts.
//
// Warps (704 threads): 0-7 transform, 8-15 epilogue 1 (two per TMEM lane quarter), 16-19 epilogue 2, 20 TMA producer, 21 MMA issuer.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "kernels.h"
#include "umma_ptx.cuh"

namespace b200 {
namespace kernels {

namespace {

constexpr int kDsXfWarps = 8, kDsEpiWarps = 8, kDsEpi2Warps = 4;
constexpr int kDsTmaWarp = kDsXfWarps + kDsEpiWarps + kDsEpi2Warps, kDsMmaWarp = kDsTmaWarp + 1;
constexpr int kDsThreads = 32 * (kDsMmaWarp + 1);
constexpr int kDsCH = 128;                     // e4m3 elements per 128-byte K chunk
constexpr int kDsW2Bytes = 9 * 32 * 128;       // 36 KB
constexpr int kDsRing = 3;                     // patch tiles in the ring
constexpr int kDsABufs = 4;                    // conv1 A tiles in tensor memory
constexpr uint32_t kDsAcc1Col = 0;             // 2 x 128 columns
constexpr uint32_t kDsAcc2Col = 256;           // 96 columns
constexpr uint32_t kDsACol = 352;              // 4 x 32 columns
constexpr int kDsTmemCols = 512;
constexpr int kDsHalfW = 28;                   // pixels per half row

// Per-channel constants: travel as a kernel parameter, copied to shared memory once per CTA and read by 128-bit broadcasts
// (scalar LDC loads straight from the parameter space cost a dependent 32-bit load per constant: measured 1.7 us per tile)
struct alignas(16) DsConsts {
    uint32_t pre_sc[256];  // folded BN1 scale/shift as packed f16x2 pairs
    uint32_t pre_sh[256];
    float s1[128], b1[128];
    float s2[32], b2[32];
};

constexpr int kDsMaxStages = 6;
constexpr int kDsOutStageBytes = 112 * 32;     // one output tile: RPT x HV x 28 pixels x 32 channels
constexpr int kDsSmemLimit = 227 * 1024;

template <int HV> struct DsCfg {               // HV = half rows per image row (2: 56 wide, 1: 28 wide)
    static constexpr int kRPT = 4 / HV;        // image rows per 128-slot tile
    static constexpr int kRS = 32 * HV;        // patch slots per image row
    static constexpr int kMaxChunks = HV == 2 ? 2 : 4;
    static constexpr int kExtBytes = 2 * kRS * 128;
    static constexpr int kPatchBytes = kDsRing * kATileBytes + kExtBytes;
    // everything but the landing stages and the resident 1x1 weights, whose split depends on the layer's K chunks
    static constexpr int kFixedBytes = 1024 + kDsW2Bytes + kPatchBytes + (int)sizeof(DsConsts) + 2 * kDsOutStageBytes + 512;
    static constexpr int Stages(int nc) {
        int s = (kDsSmemLimit - kFixedBytes - nc * 128 * kRowBytes) / kATileBytes;
        return s > kDsMaxStages ? kDsMaxStages : s;
    }
    static constexpr int SmemBytes(int nc) { return kFixedBytes + (Stages(nc) + nc) * kATileBytes; }
    static_assert(Stages(kMaxChunks) >= 3, "landing ring too shallow");
};

struct DsParams {
    void* buf;            // block buffer, NHWC e4m3
    int pitch, n, H, W;
    int Cin, c_off_out;
    int pre_relu, relu1, relu2;
    int gpi;              // row groups per image (H / RPT)
    int total_groups;     // n * gpi
    int stages;           // landing stages (what the shared memory left by the resident weights allows)
    unsigned long long* trace;  // debug: [item][16] globaltimer stamps of CTA 0 (B200_DS_TRACE), else null
};

struct DsGeom {
    int ch_base, k_lo, k_hi;
};
__device__ __forceinline__ DsGeom DsGeomOf(int c, int Cin) {
    DsGeom g;
    g.ch_base = c * kDsCH; g.k_lo = 0; g.k_hi = kDsCH;
    if (g.ch_base + kDsCH > Cin) {
        if (Cin >= kDsCH) { g.k_lo = g.ch_base + kDsCH - Cin; g.ch_base = Cin - kDsCH; }
        else g.k_hi = Cin;
    }
    return g;
}

__device__ __forceinline__ void DsStamp(const DsParams& p, uint32_t k, int ev) {
    if (p.trace && blockIdx.x == 1 && k < 48 && (threadIdx.x & 31) == 0) {
        unsigned long long tm;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tm));
        atomicMax(&p.trace[k * 16 + ev], tm);
    }
}

// The order of work of one CTA, identical for every role: conv1 tile k of every strip, each followed by the conv2 tile of the
// PREVIOUS item (so that epilogue 1 of tile k-1 has a whole tile time before the tensor pipe needs its patch).
//   on_conv1(k, img, y_start)       conv1 tile k covers image rows y_start .. y_start + RPT - 1 (may lie outside the image)
//   on_conv2(k, img, out_group)     conv2 over the patch tiles (k-1, k) -> output rows out_group*RPT .. +RPT-1
template <int RPT, typename F1, typename F2>
__device__ __forceinline__ void DsWalk(int g0, int g1, int gpi, F1&& on_conv1, F2&& on_conv2) {
    uint32_t k = 0, pend_k = 0;
    bool pend = false;
    int pend_img = 0, pend_og = 0;
    for (int g = g0; g < g1;) {
        const int img = g / gpi, rg = g - img * gpi;
        const int ng = (g1 - g) < (gpi - rg) ? (g1 - g) : (gpi - rg);
        for (int t = 0; t <= ng; ++t, ++k) {
            on_conv1(k, img, (rg + t) * RPT - 1);
            if (pend) on_conv2(pend_k, pend_img, pend_og);
            pend = t >= 1; pend_k = k; pend_img = img; pend_og = rg + t - 1;
        }
        g += ng;
    }
    if (pend) on_conv2(pend_k, pend_img, pend_og);
}

// TSA: the conv1 A operand goes through tensor memory (row-owning transform, tcgen05.st; best with ONE K chunk);
// !TSA: the transform rewrites the landed tile in place (a warp owns one 16-byte piece of all 128 rows, as in kernels_conv1x1.cu) and
// the conv1 MMAs read A from shared memory - a chunk then costs a fifth of the row-owning transform's latency chain.
template <int HV, bool TSA>
__global__ void __launch_bounds__(kDsThreads, 1)
dense_stream_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w1,
                    const __grid_constant__ CUtensorMap tmap_w2, const __grid_constant__ CUtensorMap tmap_out,
                    const __grid_constant__ DsConsts cst, const DsParams p) {
    using MmaT = __nv_fp8_e4m3;
    using ME = MmaElem<MmaT>;
    using Cfg = DsCfg<HV>;
    constexpr int RPT = Cfg::kRPT;
    constexpr int RS = Cfg::kRS;
    constexpr int EPV = ME::kPerVec;  // 16

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int Cin = p.Cin;
    const int nc = (Cin + kDsCH - 1) / kDsCH;
    const int NS = p.stages;
    uint8_t* s_raw = smem;
    uint8_t* s_w1 = s_raw + NS * kATileBytes;
    uint8_t* s_w2 = s_w1 + nc * 128 * kRowBytes;
    uint8_t* s_patch = s_w2 + kDsW2Bytes;
    uint32_t* s_cst = reinterpret_cast<uint32_t*>(s_patch + Cfg::kPatchBytes);
    uint8_t* s_out = reinterpret_cast<uint8_t*>(s_cst) + sizeof(DsConsts);   // [2][112 pixels][32 B]
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(s_out + 2 * kDsOutStageBytes);
    uint64_t* raw_empty = raw_full + kDsMaxStages;
    uint64_t* a_full = raw_empty + kDsMaxStages;
    uint64_t* a_empty = a_full + kDsABufs;
    uint64_t* acc1_full = a_empty + kDsABufs;   // [2]
    uint64_t* acc1_empty = acc1_full + 2;       // [2]
    uint64_t* patch_full = acc1_empty + 2;      // [2]
    uint64_t* acc2_full = patch_full + 2;
    uint64_t* acc2_empty = acc2_full + 1;
    uint64_t* w_bar = acc2_empty + 1;
    uint64_t* xf_full = w_bar + 1;                // [kDsMaxStages], !TSA only
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(xf_full + kDsMaxStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // this CTA's share of the row groups
    const int g0 = (int)(((long long)blockIdx.x * p.total_groups) / gridDim.x);
    const int g1 = (int)(((long long)(blockIdx.x + 1) * p.total_groups) / gridDim.x);

    if (warp == kDsTmaWarp && lane == 0) {
        for (int s = 0; s < NS; ++s) {
            MbarInit(&raw_full[s], 1);
            MbarInit(&raw_empty[s], TSA ? kDsXfWarps : 1);   // TSA: freed by the transform warps; else by the MMA commit
            MbarInit(&xf_full[s], kDsXfWarps);
        }
        for (int s = 0; s < kDsABufs; ++s) {
            MbarInit(&a_full[s], kDsXfWarps);
            MbarInit(&a_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            MbarInit(&acc1_full[s], 1);
            MbarInit(&acc1_empty[s], kDsEpiWarps);
            MbarInit(&patch_full[s], kDsEpiWarps);
        }
        MbarInit(acc2_full, 1);
        MbarInit(acc2_empty, kDsEpi2Warps);
        MbarInit(w_bar, 1);
        FenceBarrierInit();
        PrefetchTensorMap(&tmap_x);
        PrefetchTensorMap(&tmap_w1);
        PrefetchTensorMap(&tmap_w2);
        PrefetchTensorMap(&tmap_out);
    }
    if (warp == kDsMmaWarp) TmemAlloc(tmem_slot, kDsTmemCols);
    {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(&cst);
        for (int i = threadIdx.x; i < (int)(sizeof(DsConsts) / 4); i += kDsThreads) s_cst[i] = src[i];
    }
    const uint32_t cst_addr = SmemAddr(s_cst);
    TcFenceBefore();
    __syncthreads();
    TcFenceAfter();
    const uint32_t tmem_base = *tmem_slot;
    GridDepLaunch();

    if (warp == kDsTmaWarp) {
        // =========================================================== TMA producer
        if (ElectOne()) {  // the weights do not depend on the previous kernel: load them before the dependency wait
            MbarArriveExpectTx(w_bar, (uint32_t)(nc * 128 * kRowBytes + kDsW2Bytes));
            for (int c = 0; c < nc; ++c) TmaLoad2D(s_w1 + c * 128 * kRowBytes, &tmap_w1, w_bar, DsGeomOf(c, Cin).ch_base, 0);
            for (int t = 0; t < 9; ++t) TmaLoad2D(s_w2 + t * 32 * 128, &tmap_w2, w_bar, t * kDsCH, 0);
        }
        __syncwarp();
        GridDepWait();
        int stage = 0;
        uint32_t phase = 0;
        DsWalk<RPT>(g0, g1, p.gpi,
            [&](uint32_t k, int img, int y_start) {
                for (int c = 0; c < nc; ++c) {
                    MbarWaitWarp(&raw_empty[stage], phase ^ 1u);
                    if (c == 0) DsStamp(p, k, 0);
                    if (ElectOne()) {
                        MbarArriveExpectTx(&raw_full[stage], (uint32_t)kATileBytes);
                        // box {128 B of channels, 32 slots, HV half rows, RPT rows, 1 image}; slot 0 of a half row is the pixel left of it
                        TmaLoad5D(s_raw + stage * kATileBytes, &tmap_x, &raw_full[stage], DsGeomOf(c, Cin).ch_base, 0, 0, y_start, img);
                    }
                    __syncwarp();
                    if (++stage == NS) { stage = 0; phase ^= 1u; }
                }
            },
            [&](uint32_t, int, int) {});
    } else if (warp == kDsMmaWarp) {
        // =========================================================== MMA issuer
        constexpr uint32_t idesc1 = MakeInstrDesc(ME::kFmt, 128);
        constexpr uint32_t idesc2 = MakeInstrDesc(ME::kFmt, 96);
        const uint64_t w1_desc = MakeSmemDesc(SmemAddr(s_w1));
        const uint64_t w2_desc = MakeSmemDesc(SmemAddr(s_w2));
        const uint32_t patch_addr = SmemAddr(s_patch);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        MbarWaitWarp(w_bar, 0);
        int ab = 0, mstage = 0;
        uint32_t aphase = 0, mphase = 0, d = 0;
        const uint64_t raw_desc = MakeSmemDesc(SmemAddr(s_raw));
        DsWalk<RPT>(g0, g1, p.gpi,
            [&](uint32_t k, int, int) {
                // conv1 of tile k: its accumulator must have been drained by epilogue 1 of tile k-2
                DsStamp(p, k, 4);
                MbarWaitWarp(&acc1_empty[k & 1u], ((k >> 1) & 1u) ^ 1u);
                TcFenceAfter();
                const uint32_t d1 = tmem_u + kDsAcc1Col + (k & 1u) * 128;
                for (int c = 0; c < nc; ++c) {
                    const DsGeom g = DsGeomOf(c, Cin);
                    const int ks_lo = g.k_lo / ME::kStepK, ks_hi = g.k_hi / ME::kStepK;
                    const uint64_t b_desc = w1_desc + (uint64_t)((uint32_t)c * ((128 * kRowBytes) >> 4));
                    if (TSA) {
                        MbarWaitWarp(&a_full[ab], aphase);
                        TcFenceAfter();
                        if (ElectOne()) {
#pragma unroll
                            for (int ks = 0; ks < kDsCH / ME::kStepK; ++ks)
                                if (ks >= ks_lo && ks < ks_hi)
                                    UmmaTS(d1, tmem_u + kDsACol + ab * 32 + ks * 8, b_desc + (uint64_t)(2 * ks), idesc1, (c > 0 || ks > ks_lo) ? 1u : 0u);
                            UmmaCommit(&a_empty[ab]);
                            if (c == nc - 1) UmmaCommit(&acc1_full[k & 1u]);
                        }
                    } else {
                        MbarWaitWarp(&xf_full[mstage], mphase);
                        TcFenceAfter();
                        if (ElectOne()) {
                            const uint64_t a_desc = raw_desc + (uint64_t)((uint32_t)mstage * (kATileBytes >> 4));
#pragma unroll
                            for (int ks = 0; ks < kDsCH / ME::kStepK; ++ks)
                                if (ks >= ks_lo && ks < ks_hi)
                                    UmmaSS<ME::kKind>(d1, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc1, (c > 0 || ks > ks_lo) ? 1u : 0u);
                            UmmaCommit(&raw_empty[mstage]);
                            if (c == nc - 1) UmmaCommit(&acc1_full[k & 1u]);
                        }
                        if (++mstage == NS) { mstage = 0; mphase ^= 1u; }
                    }
                    __syncwarp();
                    if (c == nc - 1) DsStamp(p, k, 5);
                    if (++ab == kDsABufs) { ab = 0; aphase ^= 1u; }
                }
            },
            [&](uint32_t k, int, int) {
                // conv2 over the patch tiles (k-1, k): both written once epilogue 1 of tile k has arrived
                MbarWaitWarp(&patch_full[k & 1u], (k >> 1) & 1u);
                MbarWaitWarp(acc2_empty, (d & 1u) ^ 1u);
                TcFenceAfter();
                if (ElectOne()) {
                    const uint32_t first = (k + kDsRing - 1) % kDsRing;  // ring tile of k-1
#pragma unroll
                    for (int fr = 0; fr < 3; ++fr) {
                        // D[m][fs*32 + o] += A[m + fr*RS] . W(fr, fs)[o]
                        const uint64_t a_desc = MakeSmemDesc(patch_addr + (first * 128 + fr * RS) * 128);
                        const uint64_t b_desc = w2_desc + (uint64_t)(fr * 3 * (32 * 128 / 16));
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            UmmaSS<ME::kKind>(tmem_u + kDsAcc2Col, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc2, (fr | ks) ? 1u : 0u);
                    }
                    UmmaCommit(acc2_full);
                }
                __syncwarp();
                DsStamp(p, k, 6);
                ++d;
            });
    } else if (warp < kDsXfWarps && !TSA) {
        // =========================================================== transform warps, in place: warp tw owns 16-byte piece tw of all 128 rows
        const int tw = warp;
        uint32_t off[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = i * 32 + lane;
            off[i] = (uint32_t)(row * kRowBytes + ((tw ^ (row & 7)) << 4));
        }
        const uint32_t raw_base = SmemAddr(s_raw);
        const bool relu = p.pre_relu != 0;
        int stage = 0;
        uint32_t phase = 0;
        DsWalk<RPT>(g0, g1, p.gpi,
            [&](uint32_t k, int, int) {
                for (int c = 0; c < nc; ++c) {
                    const DsGeom g = DsGeomOf(c, Cin);
                    const int p_lo = g.k_lo / EPV, p_hi = g.k_hi / EPV;
                    const uint32_t a_base = raw_base + stage * kATileBytes;
                    const bool mine = tw >= p_lo && tw < p_hi;
                    uint32_t sc[8], sh[8];
                    if (mine) {
                        const uint32_t ca = cst_addr + (uint32_t)((g.ch_base + tw * EPV) * 2);
                        const uint4 s0 = LdsV4(ca), s1 = LdsV4(ca + 16), h0 = LdsV4(ca + 1024), h1 = LdsV4(ca + 1040);
                        sc[0] = s0.x; sc[1] = s0.y; sc[2] = s0.z; sc[3] = s0.w; sc[4] = s1.x; sc[5] = s1.y; sc[6] = s1.z; sc[7] = s1.w;
                        sh[0] = h0.x; sh[1] = h0.y; sh[2] = h0.z; sh[3] = h0.w; sh[4] = h1.x; sh[5] = h1.y; sh[6] = h1.z; sh[7] = h1.w;
                    }
                    MbarWaitWarp(&raw_full[stage], phase);
                    if (c == 0 && warp == 0) DsStamp(p, k, 1);
                    if (mine) {
                        uint4 v[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) v[i] = LdsV4(a_base + off[i]);
#pragma unroll
                        for (int i = 0; i < 4; ++i) v[i] = relu ? ProloguePiece<MmaT, true>(v[i], sc, sh) : ProloguePiece<MmaT, false>(v[i], sc, sh);
#pragma unroll
                        for (int i = 0; i < 4; ++i) StsV4(a_base + off[i], v[i]);
                    }
                    FenceProxyAsync();
                    __syncwarp();
                    if (lane == 0) MbarArrive(&xf_full[stage]);
                    if (c == nc - 1 && warp == 0) DsStamp(p, k, 3);
                    if (++stage == NS) { stage = 0; phase ^= 1u; }
                }
            },
            [&](uint32_t, int, int) {});
    } else if (warp < kDsXfWarps) {
        // =========================================================== transform warps: raw tile (smem) -> BN1 + ReLU -> A tile (tmem)
        // warp w owns rows 32*(w&3)..+31 (its TMEM lane quarter) and the 16-byte pieces 4*(w>>2)..+3 of them
        const int q = warp & 3, hh = warp >> 2;
        const int row = q * 32 + lane;
        uint32_t off[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) off[i] = (uint32_t)(row * kRowBytes + (((4 * hh + i) ^ (row & 7)) << 4));
        const uint32_t raw_base = SmemAddr(s_raw);
        const bool relu = p.pre_relu != 0;
        int stage = 0, ab = 0;
        uint32_t phase = 0, aphase = 0;
        DsWalk<RPT>(g0, g1, p.gpi,
            [&](uint32_t k, int, int) {
                for (int c = 0; c < nc; ++c) {
                    const DsGeom g = DsGeomOf(c, Cin);
                    const int p_lo = g.k_lo / EPV, p_hi = g.k_hi / EPV;
                    const uint32_t a_base = raw_base + stage * kATileBytes;
                    const bool any = 4 * hh + 3 >= p_lo && 4 * hh < p_hi;  // this warp's pieces carry multiplied channels
                    MbarWaitWarp(&raw_full[stage], phase);
                    if (c == 0 && warp == 0) DsStamp(p, k, 1);
                    uint4 v[4];
                    if (any) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) v[i] = LdsV4(a_base + off[i]);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const int pc = 4 * hh + i;
                            if (pc >= p_lo && pc < p_hi) {
                                const uint32_t ca = cst_addr + (uint32_t)((g.ch_base + pc * EPV) * 2);  // f16x2 pairs: 2 bytes per channel
                                const uint4 s0 = LdsV4(ca), s1 = LdsV4(ca + 16), h0 = LdsV4(ca + 1024), h1 = LdsV4(ca + 1040);
                                const uint32_t sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                                const uint32_t sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                                v[i] = relu ? ProloguePiece<MmaT, true>(v[i], sc, sh) : ProloguePiece<MmaT, false>(v[i], sc, sh);
                            }
                        }
                    }
                    // the landing buffer can take the next tile as soon as every warp has consumed its share
                    __syncwarp();
                    if (lane == 0) MbarArrive(&raw_empty[stage]);
                    if (c == nc - 1 && warp == 0) DsStamp(p, k, 2);
                    MbarWaitWarp(&a_empty[ab], aphase ^ 1u);
                    if (any) {
                        TcFenceAfter();
                        const uint32_t r[16] = {v[0].x, v[0].y, v[0].z, v[0].w, v[1].x, v[1].y, v[1].z, v[1].w,
                                                v[2].x, v[2].y, v[2].z, v[2].w, v[3].x, v[3].y, v[3].z, v[3].w};
                        TmemStore16(tmem_base + ((uint32_t)(q * 32) << 16) + kDsACol + ab * 32 + hh * 16, r);
                        TmemStoreWait();
                        TcFenceBefore();
                    }
                    __syncwarp();
                    if (lane == 0) MbarArrive(&a_full[ab]);
                    if (c == nc - 1 && warp == 0) DsStamp(p, k, 3);
                    if (++stage == NS) { stage = 0; phase ^= 1u; }
                    if (++ab == kDsABufs) { ab = 0; aphase ^= 1u; }
                }
            },
            [&](uint32_t, int, int) {});
    } else if (warp < kDsXfWarps + kDsEpiWarps) {
        // =========================================================== epilogue 1 warps: TMEM lane quarter q, column half h
        // conv1 accumulator -> BN2 + ReLU -> e4m3 -> patch tile k (zeros outside the image)
        const int q = warp & 3, h = (warp - kDsXfWarps) >> 2;
        const int slot = q * 32 + lane;                    // patch slot inside a tile = accumulator row
        const int r_in_tile = q / HV, half = q % HV;
        const int x = kDsHalfW * half - 1 + lane;          // image column this slot holds
        const bool x_ok = x >= 0 && x < p.W;
        const uint32_t patch_base = SmemAddr(s_patch);
        const uint32_t sw = (uint32_t)(slot & 7);
        const bool relu1 = p.relu1 != 0;
        DsWalk<RPT>(g0, g1, p.gpi,
            [&](uint32_t k, int, int y_start) {
                const int y = y_start + r_in_tile;
                const bool ok = x_ok && y >= 0 && y < p.H;
                const uint32_t ring = k % kDsRing;
                const uint32_t addr = patch_base + (ring * 128 + (uint32_t)slot) * 128;
                const bool dup = ring == 0 && slot < 2 * RS;   // ring tile 0's first rows are mirrored behind the ring
                if (warp == 8) DsStamp(p, k, 7);
                MbarWaitWarp(&acc1_full[k & 1u], (k >> 1) & 1u);
                TcFenceAfter();
                if (warp == 8) DsStamp(p, k, 8);
                const uint32_t t1 = tmem_base + ((uint32_t)(q * 32) << 16) + kDsAcc1Col + (k & 1u) * 128 + (uint32_t)h * 64;
#pragma unroll
                for (int ci = 0; ci < 2; ++ci) {
                    const int cg = 2 * h + ci;
                    uint32_t r[32];
                    TmemLoad32(t1 + ci * 32, r);
                    TmemLoadWait();
                    if (ci == 1) {  // the accumulator is in registers
                        TcFenceBefore();
                        __syncwarp();
                        if (lane == 0) MbarArrive(&acc1_empty[k & 1u]);
                    }
                    uint32_t w[8];
                    const uint32_t scp = cst_addr + 2048 + (uint32_t)cg * 128, bip = scp + 512;
                    if (relu1) EpiloguePack32Smem<MmaT, true>(r, scp, bip, w);
                    else EpiloguePack32Smem<MmaT, false>(r, scp, bip, w);
                    if (!ok) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) w[i] = 0u;
                    }
                    const uint32_t o0 = (((uint32_t)(2 * cg)) ^ sw) << 4, o1 = (((uint32_t)(2 * cg + 1)) ^ sw) << 4;
                    StsV4(addr + o0, make_uint4(w[0], w[1], w[2], w[3]));
                    StsV4(addr + o1, make_uint4(w[4], w[5], w[6], w[7]));
                    if (dup) {
                        StsV4(addr + kDsRing * kATileBytes + o0, make_uint4(w[0], w[1], w[2], w[3]));
                        StsV4(addr + kDsRing * kATileBytes + o1, make_uint4(w[4], w[5], w[6], w[7]));
                    }
                }
                FenceProxyAsync();  // patch writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) MbarArrive(&patch_full[k & 1u]);
                if (warp == 8) DsStamp(p, k, 9);
            },
            [&](uint32_t, int, int) {});
    } else {
        // =========================================================== epilogue 2 warps: one per TMEM lane quarter
        // conv2 accumulator -> add the three tap columns -> scale -> e4m3 -> the layer's 32-channel slice (through a TMA store)
        const int q = warp & 3;
        const int half = q % HV;
        const int x = kDsHalfW * half - 1 + lane;
        const bool out_lane = lane >= 1 && lane <= kDsHalfW && x < p.W;  // this lane owns an output pixel of its half row
        const bool relu2 = p.relu2 != 0;
        const bool leader = warp == kDsXfWarps + kDsEpiWarps && lane == 0;
        GridDepWait();  // the stores may alias buffers the previous kernel still reads
        uint32_t d = 0;
        DsWalk<RPT>(g0, g1, p.gpi,
            [&](uint32_t, int, int) {},
            [&](uint32_t kk, int img, int og) {
                MbarWaitWarp(acc2_full, d & 1u);
                TcFenceAfter();
                if (q == 0) DsStamp(p, kk, 10);
                uint32_t w[8];
#pragma unroll
                for (int hc = 0; hc < 2; ++hc) {  // 16 output channels at a time
                    uint32_t r0[16], r1[16], r2[16];
                    const uint32_t t2 = tmem_base + ((uint32_t)(q * 32) << 16) + kDsAcc2Col + (uint32_t)hc * 16;
                    TmemLoad16(t2, r0);
                    TmemLoad16(t2 + 32, r1);
                    TmemLoad16(t2 + 64, r2);
                    TmemLoadWait();
                    if (hc == 1) {  // the accumulator is in registers
                        TcFenceBefore();
                        __syncwarp();
                        if (lane == 0) MbarArrive(acc2_empty);
                    }
                    // out[m] = D[m - 1][fs = 0] + D[m][fs = 1] + D[m + 1][fs = 2]
                    float v[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const float a0 = __shfl_up_sync(0xffffffffu, __uint_as_float(r0[c]), 1);
                        const float a2 = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[c]), 1);
                        v[c] = (a0 + __uint_as_float(r1[c])) + a2;
                    }
                    const uint32_t s2p = cst_addr + 3072 + (uint32_t)hc * 64, b2p = s2p + 128;
#pragma unroll
                    for (int c = 0; c < 16; c += 4) {
                        const float4 s4 = LdsF4(s2p + c * 4), b4 = LdsF4(b2p + c * 4);
                        const float2 a = Fma2(make_float2(v[c], v[c + 1]), make_float2(s4.x, s4.y), make_float2(b4.x, b4.y));
                        const float2 b = Fma2(make_float2(v[c + 2], v[c + 3]), make_float2(s4.z, s4.w), make_float2(b4.z, b4.w));
                        w[hc * 4 + c / 4] = relu2 ? (CvtE4m3x2<true>(a.x, a.y) | (CvtE4m3x2<true>(b.x, b.y) << 16))
                                                  : (CvtE4m3x2<false>(a.x, a.y) | (CvtE4m3x2<false>(b.x, b.y) << 16));
                    }
                }
                // the tile leaves through a TMA store of a small staging buffer: thread-level global stores make the next proxy
                // fence of the storing warp wait for them to drain (measured: 1.3 us per tile when the patch-writing warps stored)
                uint8_t* stg = s_out + (d & 1u) * kDsOutStageBytes;
                ++d;
                if (leader) BulkWaitRead<1>();  // the store that last read this staging buffer is done
                NamedBarSync(1, kDsEpi2Warps * 32);
                if (out_lane) {
                    const uint32_t sa = SmemAddr(stg) + (uint32_t)((q * kDsHalfW + lane - 1) * 32);
                    StsV4(sa, make_uint4(w[0], w[1], w[2], w[3]));
                    StsV4(sa + 16, make_uint4(w[4], w[5], w[6], w[7]));
                }
                FenceProxyAsync();
                NamedBarSync(2, kDsEpi2Warps * 32);
                if (leader) {
                    TmaStore5D(&tmap_out, stg, 0, 0, 0, og * RPT, img);
                    BulkCommit();
                }
                if (q == 0) DsStamp(p, kk, 11);
            });
        if (leader) BulkWait<0>();
    }

    TcFenceBefore();
    __syncthreads();
    if (warp == kDsMmaWarp) {
        TcFenceAfter();
        TmemDealloc(tmem_base, kDsTmemCols);
    }
}

template <int HV, bool TSA>
cudaError_t LaunchDs(const CUtensorMap& tx, const CUtensorMap& tw1, const CUtensorMap& tw2, const CUtensorMap& tout, const DsConsts& cst,
                     DsParams p, cudaStream_t stream) {
    using Cfg = DsCfg<HV>;
    auto kern = dense_stream_kernel<HV, TSA>;
    static int sm_count[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!sm_count[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kDsSmemLimit);
        if (e != cudaSuccess) return e;
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sm_count[dev] = n > 0 ? n : 148;
    }
    const int grid = p.total_groups < sm_count[dev] ? p.total_groups : sm_count[dev];
    const int nc = (p.Cin + kDsCH - 1) / kDsCH;
    p.stages = Cfg::Stages(nc);
    cudaError_t le = LaunchPdl(kern, grid, kDsThreads, Cfg::SmemBytes(nc), stream, tx, tw1, tw2, tout, cst, p);
    CountLaunch();
    return le;
}

}  // namespace

bool DenseLayerStreamSupported(int H, int W, int Cin, int pitch) {
    if (W != 28 && W != 56) return false;
    const int hv = W / kDsHalfW, rpt = 4 / hv;
    if (H < rpt || H % rpt != 0) return false;
    if (Cin < 32 || Cin % 32 != 0 || Cin > (hv == 2 ? 2 : 4) * kDsCH) return false;
    if (pitch % 16 != 0 || Cin + 32 > pitch) return false;
    if (Cin < kDsCH && pitch < kDsCH) return false;  // a short single chunk still reads a whole 128-byte box
    return true;
}

cudaError_t DenseLayerStreamFp8(const DenseLayerStreamArgs& a, cudaStream_t stream) {
    if (!DenseLayerStreamSupported(a.H, a.W, a.Cin, a.pitch) || !a.w1_map || !a.w2_map || !a.buf) return cudaErrorInvalidValue;
    if (!a.pre_scale || !a.pre_shift || !a.s1 || !a.s2) return cudaErrorInvalidValue;
    if (a.n <= 0) return cudaSuccess;
    const int hv = a.W / kDsHalfW, rpt = 4 / hv;
    DsParams p;
    p.buf = a.buf; p.pitch = a.pitch; p.n = a.n; p.H = a.H; p.W = a.W;
    p.Cin = a.Cin; p.c_off_out = a.c_off_out;
    p.pre_relu = a.pre_relu; p.relu1 = a.relu1; p.relu2 = a.relu2;
    p.gpi = a.H / rpt;
    p.total_groups = a.n * p.gpi;
    p.trace = nullptr;
    static unsigned long long* trace_buf = nullptr;
    if (getenv("B200_DS_TRACE")) {
        if (!trace_buf) cudaMalloc(&trace_buf, 48 * 16 * 8);
        cudaMemsetAsync(trace_buf, 0, 48 * 16 * 8, stream);
        p.trace = trace_buf;
    }
    DsConsts cst;
    memset(&cst, 0, sizeof(cst));
    for (int c = 0; c < a.Cin; ++c) {
        const uint32_t hs = __half_as_ushort(__float2half_rn(a.pre_scale[c])), ht = __half_as_ushort(__float2half_rn(a.pre_shift[c]));
        cst.pre_sc[c / 2] |= hs << (16 * (c & 1));
        cst.pre_sh[c / 2] |= ht << (16 * (c & 1));
    }
    for (int c = 0; c < 128; ++c) { cst.s1[c] = a.s1[c]; cst.b1[c] = a.b1 ? a.b1[c] : 0.f; }
    for (int c = 0; c < 32; ++c) { cst.s2[c] = a.s2[c]; cst.b2[c] = a.b2 ? a.b2[c] : 0.f; }
    // input as [image][row][half row][slot = pixel - 28*half + 1][channel]: the half rows overlap by two pixels (3x3 halo) and the
    // base is shifted one pixel to the left, so slot 0 of the left half reads the pixel in front of the row (its conv1 result is
    // replaced by zeros in epilogue 1; the engine's arena has a guard in front of the first buffer)
    TensorMap tx;
    const uint64_t px = (uint64_t)a.pitch;
    const uint64_t dims[5] = {(uint64_t)a.pitch, 32, (uint64_t)hv, (uint64_t)a.H, (uint64_t)a.n};
    const uint64_t strides[4] = {px, (uint64_t)kDsHalfW * px, (uint64_t)a.W * px, (uint64_t)a.H * a.W * px};
    const uint32_t box[5] = {128u, 32u, (uint32_t)hv, (uint32_t)rpt, 1u};
    if (MakeTensorMap(&tx, reinterpret_cast<const uint8_t*>(a.buf) - a.pitch, 1, 5, dims, strides, box, true) != 0) return cudaErrorInvalidValue;
    // output as [image][row][half row][28 pixels][the layer's 32 channels]: one box = one tile's pixels, nothing else
    TensorMap to;
    {
        const uint64_t odims[5] = {32, (uint64_t)kDsHalfW, (uint64_t)hv, (uint64_t)a.H, (uint64_t)a.n};
        const uint32_t obox[5] = {32u, (uint32_t)kDsHalfW, (uint32_t)hv, (uint32_t)rpt, 1u};
        if (MakeTensorMap(&to, reinterpret_cast<uint8_t*>(a.buf) + a.c_off_out, 1, 5, odims, strides, obox, false) != 0) return cudaErrorInvalidValue;
    }
    const CUtensorMap& tout = *reinterpret_cast<const CUtensorMap*>(&to);
    const CUtensorMap& t = *reinterpret_cast<const CUtensorMap*>(&tx);
    const CUtensorMap& tw1 = *reinterpret_cast<const CUtensorMap*>(a.w1_map);
    const CUtensorMap& tw2 = *reinterpret_cast<const CUtensorMap*>(a.w2_map);
    // one K chunk: A through tensor memory; more: in-place transform + shared-memory A (B200_ENGINE_LAYERFUSE_TSA=0/1 forces one)
    static const int tsa_env = [] { const char* e = getenv("B200_ENGINE_LAYERFUSE_TSA"); return e ? atoi(e) : -1; }();
    const bool tsa = tsa_env >= 0 ? tsa_env != 0 : a.Cin <= kDsCH;
    cudaError_t le = hv == 2 ? (tsa ? LaunchDs<2, true>(t, tw1, tw2, tout, cst, p, stream) : LaunchDs<2, false>(t, tw1, tw2, tout, cst, p, stream))
                             : (tsa ? LaunchDs<1, true>(t, tw1, tw2, tout, cst, p, stream) : LaunchDs<1, false>(t, tw1, tw2, tout, cst, p, stream));
    if (p.trace && le == cudaSuccess) {  // debug only: dump the timeline of one CTA
        cudaStreamSynchronize(stream);
        static unsigned long long host[48 * 16];
        cudaMemcpy(host, trace_buf, sizeof(host), cudaMemcpyDeviceToHost);
        static int dumps = 0;
        const int want = getenv("B200_DS_TRACE_LAUNCH") ? atoi(getenv("B200_DS_TRACE_LAUNCH")) : 5;
        if (dumps++ == want) {
            const char* names[12] = {"tma0", "xf_raw", "xf_rdone", "xf_done", "mma_c_at", "mma_c_iss", "mma_d_iss", "e1_at", "e1_go", "e1_done", "e2_go", "e2_done"};
            unsigned long long t0 = ~0ull;
            for (int i = 0; i < 48 * 16; ++i) if (host[i] && host[i] < t0) t0 = host[i];
            fprintf(stderr, "dense stream trace H=%d Cin=%d (ns since first stamp)\n", a.H, a.Cin);
            fprintf(stderr, "  k ");
            for (int e = 0; e < 12; ++e) fprintf(stderr, " %9s", names[e]);
            fprintf(stderr, "\n");
            for (int k = 0; k < 40; ++k) {
                fprintf(stderr, " %2d ", k);
                for (int e = 0; e < 12; ++e) fprintf(stderr, " %9lld", host[k * 16 + e] ? (long long)(host[k * 16 + e] - t0) : -1LL);
                fprintf(stderr, "\n");
            }
        }
    }
    return le;
}

}  // namespace kernels
}  // namespace b200
