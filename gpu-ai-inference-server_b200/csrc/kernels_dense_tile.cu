// kernels_dense_tile.cu — one DenseNet dense LAYER (BN-ReLU-Conv1x1(->128)-BN-ReLU-Conv3x3(->32)) fused into one persistent
// kernel for the LARGE-image blocks (56x56 and 28x28), where a CTA cannot own whole images: the image is cut into 14x14
// output tiles and the 128-channel bottleneck tensor of a tile's 16x16 halo patch is recomputed per tile (1.31x conv1
// work) and lives only in shared memory.  Per layer this removes the bottleneck tensor's round trip through HBM/L2
// (128 B written + ~147 B re-read per pixel against 32 B of real output) and one of the two kernel launches.  Replaces, for
// those blocks, the Conv/BatchNormalization/Relu/Concat nodes ONNX Runtime executes one by one inside
// `Ort::Session::Run` (reference inference_engine/src/model.cpp:1264-1270).
//
// One launch = one layer over every tile of the batch ("unit" = one 14x14 tile of one image).  Within a launch no CTA reads
// what another writes (a layer reads channels [0, Cin) and writes [Cin, Cin + 32)), so units are independent and the CTA
// pipelines them:
//   phase A  conv1 of unit k+1: two 4-D TMA boxes {128 B, 16 px, 8 rows} per K chunk land the halo patch (out-of-image
//            pixels zero-filled) + the weight chunk -> transform warps apply BN1+ReLU in place -> tcgen05.mma into 2x128
//            TMEM columns.
//   epi 1    TMEM -> BN2+ReLU -> e4m3 -> 128-byte-swizzled patch in shared memory (double buffered); patch pixels outside
//            the image are written as zeros (they are the 3x3 conv's padding).
//   phase B  conv2 of unit k: nine row-shifted UMMA views of the patch x the resident 3x3 weights -> 2x32 TMEM columns
//            (double buffered).
//   epi 2    TMEM -> scale -> e4m3 -> 32-byte stores into the block buffer's channel slice.
// The tensor pipe executes A(k+1) ahead of B(k), so epilogue 1 of unit k+1 and the loads of unit k+2 run under B(k).
//
// Warps (576 threads): 0-7 transform, 8-15 epilogue (two per TMEM lane quarter), 16 TMA producer, 17 MMA issuer.
//
// The conv1 A operand goes through TENSOR memory (tcgen05.mma with A in TMEM): the transform warps read the landed raw tile
// from shared memory once, apply BN1+ReLU in registers and tcgen05.st the result into a 4-deep ring of 32-column A tiles; the
// landing buffer is released as soon as it has been read.  conv1 weights (Cin <= 256) and conv2 weights are resident.
//
// STATUS (measured on B200, bs256, e4m3): bit-compatible with the two-kernel path (same parity tests pass) but SLOWER - block 1
// 941 us against 735 us, first five layers of block 2 237 us against 241 us - so the engine only uses it when
// B200_ENGINE_TILEFUSE=1.  Why (B200_DENSE_DBG ablations + B200_DENSE_TRACE timelines): 141 KB of shared memory is pinned by
// the two patches, the 3x3 weights and the 1x1 weights, which leaves a 5 x 16 KB landing ring; with ~3 us of loaded TMA latency
// that caps the stream at ~27 KB/us per SM, half of what the layer needs (the kernel with ALL math removed still takes 613 us for
// block 1).  The two-kernel schedule spends the same shared memory on 5-7 deep rings per kernel instead.
#include <cstdlib>

#include "kernels.h"
#include "umma_ptx.cuh"

namespace b200 {
namespace kernels {

namespace {

constexpr int kDtThreads = 576;
constexpr int kDtXfWarps = 8;
constexpr int kDtEpiWarps = 8;
constexpr int kDtTile = 14;                           // output tile edge
constexpr int kDtPW = 16;                             // patch edge (tile + halo)
constexpr int kDtMargin = 24;                         // patch slots in front of / between / behind the patches (tap shifts reach +-17)
constexpr int kDtPatchSlots = kDtMargin + 256 + kDtMargin + 256 + kDtMargin;
constexpr int kDtPatchBytes = kDtPatchSlots * 128;    // 73 KB: two patches
constexpr int kDtW2Bytes = 9 * 32 * 128;              // 36 KB
constexpr int kDtCH = 128;                            // e4m3 elements per 128-byte K chunk
constexpr int kDtMaxChunks = 2;                       // conv1 weights are resident: Cin <= 256
constexpr int kDtW1Bytes = kDtMaxChunks * 128 * kRowBytes;  // 32 KB
constexpr int kDtRaw = 5;                             // landing ring: [128 patch pixels][128 B] raw tiles
constexpr int kDtABufs = 4;                           // transformed A tiles in tensor memory (32 columns each)
constexpr int kDtVecBytes = (128 + 128 + 32 + 32) * 4;
constexpr int kDtBnBytes = 2 * (kDtMaxChunks * kDtCH / 2) * 4;  // folded BN1 scale | shift as f16x2 pairs
constexpr int kDtSmemBytes = 1024 + kDtRaw * kATileBytes + kDtW1Bytes + kDtW2Bytes + kDtPatchBytes + kDtVecBytes + kDtBnBytes + 512;
constexpr int kDtTmemCols = 512;                      // 256 (conv1) + 2 x 64 (conv2, double buffered) + 4 x 32 (A tiles)
static_assert(kDtSmemBytes <= 232448, "dense tile kernel exceeds the 227 KB shared-memory limit");

struct DtParams {
    const DenseLayerDesc* layer;   // device pointer to THIS layer's descriptor
    void* buf;                     // block buffer, NHWC e4m3
    int pitch;
    int n, H, W;
    int tiles_x, tiles_per_image, num_units;
    int dbg;                       // debug (B200_DENSE_DBG bitmask): 1 skip conv1 MMAs, 2 skip conv2 MMAs, 4 skip transform, 8 skip epilogue-1 math
    unsigned long long* trace;     // debug (B200_DENSE_TRACE): [unit][16] globaltimer stamps of CTA 0, else null
};

__device__ __forceinline__ void DtStamp(const DtParams& p, int k, int ev) {
    if (p.trace && blockIdx.x == 0 && k < 24 && (threadIdx.x & 31) == 0) {
        unsigned long long tm;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tm));
        p.trace[k * 16 + ev] = tm;
    }
}

__device__ __forceinline__ void TmaLoad2DGlobalMapT(void* smem_dst, const void* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(SmemAddr(smem_dst)), "l"((uint64_t)map), "r"(SmemAddr(bar)), "r"(c0), "r"(c1)
        : "memory");
}

struct DtGeom {
    int ch_base, k_lo, k_hi;
};
// K chunk c of a Cin-channel layer: the last chunk is placed at Cin - 128 (overlapping its predecessor, an L2 hit) and only its
// new K steps are multiplied; Cin < 128 uses the first K steps of the single chunk.
__device__ __forceinline__ DtGeom DtGeomOf(int c, int Cin) {
    DtGeom g;
    g.ch_base = c * kDtCH; g.k_lo = 0; g.k_hi = kDtCH;
    if (g.ch_base + kDtCH > Cin) {
        if (Cin >= kDtCH) { g.k_lo = g.ch_base + kDtCH - Cin; g.ch_base = Cin - kDtCH; }
        else g.k_hi = Cin;
    }
    return g;
}

// D[tmem] (+)= A[tmem] * B[smem]^T (kind::f8f6f4): the A tile lives in tensor memory, row = lane, K packed 4 e4m3 per column.
__device__ __forceinline__ void UmmaTS(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 16 consecutive 32-bit columns of this thread's lane
__device__ __forceinline__ void TmemStore16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void TmemStoreWait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(kDtThreads, 1)
dense_tile_kernel(const __grid_constant__ CUtensorMap tmap_x, const DtParams p) {
    using MmaT = __nv_fp8_e4m3;
    using ME = MmaElem<MmaT>;
    constexpr int EPV = ME::kPerVec;   // 16

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w1 = smem + kDtRaw * kATileBytes;
    uint8_t* s_w2 = s_w1 + kDtW1Bytes;
    uint8_t* s_patch = s_w2 + kDtW2Bytes;
    float* s_vec = reinterpret_cast<float*>(s_patch + kDtPatchBytes);  // s1 128 | b1 128 | s2 32 | b2 32
    uint32_t* s_bn = reinterpret_cast<uint32_t*>(s_vec + 320);         // BN1 scale pairs [128] | shift pairs [128]
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(s_bn + kDtBnBytes / 4);
    uint64_t* raw_empty = raw_full + kDtRaw;
    uint64_t* a_full = raw_empty + kDtRaw;
    uint64_t* a_empty = a_full + kDtABufs;
    uint64_t* w_full = a_empty + kDtABufs;
    uint64_t* acc1_full = w_full + 1;        // [t]
    uint64_t* acc1_empty = acc1_full + 2;    // [t]
    uint64_t* patch_full = acc1_empty + 2;   // [2]
    uint64_t* acc2_full = patch_full + 2;    // [buf][t]
    uint64_t* acc2_empty = acc2_full + 4;    // [buf][t]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc2_empty + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const DenseLayerDesc* L = p.layer;
    const int num_my = ((int)blockIdx.x < p.num_units) ? (p.num_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int Cin = L->Cin;
    const int nc = (Cin + kDtCH - 1) / kDtCH;   // 1 or 2
    const int items_per_unit = 2 * nc;          // (M tile, chunk) pairs, tile-major

    if (warp == 16 && lane == 0) {
        for (int s = 0; s < kDtRaw; ++s) {
            MbarInit(&raw_full[s], 1);
            MbarInit(&raw_empty[s], kDtXfWarps);
        }
        for (int s = 0; s < kDtABufs; ++s) {
            MbarInit(&a_full[s], kDtXfWarps);
            MbarInit(&a_empty[s], 1);
        }
        MbarInit(w_full, 1);
        for (int t = 0; t < 2; ++t) {
            MbarInit(&acc1_full[t], 1);
            MbarInit(&acc1_empty[t], kDtEpiWarps);
        }
        for (int b = 0; b < 2; ++b) MbarInit(&patch_full[b], kDtEpiWarps);
        for (int i = 0; i < 4; ++i) {
            MbarInit(&acc2_full[i], 1);
            MbarInit(&acc2_empty[i], kDtEpiWarps / 2);
        }
        FenceBarrierInit();
        PrefetchTensorMap(&tmap_x);
    }
    if (warp == 17) TmemAlloc(tmem_slot, kDtTmemCols);
    // margins are only ever read into discarded accumulator rows, but keep them finite
    for (int i = threadIdx.x; i < kDtPatchBytes / 16; i += kDtThreads) reinterpret_cast<uint4*>(s_patch)[i] = make_uint4(0u, 0u, 0u, 0u);
    // per-channel vectors (weights: not produced by the previous kernel, safe before the dependency wait)
    if (threadIdx.x < 128) {
        s_vec[threadIdx.x] = L->s1[threadIdx.x];
        s_vec[128 + threadIdx.x] = L->b1 ? L->b1[threadIdx.x] : 0.f;
        if (threadIdx.x < 32) {
            s_vec[256 + threadIdx.x] = L->s2[threadIdx.x];
            s_vec[288 + threadIdx.x] = L->b2 ? L->b2[threadIdx.x] : 0.f;
        }
        s_bn[threadIdx.x] = 2 * (int)threadIdx.x < Cin ? L->pre_scale[threadIdx.x] : 0u;
        s_bn[128 + threadIdx.x] = 2 * (int)threadIdx.x < Cin ? L->pre_shift[threadIdx.x] : 0u;
    }
    FenceProxyAsync();
    TcFenceBefore();
    __syncthreads();
    TcFenceAfter();
    const uint32_t tmem_base = *tmem_slot;
    constexpr uint32_t acc1_col = 0, acc2_col = 256, a_col = 384;
    GridDepLaunch();

    if (warp == 16) {
        // =========================================================== TMA producer
        if (ElectOne()) {
            MbarArriveExpectTx(w_full, (uint32_t)(kDtW2Bytes + nc * 128 * kRowBytes));
            for (int t = 0; t < 9; ++t) TmaLoad2DGlobalMapT(s_w2 + t * 32 * 128, &L->w2, w_full, t * kDtCH, 0);
            for (int c = 0; c < nc; ++c) TmaLoad2DGlobalMapT(s_w1 + c * 128 * kRowBytes, &L->w1, w_full, DtGeomOf(c, Cin).ch_base, 0);
        }
        __syncwarp();
        GridDepWait();
        int stage = 0;
        uint32_t phase = 0;
        int k = 0;
        for (int u = blockIdx.x; u < p.num_units; u += gridDim.x, ++k) {
            const int img = u / p.tiles_per_image, tt = u - img * p.tiles_per_image;
            const int ty = tt / p.tiles_x, tx = tt - ty * p.tiles_x;
            const int x0 = tx * kDtTile - 1, y0 = ty * kDtTile - 1;
            for (int it = 0; it < items_per_unit; ++it) {
                const int t = it >= nc ? 1 : 0, c = it - t * nc;
                MbarWaitWarp(&raw_empty[stage], phase ^ 1u);
                if (it == 0) DtStamp(p, k, 0);
                if (ElectOne()) {
                    MbarArriveExpectTx(&raw_full[stage], (uint32_t)kATileBytes);
                    TmaLoad4D(smem + stage * kATileBytes, &tmap_x, &raw_full[stage], DtGeomOf(c, Cin).ch_base, x0, y0 + 8 * t, img);
                }
                __syncwarp();
                if (it == items_per_unit - 1) DtStamp(p, k, 1);
                if (++stage == kDtRaw) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 17) {
        // =========================================================== MMA issuer
        constexpr uint32_t idesc1 = MakeInstrDesc(ME::kFmt, 128);
        constexpr uint32_t idesc2 = MakeInstrDesc(ME::kFmt, 32);
        const uint64_t w1_desc = MakeSmemDesc(SmemAddr(s_w1));
        const uint64_t w2_desc = MakeSmemDesc(SmemAddr(s_w2));
        const uint32_t patch_addr0 = SmemAddr(s_patch) + kDtMargin * 128;
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        // The issuing thread is back-pressured by the tensor pipe (72 conv2 MMAs block it for ~2.3 us), so it works as a small
        // scheduler: conv1 items (whose accumulators the epilogue is waiting for) go first whenever their operands are ready,
        // conv2 is issued in slices of 12 MMAs in between.  conv1 of unit ka may only be issued once conv2 of unit ka-2 has been
        // issued in full: epilogue 1 of unit ka overwrites the patch conv2 of unit ka-2 read, and the commit that releases it
        // covers every MMA issued before it.
        int ab = 0;
        uint32_t aphase = 0;
        int ka = 0, ia = 0;   // next conv1 item: unit ka, item ia (tile-major)
        int kb = 0, sb = 0;   // next conv2 slice: unit kb, slice sb = tile * 3 + tap row
        MbarWaitWarp(w_full, 0);
        while (ka < num_my || kb < num_my) {
            bool did = false;
            if (ka < num_my && ka <= kb + 1) {
                const int t = ia >= nc ? 1 : 0, c = ia - t * nc;
                int ready = 1;
                if (lane == 0) {
                    ready = MbarTest(&a_full[ab], aphase);
                    if (ready && c == 0) ready = MbarTest(&acc1_empty[t], ((uint32_t)ka & 1u) ^ 1u);
                }
                ready = __shfl_sync(0xffffffffu, ready, 0);
                if (ready) {
                    TcFenceAfter();
                    const DtGeom g = DtGeomOf(c, Cin);
                    const int ks_lo = g.k_lo / ME::kStepK, ks_hi = g.k_hi / ME::kStepK;
                    const uint64_t b_desc = w1_desc + (uint64_t)(c * ((128 * kRowBytes) >> 4));
                    if (ElectOne()) {
#pragma unroll
                        for (int ks = 0; ks < kDtCH / ME::kStepK; ++ks)
                            if (ks >= ks_lo && ks < ks_hi && !(p.dbg & 1))
                                UmmaTS(tmem_u + acc1_col + t * 128, tmem_u + a_col + ab * 32 + ks * 8, b_desc + (uint64_t)(2 * ks), idesc1,
                                       (c > 0 || ks > ks_lo) ? 1u : 0u);
                        UmmaCommit(&a_empty[ab]);
                        if (c == nc - 1) UmmaCommit(&acc1_full[t]);
                    }
                    __syncwarp();
                    if (++ab == kDtABufs) { ab = 0; aphase ^= 1u; }
                    if (++ia == items_per_unit) { DtStamp(p, ka, 3); ia = 0; ++ka; }
                    did = true;
                }
            }
            if (!did && kb < num_my) {
                const int t = sb >= 3 ? 1 : 0, fr = sb - 3 * t;
                const uint32_t b = (uint32_t)kb & 1u, ph = ((uint32_t)kb >> 1) & 1u;
                int ready = 1;
                if (lane == 0) {
                    if (sb == 0) ready = MbarTest(&patch_full[b], ph);
                    if (ready && fr == 0) ready = MbarTest(&acc2_empty[b * 2 + t], ph ^ 1u);
                }
                ready = __shfl_sync(0xffffffffu, ready, 0);
                if (ready) {
                    TcFenceAfter();
                    if (sb == 0) DtStamp(p, kb, 6);
                    const uint32_t patch_addr = patch_addr0 + b * (uint32_t)((256 + kDtMargin) * 128);
                    if (ElectOne()) {
#pragma unroll
                        for (int fs = 0; fs < 3; ++fs) {
                            const int shift = (fr - 1) * kDtPW + (fs - 1);
                            const uint64_t a_desc = MakeSmemDesc(patch_addr + (uint32_t)((t * 128 + shift) * 128));
                            const uint64_t b_desc = w2_desc + (uint64_t)((fr * 3 + fs) * (32 * 128 / 16));
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks)
                                if (!(p.dbg & 2)) UmmaSS<ME::kKind>(tmem_u + acc2_col + b * 64 + t * 32, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc2,
                                                  (fr | fs | ks) ? 1u : 0u);
                        }
                        if (fr == 2) UmmaCommit(&acc2_full[b * 2 + t]);
                    }
                    __syncwarp();
                    if (++sb == 6) { DtStamp(p, kb, 7); sb = 0; ++kb; }
                    did = true;
                }
            }
            if (!did) __nanosleep(40);
        }
    } else if (warp < kDtXfWarps) {
        // =========================================================== transform warps: raw tile (smem) -> BN1 + ReLU -> A tile (tmem)
        // warp w owns rows 32*(w&3)..+31 (its TMEM lane quarter) and 16-byte pieces 4*(w>>2)..+3 of them
        const int q = warp & 3, hh = warp >> 2;
        const int row = q * 32 + lane;
        uint32_t off[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) off[i] = (uint32_t)(row * kRowBytes + (((4 * hh + i) ^ (row & 7)) << 4));
        const uint32_t smem_base = SmemAddr(smem);
        const uint32_t bn_addr = SmemAddr(s_bn);
        const bool relu = L->pre_relu != 0;
        int stage = 0, ab = 0;
        uint32_t phase = 0, aphase = 0;
        for (int k = 0; k < num_my; ++k) {
            for (int it = 0; it < items_per_unit; ++it) {
                const int c = it >= nc ? it - nc : it;
                const DtGeom g = DtGeomOf(c, Cin);
                const int p_lo = g.k_lo / EPV, p_hi = g.k_hi / EPV;
                const uint32_t a_base = smem_base + stage * kATileBytes;
                MbarWaitWarp(&raw_full[stage], phase);
                uint4 v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = LdsV4(a_base + off[i]);
                if (!(p.dbg & 4)) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int pc = 4 * hh + i;
                        if (pc >= p_lo && pc < p_hi) {
                            const uint32_t ca = bn_addr + (uint32_t)((g.ch_base + pc * EPV) * 2);  // f16x2 pair index = channel / 2
                            const uint4 s0 = LdsV4(ca), s1 = LdsV4(ca + 16), h0 = LdsV4(ca + 512), h1 = LdsV4(ca + 528);
                            const uint32_t sc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                            const uint32_t sh[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
                            v[i] = relu ? ProloguePiece<MmaT, true>(v[i], sc, sh) : ProloguePiece<MmaT, false>(v[i], sc, sh);
                        }
                    }
                }
                // the landing buffer can take the next tile as soon as every warp has read its share
                __syncwarp();
                if (lane == 0) MbarArrive(&raw_empty[stage]);
                MbarWaitWarp(&a_empty[ab], aphase ^ 1u);
                TcFenceAfter();
                const uint32_t r[16] = {v[0].x, v[0].y, v[0].z, v[0].w, v[1].x, v[1].y, v[1].z, v[1].w,
                                        v[2].x, v[2].y, v[2].z, v[2].w, v[3].x, v[3].y, v[3].z, v[3].w};
                TmemStore16(tmem_base + ((uint32_t)(q * 32) << 16) + a_col + ab * 32 + hh * 16, r);
                TmemStoreWait();
                TcFenceBefore();
                __syncwarp();
                if (lane == 0) MbarArrive(&a_full[ab]);
                if (it == items_per_unit - 1 && warp == 7) DtStamp(p, k, 2);
                if (++stage == kDtRaw) { stage = 0; phase ^= 1u; }
                if (++ab == kDtABufs) { ab = 0; aphase ^= 1u; }
            }
        }
    } else {
        // =========================================================== epilogue warps: lane quarter q, column half / tile h
        const int q = warp & 3, h = (warp - kDtXfWarps) >> 2;
        const int row = q * 32 + lane;
        uint8_t* buf = reinterpret_cast<uint8_t*>(p.buf);
        const uint32_t vaddr = SmemAddr(s_vec);
        const uint32_t patch0 = SmemAddr(s_patch) + kDtMargin * 128;
        const bool relu1 = L->relu1 != 0, relu2 = L->relu2 != 0;
        const int c_off_out = L->c_off_out;
        GridDepWait();

        // conv1 accumulators of the CTA's k-th unit `u` -> BN2 + ReLU -> e4m3 -> patch[k & 1]
        auto epi1 = [&](int k, int u) {
            const int img = u / p.tiles_per_image, tt = u - img * p.tiles_per_image;
            const int ty = tt / p.tiles_x, tx = tt - ty * p.tiles_x;
            const int x0 = tx * kDtTile - 1, y0 = ty * kDtTile - 1;
            const uint32_t pbase = patch0 + ((uint32_t)k & 1u) * (uint32_t)((256 + kDtMargin) * 128);
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                MbarWaitWarp(&acc1_full[t], (uint32_t)k & 1u);
                TcFenceAfter();
                if (warp == 12 && t == 0) DtStamp(p, k, 4);
                const int m = t * 128 + row;
                const int y = y0 + (m >> 4), x = x0 + (m & 15);
                const bool inside = y >= 0 && y < p.H && x >= 0 && x < p.W;
                const uint32_t slot_addr = pbase + (uint32_t)m * 128;
                const uint32_t sw = (uint32_t)m & 7u;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    if (p.dbg & 8) break;
                    const int cg = 2 * h + j;
                    uint32_t r[32];
                    TmemLoad32(tmem_base + ((uint32_t)(q * 32) << 16) + acc1_col + t * 128 + cg * 32, r);
                    TmemLoadWait();
                    uint32_t w[8];
                    if (relu1) EpiloguePack32Smem<MmaT, true>(r, vaddr + cg * 128, vaddr + 512 + cg * 128, w);
                    else EpiloguePack32Smem<MmaT, false>(r, vaddr + cg * 128, vaddr + 512 + cg * 128, w);
                    if (!inside) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) w[i] = 0u;
                    }
                    StsV4(slot_addr + (((2 * cg) ^ sw) << 4), make_uint4(w[0], w[1], w[2], w[3]));
                    StsV4(slot_addr + (((2 * cg + 1) ^ sw) << 4), make_uint4(w[4], w[5], w[6], w[7]));
                }
                TcFenceBefore();
                __syncwarp();
                if (lane == 0) MbarArrive(&acc1_empty[t]);  // conv1 of the next unit may refill this tile's accumulator
            }
            FenceProxyAsync();  // patch writes -> visible to the tensor core
            __syncwarp();
            if (lane == 0) MbarArrive(&patch_full[k & 1]);
            if (warp == 12) DtStamp(p, k, 5);
        };

        int u = blockIdx.x;
        if (num_my > 0) epi1(0, u);
        for (int k = 0; k < num_my; ++k, u += gridDim.x) {
            if (k + 1 < num_my) epi1(k + 1, u + gridDim.x);
            // ---- epilogue 2: M tile `h` of unit k -> the layer's 32-channel slice of the block buffer
            const int img = u / p.tiles_per_image, tt = u - img * p.tiles_per_image;
            const int ty = tt / p.tiles_x, tx = tt - ty * p.tiles_x;
            const uint32_t b = (uint32_t)k & 1u, ph = ((uint32_t)k >> 1) & 1u;
            const int s = h * 128 + row;
            const int yy = s >> 4, xx = s & 15;
            const bool valid = yy >= 1 && yy <= kDtTile && xx >= 1 && xx <= kDtTile;
            MbarWaitWarp(&acc2_full[b * 2 + h], ph);
            TcFenceAfter();
            if (warp == 12) DtStamp(p, k, 8);
            uint32_t r[32];
            TmemLoad32(tmem_base + ((uint32_t)(q * 32) << 16) + acc2_col + b * 64 + h * 32, r);
            TmemLoadWait();
            TcFenceBefore();
            __syncwarp();
            if (lane == 0) MbarArrive(&acc2_empty[b * 2 + h]);
            uint32_t w[8];
            if (relu2) EpiloguePack32Smem<MmaT, true>(r, vaddr + 1024, vaddr + 1152, w);
            else EpiloguePack32Smem<MmaT, false>(r, vaddr + 1024, vaddr + 1152, w);
            if (valid) {
                const int y = ty * kDtTile + yy - 1, x = tx * kDtTile + xx - 1;
                uint8_t* dst = buf + (((size_t)img * p.H + y) * p.W + x) * p.pitch + c_off_out;
                *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(dst + 16) = make_uint4(w[4], w[5], w[6], w[7]);
            }
            if (warp == 12) DtStamp(p, k, 9);
        }
    }

    TcFenceBefore();
    __syncthreads();
    if (warp == 17) {
        TcFenceAfter();
        TmemDealloc(tmem_base, kDtTmemCols);
    }
}

}  // namespace

bool DenseTileGeometry(int H, int W) { return H == W && H > kDtTile && H % kDtTile == 0; }
int DenseTileMaxCin() { return kDtMaxChunks * kDtCH; }

cudaError_t DenseTileFp8(const DenseBlockArgs& a, int layer, cudaStream_t stream) {
    if (!DenseTileGeometry(a.H, a.W) || layer < 0 || layer >= a.num_layers || !a.layers_dev) return cudaErrorInvalidValue;
    // (the caller guarantees Cin <= DenseTileMaxCin() for every layer of a tiled run: the descriptors live in device memory)
    if (a.n <= 0) return cudaSuccess;
    DtParams p;
    p.layer = a.layers_dev + layer;
    p.buf = a.buf; p.pitch = a.pitch; p.n = a.n; p.H = a.H; p.W = a.W;
    p.tiles_x = a.W / kDtTile;
    p.tiles_per_image = p.tiles_x * (a.H / kDtTile);
    p.num_units = a.n * p.tiles_per_image;
    p.trace = nullptr;
    { const char* d = getenv("B200_DENSE_DBG"); p.dbg = d ? atoi(d) : 0; }
    static unsigned long long* trace_buf = nullptr;
    const char* tr = getenv("B200_DENSE_TRACE");
    const bool tracing = tr && atoi(tr) == 100 + layer && a.H == 56;
    if (tracing) {
        if (!trace_buf) cudaMalloc(&trace_buf, 24 * 16 * 8);
        cudaMemsetAsync(trace_buf, 0, 24 * 16 * 8, stream);
        p.trace = trace_buf;
    }
    TensorMap tx;
    const uint64_t dims[4] = {(uint64_t)a.pitch, (uint64_t)a.W, (uint64_t)a.H, (uint64_t)a.n};
    const uint64_t strides[3] = {(uint64_t)a.pitch, (uint64_t)a.W * a.pitch, (uint64_t)a.H * a.W * a.pitch};
    const uint32_t box[4] = {128u, (uint32_t)kDtPW, 8u, 1u};
    if (MakeTensorMap(&tx, a.buf, 1, 4, dims, strides, box, true) != 0) return cudaErrorInvalidValue;
    static int sm_count[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!sm_count[dev]) {
        cudaError_t e = cudaFuncSetAttribute(dense_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDtSmemBytes);
        if (e != cudaSuccess) return e;
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sm_count[dev] = n > 0 ? n : 148;
    }
    const int grid = p.num_units < sm_count[dev] ? p.num_units : sm_count[dev];
    cudaError_t le = LaunchPdl(dense_tile_kernel, grid, kDtThreads, kDtSmemBytes, stream, *reinterpret_cast<const CUtensorMap*>(&tx), p);
    CountLaunch();
    if (tracing && le == cudaSuccess) {  // debug only: timeline of CTA 0's first units
        cudaStreamSynchronize(stream);
        static unsigned long long host[24 * 16];
        cudaMemcpy(host, trace_buf, sizeof(host), cudaMemcpyDeviceToHost);
        static int dumps = 0;
        if (dumps++ < 1) {
            const char* names[10] = {"tma_c0", "tma_cN", "xf_done", "A_issued", "ep_acc1", "ep1_done", "B_start", "B_issued", "ep_acc2", "ep2_done"};
            const unsigned long long t0 = host[0];
            fprintf(stderr, "dense tile trace H=%d layer=%d Cin-chunks (ns since first TMA)\n", a.H, layer);
            for (int k = 0; k < 12; ++k) {
                fprintf(stderr, " unit %2d:", k);
                for (int e = 0; e < 10; ++e) fprintf(stderr, " %s=%lld", names[e], host[k * 16 + e] ? (long long)(host[k * 16 + e] - t0) : -1LL);
                fprintf(stderr, "\n");
            }
        }
    }
    return le;
}

}  // namespace kernels
}  // namespace b200
