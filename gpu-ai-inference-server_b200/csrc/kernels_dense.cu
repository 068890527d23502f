// kernels_dense.cu — a whole DenseNet dense block (all of its BN-ReLU-Conv1x1-BN-ReLU-Conv3x3 layers) in ONE
// persistent kernel, for the small-image blocks (14x14 and 7x7) where a layer-per-kernel schedule is pure launch and
// pipeline-fill latency.  Replaces, for those blocks, the ~80 Conv/BatchNormalization/Relu/Concat nodes ONNX Runtime
// executes one by one inside `Ort::Session::Run` (reference inference_engine/src/model.cpp:1264-1270).
//
// Why it needs no grid-wide synchronisation: images are independent and a dense layer only ever reads channels of
// the SAME pixels' neighbourhood, so a CTA that owns whole images can run every layer of the block on them before
// any other CTA has finished layer 1.  The only cross-layer hand-over is inside the CTA: the 32 channels a layer
// appends to the block buffer (global memory, concat in place) are re-read by the next layer's LAST K chunk.
//
// Per (image group, layer) "unit":
//   phase A  1x1 conv: TMA box loads of the block-buffer rows (+ the layer's weight chunk) -> transform warps apply
//            the layer's folded BN1+ReLU in place -> tcgen05.mma into the conv1 accumulators (MT1 x 128 TMEM columns).
//   epi 1    TMEM -> scale/bias(BN2)+ReLU -> e4m3 -> written straight into a zero-ringed, 128-byte-swizzled patch
//            in SHARED memory ([slot = padded pixel][128 channels]); the bottleneck tensor never leaves the SM.
//   phase B  3x3 conv: nine row-shifted UMMA views of the patch x the TMA-loaded 3x3 weights -> 2 x 32 TMEM columns.
//            14x14 images (padded row = 16 slots): the three taps of a filter row are stacked along N instead (N = 96, three
//            row-shifted views, see kernels_conv3x3.cu) and epilogue 2 adds the column groups across neighbouring lanes.
//   epi 2    TMEM -> per-channel scale -> e4m3 -> 32-byte stores into the block buffer's channel slice.
// Phase A of layer l+1 (all K chunks but the last) overlaps phase B / epi 2 of layer l.
//
// Warps (448 threads): 0 TMA producer, 1 MMA issuer, 2-9 transform, 10-13 epilogue.
#include <cstdlib>

#include "kernels.h"
#include "umma_ptx.cuh"

namespace b200 {
namespace kernels {

namespace {

constexpr int kDbThreads = 448;
constexpr int kDbXfWarps = 8;
constexpr int kDbEpiThreads = 128;
constexpr int kDbMargin = 24;                        // patch slots in front of / behind the images (tap shifts reach +-(PW+1))
constexpr int kDbPatchSlots = kDbMargin + 256 + kDbMargin;
constexpr int kDbPatchBytes = kDbPatchSlots * 128;   // 38 KB
constexpr int kDbW2Bytes = 9 * 32 * 128;             // 36 KB
constexpr int kDbCH = 128;                           // e4m3 elements per 128-byte K chunk

template <int MT1> struct DbCfg {
    static constexpr int kStageBytes = MT1 * kATileBytes + 128 * kRowBytes;  // A tiles + the [128][128 B] weight chunk
    static constexpr int kStages = MT1 == 2 ? 3 : 4;
    static constexpr int kVecBytes = 2 * (128 + 128 + 32 + 32) * 4;          // s1, b1, s2, b2, double buffered
    static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kDbW2Bytes + kDbPatchBytes + kVecBytes + 512;
    static constexpr int kTmemCols = MT1 == 2 ? 512 : 256;                   // MT1*128 (conv1) + 2 conv2 accumulators, power of two
    static constexpr int kAcc2Stride = MT1 == 2 ? 128 : 32;                  // 96 columns used when the taps are stacked
};

struct DbParams {
    const DenseLayerDesc* layers;  // device array
    int num_layers;
    void* buf;                     // block buffer, NHWC e4m3
    int pitch;                     // channels per pixel of the block buffer
    int n, H, W;
    int ipc;                       // images per CTA pass
    int num_groups;
    int interleave;                // a CTA that owns two image groups alternates between them layer by layer
    unsigned long long* trace;  // debug: [unit][16] globaltimer stamps of CTA 0 (B200_DENSE_TRACE), else null
};

// One elected lane; warp converged.
__device__ __forceinline__ void TmaLoad2DGlobalMap(void* smem_dst, const void* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(SmemAddr(smem_dst)), "l"((uint64_t)map), "r"(SmemAddr(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void TmaLoad3DGlobalMap(void* smem_dst, const void* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(SmemAddr(smem_dst)), "l"((uint64_t)map), "r"(SmemAddr(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void FenceProxyAsyncGlobal() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

__device__ __forceinline__ void Stamp(const DbParams& p, uint32_t k, int ev) {
    if (p.trace && blockIdx.x == 0 && k < 64 && (threadIdx.x & 31) == 0) {
        unsigned long long tm;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tm));
        atomicMax(&p.trace[k * 16 + ev], tm);
    }
}

struct DbGeom {
    int ch_base, k_lo, k_hi;
};
__device__ __forceinline__ DbGeom DbGeomOf(int c, int Cin) {
    DbGeom g;
    g.ch_base = c * kDbCH; g.k_lo = 0; g.k_hi = kDbCH;
    if (g.ch_base + kDbCH > Cin) {
        if (Cin >= kDbCH) { g.k_lo = g.ch_base + kDbCH - Cin; g.ch_base = Cin - kDbCH; }
        else g.k_hi = Cin;
    }
    return g;
}

// The order of the (image group, layer) units of one CTA, identical for every role.  A unit's last K chunk needs the 32 channels
// the previous layer of the SAME images appended (global-memory round trip: stores, fences, TMA reload); a CTA that owns two
// groups therefore alternates between them layer by layer, so that this hand-over hides behind the other group's unit.
//   f(k, grp, l, slot): k = running unit index of this CTA, slot = 0 / 1 = which of the two interleaved groups
template <typename F>
__device__ __forceinline__ void DbWalk(const DbParams& p, F&& f) {
    uint32_t k = 0;
    const int stride = (int)gridDim.x;
    if (p.interleave) {
        for (int g = blockIdx.x; g < p.num_groups; g += 2 * stride) {
            const int gb = g + stride;
            const bool two = gb < p.num_groups;
            for (int l = 0; l < p.num_layers; ++l) {
                f(k++, g, l, 0);
                if (two) f(k++, gb, l, 1);
            }
        }
    } else {
        for (int g = blockIdx.x; g < p.num_groups; g += stride)
            for (int l = 0; l < p.num_layers; ++l) f(k++, g, l, 0);
    }
}
// Layer of the unit that FOLLOWS unit (grp, l, slot) in DbWalk's order, or -1 after the last one.
__device__ __forceinline__ int DbNextLayer(const DbParams& p, int grp, int l, int slot) {
    const int stride = (int)gridDim.x;
    if (p.interleave) {
        const int ga = slot == 0 ? grp : grp - stride;            // first group of the pair
        const bool two = ga + stride < p.num_groups;
        if (slot == 0 && two) return l;                              // the other group, same layer
        if (l + 1 < p.num_layers) return l + 1;
        return ga + 2 * stride < p.num_groups ? 0 : -1;              // next pair starts at layer 0
    }
    if (l + 1 < p.num_layers) return l + 1;
    return grp + stride < p.num_groups ? 0 : -1;
}

template <int MT1>
__global__ void __launch_bounds__(kDbThreads, 1)
dense_block_kernel(const __grid_constant__ CUtensorMap tmap_x, const DbParams p) {   // tmap_x: box {128 B, MT1 * 128 rows}
    using MmaT = __nv_fp8_e4m3;
    using ME = MmaElem<MmaT>;
    using Cfg = DbCfg<MT1>;
    constexpr int NS = Cfg::kStages;
    constexpr int EPV = ME::kPerVec;   // 16
    constexpr int kPairs = EPV / 2;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w2 = smem + NS * Cfg::kStageBytes;
    uint8_t* s_patch = s_w2 + kDbW2Bytes;
    float* s_vec = reinterpret_cast<float*>(s_patch + kDbPatchBytes);  // [2][s1 128 | b1 128 | s2 32 | b2 32]
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(s_vec + 2 * 320);
    uint64_t* xf_full = raw_full + NS;
    uint64_t* empty_bar = xf_full + NS;
    uint64_t* w2_full = empty_bar + NS;
    uint64_t* w2_empty = w2_full + 1;
    uint64_t* acc1_full = w2_empty + 1;
    uint64_t* acc1_empty = acc1_full + 1;
    uint64_t* patch_full = acc1_empty + 1;
    uint64_t* acc2_full = patch_full + 1;   // [2]
    uint64_t* acc2_empty = acc2_full + 2;   // [2]
    uint64_t* out_ready = acc2_empty + 2;   // [2]: one per interleaved group slot
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(out_ready + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the single-thread roles sit on the highest warp ids (the issue arbiter prefers high warp ids; see kernels_conv1x1.cu)
    constexpr int kNW = kDbThreads / 32;
    const int wrole = warp >= kNW - 2 ? warp - (kNW - 2) : warp + 2;  // 0 TMA, 1 MMA, 2..9 transform, 10..13 epilogue
    const int HW = p.H * p.W, PW = p.W + 2, SL = (p.H + 2) * PW;  // slots per image (with the zero ring)

    if (wrole == 0 && lane == 0) {
        for (int s = 0; s < NS; ++s) {
            MbarInit(&raw_full[s], 1);
            MbarInit(&xf_full[s], kDbXfWarps);
            MbarInit(&empty_bar[s], 1);
        }
        MbarInit(w2_full, 1);
        MbarInit(w2_empty, 1);
        MbarInit(acc1_full, 1);
        MbarInit(acc1_empty, kDbEpiThreads / 32);
        MbarInit(patch_full, kDbEpiThreads / 32);
        for (int t = 0; t < 2; ++t) {
            MbarInit(&acc2_full[t], 1);
            MbarInit(&acc2_empty[t], kDbEpiThreads / 32);
        }
        MbarInit(&out_ready[0], kDbEpiThreads / 32);
        MbarInit(&out_ready[1], kDbEpiThreads / 32);
        FenceBarrierInit();
        PrefetchTensorMap(&tmap_x);
    }
    if (wrole == 1) TmemAlloc(tmem_slot, Cfg::kTmemCols);
    // zero the patch once: the ring around every image is the 3x3 conv's zero padding and is never written again
    for (int i = threadIdx.x; i < kDbPatchBytes / 16; i += kDbThreads) reinterpret_cast<uint4*>(s_patch)[i] = make_uint4(0u, 0u, 0u, 0u);
    FenceProxyAsync();
    TcFenceBefore();
    __syncthreads();
    TcFenceAfter();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t acc1_col = 0, acc2_col = MT1 * 128;
    // filter-row taps stacked along N: needs the padded row to be exactly 16 slots so that a pixel's left/right neighbours
    // sit in the same warp's TMEM lanes
    const bool stack_taps = MT1 == 2 && p.W == 14;
    GridDepLaunch();

    if (wrole == 0) {
        // =========================================================== TMA producer
        GridDepWait();
        int stage = 0;
        uint32_t phase = 0;
        uint32_t units_of[2] = {0u, 0u};  // units finished per group slot = phases of out_ready[slot]
        DbWalk(p, [&](uint32_t k, int grp, int l, int slot) {
            const int row0 = grp * p.ipc * HW;
            const DenseLayerDesc* L = p.layers + l;
            const int Cin = L->Cin;
            const int nc = (Cin + kDbCH - 1) / kDbCH;
            for (int c = 0; c < nc; ++c) {
                if (c == nc - 1) {
                    // the last K chunk holds the 32 channels the previous layer of THESE images has just appended
                    MbarWaitWarp(&out_ready[slot], (units_of[slot] & 1u) ^ 1u);
                    Stamp(p, k, 1);
                }
                MbarWaitWarp(&empty_bar[stage], phase ^ 1u);
                if (c == nc - 1) Stamp(p, k, 2);
                if (ElectOne()) {
                    const DbGeom g = DbGeomOf(c, Cin);
                    uint8_t* dst = smem + stage * Cfg::kStageBytes;
                    MbarArriveExpectTx(&raw_full[stage], (uint32_t)Cfg::kStageBytes);
                    // ONE box for all MT1 A tiles (a TMA instruction costs the unit ~700 cycles whatever its size, tools/ubench/tma_rate.cu)
                    TmaLoad2D(dst, &tmap_x, &raw_full[stage], g.ch_base, row0);
                    TmaLoad2DGlobalMap(dst + MT1 * kATileBytes, &L->w1, &raw_full[stage], g.ch_base, 0);
                }
                __syncwarp();
                if (++stage == NS) { stage = 0; phase ^= 1u; }
            }
            // 3x3 weights of this layer, AFTER the last A chunk: their buffer is free only once phase B of the previous unit has
            // run, and with two interleaved groups that is later than the moment the last chunk can be fetched
            MbarWaitWarp(w2_empty, (k & 1u) ^ 1u);
            if (ElectOne()) {
                MbarArriveExpectTx(w2_full, (uint32_t)kDbW2Bytes);
                TmaLoad3DGlobalMap(s_w2, &L->w2, w2_full, 0, 0, 0);   // all nine taps: box {128 B, 32 rows, 9 taps}
            }
            __syncwarp();
            ++units_of[slot];
        });
    } else if (wrole == 1) {
        // =========================================================== MMA issuer
        constexpr uint32_t idesc1 = MakeInstrDesc(ME::kFmt, 128);
        constexpr uint32_t idesc2 = MakeInstrDesc(ME::kFmt, 32);
        const uint64_t stage_desc = MakeSmemDesc(SmemAddr(smem));
        const uint64_t w2_desc = MakeSmemDesc(SmemAddr(s_w2));
        const uint32_t patch_addr = SmemAddr(s_patch) + kDbMargin * 128;
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        int stage = 0;
        uint32_t phase = 0;
        DbWalk(p, [&](uint32_t k, int, int l, int) {
            {
                const int Cin = p.layers[l].Cin;
                const int nc = (Cin + kDbCH - 1) / kDbCH;
                // ---- phase A: the conv1 accumulators must have been drained by epilogue 1 of the previous unit
                MbarWaitWarp(acc1_empty, (k & 1u) ^ 1u);
                TcFenceAfter();
                for (int c = 0; c < nc; ++c) {
                    const DbGeom g = DbGeomOf(c, Cin);
                    const int ks_lo = g.k_lo / ME::kStepK, ks_hi = g.k_hi / ME::kStepK;
                    const uint64_t a_desc = stage_desc + (uint64_t)((uint32_t)stage * (Cfg::kStageBytes >> 4));
                    const uint64_t b_desc = a_desc + (uint64_t)((MT1 * kATileBytes) >> 4);
                    if (c == nc - 1) Stamp(p, k, 14);
                    MbarWaitWarp(&xf_full[stage], phase);
                    TcFenceAfter();
                    if (c == nc - 1) Stamp(p, k, 15);
                    if (ElectOne()) {
#pragma unroll
                        for (int t = 0; t < MT1; ++t) {
#pragma unroll
                            for (int ks = 0; ks < kDbCH / ME::kStepK; ++ks)
                                if (ks >= ks_lo && ks < ks_hi)
                                    UmmaSS<ME::kKind>(tmem_u + acc1_col + t * 128, a_desc + (uint64_t)(t * (kATileBytes >> 4) + 2 * ks),
                                                      b_desc + (uint64_t)(2 * ks), idesc1, (c > 0 || ks > ks_lo) ? 1u : 0u);
                        }
                        if (c == nc - 1 && p.trace && blockIdx.x == 0 && k < 64) {
                            unsigned long long tm;
                            asm volatile("mov.u64 %0, %globaltimer;" : "=l"(tm));
                            p.trace[k * 16 + 0] = tm;
                        }
                        UmmaCommit(&empty_bar[stage]);
                        if (c == nc - 1) UmmaCommit(acc1_full);
                    }
                    __syncwarp();
                    if (c == nc - 1) Stamp(p, k, 5);
                    if (++stage == NS) { stage = 0; phase ^= 1u; }
                }
                // ---- phase B: 3x3 conv over the patch epilogue 1 is writing
                MbarWaitWarp(w2_full, k & 1u);
                MbarWaitWarp(patch_full, k & 1u);
                TcFenceAfter();
                Stamp(p, k, 6);
#pragma unroll 1
                for (int t = 0; t < 2; ++t) {
                    MbarWaitWarp(&acc2_empty[t], (k & 1u) ^ 1u);
                    TcFenceAfter();
                    if (ElectOne()) {
                        const uint32_t d2 = tmem_u + acc2_col + t * Cfg::kAcc2Stride;
                        if (stack_taps) {
                            constexpr uint32_t idesc96 = MakeInstrDesc(ME::kFmt, 96);
#pragma unroll
                            for (int fr = 0; fr < 3; ++fr) {
                                // D[m][fs*32 + o] = A[m + (fr-1)*PW] . W(fr, fs)[o]: the contribution to output slot m - (fs - 1)
                                const uint64_t a_desc = MakeSmemDesc(patch_addr + (uint32_t)((t * 128 + (fr - 1) * (p.W + 2)) * 128));
                                const uint64_t b_desc = w2_desc + (uint64_t)(fr * 3 * (32 * 128 / 16));
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)
                                    UmmaSS<ME::kKind>(d2, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc96, (fr | ks) ? 1u : 0u);
                            }
                        } else {
#pragma unroll
                            for (int tap = 0; tap < 9; ++tap) {
                                const int shift = (tap / 3 - 1) * (p.W + 2) + (tap % 3 - 1);
                                const uint64_t a_desc = MakeSmemDesc(patch_addr + (uint32_t)((t * 128 + shift) * 128));
                                const uint64_t b_desc = w2_desc + (uint64_t)(tap * (32 * 128 / 16));
#pragma unroll
                                for (int ks = 0; ks < 4; ++ks)
                                    UmmaSS<ME::kKind>(d2, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc2, (tap | ks) ? 1u : 0u);
                            }
                        }
                        UmmaCommit(&acc2_full[t]);
                        if (t == 1) UmmaCommit(w2_empty);
                    }
                    __syncwarp();
                }
                Stamp(p, k, 7);
            }
        });
    } else if (wrole < 2 + kDbXfWarps) {
        // =========================================================== transform warps: in-place BN1 + ReLU on the A tiles
        const int tw = wrole - 2;
        uint32_t off_full[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = i * 32 + lane;
            off_full[i] = (uint32_t)(row * kRowBytes + ((tw ^ (row & 7)) << 4));
        }
        const uint32_t smem_base = SmemAddr(smem);
        int stage = 0;
        uint32_t phase = 0;
        DbWalk(p, [&](uint32_t k, int, int l, int) {
            {
                const DenseLayerDesc* L = p.layers + l;
                const int Cin = L->Cin;
                const int nc = (Cin + kDbCH - 1) / kDbCH;
                const uint32_t* gsc = L->pre_scale;
                const uint32_t* gsh = L->pre_shift;
                const bool relu = L->pre_relu != 0;
                for (int c = 0; c < nc; ++c) {
                    const DbGeom g = DbGeomOf(c, Cin);
                    const int p_lo = g.k_lo / EPV, p_hi = g.k_hi / EPV;
                    const uint32_t a_base = smem_base + stage * Cfg::kStageBytes;
                    // this warp owns 16-byte piece `tw` of every row; its 16 channels' packed constants
                    const bool mine = tw >= p_lo && tw < p_hi;
                    uint32_t sc[kPairs], sh[kPairs];
                    if (mine) {
                        const int ch0 = g.ch_base + tw * EPV;
                        const uint4 a0 = __ldg(reinterpret_cast<const uint4*>(gsc + ch0 / 2)), a1 = __ldg(reinterpret_cast<const uint4*>(gsc + ch0 / 2 + 4));
                        const uint4 b0 = __ldg(reinterpret_cast<const uint4*>(gsh + ch0 / 2)), b1 = __ldg(reinterpret_cast<const uint4*>(gsh + ch0 / 2 + 4));
                        sc[0] = a0.x; sc[1] = a0.y; sc[2] = a0.z; sc[3] = a0.w; sc[4] = a1.x; sc[5] = a1.y; sc[6] = a1.z; sc[7] = a1.w;
                        sh[0] = b0.x; sh[1] = b0.y; sh[2] = b0.z; sh[3] = b0.w; sh[4] = b1.x; sh[5] = b1.y; sh[6] = b1.z; sh[7] = b1.w;
                    }
                    MbarWaitWarp(&raw_full[stage], phase);
                    if (c == nc - 1 && tw == 7) Stamp(p, k, 3);
                    if (mine) {
#pragma unroll
                        for (int t = 0; t < MT1; ++t) {
                            uint4 v[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) v[i] = LdsV4(a_base + t * kATileBytes + off_full[i]);
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                v[i] = relu ? ProloguePiece<MmaT, true>(v[i], sc, sh) : ProloguePiece<MmaT, false>(v[i], sc, sh);
#pragma unroll
                            for (int i = 0; i < 4; ++i) StsV4(a_base + t * kATileBytes + off_full[i], v[i]);
                        }
                    }
                    FenceProxyAsync();
                    __syncwarp();
                    if (lane == 0) MbarArrive(&xf_full[stage]);  // one arrive per warp: 256 per-thread arrives serialise on the barrier word
                    if (c == nc - 1) Stamp(p, k, 4);
                    if (++stage == NS) { stage = 0; phase ^= 1u; }
                }
            }
        });
    } else {
        // =========================================================== epilogue warps (TMEM lane quarter = warp & 3)
        const int q = warp & 3;
        const int et = (wrole - (2 + kDbXfWarps)) * 32 + lane;  // 0..127
        const int row = q * 32 + lane;                          // accumulator row this thread reads
        uint8_t* buf = reinterpret_cast<uint8_t*>(p.buf);
        // fixed per-thread maps.  epilogue 1: accumulator row -> patch slot.  epilogue 2: slot -> pixel of the group.
        uint32_t slot_addr[MT1];
        bool slot_ok[MT1];
#pragma unroll
        for (int t = 0; t < MT1; ++t) {
            const int pi = t * 128 + row;
            const int img = pi / HW, rem = pi - img * HW, y = rem / p.W, x = rem - y * p.W;
            slot_ok[t] = pi < p.ipc * HW;
            const int slot = kDbMargin + img * SL + (y + 1) * PW + (x + 1);
            slot_addr[t] = SmemAddr(s_patch) + (uint32_t)slot * 128;
        }
        int pix_rel[2];  // pixel index inside the group, or -1
#pragma unroll
        for (int t = 0; t < 2; ++t) {
            const int s = t * 128 + row;
            const int img = s / SL, rem = s - img * SL, yy = rem / PW, xx = rem - yy * PW;
            pix_rel[t] = (img < p.ipc && yy >= 1 && yy <= p.H && xx >= 1 && xx <= p.W) ? img * HW + (yy - 1) * p.W + (xx - 1) : -1;
        }
        GridDepWait();
        DbWalk(p, [&](uint32_t k, int grp, int l, int slot) {
            const int img0 = grp * p.ipc;
            const int n_valid = (p.n - img0 < p.ipc ? p.n - img0 : p.ipc) * HW;  // pixels of this group that exist
            {
                const DenseLayerDesc* L = p.layers + l;
                float* vec = s_vec + (k & 1u) * 320;
                // per-channel vectors of a layer -> shared memory, double buffered by unit parity.  The vectors of unit k + 1 are
                // fetched while this unit waits for its 3x3 MMAs (below): at the start of a unit the global-load latency would sit on
                // the epilogue warps' chain
                auto load_vec = [&](const DenseLayerDesc* Lx, float* v) {
                    v[et] = Lx->s1[et];
                    v[128 + et] = Lx->b1 ? Lx->b1[et] : 0.f;
                    if (et < 32) {
                        v[256 + et] = Lx->s2[et];
                        v[288 + et] = Lx->b2 ? Lx->b2[et] : 0.f;
                    }
                };
                if (k == 0) load_vec(L, vec);
                NamedBarSync(1, kDbEpiThreads);
                const uint32_t vaddr = SmemAddr(vec);
                // ---- epilogue 1: conv1 accumulators -> BN2 + ReLU -> e4m3 -> swizzled patch rows
                MbarWaitWarp(acc1_full, k & 1u);
                TcFenceAfter();
                if (q == 2) Stamp(p, k, 8);
#pragma unroll
                for (int t = 0; t < MT1; ++t) {
#pragma unroll 1
                    for (int cg = 0; cg < 4; ++cg) {
                        uint32_t r[32];
                        TmemLoad32(tmem_base + ((uint32_t)(q * 32) << 16) + acc1_col + t * 128 + cg * 32, r);
                        TmemLoadWait();
                        uint32_t w[8];
                        if (L->relu1) EpiloguePack32Smem<MmaT, true>(r, vaddr + cg * 128, vaddr + 512 + cg * 128, w);
                        else EpiloguePack32Smem<MmaT, false>(r, vaddr + cg * 128, vaddr + 512 + cg * 128, w);
                        if (slot_ok[t]) {
                            const uint32_t sw = (slot_addr[t] >> 7) & 7u;
                            StsV4(slot_addr[t] + (((2 * cg) ^ sw) << 4), make_uint4(w[0], w[1], w[2], w[3]));
                            StsV4(slot_addr[t] + (((2 * cg + 1) ^ sw) << 4), make_uint4(w[4], w[5], w[6], w[7]));
                        }
                    }
                }
                TcFenceBefore();
                FenceProxyAsync();  // patch writes -> visible to the tensor core
                __syncwarp();
                if (lane == 0) {
                    MbarArrive(acc1_empty);
                    MbarArrive(patch_full);
                }
                if (q == 2) Stamp(p, k, 9);
                {   // the next unit's vectors (its buffer was last read by epilogue 2 of unit k - 1, long finished)
                    const int ln = DbNextLayer(p, grp, l, slot);
                    if (ln >= 0) load_vec(p.layers + ln, s_vec + ((k + 1u) & 1u) * 320);
                }
                // ---- epilogue 2: conv2 accumulators -> scale -> e4m3 -> the layer's 32-channel slice of the block buffer
#pragma unroll 1
                for (int t = 0; t < 2; ++t) {
                    MbarWaitWarp(&acc2_full[t], k & 1u);
                    TcFenceAfter();
                    if (q == 2) Stamp(p, k, 10 + t);
                    uint32_t r[32];
                    const uint32_t t2 = tmem_base + ((uint32_t)(q * 32) << 16) + acc2_col + t * Cfg::kAcc2Stride;
                    if (stack_taps) {
                        uint32_t r0[32], r2[32];
                        TmemLoad32(t2, r0);
                        TmemLoad32(t2 + 32, r);
                        TmemLoad32(t2 + 64, r2);
                        TmemLoadWait();
                        // out[m] = D[m - 1][fs = 0] + D[m][fs = 1] + D[m + 1][fs = 2]; a padded row is 16 lanes, stored pixels are x = 1..14
#pragma unroll
                        for (int c = 0; c < 32; ++c) {
                            const float a0 = __shfl_up_sync(0xffffffffu, __uint_as_float(r0[c]), 1);
                            const float a2 = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[c]), 1);
                            r[c] = __float_as_uint((a0 + __uint_as_float(r[c])) + a2);
                        }
                    } else {
                        TmemLoad32(t2, r);
                        TmemLoadWait();
                    }
                    TcFenceBefore();
                    __syncwarp();
                    if (lane == 0) MbarArrive(&acc2_empty[t]);
                    uint32_t w[8];
                    if (L->relu2) EpiloguePack32Smem<MmaT, true>(r, vaddr + 1024, vaddr + 1152, w);
                    else EpiloguePack32Smem<MmaT, false>(r, vaddr + 1024, vaddr + 1152, w);
                    if (pix_rel[t] >= 0 && pix_rel[t] < n_valid) {
                        uint8_t* dst = buf + ((size_t)img0 * HW + pix_rel[t]) * p.pitch + L->c_off_out;
                        *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
                        *reinterpret_cast<uint4*>(dst + 16) = make_uint4(w[4], w[5], w[6], w[7]);
                    }
                }
                // the next layer's last K chunk re-reads these channels through TMA (async proxy)
                if (q == 2) Stamp(p, k, 12);
                __threadfence();
                FenceProxyAsyncGlobal();
                __syncwarp();
                if (lane == 0) MbarArrive(&out_ready[slot]);
                if (q == 2) Stamp(p, k, 13);
            }
        });
    }

    TcFenceBefore();
    __syncthreads();
    if (wrole == 1) {
        TcFenceAfter();
        TmemDealloc(tmem_base, Cfg::kTmemCols);
    }
}

template <int MT1>
cudaError_t LaunchDb(const CUtensorMap& tx, const DbParams& p, cudaStream_t stream) {
    using Cfg = DbCfg<MT1>;
    auto kern = dense_block_kernel<MT1>;
    static int sm_count[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!sm_count[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) return e;
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sm_count[dev] = n > 0 ? n : 148;
    }
    const int grid = p.num_groups < sm_count[dev] ? p.num_groups : sm_count[dev];
    cudaError_t le = LaunchPdl(kern, grid, kDbThreads, Cfg::kSmemBytes, stream, tx, p);
    CountLaunch();
    return le;
}

}  // namespace

bool DenseBlockGeometry(int H, int W, int* ipc, int* mt1) {
    if (H < 1 || W < 1 || H != W) return false;
    const int HW = H * W, SL = (H + 2) * (W + 2);
    if (W + 3 > kDbMargin) return false;
    int best = 0;
    for (int i = 1; i <= 8; ++i)
        if (i * HW <= 256 && i * SL <= 256) best = i;
    if (const char* e = getenv("B200_DENSE_IPC")) {  // experiment: fewer images per CTA pass = more CTAs at work
        const int v = atoi(e);
        if (v >= 1 && v <= best) best = v;
    }
    if (!best) return false;
    *ipc = best;
    *mt1 = (best * HW + 127) / 128;
    return true;
}

cudaError_t DenseBlockFp8(const DenseBlockArgs& a, cudaStream_t stream) {
    int ipc = 0, mt1 = 0;
    if (!DenseBlockGeometry(a.H, a.W, &ipc, &mt1) || a.num_layers <= 0 || !a.layers_dev) return cudaErrorInvalidValue;
    if (a.n <= 0) return cudaSuccess;
    // Images per CTA pass: the kernel is a per-unit latency chain, so what counts is the number of passes a CTA makes
    // (waves) times the length of the chain, which grows with the conv1 M tiles.  Measured at bs256, 7x7: 3 images per pass
    // (86 CTAs, 2 M tiles) 195 us, 2 per pass (128 CTAs, 1 M tile) 144 us, 1 per pass (2 waves) 276 us.
    if (!getenv("B200_DENSE_IPC")) {
        int sms = 148;
        { int dev = 0, v = 0; cudaGetDevice(&dev); if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && v > 0) sms = v; }
        const int HW = a.H * a.W;
        double best_cost = 1e30;
        int best_ipc = ipc;
        for (int i = 1; i <= ipc; ++i) {
            const int groups = (a.n + i - 1) / i, waves = (groups + sms - 1) / sms, tiles = (i * HW + 127) / 128;
            const double cost = waves * (6.0 + 1.8 * tiles);
            if (cost < best_cost - 1e-9) { best_cost = cost; best_ipc = i; }
        }
        ipc = best_ipc;
        mt1 = (ipc * HW + 127) / 128;
    }
    DbParams p;
    p.layers = a.layers_dev; p.num_layers = a.num_layers;
    p.buf = a.buf; p.pitch = a.pitch; p.n = a.n; p.H = a.H; p.W = a.W;
    p.ipc = ipc;
    p.num_groups = (a.n + ipc - 1) / ipc;
    { static const bool il = [] { const char* e = getenv("B200_DENSE_INTERLEAVE"); return !(e && e[0] == '0'); }(); p.interleave = il ? 1 : 0; }
    p.trace = nullptr;
    static unsigned long long* trace_buf = nullptr;
    if (getenv("B200_DENSE_TRACE")) {
        if (!trace_buf) { cudaMalloc(&trace_buf, 64 * 16 * 8); }
        cudaMemsetAsync(trace_buf, 0, 64 * 16 * 8, stream);
        p.trace = trace_buf;
    }
    TensorMap tx;
    const uint64_t rows = (uint64_t)a.n * a.H * a.W;
    const uint64_t dims[2] = {(uint64_t)a.pitch, rows};
    const uint64_t strides[1] = {(uint64_t)a.pitch};
    const uint32_t box[2] = {128u, (uint32_t)(mt1 * kTileM)};
    if (MakeTensorMap(&tx, a.buf, 1, 2, dims, strides, box, true) != 0) return cudaErrorInvalidValue;
    const CUtensorMap& t = *reinterpret_cast<const CUtensorMap*>(&tx);
    cudaError_t le = mt1 == 2 ? LaunchDb<2>(t, p, stream) : LaunchDb<1>(t, p, stream);
    if (p.trace && le == cudaSuccess) {  // debug only: dump the first units' timeline of CTA 0
        cudaStreamSynchronize(stream);
        static unsigned long long host[64 * 16];
        cudaMemcpy(host, trace_buf, sizeof(host), cudaMemcpyDeviceToHost);
        static int dumps = 0;
        if (dumps++ < 2) {
            const char* names[16] = {"mma_issued", "tma_outrdy", "tma_tail", "xf_tail_in", "xf_tail_out", "mma_A_done", "mma_patch", "mma_B_done",
                                     "ep_acc1", "ep_patch", "ep_acc2_0", "ep_acc2_1", "ep_prefence", "ep_outrdy", "mma_at_tail", "mma_tail_seen"};
            unsigned long long t0 = host[1];
            fprintf(stderr, "dense trace H=%d layers=%d (ns since first stamp)\n", a.H, a.num_layers);
            for (int k = 0; k < 6 && k < a.num_layers; ++k) {
                fprintf(stderr, " unit %d:", k);
                for (int e = 0; e < 16; ++e) fprintf(stderr, " %s=%lld", names[e], host[k * 16 + e] ? (long long)(host[k * 16 + e] - t0) : -1LL);
                fprintf(stderr, "\n");
            }
        }
    }
    return le;
}

}  // namespace kernels
}  // namespace b200
