// kernels_conv3x3.cu — 3x3 / stride 1 / pad 1 convolution of a 128-byte-per-pixel bottleneck (Cin = 128 e4m3 or
// 64..128 bf16 channels) into <= 32 output channels: the second conv of every DenseNet dense layer.
//
// Replaces the Conv node ONNX Runtime would execute inside `Ort::Session::Run` (reference
// inference_engine/src/model.cpp:1264-1270) for `/features/denseblockX/denselayerY/conv2/Conv`.
//
// A CTA owns TH x TW output pixels of one image.  ONE 4-D TMA box {128 B of channels, 16 pixels, TH+2 rows, 1 image}
// lands the zero-padded input patch (out-of-image coordinates are zero-filled by the TMA unit = the conv padding) as
// [slot = row*16 + x][128 B] in SWIZZLE_128B layout, which is exactly the K-major UMMA operand layout with one
// pixel per row.  The nine filter taps are nine ROW-SHIFTED views of that tile: output slot m reads patch slot
// m + fr*16 + fs, i.e. the same descriptor with its start address advanced by (fr*16 + fs) * 128 bytes.  (Measured
// on B200: the tensor core applies the 128-byte XOR swizzle to ABSOLUTE shared-memory address bits, exactly like
// the TMA unit that wrote the tile, so a start address that is not a multiple of 8 rows needs no fix-up; the
// descriptor's matrix-base-offset field must stay 0 - setting it to the row phase gives wrong results.)
// N = 32 output channels would make every MMA A-fetch bound (4 KB of A per 16 cycles of math, measured: tensor pipe 30 % active),
// so the three taps of one filter ROW are stacked along N: the weight tile of filter row fr is [fs*32 + o][128 B] (96 rows) and
// ONE MMA per (fr, K step) with the A view shifted by fr*16 slots computes, for every patch slot m,
//     D[m][fs*32 + o] += A[m + fr*16] . W(fr, fs)[o]          which is the tap's contribution to OUTPUT slot m - fs.
// The epilogue adds the three column groups across neighbouring TMEM lanes (two warp shuffles per channel; a warp owns two
// 16-slot patch rows, and m + 2 stays inside the row for every stored x <= 13): out[m] = D[m][0:32] + D[m+1][32:64] + D[m+2][64:96].
// A is fetched 3x per K step instead of 9x (12 MMAs of N = 96 per 128-byte plane instead of 36 of N = 32).
// No thread ever touches the activations: the producer
// is one elected lane issuing TMA, so the whole CTA is 6 warps (TMA, MMA, 4 epilogue).  The 9 weight tiles
// ([32][128 B] each) are TMA-loaded once and stay resident; patches stream through a ring of kBufs buffers.
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "umma_ptx.cuh"

namespace b200 {
namespace kernels {

namespace {

constexpr int kC3Threads = 192;
constexpr int kC3PW = 16;                          // patch width in slots (TW <= 14)
constexpr int kC3PlaneBytes = 10 * kC3PW * 128;    // (TH+2 <= 10) rows x 16 slots x 128 B = 20 KB per 128-byte K plane
constexpr int kC3BN = 32;
constexpr int kC3NS = 3 * kC3BN;                   // MMA N: the three taps of a filter row stacked
constexpr int kC3Acc = 4;                          // TMEM accumulators
constexpr int kC3AccStride = 128;                  // columns between accumulators (96 used)

template <int CPT> struct C3Cfg {                  // CPT = 128-byte K planes per pixel (1: e4m3, 2: bf16)
    static constexpr int kBufs = CPT == 1 ? 6 : 3;
    static constexpr int kWeightBytes = 9 * CPT * kC3BN * 128;
    static constexpr int kBufBytes = CPT * kC3PlaneBytes;
    static constexpr int kSmemBytes = 1024 + kWeightBytes + kBufs * kBufBytes + 1024 /*junk-row overreach*/ + 256;
};

struct C3Params {
    void* out;
    const float* out_scale;
    const float* bias;
    int post_relu;
    int H, W, out_pitch, out_coff, Cout, in_coff;
    int n, TH, TW, tiles_x, tiles_y, num_tiles;
    int ksteps;  // K steps (32 B) per 128-byte plane that carry data (4 when the plane is full)
};

template <typename MmaT, typename OutT>
__global__ void __launch_bounds__(kC3Threads, 1)
conv3x3_tma_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_in, const C3Params p) {
    using ME = MmaElem<MmaT>;
    constexpr int CPT = sizeof(MmaT) == 1 ? 1 : 2;
    using Cfg = C3Cfg<CPT>;
    constexpr int NB = Cfg::kBufs;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w = smem;
    uint8_t* s_patch = smem + Cfg::kWeightBytes;
    uint64_t* w_bar = reinterpret_cast<uint64_t*>(s_patch + NB * Cfg::kBufBytes + 1024);
    uint64_t* patch_full = w_bar + 1;
    uint64_t* patch_empty = patch_full + NB;
    uint64_t* tmem_full = patch_empty + NB;
    uint64_t* tmem_empty = tmem_full + kC3Acc;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + kC3Acc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the single-thread roles sit on the highest warp ids (the issue arbiter prefers high warp ids; see kernels_conv1x1.cu)
    constexpr int kNW = kC3Threads / 32;
    const int wrole = warp >= kNW - 2 ? warp - (kNW - 2) : warp + 2;  // 0 TMA, 1 MMA, 2..5 epilogue

    if (wrole == 0 && lane == 0) {
        MbarInit(w_bar, 1);
        for (int b = 0; b < NB; ++b) {
            MbarInit(&patch_full[b], 1);
            MbarInit(&patch_empty[b], 1);
        }
        for (int a = 0; a < kC3Acc; ++a) {
            MbarInit(&tmem_full[a], 1);
            MbarInit(&tmem_empty[a], 128);
        }
        FenceBarrierInit();
        PrefetchTensorMap(&tmap_w);
        PrefetchTensorMap(&tmap_in);
    }
    if (wrole == 1) TmemAlloc(tmem_slot, kC3Acc * kC3AccStride);
    TcFenceBefore();
    __syncthreads();
    TcFenceAfter();
    const uint32_t tmem_base = *tmem_slot;
    GridDepLaunch();

    if (wrole == 0) {
        // =========================================================== TMA producer: weights once, then the patch ring
        if (ElectOne()) {
            MbarArriveExpectTx(w_bar, (uint32_t)Cfg::kWeightBytes);
            // packed K order is (tap, plane); resident order is (filter row, plane, fs) so that the three taps of a row form one
            // contiguous [96][128 B] tile
            for (int tap = 0; tap < 9; ++tap)
                for (int j = 0; j < CPT; ++j)
                    TmaLoad2D(s_w + (((tap / 3) * CPT + j) * 3 + tap % 3) * kC3BN * 128, &tmap_w, w_bar, (tap * CPT + j) * ME::kChunk, 0);
        }
        __syncwarp();
        GridDepWait();  // the patches are the previous kernel's output
        const uint32_t box_bytes = (uint32_t)(CPT * (p.TH + 2) * kC3PW * 128);
        uint32_t k = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++k) {
            const uint32_t buf = k % NB, par = ((k / NB) & 1u) ^ 1u;
            MbarWait(&patch_empty[buf], par);
            if (ElectOne()) {
                const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, img = tile / (p.tiles_x * p.tiles_y);
                MbarArriveExpectTx(&patch_full[buf], box_bytes);
#pragma unroll
                for (int j = 0; j < CPT; ++j)
                    TmaLoad4D(s_patch + buf * Cfg::kBufBytes + j * kC3PlaneBytes, &tmap_in, &patch_full[buf], p.in_coff + j * ME::kChunk,
                              tx * p.TW - 1, ty * p.TH - 1, img);
            }
            __syncwarp();
        }
    } else if (wrole == 1) {
        // =========================================================== MMA issuer
        constexpr uint32_t idesc = MakeInstrDesc(ME::kFmt, kC3NS);
        const uint64_t b_base = MakeSmemDesc(SmemAddr(s_w));
        const uint32_t patch_addr = SmemAddr(s_patch);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        MbarWait(w_bar, 0);
        uint32_t k = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++k) {
            const uint32_t buf = k % NB, ph = (k / NB) & 1u;
            const uint32_t acc = k % kC3Acc, acc_ph = (k / kC3Acc) & 1u;
            MbarWait(&tmem_empty[acc], acc_ph ^ 1u);
            MbarWait(&patch_full[buf], ph);
            TcFenceAfter();
            if (ElectOne()) {
                const uint32_t d_addr = tmem_u + acc * kC3AccStride;
                const uint32_t a_buf = patch_addr + buf * Cfg::kBufBytes;
#pragma unroll
                for (int fr = 0; fr < 3; ++fr) {
#pragma unroll
                    for (int j = 0; j < CPT; ++j) {
                        // rows shifted by fr*16 slots of 128 B; the fs shift lives in the epilogue
                        const uint64_t a_desc = MakeSmemDesc(a_buf + j * kC3PlaneBytes + fr * kC3PW * 128);
                        const uint64_t b_desc = b_base + (uint64_t)(((fr * CPT + j) * 3) * (kC3BN * 128 / 16));
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            if (ks < p.ksteps) UmmaSS<ME::kKind>(d_addr, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc, (fr | j | ks) ? 1u : 0u);
                    }
                }
                UmmaCommit(&patch_empty[buf]);
                UmmaCommit(&tmem_full[acc]);
            }
            __syncwarp();
        }
    } else {
        // =========================================================== epilogue (warps 2..5 -> TMEM lane quarters 2,3,0,1)
        const int e = warp & 3;
        OutT* out = reinterpret_cast<OutT*>(p.out);
        float sc[32], bi[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            sc[q] = q < p.Cout ? p.out_scale[q] : 0.f;
            bi[q] = (p.bias && q < p.Cout) ? p.bias[q] : 0.f;
        }
        const int mrow = e * 32 + lane;
        const int y = mrow / kC3PW, x = mrow - y * kC3PW;
        GridDepWait();  // stores may alias buffers the previous kernel still reads
        uint32_t k = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++k) {
            const uint32_t acc = k % kC3Acc, acc_phase = (k / kC3Acc) & 1u;
            const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, img = tile / (p.tiles_x * p.tiles_y);
            MbarWait(&tmem_full[acc], acc_phase);
            TcFenceAfter();
            uint32_t r[32];
            {
                uint32_t r1[32], r2[32];
                const uint32_t t_addr = tmem_base + ((uint32_t)(e * 32) << 16) + acc * kC3AccStride;
                TmemLoad32(t_addr, r);
                TmemLoad32(t_addr + 32, r1);
                TmemLoad32(t_addr + 64, r2);
                TmemLoadWait();
                TcFenceBefore();
                MbarArrive(&tmem_empty[acc]);  // the accumulator is in registers: release it before the math and the stores
                // out[m] = D[m][fs = 0] + D[m + 1][fs = 1] + D[m + 2][fs = 2]  (lanes 14, 15 of a 16-slot row are never stored)
#pragma unroll
                for (int c = 0; c < 32; ++c) {
                    const float a1 = __shfl_down_sync(0xffffffffu, __uint_as_float(r1[c]), 1);
                    const float a2 = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[c]), 2);
                    r[c] = __float_as_uint((__uint_as_float(r[c]) + a1) + a2);
                }
            }
            const int oy = ty * p.TH + y, ox = tx * p.TW + x;
            if (y < p.TH && x < p.TW && oy < p.H && ox < p.W) {
                constexpr int kWords = 32 * (int)sizeof(OutT) / 4;
                uint32_t w[kWords];
                if (p.post_relu) EpiloguePack32<OutT, true>(r, sc, bi, w);
                else EpiloguePack32<OutT, false>(r, sc, bi, w);
                OutT* orow = out + ((size_t)(img * p.H + oy) * p.W + ox) * p.out_pitch + p.out_coff;
                constexpr int kPer = 16 / (int)sizeof(OutT);  // channels per 16-byte store
#pragma unroll
                for (int q = 0; q < kWords / 4; ++q)
                    if (q * kPer < p.Cout) *reinterpret_cast<uint4*>(orow + q * kPer) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
            }
            __syncwarp();
        }
    }
    TcFenceBefore();
    __syncthreads();
    if (wrole == 1) {
        TcFenceAfter();
        TmemDealloc(tmem_base, kC3Acc * kC3AccStride);
    }
}

template <typename MmaT, typename OutT>
cudaError_t LaunchC3(const CUtensorMap& tw, const CUtensorMap& tin, const C3Params& p, cudaStream_t stream) {
    using Cfg = C3Cfg<sizeof(MmaT) == 1 ? 1 : 2>;
    auto kern = conv3x3_tma_kernel<MmaT, OutT>;
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
    const int grid = p.num_tiles < sm_count[dev] ? p.num_tiles : sm_count[dev];
    cudaError_t le = LaunchPdl(kern, grid, kC3Threads, Cfg::kSmemBytes, stream, tw, tin, p);
    CountLaunch();
    return le;
}


// ---------------------------------------------------------------------------------------------------------------- CTA pairs
// The same convolution with the two CTAs of a cluster executing ONE tcgen05.mma of M = 256 per (filter row, K step)
// (cta_group::2): each CTA supplies its own patch view (128 rows of A) and only HALF of the stacked weight tile (48 of the 96
// rows, at the same shared-memory offset in both CTAs).  With both operands in shared memory a dispatch is operand-fetch bound at
// 64 B/clk (tools/ubench/mma_issue.cu): 4 KB of A + 3 KB of B = 91 cycles alone, 4 KB + 1.5 KB = 65 cycles as a pair.
// The leader (cluster rank 0) issues the MMAs for both; its peer forwards "my patch has landed" and "my accumulator is drained"
// to the leader's barriers, the commit multicasts "patch free" / "accumulator full" to both CTAs.  Both CTAs walk the same number
// of steps (a peer without a tile left repeats the last one and stores nothing).  e4m3 only (one 128-byte K plane per pixel).
constexpr int kC3PairHalfRows = kC3NS / 2;          // 48 weight rows per CTA and filter row
constexpr int kC3PairWeightBytes = 3 * kC3PairHalfRows * 128;
constexpr int kC3PairBufs = 6;
constexpr int kC3PairSmemBytes = 1024 + 18 * 1024 /*weights, padded*/ + kC3PairBufs * kC3PlaneBytes + 1024 + 256;

constexpr int kC3PairThreads = 320;   // 8 epilogue warps (two per TMEM lane quarter, 16 output channels each), TMA, MMA

__global__ void __launch_bounds__(kC3PairThreads, 1)
conv3x3_pair_kernel(const __grid_constant__ CUtensorMap tmap_w16, const __grid_constant__ CUtensorMap tmap_in, const C3Params p) {
    using MmaT = __nv_fp8_e4m3;
    using OutT = __nv_fp8_e4m3;
    using ME = MmaElem<MmaT>;
    constexpr int NB = kC3PairBufs;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w = smem;                              // [fr][48 rows][128 B]
    uint8_t* s_patch = smem + 18 * 1024;
    uint64_t* w_bar = reinterpret_cast<uint64_t*>(s_patch + NB * kC3PlaneBytes + 1024);
    uint64_t* patch_full = w_bar + 1;
    uint64_t* patch_empty = patch_full + NB;
    uint64_t* peer_full = patch_empty + NB;           // leader only: the peer's patch of this buffer has landed
    uint64_t* tmem_full = peer_full + NB;
    uint64_t* tmem_empty = tmem_full + kC3Acc;        // leader only: both CTAs' epilogues have drained the accumulator (16 warps)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + kC3Acc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kNW = kC3PairThreads / 32;
    const int wrole = warp >= kNW - 2 ? warp - (kNW - 2) : warp + 2;  // 0 TMA, 1 MMA, 2..9 epilogue
    const uint32_t rank = ClusterCtaRank();
    const bool is_leader = rank == 0;
    // steps of this pair: the tiles of the even CTA decide; the odd CTA may be one tile short at the very end
    const int base = (int)(blockIdx.x & ~1u);
    const int steps = base < p.num_tiles ? (p.num_tiles - base + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (wrole == 0 && lane == 0) {
        MbarInit(w_bar, 1);
        for (int b = 0; b < NB; ++b) {
            MbarInit(&patch_full[b], 1);
            MbarInit(&patch_empty[b], 1);
            MbarInit(&peer_full[b], 1);
        }
        for (int a = 0; a < kC3Acc; ++a) {
            MbarInit(&tmem_full[a], 1);
            MbarInit(&tmem_empty[a], 16);
        }
        FenceBarrierInit();
        PrefetchTensorMap(&tmap_w16);
        PrefetchTensorMap(&tmap_in);
    }
    if (wrole == 1) TmemAlloc2(tmem_slot, kC3Acc * kC3AccStride);
    TcFenceBefore();
    __syncthreads();
    ClusterSync();   // the peer's barriers exist before anything arrives on them remotely
    TcFenceAfter();
    const uint32_t tmem_base = *tmem_slot;
    GridDepLaunch();

    if (wrole == 0) {
        // =========================================================== TMA producer: this CTA's half of the weights, then the patch ring
        if (ElectOne()) {
            MbarArriveExpectTx(w_bar, (uint32_t)kC3PairWeightBytes);
            // stacked rows of filter row fr: [fs*32 + o]; this CTA holds rows 48*rank .. 48*rank + 47 in pieces of 16 output channels
            for (int fr = 0; fr < 3; ++fr)
                for (int piece = 0; piece < 3; ++piece) {
                    const int row = (int)rank * kC3PairHalfRows + piece * 16;   // stacked row of the piece
                    const int fs = row / 32, o0 = row % 32;
                    TmaLoad2D(s_w + (fr * kC3PairHalfRows + piece * 16) * 128, &tmap_w16, w_bar, (fr * 3 + fs) * ME::kChunk, o0);
                }
        }
        __syncwarp();
        GridDepWait();
        const uint32_t box_bytes = (uint32_t)((p.TH + 2) * kC3PW * 128);
        for (int k = 0; k < steps; ++k) {
            const uint32_t buf = (uint32_t)k % NB, par = (((uint32_t)k / NB) & 1u) ^ 1u;
            int tile = (int)blockIdx.x + k * (int)gridDim.x;
            if (tile >= p.num_tiles) tile = p.num_tiles - 1;
            MbarWait(&patch_empty[buf], par);
            if (ElectOne()) {
                const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, img = tile / (p.tiles_x * p.tiles_y);
                MbarArriveExpectTx(&patch_full[buf], box_bytes);
                TmaLoad4D(s_patch + buf * kC3PlaneBytes, &tmap_in, &patch_full[buf], p.in_coff, tx * p.TW - 1, ty * p.TH - 1, img);
            }
            __syncwarp();
        }
    } else if (wrole == 1) {
        if (!is_leader) {
            // =========================================================== peer: forward "my patch has landed" to the leader
            for (int k = 0; k < steps; ++k) {
                const uint32_t buf = (uint32_t)k % NB, ph = ((uint32_t)k / NB) & 1u;
                MbarWait(&patch_full[buf], ph);
                if (ElectOne()) MbarArriveCluster(MapaShared(SmemAddr(&peer_full[buf]), 0));
                __syncwarp();
            }
        } else {
            // =========================================================== leader: MMA issuer for the pair
            constexpr uint32_t idesc = MakeInstrDescM(ME::kFmt, kC3NS, 256);
            const uint64_t b_base = MakeSmemDesc(SmemAddr(s_w));
            const uint32_t patch_addr = SmemAddr(s_patch);
            const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
            MbarWait(w_bar, 0);
            for (int k = 0; k < steps; ++k) {
                const uint32_t buf = (uint32_t)k % NB, ph = ((uint32_t)k / NB) & 1u;
                const uint32_t acc = (uint32_t)k % kC3Acc, acc_ph = ((uint32_t)k / kC3Acc) & 1u;
                MbarWait(&tmem_empty[acc], acc_ph ^ 1u);
                MbarWait(&patch_full[buf], ph);
                MbarWait(&peer_full[buf], ph);
                TcFenceAfter();
                if (ElectOne()) {
                    const uint32_t d_addr = tmem_u + acc * kC3AccStride;
                    const uint32_t a_buf = patch_addr + buf * kC3PlaneBytes;
#pragma unroll
                    for (int fr = 0; fr < 3; ++fr) {
                        const uint64_t a_desc = MakeSmemDesc(a_buf + fr * kC3PW * 128);
                        const uint64_t b_desc = b_base + (uint64_t)(fr * (kC3PairHalfRows * 128 / 16));
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            if (ks < p.ksteps) UmmaSS2Fp8(d_addr, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc, (fr | ks) ? 1u : 0u);
                    }
                    UmmaCommit2(&patch_empty[buf]);
                    UmmaCommit2(&tmem_full[acc]);
                }
                __syncwarp();
            }
        }
    } else {
        // =========================================================== epilogue: warp -> TMEM lane quarter e = warp & 3, channel half h
        // (the 4-warp epilogue of the single-CTA kernel needs ~0.6 us per tile and would hide the pair's faster MMAs)
        const int e = warp & 3, h = warp >> 2;
        OutT* out = reinterpret_cast<OutT*>(p.out);
        float sc[16], bi[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            sc[q] = h * 16 + q < p.Cout ? p.out_scale[h * 16 + q] : 0.f;
            bi[q] = (p.bias && h * 16 + q < p.Cout) ? p.bias[h * 16 + q] : 0.f;
        }
        const int mrow = e * 32 + lane;
        const int y = mrow / kC3PW, x = mrow - y * kC3PW;
        const uint32_t leader_empty = MapaShared(SmemAddr(tmem_empty), 0);
        GridDepWait();
        for (int k = 0; k < steps; ++k) {
            const uint32_t acc = (uint32_t)k % kC3Acc, acc_phase = ((uint32_t)k / kC3Acc) & 1u;
            const int tile = (int)blockIdx.x + k * (int)gridDim.x;
            const bool live = tile < p.num_tiles;
            const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, img = tile / (p.tiles_x * p.tiles_y);
            MbarWait(&tmem_full[acc], acc_phase);
            TcFenceAfter();
            uint32_t r0[16], r1[16], r2[16];
            const uint32_t t_addr = tmem_base + ((uint32_t)(e * 32) << 16) + acc * kC3AccStride + (uint32_t)h * 16;
            TmemLoad16(t_addr, r0);
            TmemLoad16(t_addr + 32, r1);
            TmemLoad16(t_addr + 64, r2);
            TmemLoadWait();
            TcFenceBefore();
            __syncwarp();
            if (lane == 0) MbarArriveCluster(leader_empty + acc * 8);  // one arrive per warp, on the LEADER's barrier
            // out[m] = D[m][fs = 0] + D[m + 1][fs = 1] + D[m + 2][fs = 2]
            float v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const float a1 = __shfl_down_sync(0xffffffffu, __uint_as_float(r1[c]), 1);
                const float a2 = __shfl_down_sync(0xffffffffu, __uint_as_float(r2[c]), 2);
                v[c] = (__uint_as_float(r0[c]) + a1) + a2;
            }
            const int oy = ty * p.TH + y, ox = tx * p.TW + x;
            if (live && y < p.TH && x < p.TW && oy < p.H && ox < p.W) {
                uint32_t w[4];
#pragma unroll
                for (int c = 0; c < 16; c += 4) {
                    const float2 a = Fma2(make_float2(v[c], v[c + 1]), make_float2(sc[c], sc[c + 1]), make_float2(bi[c], bi[c + 1]));
                    const float2 b = Fma2(make_float2(v[c + 2], v[c + 3]), make_float2(sc[c + 2], sc[c + 3]), make_float2(bi[c + 2], bi[c + 3]));
                    w[c / 4] = p.post_relu ? (CvtE4m3x2<true>(a.x, a.y) | (CvtE4m3x2<true>(b.x, b.y) << 16))
                                           : (CvtE4m3x2<false>(a.x, a.y) | (CvtE4m3x2<false>(b.x, b.y) << 16));
                }
                OutT* orow = out + ((size_t)(img * p.H + oy) * p.W + ox) * p.out_pitch + p.out_coff + h * 16;
                *reinterpret_cast<uint4*>(orow) = make_uint4(w[0], w[1], w[2], w[3]);
            }
            __syncwarp();
        }
    }
    TcFenceBefore();
    __syncthreads();
    ClusterSync();   // neither CTA may release tensor memory (or exit with remote arrives pending) before both are done
    if (wrole == 1) {
        TcFenceAfter();
        TmemDealloc2(tmem_base, kC3Acc * kC3AccStride);
    }
}

cudaError_t LaunchC3Pair(const CUtensorMap& tw16, const CUtensorMap& tin, const C3Params& p, cudaStream_t stream) {
    auto kern = conv3x3_pair_kernel;
    static int sm_count[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!sm_count[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kC3PairSmemBytes);
        if (e != cudaSuccess) return e;
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sm_count[dev] = n > 0 ? n : 148;
    }
    // a persistent kernel must be fully resident: not every TPC of the chip has both SMs enabled, so ask how many pairs fit
    static int max_pairs[64] = {0};
    if (!max_pairs[dev]) {
        cudaLaunchConfig_t q = {};
        q.gridDim = dim3((unsigned)(sm_count[dev] & ~1));
        q.blockDim = dim3((unsigned)kC3Threads);
        q.dynamicSmemBytes = kC3PairSmemBytes;
        cudaLaunchAttribute qa[1];
        qa[0].id = cudaLaunchAttributeClusterDimension;
        qa[0].val.clusterDim.x = 2; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
        q.attrs = qa; q.numAttrs = 1;
        int n = 0;
        if (cudaOccupancyMaxActiveClusters(&n, kern, &q) != cudaSuccess || n <= 0) n = sm_count[dev] / 2;
        max_pairs[dev] = n;
        if (getenv("B200_ENGINE_VERBOSE")) fprintf(stderr, "conv3x3 pair kernel: %d CTA pairs resident on %d SMs\n", n, sm_count[dev]);
    }
    int grid = (p.num_tiles + 1) & ~1;
    if (grid > 2 * max_pairs[dev]) grid = 2 * max_pairs[dev];
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)kC3PairThreads);
    cfg.dynamicSmemBytes = kC3PairSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    attr[1].id = cudaLaunchAttributeClusterDimension;
    attr[1].val.clusterDim.x = 2;
    attr[1].val.clusterDim.y = 1;
    attr[1].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    cudaError_t le = cudaLaunchKernelEx(&cfg, kern, tw16, tin, p);
    CountLaunch();
    return le;
}

}  // namespace

bool Conv3x3TmaSupported(const ConvArgs& a) {
    const DType it = a.in.dtype, ot = a.out.dtype;
    if (it != ot || (it != DType::BF16 && it != DType::FP8)) return false;
    if (!(a.R == 3 && a.S == 3 && a.stride == 1 && a.pad == 1) || a.pre_scale || a.pool2 || a.stem_nchw) return false;
    const int esz = (int)DTypeSize(it);
    const int step_k = 32 / esz;
    if (a.Cout > 32 || a.Cout % (16 / esz) != 0) return false;
    // the patch box is 128 bytes of channels per K plane: e4m3 takes Cin <= 128 (one plane, partial K steps allowed),
    // bf16 exactly Cin == 128 (two planes)
    if (a.Cin % step_k != 0 || a.Cin > 128) return false;
    if (esz == 2 && a.Cin != 128) return false;
    if ((a.in.pitch * esz) % 16 != 0 || (a.in.c_off * esz) % 16 != 0 || (a.out.pitch * esz) % 16 != 0 || (a.out.c_off * esz) % 16 != 0) return false;
    if (reinterpret_cast<uintptr_t>(a.in.base) % 16 != 0) return false;
    // the box always reads 128 bytes of channels: it must stay inside the pixel
    const int box_ch = 128 / esz;
    const int planes = a.Cin * esz > 128 ? 2 : 1;
    if (a.in.c_off + planes * box_ch > a.in.pitch) return false;
    return a.in.H >= 1 && a.in.W >= 1 && a.in.H == a.out.H && a.in.W == a.out.W;
}

cudaError_t Conv3x3Tma(const ConvArgs& a, const UmmaWeights& w, cudaStream_t stream) {
    if (!Conv3x3TmaSupported(a) || !w.tensor_map) return cudaErrorInvalidValue;
    if (a.n <= 0) return cudaSuccess;
    const DType it = a.in.dtype;
    const int esz = (int)DTypeSize(it);
    C3Params p;
    p.out = a.out.base; p.out_scale = w.out_scale; p.bias = a.bias; p.post_relu = a.post_relu;
    p.H = a.in.H; p.W = a.in.W; p.out_pitch = a.out.pitch; p.out_coff = a.out.c_off; p.Cout = a.Cout; p.in_coff = a.in.c_off;
    p.n = a.n;
    p.TW = p.W % 14 == 0 ? 14 : (p.W < 14 ? p.W : (p.W % 13 == 0 ? 13 : (p.W % 12 == 0 ? 12 : 14)));
    p.TH = p.H % 8 == 0 ? 8 : (p.H % 7 == 0 ? 7 : (p.H < 8 ? p.H : 8));
    p.tiles_x = (p.W + p.TW - 1) / p.TW;
    p.tiles_y = (p.H + p.TH - 1) / p.TH;
    p.num_tiles = a.n * p.tiles_x * p.tiles_y;
    const int plane_elems = 128 / esz;
    const int last = a.Cin % plane_elems;
    p.ksteps = last == 0 ? 4 : last * esz / 32;
    TensorMap tin;
    const uint64_t dims[4] = {(uint64_t)a.in.pitch, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)a.n};
    const uint64_t strides[3] = {(uint64_t)a.in.pitch * esz, (uint64_t)p.W * a.in.pitch * esz, (uint64_t)p.H * p.W * a.in.pitch * esz};
    const uint32_t box[4] = {(uint32_t)plane_elems, (uint32_t)kC3PW, (uint32_t)(p.TH + 2), 1u};
    if (MakeTensorMap(&tin, a.in.base, esz, 4, dims, strides, box, true) != 0) return cudaErrorInvalidValue;
    const CUtensorMap& tw = *reinterpret_cast<const CUtensorMap*>(w.tensor_map);
    const CUtensorMap& ti = *reinterpret_cast<const CUtensorMap*>(&tin);
    if (it == DType::BF16) return LaunchC3<__nv_bfloat16, __nv_bfloat16>(tw, ti, p, stream);
    // CTA pairs (cta_group::2): B200_ENGINE_C3PAIR=1; needs whole 32-channel outputs and at least one tile per CTA of a pair
    static const bool pair_enabled = [] { const char* e = getenv("B200_ENGINE_C3PAIR"); return e && e[0] == '1'; }();
    if (pair_enabled && a.Cout == 32 && p.num_tiles >= 2 && w.w && w.Cout_pad == 32) {
        TensorMap tw16;
        const uint64_t wdims[2] = {(uint64_t)w.K_pad, 32};
        const uint64_t wstrides[1] = {(uint64_t)w.K_pad};
        const uint32_t wbox[2] = {128u, 16u};
        if (MakeTensorMap(&tw16, w.w, 1, 2, wdims, wstrides, wbox, true) != 0) return cudaErrorInvalidValue;
        return LaunchC3Pair(*reinterpret_cast<const CUtensorMap*>(&tw16), ti, p, stream);
    }
    return LaunchC3<__nv_fp8_e4m3, __nv_fp8_e4m3>(tw, ti, p, stream);
}

}  // namespace kernels
}  // namespace b200
