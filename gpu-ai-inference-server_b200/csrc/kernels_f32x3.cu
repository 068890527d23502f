// kernels_f32x3.cu — FP32 reference mode on the tcgen05 tensor cores: every fp32 operand is split into two bf16 terms,
//     a = a0 + a1,   a0 = bf16_rn(a),  a1 = bf16_rn(a - a0)          (16 significant bits together)
// and a product is evaluated as three bf16 MMAs with fp32 accumulation in tensor memory,
//     a*w  ~=  a0*w0 + a0*w1 + a1*w0                                    (dropped terms: a1*w1 and the residuals, ~2^-17 relative)
// Activations stay plain fp32 NHWC in HBM (the arena, the memory-bound kernels and the exact FFMA fallback are unchanged):
// a 128-byte TMA row = 32 fp32 channels, and the transform warps rewrite it IN PLACE as the 128-byte K-major row
// [a0 of the 32 channels | a1 of the 32 channels] (64 bf16), after applying this layer's folded BatchNorm+ReLU in fp32.
// Weights are packed once at load time in the same paired layout ([w0 x32 | w1 x32] per 32 input channels), so the three
// products are three (A K-step, B K-step) pairings of ONE A tile and ONE B tile:
//     (a0,w0): A steps 0,1 x B steps 0,1    (a0,w1): A steps 0,1 x B steps 2,3    (a1,w0): A steps 2,3 x B steps 0,1
// Against kind::tf32 with hi/lo splits (3xTF32) this needs half the shared memory per channel (4 B instead of 8 B) and half
// the MMA issue slots (bf16 K = 16 per instruction at twice the tf32 rate).  Measured/simulated error of the whole DenseNet-121
// forward: 4e-5 of max|logit| (tools/sim_precision.py bf16x3) against the 1e-3 gate; plain TF32 gives 1.4e-3.
//
// Replaces, for FP32 mode, the Conv nodes ONNX Runtime executes inside `Ort::Session::Run`
// (reference inference_engine/src/model.cpp:1264-1270; CUDA EP options :884-892).
#include <cstdlib>

#include "kernels.h"
#include "umma_ptx.cuh"

namespace b200 {
namespace kernels {

namespace {

// (lo, hi) fp32 -> packed bf16x2 leading terms (w0) and packed bf16x2 residual terms (w1); element `lo` in the low half
__device__ __forceinline__ void SplitPair(float lo, float hi, uint32_t& w0, uint32_t& w1) {
    w0 = CvtBf16x2<false>(lo, hi);
    const float r_lo = lo - __uint_as_float(w0 << 16);
    const float r_hi = hi - __uint_as_float(w0 & 0xFFFF0000u);
    w1 = CvtBf16x2<false>(r_lo, r_hi);
}

// The six (A K-step, B K-step) pairings of one 32-channel chunk, smallest terms first.
__device__ __forceinline__ void IssueSplitChunk(uint32_t d_addr, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool first) {
    UmmaSS<0>(d_addr, a_desc + 4, b_desc + 0, idesc, first ? 0u : 1u);  // a1 * w0
    UmmaSS<0>(d_addr, a_desc + 6, b_desc + 2, idesc, 1u);
    UmmaSS<0>(d_addr, a_desc + 0, b_desc + 4, idesc, 1u);              // a0 * w1
    UmmaSS<0>(d_addr, a_desc + 2, b_desc + 6, idesc, 1u);
    UmmaSS<0>(d_addr, a_desc + 0, b_desc + 0, idesc, 1u);              // a0 * w0
    UmmaSS<0>(d_addr, a_desc + 2, b_desc + 2, idesc, 1u);
}

// One half-row (16 fp32 channels = four 16-byte pieces) of a SWIZZLE_128B tile: optional y = relu?(x*scale+shift), split,
// written back as two a0 pieces and two a1 pieces of the same row.  `half` selects channels [16*half, 16*half+16).
// All 32 lanes of the warp execute this together (lanes l and l^16 share a row: loads of both complete before either stores).
template <bool PRE>
__device__ __forceinline__ void SplitHalfRow(uint32_t row_addr, int row, int half, uint32_t sc_addr, uint32_t sh_addr, bool relu) {
    float4 v[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = LdsF4(row_addr + ((((4 * half + i) ^ (row & 7))) << 4));
    if (PRE) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float4 s = LdsF4(sc_addr + (16 * half + 4 * i) * 4), t = LdsF4(sh_addr + (16 * half + 4 * i) * 4);
            v[i].x = fmaf(v[i].x, s.x, t.x); v[i].y = fmaf(v[i].y, s.y, t.y);
            v[i].z = fmaf(v[i].z, s.z, t.z); v[i].w = fmaf(v[i].w, s.w, t.w);
            if (relu) {
                v[i].x = fmaxf(v[i].x, 0.f); v[i].y = fmaxf(v[i].y, 0.f);
                v[i].z = fmaxf(v[i].z, 0.f); v[i].w = fmaxf(v[i].w, 0.f);
            }
        }
    }
    uint32_t w0[8], w1[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        SplitPair(v[i].x, v[i].y, w0[2 * i], w1[2 * i]);
        SplitPair(v[i].z, v[i].w, w0[2 * i + 1], w1[2 * i + 1]);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        StsV4(row_addr + ((((2 * half + j) ^ (row & 7))) << 4), make_uint4(w0[4 * j], w0[4 * j + 1], w0[4 * j + 2], w0[4 * j + 3]));
        StsV4(row_addr + ((((4 + 2 * half + j) ^ (row & 7))) << 4), make_uint4(w1[4 * j], w1[4 * j + 1], w1[4 * j + 2], w1[4 * j + 3]));
    }
}

// =====================================================================================================================
// 1x1 convolution:  D[M pixels][Cout] = relu?(bn(X[M][Cin])) * W^T,  out = relu?(D * scale + bias), fp32 in / fp32 out
// =====================================================================================================================
constexpr int kFsThreads = 32 * 18;
constexpr int kFsXfWarps = 8, kFsEpiWarps = 8;
constexpr int kFsCH = 32;  // fp32 channels per 128-byte chunk
constexpr int kFsMaxC = 1024;
constexpr int kFsSmemBudget = 226 * 1024;

template <int BN> struct FsCfg {
    static constexpr int kABytes = kATileBytes;
    static constexpr int kBBytes = BN * kRowBytes;
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kSlabs = BN / 32;                         // one 128-byte slab = 32 fp32 output channels
    static constexpr int kStagingBytes = kSlabs * kTileM * 128;
    static constexpr int kVecBytes = 4 * kFsMaxC * 4;
    static constexpr int kFixedBytes = 1024 + kStagingBytes + kVecBytes + 512;
    static constexpr int kStagesFit = (kFsSmemBudget - kFixedBytes) / kStageBytes;
    static constexpr int kStages = kStagesFit > 6 ? 6 : kStagesFit;
    static constexpr int kSmemBytes = kFixedBytes + kStages * kStageBytes;
    static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
    static_assert(kStages >= 3, "pipeline too shallow");
};

struct FsL1Params {
    const float* pre_scale;
    const float* pre_shift;
    const float* out_scale;
    const float* bias;
    int pre_relu, post_relu;
    int in_coff, out_coff;
    int Cin, Cout, M;
    int num_m_tiles, num_n_tiles, num_chunks;
    float out_scale_mul;
};

template <int BN>
__global__ void __launch_bounds__(kFsThreads, 1)
conv1x1_f32x3_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_in,
                     const __grid_constant__ CUtensorMap tmap_out, const FsL1Params p) {
    using Cfg = FsCfg<BN>;
    constexpr int NS = Cfg::kStages;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_staging = smem + NS * Cfg::kStageBytes;
    float* s_pre_scale = reinterpret_cast<float*>(s_staging + Cfg::kStagingBytes);
    float* s_pre_shift = s_pre_scale + kFsMaxC;
    float* s_out_scale = s_pre_shift + kFsMaxC;
    float* s_bias = s_out_scale + kFsMaxC;
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(s_bias + kFsMaxC);
    uint64_t* xf_full = raw_full + NS;
    uint64_t* empty_bar = xf_full + NS;
    uint64_t* tmem_full = empty_bar + NS;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the two single-thread roles sit on the highest warp ids (the issue arbiter prefers them; see kernels_conv1x1.cu)
    constexpr int kNW = kFsThreads / 32;
    const int wrole = warp >= kNW - 2 ? warp - (kNW - 2) : warp + 2;  // 0 TMA, 1 MMA, 2.. transform, then epilogue
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const bool has_pre = p.pre_scale != nullptr;

    if (wrole == 0 && lane == 0) {
        for (int s = 0; s < NS; ++s) {
            MbarInit(&raw_full[s], 1);
            MbarInit(&xf_full[s], kFsXfWarps);
            MbarInit(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            MbarInit(&tmem_full[a], 1);
            MbarInit(&tmem_empty[a], kFsEpiWarps);
        }
        FenceBarrierInit();
        PrefetchTensorMap(&tmap_w);
        PrefetchTensorMap(&tmap_in);
        PrefetchTensorMap(&tmap_out);
    }
    if (wrole == 1) TmemAlloc(tmem_slot, Cfg::kTmemCols);
    if (has_pre) {
        for (int i = threadIdx.x; i < p.Cin; i += kFsThreads) {
            s_pre_scale[i] = p.pre_scale[i];
            s_pre_shift[i] = p.pre_shift[i];
        }
    }
    for (int i = threadIdx.x; i < p.num_n_tiles * BN; i += kFsThreads) {
        s_out_scale[i] = i < p.Cout ? p.out_scale[i] * p.out_scale_mul : 0.f;
        s_bias[i] = (p.bias && i < p.Cout) ? p.bias[i] : 0.f;
    }
    TcFenceBefore();
    __syncthreads();
    TcFenceAfter();
    const uint32_t tmem_base = *tmem_slot;
    GridDepLaunch();

    if (wrole == 0) {
        // =========================================================== TMA producer
        GridDepWait();
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_tile = tile / p.num_n_tiles, n_tile = tile - m_tile * p.num_n_tiles;
            for (int c = 0; c < p.num_chunks; ++c) {
                MbarWaitWarp(&empty_bar[stage], phase ^ 1u);
                if (ElectOne()) {
                    uint8_t* a_dst = smem + stage * Cfg::kStageBytes;
                    MbarArriveExpectTx(&raw_full[stage], (uint32_t)Cfg::kStageBytes);
                    TmaLoad2D(a_dst, &tmap_in, &raw_full[stage], p.in_coff + c * kFsCH, m_tile * kTileM);
                    TmaLoad2D(a_dst + Cfg::kABytes, &tmap_w, &raw_full[stage], c * 64, n_tile * BN);
                }
                __syncwarp();
                if (++stage == NS) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (wrole == 1) {
        // =========================================================== MMA issuer
        constexpr uint32_t idesc = MakeInstrDesc(1 /*BF16*/, BN);
        const uint64_t stage_desc = MakeSmemDesc(SmemAddr(smem));
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        int stage = 0;
        uint32_t phase = 0, tile_iter = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const uint32_t acc = tile_iter & 1u, acc_phase = (tile_iter >> 1) & 1u;
            MbarWaitWarp(&tmem_empty[acc], acc_phase ^ 1u);
            TcFenceAfter();
            const uint32_t d_addr = tmem_u + acc * BN;
            for (int c = 0; c < p.num_chunks; ++c) {
                const uint64_t a_desc = stage_desc + (uint64_t)((uint32_t)stage * (Cfg::kStageBytes >> 4));
                const uint64_t b_desc = a_desc + (uint64_t)(Cfg::kABytes >> 4);
                MbarWaitWarp(&xf_full[stage], phase);
                TcFenceAfter();
                if (ElectOne()) {
                    IssueSplitChunk(d_addr, a_desc, b_desc, idesc, c == 0);
                    UmmaCommit(&empty_bar[stage]);
                    if (c == p.num_chunks - 1) UmmaCommit(&tmem_full[acc]);
                }
                __syncwarp();
                if (++stage == NS) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (wrole < 2 + kFsXfWarps) {
        // =========================================================== transform warps: BN + ReLU + bf16 split, in place
        const int tw = wrole - 2;
        const int row = tw * 16 + (lane & 15), half = lane >> 4;
        const uint32_t smem_base = SmemAddr(smem);
        const uint32_t sc_base = SmemAddr(s_pre_scale), sh_base = SmemAddr(s_pre_shift);
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            for (int c = 0; c < p.num_chunks; ++c) {
                const uint32_t row_addr = smem_base + stage * Cfg::kStageBytes + row * kRowBytes;
                MbarWaitWarp(&raw_full[stage], phase);
                if (has_pre) SplitHalfRow<true>(row_addr, row, half, sc_base + c * kFsCH * 4, sh_base + c * kFsCH * 4, p.pre_relu != 0);
                else SplitHalfRow<false>(row_addr, row, half, 0, 0, false);
                FenceProxyAsync();
                __syncwarp();
                if (lane == 0) MbarArrive(&xf_full[stage]);
                if (++stage == NS) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // =========================================================== epilogue: TMEM -> scale/bias(+ReLU) -> fp32 staging -> TMA store
        const int ew = wrole - (2 + kFsXfWarps);
        const int q = warp & 3;             // TMEM lane quarter this warp may access
        const int h = ew >> 2;
        constexpr int kCgs = BN / 32;
        constexpr int kCgPerWarp = kCgs >= 2 ? kCgs / 2 : 1;
        const bool leader = (wrole == 2 + kFsXfWarps) && lane == 0;
        const int row = q * 32 + lane;
        if (leader) GridDepWait();
        uint32_t tile_iter = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const int m_tile = tile / p.num_n_tiles, n_tile = tile - m_tile * p.num_n_tiles;
            const uint32_t acc = tile_iter & 1u, acc_phase = (tile_iter >> 1) & 1u;
            if (leader) BulkWaitRead<0>();  // the store that last read the staging buffer is done
            NamedBarSync(1, kFsEpiWarps * 32);
            MbarWaitWarp(&tmem_full[acc], acc_phase);
            TcFenceAfter();
            if (kCgs >= 2 || h == 0) {
#pragma unroll
                for (int ci = 0; ci < kCgPerWarp; ++ci) {
                    const int cg = h * kCgPerWarp + ci;
                    uint32_t r[32];
                    TmemLoad32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + cg * 32, r);
                    TmemLoadWait();
                    const uint32_t scp = SmemAddr(s_out_scale) + (n_tile * BN + cg * 32) * 4;
                    const uint32_t bip = SmemAddr(s_bias) + (n_tile * BN + cg * 32) * 4;
                    const uint32_t slab = SmemAddr(s_staging) + cg * (kTileM * 128) + row * 128;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float4 s4 = LdsF4(scp + i * 16), b4 = LdsF4(bip + i * 16);
                        float4 y;
                        y.x = fmaf(__uint_as_float(r[4 * i]), s4.x, b4.x);
                        y.y = fmaf(__uint_as_float(r[4 * i + 1]), s4.y, b4.y);
                        y.z = fmaf(__uint_as_float(r[4 * i + 2]), s4.z, b4.z);
                        y.w = fmaf(__uint_as_float(r[4 * i + 3]), s4.w, b4.w);
                        if (p.post_relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
                        StsV4(slab + ((i ^ (row & 7)) << 4), make_uint4(__float_as_uint(y.x), __float_as_uint(y.y), __float_as_uint(y.z), __float_as_uint(y.w)));
                    }
                }
            }
            TcFenceBefore();
            __syncwarp();
            if (lane == 0) MbarArrive(&tmem_empty[acc]);
            FenceProxyAsync();
            NamedBarSync(2, kFsEpiWarps * 32);
            if (leader) {
#pragma unroll
                for (int s = 0; s < Cfg::kSlabs; ++s)
                    TmaStore2D(&tmap_out, s_staging + s * (kTileM * 128), p.out_coff + n_tile * BN + s * 32, m_tile * kTileM);
                BulkCommit();
            }
        }
        if (leader) BulkWait<0>();
    }

    TcFenceBefore();
    __syncthreads();
    if (wrole == 1) {
        TcFenceAfter();
        TmemDealloc(tmem_base, Cfg::kTmemCols);
    }
}

int SmCount() {
    static int sm_count[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!sm_count[dev]) {
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sm_count[dev] = n > 0 ? n : 148;
    }
    return sm_count[dev];
}

template <int BN>
cudaError_t LaunchFsL1(const CUtensorMap& tw, const CUtensorMap& tin, const CUtensorMap& tout, const FsL1Params& p, cudaStream_t stream) {
    using Cfg = FsCfg<BN>;
    auto kern = conv1x1_f32x3_kernel<BN>;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    const int sms = SmCount();
    const int grid = tiles < sms ? tiles : sms;
    cudaError_t le = LaunchPdl(kern, grid, kFsThreads, Cfg::kSmemBytes, stream, tw, tin, tout, p);
    CountLaunch();
    return le;
}

// =====================================================================================================================
// 3x3 / stride 1 / pad 1 convolution of Cin <= 128 (multiple of 32) fp32 channels into <= 32 channels
// =====================================================================================================================
// Same geometry as kernels_conv3x3.cu (one 4-D TMA box lands the zero-padded patch, nine row-shifted descriptors are the nine
// taps), but a pixel is Cin*4 bytes, so the K dimension is walked in 32-channel PLANES: a pipeline stage = one plane of the
// patch ((TH+2) x 16 slots x 128 B), split in place by the transform warps.  All weights stay resident (144 KB for Cin = 128).
//
// N = 32 output channels would make every MMA A-fetch bound (4 KB of A per 16 cycles of math), so both the split terms and the
// three taps of a filter ROW are stacked along N.  The resident weight tile of (filter row fr, plane pair) has 192 rows:
//     rows   0.. 95:  w0 of taps (fr, 0), (fr, 1), (fr, 2)   [fs*32 + o]
//     rows  96..191:  w1 of the same taps                      [96 + fs*32 + o]
// and per K step two MMAs run with the A view shifted by fr*16 slots (the fs shift moves into the epilogue, see kernels_conv3x3.cu):
//     a0 x rows 0..191 (N = 192)  ->  columns fs*32+o += a0*w0,  columns 96+fs*32+o += a0*w1
//     a1 x rows 0.. 95 (N =  96)  ->  columns fs*32+o += a1*w0
// i.e. 2 A fetches per (filter row, K step) instead of 9 x 3 = 27 single products.  The epilogue computes
//     out[m][o] = sum over fs of ( D[m + fs][fs*32 + o] + D[m + fs][96 + fs*32 + o] )       (warp shuffles across TMEM lanes)
// A weight tile row is 128 B = two planes of one tap ([plane 2p | plane 2p+1]).
constexpr int kF3Threads = 32 * 16;   // 10 transform + 4 epilogue + TMA + MMA
constexpr int kF3XfWarps = 10;        // one warp per patch row (TH + 2 <= 10): a plane is split in ONE pass (two passes of 8 warps
                                      // made the transform, not the MMAs, the slowest stage: ~1200 against 912 cycles per plane)
constexpr int kF3PW = 16;
constexpr int kF3PatchBytes = 10 * kF3PW * 128;
constexpr int kF3WRows = 192;
constexpr int kF3WTile = kF3WRows * 128;                   // one (filter row, plane pair) weight tile
constexpr int kF3Acc = 2;
constexpr int kF3AccCols = 256;                            // column stride between accumulators (192 used)
constexpr int kF3MaxPlanes = 4;                            // Cin <= 128
constexpr int kF3Stages = 4;
constexpr int kF3ResBytes = 3 * (kF3MaxPlanes / 2) * kF3WTile;
constexpr int kF3SmemBytes = 1024 + kF3ResBytes + kF3Stages * kF3PatchBytes + 1024 /*junk-row overreach*/ + 512;

struct FsC3Params {
    float* out;
    const float* out_scale;
    const float* bias;
    int post_relu;
    int H, W, out_pitch, out_coff, Cout, in_coff;
    int n, TH, TW, tiles_x, tiles_y, num_tiles;
    int planes;  // Cin / 32
};

__global__ void __launch_bounds__(kF3Threads, 1)
conv3x3_f32x3_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_in, const FsC3Params p) {
    constexpr int NS = kF3Stages;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* s_w = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* smem = s_w + kF3ResBytes;
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(smem + NS * kF3PatchBytes + 1024);
    uint64_t* xf_full = raw_full + NS;
    uint64_t* empty_bar = xf_full + NS;
    uint64_t* tmem_full = empty_bar + NS;
    uint64_t* tmem_empty = tmem_full + kF3Acc;
    uint64_t* w_bar = tmem_empty + kF3Acc;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
    float* s_sc = reinterpret_cast<float*>(tmem_slot + 4);  // 16-byte aligned: (3*4 + 2*4 + 1) * 8 + 16 = 184 -> 192
    s_sc = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(s_sc) + 15) & ~(uintptr_t)15);
    float* s_bi = s_sc + 32;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kNW = kF3Threads / 32;
    const int wrole = warp >= kNW - 2 ? warp - (kNW - 2) : warp + 2;  // 0 TMA, 1 MMA, 2..11 transform, 12..15 epilogue
    const int npairs = (p.planes + 1) >> 1;

    if (wrole == 0 && lane == 0) {
        for (int s = 0; s < NS; ++s) {
            MbarInit(&raw_full[s], 1);
            MbarInit(&xf_full[s], kF3XfWarps);
            MbarInit(&empty_bar[s], 1);
        }
        for (int a = 0; a < kF3Acc; ++a) {
            MbarInit(&tmem_full[a], 1);
            MbarInit(&tmem_empty[a], 4);
        }
        MbarInit(w_bar, 1);
        FenceBarrierInit();
        PrefetchTensorMap(&tmap_w);
        PrefetchTensorMap(&tmap_in);
    }
    if (wrole == 1) TmemAlloc(tmem_slot, kF3Acc * kF3AccCols);
    if (threadIdx.x < 32) {
        s_sc[threadIdx.x] = (int)threadIdx.x < p.Cout ? p.out_scale[threadIdx.x] : 0.f;
        s_bi[threadIdx.x] = (p.bias && (int)threadIdx.x < p.Cout) ? p.bias[threadIdx.x] : 0.f;
    }
    TcFenceBefore();
    __syncthreads();
    TcFenceAfter();
    const uint32_t tmem_base = *tmem_slot;
    GridDepLaunch();

    if (wrole == 0) {
        // =========================================================== TMA producer: resident weights, then (tile, plane) stages
        if (ElectOne()) {  // the weights do not depend on the previous kernel: load them before the dependency wait
            MbarArriveExpectTx(w_bar, (uint32_t)(3 * npairs * kF3WTile));
            for (int t = 0; t < 3 * npairs; ++t) TmaLoad2D(s_w + t * kF3WTile, &tmap_w, w_bar, t * 64, 0);
        }
        __syncwarp();
        GridDepWait();
        const uint32_t stage_tx = (uint32_t)((p.TH + 2) * kF3PW * 128);
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, img = tile / (p.tiles_x * p.tiles_y);
            for (int j = 0; j < p.planes; ++j) {
                MbarWaitWarp(&empty_bar[stage], phase ^ 1u);
                if (ElectOne()) {
                    MbarArriveExpectTx(&raw_full[stage], stage_tx);
                    TmaLoad4D(smem + stage * kF3PatchBytes, &tmap_in, &raw_full[stage], p.in_coff + j * kFsCH, tx * p.TW - 1, ty * p.TH - 1, img);
                }
                __syncwarp();
                if (++stage == NS) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (wrole == 1) {
        // =========================================================== MMA issuer
        constexpr uint32_t idesc192 = MakeInstrDesc(1 /*BF16*/, 192), idesc96 = MakeInstrDesc(1, 96);
        const uint32_t smem_u = SmemAddr(smem), w_u = SmemAddr(s_w);
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        MbarWaitWarp(w_bar, 0);
        int stage = 0;
        uint32_t phase = 0, k = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++k) {
            const uint32_t acc = k % kF3Acc, acc_ph = (k / kF3Acc) & 1u;
            MbarWaitWarp(&tmem_empty[acc], acc_ph ^ 1u);
            TcFenceAfter();
            const uint32_t d_addr = tmem_u + acc * kF3AccCols;
            for (int j = 0; j < p.planes; ++j) {
                MbarWaitWarp(&xf_full[stage], phase);
                TcFenceAfter();
                if (ElectOne()) {
                    const uint32_t a_buf = smem_u + stage * kF3PatchBytes;
                    const uint32_t b_plane = w_u + (j >> 1) * kF3WTile + (j & 1) * 64;  // this plane's 64-byte half of the pair tiles
#pragma unroll
                    for (int fr = 0; fr < 3; ++fr) {
                        const uint64_t a_desc = MakeSmemDesc(a_buf + fr * kF3PW * 128);
                        const uint64_t b_desc = MakeSmemDesc(b_plane + fr * npairs * kF3WTile);
                        UmmaSS<0>(d_addr, a_desc + 0, b_desc + 0, idesc192, (j | fr) ? 1u : 0u);  // a0 x [w0 ; w1] of three taps
                        UmmaSS<0>(d_addr, a_desc + 2, b_desc + 2, idesc192, 1u);
                        UmmaSS<0>(d_addr, a_desc + 4, b_desc + 0, idesc96, 1u);                   // a1 x w0 of three taps
                        UmmaSS<0>(d_addr, a_desc + 6, b_desc + 2, idesc96, 1u);
                    }
                    UmmaCommit(&empty_bar[stage]);
                    if (j == p.planes - 1) UmmaCommit(&tmem_full[acc]);
                }
                __syncwarp();
                if (++stage == NS) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (wrole < 2 + kF3XfWarps) {
        // =========================================================== transform warps: fp32 plane -> [a0 | a1] rows, in place
        const int tw = wrole - 2;
        const int half = lane >> 4;
        const uint32_t smem_u = SmemAddr(smem);
        const int units = p.TH + 2;  // one unit = one patch row = 16 slots
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
            for (int j = 0; j < p.planes; ++j) {
                MbarWaitWarp(&raw_full[stage], phase);
                for (int u = tw; u < units; u += kF3XfWarps) {
                    const int row = u * 16 + (lane & 15);
                    SplitHalfRow<false>(smem_u + stage * kF3PatchBytes + row * 128, row, half, 0, 0, false);
                }
                FenceProxyAsync();
                __syncwarp();
                if (lane == 0) MbarArrive(&xf_full[stage]);
                if (++stage == NS) { stage = 0; phase ^= 1u; }
            }
        }
    } else {
        // =========================================================== epilogue
        const int e = warp & 3;  // TMEM lane quarter
        const uint32_t sc_addr = SmemAddr(s_sc), bi_addr = SmemAddr(s_bi);
        const int mrow = e * 32 + lane;
        const int y = mrow / kF3PW, x = mrow - y * kF3PW;
        GridDepWait();  // stores may alias buffers the previous kernel still reads
        uint32_t k = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++k) {
            const uint32_t acc = k % kF3Acc, acc_phase = (k / kF3Acc) & 1u;
            const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, img = tile / (p.tiles_x * p.tiles_y);
            MbarWaitWarp(&tmem_full[acc], acc_phase);
            TcFenceAfter();
            float r[32];
            {
                const uint32_t t_addr = tmem_base + ((uint32_t)(e * 32) << 16) + acc * kF3AccCols;
#pragma unroll
                for (int fs = 0; fs < 3; ++fs) {   // out[m] += (D[m + fs][fs*32 + o] + D[m + fs][96 + fs*32 + o])
                    uint32_t t0[32], t1[32];
                    TmemLoad32(t_addr + fs * 32, t0);
                    TmemLoad32(t_addr + 96 + fs * 32, t1);
                    TmemLoadWait();
#pragma unroll
                    for (int c = 0; c < 32; ++c) {
                        float v = __uint_as_float(t0[c]) + __uint_as_float(t1[c]);
                        if (fs) v = __shfl_down_sync(0xffffffffu, v, fs);
                        r[c] = fs ? r[c] + v : v;
                    }
                }
            }
            TcFenceBefore();
            __syncwarp();
            if (lane == 0) MbarArrive(&tmem_empty[acc]);  // the accumulator is in registers: release it before the stores
            const int oy = ty * p.TH + y, ox = tx * p.TW + x;
            if (y < p.TH && x < p.TW && oy < p.H && ox < p.W) {
                float* orow = p.out + ((size_t)(img * p.H + oy) * p.W + ox) * p.out_pitch + p.out_coff;
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    if (q * 4 < p.Cout) {
                        const float4 s4 = LdsF4(sc_addr + q * 16), b4 = LdsF4(bi_addr + q * 16);
                        float4 v;
                        v.x = fmaf(r[4 * q], s4.x, b4.x);
                        v.y = fmaf(r[4 * q + 1], s4.y, b4.y);
                        v.z = fmaf(r[4 * q + 2], s4.z, b4.z);
                        v.w = fmaf(r[4 * q + 3], s4.w, b4.w);
                        if (p.post_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                        *reinterpret_cast<float4*>(orow + q * 4) = v;
                    }
                }
            }
            __syncwarp();
        }
    }
    TcFenceBefore();
    __syncthreads();
    if (wrole == 1) {
        TcFenceAfter();
        TmemDealloc(tmem_base, kF3Acc * kF3AccCols);
    }
}

bool Is1x1(const ConvArgs& a) { return a.R == 1 && a.S == 1 && a.stride == 1 && a.pad == 0 && a.in.H == a.out.H && a.in.W == a.out.W; }
bool Is3x3(const ConvArgs& a) { return a.R == 3 && a.S == 3 && a.stride == 1 && a.pad == 1 && a.in.H == a.out.H && a.in.W == a.out.W; }

}  // namespace

bool ConvF32x3Supported(const ConvArgs& a) {
    if (a.in.dtype != DType::F32 || a.out.dtype != DType::F32 || a.pool2 || a.stem_nchw) return false;
    if (a.Cin % 32 != 0 || a.Cin < 32) return false;
    if (a.in.pitch % 4 != 0 || a.in.c_off % 4 != 0 || a.out.pitch % 4 != 0 || a.out.c_off % 4 != 0) return false;
    if (reinterpret_cast<uintptr_t>(a.in.base) % 16 != 0 || reinterpret_cast<uintptr_t>(a.out.base) % 16 != 0) return false;
    if (a.in.c_off + a.Cin > a.in.pitch) return false;
    if (Is1x1(a)) return a.Cin <= kFsMaxC && a.Cout % 32 == 0 && a.Cout >= 32 && a.Cout <= kFsMaxC;
    if (Is3x3(a)) return !a.pre_scale && a.Cin <= 32 * kF3MaxPlanes && a.Cout <= 32 && a.Cout % 4 == 0 && a.Cout >= 4 && a.in.H >= 1 && a.in.W >= 1;
    return false;
}

int F32x3TileN(const ConvArgs& a) {
    if (Is3x3(a)) return kF3WRows;  // [w0 ; w1] x three taps stacked along N
    return a.Cout % 128 == 0 ? 128 : a.Cout % 64 == 0 ? 64 : 32;
}
int F32x3PackedRows(const ConvArgs& a) {
    if (Is3x3(a)) return kF3WRows;
    const int bn = F32x3TileN(a);
    return (a.Cout + bn - 1) / bn * bn;
}
int F32x3PackedK(const ConvArgs& a) {
    if (Is3x3(a)) return 3 * ((a.Cin / 32 + 1) / 2) * 64;
    return a.R * a.S * a.Cin * 2;
}

// (row, bf16 column) of split term `term` (0: w0 = bf16(w), 1: w1 = bf16(w - w0)) of weight (output o, filter tap, input c).
// 1x1: row o, the two terms 32 columns apart inside the 64-column group of the channel's 32-channel chunk.
// 3x3: row term*96 + fs*32 + o, column group (filter row, plane pair), 32 columns per plane of the pair.
void F32x3WeightPos(const ConvArgs& a, int o, int tap, int c, int term, int* row, int* col) {
    if (Is3x3(a)) {
        const int plane = c / 32, npairs = (a.Cin / 32 + 1) / 2, fr = tap / 3, fs = tap % 3;
        *row = term * 96 + fs * 32 + o;
        *col = (fr * npairs + plane / 2) * 64 + (plane & 1) * 32 + (c % 32);
        return;
    }
    *row = o;
    *col = (tap * (a.Cin / 32) + c / 32) * 64 + term * 32 + (c % 32);
}

cudaError_t ConvF32x3(const ConvArgs& a, const UmmaWeights& w, cudaStream_t stream) {
    if (!ConvF32x3Supported(a) || !w.tensor_map) return cudaErrorInvalidValue;
    if (a.n <= 0) return cudaSuccess;
    const CUtensorMap& tw = *reinterpret_cast<const CUtensorMap*>(w.tensor_map);
    if (Is1x1(a)) {
        FsL1Params p;
        p.pre_scale = a.pre_scale; p.pre_shift = a.pre_shift; p.out_scale = w.out_scale; p.bias = a.bias;
        p.pre_relu = a.pre_relu; p.post_relu = a.post_relu;
        p.in_coff = a.in.c_off; p.out_coff = a.out.c_off;
        p.Cin = a.Cin; p.Cout = a.Cout;
        p.M = a.n * a.out.H * a.out.W;
        const int bn = F32x3TileN(a);
        p.num_n_tiles = a.Cout / bn;
        p.num_chunks = a.Cin / kFsCH;
        p.num_m_tiles = (p.M + kTileM - 1) / kTileM;
        p.out_scale_mul = a.out_mul;
        TensorMap tin, tout;
        {
            const uint64_t dims[2] = {(uint64_t)a.in.pitch, (uint64_t)p.M};
            const uint64_t strides[1] = {(uint64_t)a.in.pitch * 4};
            const uint32_t box[2] = {(uint32_t)kFsCH, (uint32_t)kTileM};
            if (MakeTensorMap(&tin, a.in.base, 4, 2, dims, strides, box, true) != 0) return cudaErrorInvalidValue;
        }
        {
            const uint64_t dims[2] = {(uint64_t)a.out.pitch, (uint64_t)p.M};
            const uint64_t strides[1] = {(uint64_t)a.out.pitch * 4};
            const uint32_t box[2] = {32u, (uint32_t)kTileM};
            if (MakeTensorMap(&tout, a.out.base, 4, 2, dims, strides, box, true) != 0) return cudaErrorInvalidValue;
        }
        const CUtensorMap& ti = *reinterpret_cast<const CUtensorMap*>(&tin);
        const CUtensorMap& to = *reinterpret_cast<const CUtensorMap*>(&tout);
        if (bn == 128) return LaunchFsL1<128>(tw, ti, to, p, stream);
        if (bn == 64) return LaunchFsL1<64>(tw, ti, to, p, stream);
        return LaunchFsL1<32>(tw, ti, to, p, stream);
    }
    FsC3Params p;
    p.out = (float*)a.out.base; p.out_scale = w.out_scale; p.bias = a.bias; p.post_relu = a.post_relu;
    p.H = a.in.H; p.W = a.in.W; p.out_pitch = a.out.pitch; p.out_coff = a.out.c_off; p.Cout = a.Cout; p.in_coff = a.in.c_off;
    p.n = a.n;
    p.TW = p.W % 14 == 0 ? 14 : (p.W < 14 ? p.W : (p.W % 13 == 0 ? 13 : (p.W % 12 == 0 ? 12 : 14)));
    p.TH = p.H % 8 == 0 ? 8 : (p.H % 7 == 0 ? 7 : (p.H < 8 ? p.H : 8));
    p.tiles_x = (p.W + p.TW - 1) / p.TW;
    p.tiles_y = (p.H + p.TH - 1) / p.TH;
    p.num_tiles = a.n * p.tiles_x * p.tiles_y;
    p.planes = a.Cin / kFsCH;
    TensorMap tin;
    const uint64_t dims[4] = {(uint64_t)a.in.pitch, (uint64_t)p.W, (uint64_t)p.H, (uint64_t)a.n};
    const uint64_t strides[3] = {(uint64_t)a.in.pitch * 4, (uint64_t)p.W * a.in.pitch * 4, (uint64_t)p.H * p.W * a.in.pitch * 4};
    const uint32_t box[4] = {(uint32_t)kFsCH, (uint32_t)kF3PW, (uint32_t)(p.TH + 2), 1u};
    if (MakeTensorMap(&tin, a.in.base, 4, 4, dims, strides, box, true) != 0) return cudaErrorInvalidValue;
    const CUtensorMap& ti = *reinterpret_cast<const CUtensorMap*>(&tin);
    const int sms = SmCount();
    const int grid = p.num_tiles < sms ? p.num_tiles : sms;
    static bool attr_set[64] = {false};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(conv3x3_f32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kF3SmemBytes);
        if (e != cudaSuccess) return e;
        attr_set[dev] = true;
    }
    cudaError_t le = LaunchPdl(conv3x3_f32x3_kernel, grid, kF3Threads, kF3SmemBytes, stream, tw, ti, p);
    CountLaunch();
    return le;
}

}  // namespace kernels
}  // namespace b200
