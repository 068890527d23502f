// kernels_conv1x1.cu — 1x1 convolution (pointwise GEMM) with a fused per-input-channel BatchNorm+ReLU prologue and a
// per-output-channel scale/bias(+ReLU) epilogue: the first conv of every DenseNet dense layer,
//     D[M = n*H*W pixels][Cout] = relu(bn1(X[M][Cin])) * W[Cout][Cin]^T,   out = relu?(D * scale + bias)
// (reference: the Conv/BatchNormalization/Relu nodes ONNX Runtime executes inside `Ort::Session::Run`,
// inference_engine/src/model.cpp:1264-1270).
//
// Data never passes through the LSU on its way in or out:
//   warp 0       TMA producer: A box [128 pixels][128 B of channels] straight out of the NHWC block buffer (the
//                channel slice of the concat-in-place layout is just a box coordinate) + the weight box, SWIZZLE_128B.
//   warps 2-9    transform: the folded BatchNorm + ReLU of THIS layer applied in place on the landed A tile
//                (ld.shared.v4 -> packed f16x2/bf16x2 FMA.relu -> st.shared.v4, conflict-free under the swizzle).
//   warp 1       MMA issuer (tcgen05.mma M=128, N=BN, fp32 accumulate in TMEM, two accumulators).
//   warps 10-17  epilogue: tcgen05.ld -> packed FFMA2 scale/bias -> cvt(.relu) -> swizzled staging tile -> one TMA
//                store per 128-byte column slab (rows past M are clipped by the tensor map).
// K tail: when Cin is not a multiple of the 128-byte chunk the last box is placed at channel Cin - chunk, i.e. it
// OVERLAPS the previous chunk (an L2 hit, no extra HBM bytes) and only its last K steps are multiplied.
#include <cstdlib>

#include "kernels.h"
#include "umma_ptx.cuh"

namespace b200 {
namespace kernels {

namespace {

constexpr int kL1Threads = 32 * 18;
constexpr int kL1XfWarps = 8, kL1EpiWarps = 8;
constexpr int kL1MaxCin = 1024;
constexpr int kL1MaxCout = 1024;
constexpr int kL1ResChunks = 4;  // resident-weight mode: up to 4 K chunks (512 e4m3 / 256 bf16 input channels)
constexpr int kL1SmemBudget = 226 * 1024;

// RESB: the whole weight matrix (one N tile, <= kL1ResChunks K chunks) is TMA-loaded once per CTA - before the
// programmatic-dependency wait, so it overlaps the previous kernel - and the pipeline stages carry only A tiles:
// half the L2->SM traffic per tile and twice the pipeline depth in the same shared memory.
// POOL (transition layers): the 2x2 average pool that follows the 1x1 conv is commuted in front of it (both are linear);
// a stage then carries FOUR raw planes - the (dy, dx) pixels of every output pixel, each landed by its own 5-D TMA box -
// and the transform warps write sum(relu(bn(x))) over the four planes into plane 0, the A operand.
template <int BN, int OUT_ESZ, bool RESB, bool POOL> struct L1Cfg {
    static constexpr int kAPlanes = POOL ? 4 : 1;
    static constexpr int kABytes = kAPlanes * kATileBytes;
    static constexpr int kStageBytes = kABytes + (RESB ? 0 : BN * kRowBytes);
    static constexpr int kWBytes = RESB ? kL1ResChunks * BN * kRowBytes : 0;
    static constexpr int kSlabs = BN * OUT_ESZ / 128;            // 128-byte column slabs of the output tile
    static constexpr int kStagingBytes = kSlabs * kTileM * 128;  // one output tile
    static constexpr int kVecBytes = kL1MaxCin * 2 * 2 + kL1MaxCout * 4 * 2;  // packed prologue pairs + fp32 scale/bias
    static constexpr int kStagingBufs = (POOL && OUT_ESZ == 2) ? 1 : 2;  // output tiles in flight towards the TMA store
    static constexpr int kFixedBytes = 1024 + kWBytes + kStagingBufs * kStagingBytes + kVecBytes + 512;
    static constexpr int kStagesFit = (kL1SmemBudget - kFixedBytes) / kStageBytes;
    static constexpr int kStages = kStagesFit > 8 ? 8 : kStagesFit;
    static constexpr int kSmemBytes = kFixedBytes + kStages * kStageBytes;
    static constexpr int kTmemCols = 2 * BN < 32 ? 32 : 2 * BN;
    static_assert(kStages >= 2, "pipeline too shallow");
};

struct L1Params {
    const float* pre_scale;
    const float* pre_shift;
    const float* out_scale;
    const float* bias;
    int pre_relu, post_relu;
    int in_coff, out_coff;  // element offsets of the channel slices
    int Cin, Cout, M;
    int num_m_tiles, num_n_tiles, num_chunks;
    int tile_rows;        // output pixels per M tile (128, or k*Wo whole output rows in POOL mode)
    int pool_k;           // POOL: output rows per M tile
    float out_scale_mul;  // POOL: 0.25 (the average), folded into the epilogue scale
};

// Epilogue constants of one 128-column N tile as a kernel parameter (CSTP): with a compile-time column group the scale / bias become
// constant-bank operands of the FMAs instead of 16-byte broadcast loads from shared memory - those cost 2 wavefronts each, 512 per
// tile, on a kernel whose shared-memory port is the busiest resource (measured: -36 us per forward in block 1 with immediates).
struct alignas(16) L1Consts {
    float s[128];
    float b[128];
};

// Chunk geometry shared by the producer, the transform warps and the MMA issuer.
struct ChunkGeom {
    int ch_base;   // first channel of the box
    int k_lo, k_hi;  // valid element range inside the box
};
template <int CH> __device__ __forceinline__ ChunkGeom GeomOf(int c, int Cin) {
    ChunkGeom g;
    g.ch_base = c * CH;
    g.k_lo = 0;
    g.k_hi = CH;
    if (g.ch_base + CH > Cin) {
        if (Cin >= CH) {  // overlapped tail: box ends exactly at Cin
            g.k_lo = g.ch_base + CH - Cin;
            g.ch_base = Cin - CH;
        } else {          // single short chunk
            g.k_hi = Cin;
        }
    }
    return g;
}

template <typename MmaT, typename OutT, int BN, bool RESB, bool POOL, bool CSTP = false>
__global__ void __launch_bounds__(kL1Threads, 1)
conv1x1_tma_kernel(const __grid_constant__ CUtensorMap tmap_w, const __grid_constant__ CUtensorMap tmap_in,
                   const __grid_constant__ CUtensorMap tmap_out, const L1Params p, const __grid_constant__ L1Consts cst) {
    using ME = MmaElem<MmaT>;
    using Cfg = L1Cfg<BN, (int)sizeof(OutT), RESB, POOL>;
    constexpr int NS = Cfg::kStages;
    constexpr int CH = ME::kChunk;
    constexpr int EPV = ME::kPerVec;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w = smem + NS * Cfg::kStageBytes;  // resident weights (RESB), 1024-byte aligned
    uint8_t* s_staging = s_w + Cfg::kWBytes;
    uint32_t* s_pre_scale = reinterpret_cast<uint32_t*>(s_staging + Cfg::kStagingBufs * Cfg::kStagingBytes);  // packed pairs
    uint32_t* s_pre_shift = s_pre_scale + kL1MaxCin / 2;
    float* s_out_scale = reinterpret_cast<float*>(s_pre_shift + kL1MaxCin / 2);
    float* s_bias = s_out_scale + kL1MaxCout;
    uint64_t* raw_full = reinterpret_cast<uint64_t*>(s_bias + kL1MaxCout);
    uint64_t* xf_full = raw_full + NS;
    uint64_t* empty_bar = xf_full + NS;
    uint64_t* tmem_full = empty_bar + NS;
    uint64_t* tmem_empty = tmem_full + 2;
    uint64_t* w_bar = tmem_empty + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_bar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Role map: the two single-thread roles sit on the HIGHEST warp ids.  The SM's issue arbiter prefers the highest
    // warp id of an SMSP (measured timeline: with the MMA issuer on warp 1 it was starved for ~1 us by the poll loops of
    // higher-numbered waiting warps - a priority inversion, since those warps were waiting for the MMA).
    constexpr int kNW = kL1Threads / 32;
    const int wrole = warp >= kNW - 2 ? warp - (kNW - 2) : warp + 2;  // 0 TMA, 1 MMA, 2.. transform, then epilogue
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;
    const bool has_pre = p.pre_scale != nullptr;

    if (wrole == 0 && lane == 0) {
        for (int s = 0; s < NS; ++s) {
            MbarInit(&raw_full[s], 1);
            MbarInit(&xf_full[s], kL1XfWarps);
            MbarInit(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            MbarInit(&tmem_full[a], 1);
            MbarInit(&tmem_empty[a], kL1EpiWarps);
        }
        MbarInit(w_bar, 1);
        FenceBarrierInit();
        PrefetchTensorMap(&tmap_w);
        PrefetchTensorMap(&tmap_in);
        PrefetchTensorMap(&tmap_out);
    }
    if (wrole == 1) TmemAlloc(tmem_slot, Cfg::kTmemCols);
    if (has_pre) {
        for (int i = threadIdx.x; i < (p.Cin + 1) / 2; i += kL1Threads) {
            const int c0 = 2 * i, c1 = 2 * i + 1;
            s_pre_scale[i] = PackPair<MmaT>(p.pre_scale[c0], c1 < p.Cin ? p.pre_scale[c1] : 0.f);
            s_pre_shift[i] = PackPair<MmaT>(p.pre_shift[c0], c1 < p.Cin ? p.pre_shift[c1] : 0.f);
        }
    }
    for (int i = threadIdx.x; i < p.num_n_tiles * BN; i += kL1Threads) {
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
        const ChunkGeom gt = GeomOf<CH>(p.num_chunks - 1, p.Cin);  // only the last chunk can be irregular
        if (RESB) {  // weights do not depend on the previous kernel: load them before the dependency wait
            if (ElectOne()) {
                MbarArriveExpectTx(w_bar, (uint32_t)(p.num_chunks * BN * kRowBytes));
                for (int c = 0; c < p.num_chunks; ++c)
                    TmaLoad2D(s_w + c * BN * kRowBytes, &tmap_w, w_bar, c == p.num_chunks - 1 ? gt.ch_base : c * CH, 0);
            }
            __syncwarp();
        }
        GridDepWait();
        int stage = 0;
        uint32_t phase = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int m_tile = tile / p.num_n_tiles, n_tile = tile - m_tile * p.num_n_tiles;
            for (int c = 0; c < p.num_chunks; ++c) {
                MbarWaitWarp(&empty_bar[stage], phase ^ 1u);
                if (ElectOne()) {
                    const int ch_base = c == p.num_chunks - 1 ? gt.ch_base : c * CH;
                    uint8_t* a_dst = smem + stage * Cfg::kStageBytes;
                    if (POOL) {
                        MbarArriveExpectTx(&raw_full[stage], (uint32_t)(4 * p.tile_rows * kRowBytes + (RESB ? 0 : BN * kRowBytes)));
#pragma unroll
                        for (int pl = 0; pl < 4; ++pl)
                            TmaLoad5D(a_dst + pl * kATileBytes, &tmap_in, &raw_full[stage], p.in_coff + ch_base, 0, m_tile * p.pool_k, pl & 1, pl >> 1);
                    } else {
                        MbarArriveExpectTx(&raw_full[stage], (uint32_t)Cfg::kStageBytes);
                        TmaLoad2D(a_dst, &tmap_in, &raw_full[stage], p.in_coff + ch_base, m_tile * kTileM);
                    }
                    if (!RESB) TmaLoad2D(a_dst + Cfg::kABytes, &tmap_w, &raw_full[stage], ch_base, n_tile * BN);
                }
                __syncwarp();
                if (++stage == NS) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (wrole == 1) {
        // =========================================================== MMA issuer
        constexpr uint32_t idesc = MakeInstrDesc(ME::kFmt, BN);
        const uint64_t stage_desc = MakeSmemDesc(SmemAddr(smem));
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        uint64_t* ready = has_pre ? xf_full : raw_full;
        const ChunkGeom gt = GeomOf<CH>(p.num_chunks - 1, p.Cin);
        const int t_lo = gt.k_lo / ME::kStepK, t_hi = gt.k_hi / ME::kStepK;
        int stage = 0;
        uint32_t phase = 0, tile_iter = 0;
        const uint64_t w_desc = MakeSmemDesc(SmemAddr(s_w));
        if (RESB) MbarWaitWarp(w_bar, 0);
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const uint32_t acc = tile_iter & 1u, acc_phase = (tile_iter >> 1) & 1u;
            MbarWaitWarp(&tmem_empty[acc], acc_phase ^ 1u);
            TcFenceAfter();
            const uint32_t d_addr = tmem_u + acc * BN;
            for (int c = 0; c < p.num_chunks; ++c) {
                const bool last = c == p.num_chunks - 1;
                const int ks_lo = last ? t_lo : 0, ks_hi = last ? t_hi : CH / ME::kStepK;
                const uint64_t a_desc = stage_desc + (uint64_t)((uint32_t)stage * (Cfg::kStageBytes >> 4));
                const uint64_t b_desc = RESB ? w_desc + (uint64_t)((uint32_t)c * (BN * kRowBytes >> 4)) : a_desc + (uint64_t)(Cfg::kABytes >> 4);
                MbarWaitWarp(&ready[stage], phase);
                TcFenceAfter();
                if (ElectOne()) {
#pragma unroll
                    for (int ks = 0; ks < CH / ME::kStepK; ++ks)
                        if (ks >= ks_lo && ks < ks_hi)
                            UmmaSS<ME::kKind>(d_addr, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc, (c > 0 || ks > ks_lo) ? 1u : 0u);
                    UmmaCommit(&empty_bar[stage]);
                    if (last) UmmaCommit(&tmem_full[acc]);
                }
                __syncwarp();
                if (++stage == NS) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (wrole < 2 + kL1XfWarps) {
        // =========================================================== transform warps: in-place BN + ReLU on the A tile
        if (has_pre) {
            const int tw = wrole - 2;
            constexpr int kPairs = EPV / 2;
            constexpr int kMaxU = 4;  // (32 rows x one 16-byte piece) units per warp and chunk
            // Full chunks: this warp owns piece `tw` of all 128 rows.  The (possibly partial) last chunk has vp < 8 valid
            // pieces and a different unit -> (rows, piece) map; both maps are fixed per launch and computed once.
            uint32_t off_full[kMaxU], off_tail[kMaxU];
            int ch_tail[kMaxU];
            const ChunkGeom gt = GeomOf<CH>(p.num_chunks - 1, p.Cin);
            const int vp_t = (gt.k_hi - gt.k_lo) / EPV, p_lo_t = gt.k_lo / EPV;
#pragma unroll
            for (int i = 0; i < kMaxU; ++i) {
                const int row = i * 32 + lane;
                off_full[i] = (uint32_t)(row * kRowBytes + ((tw ^ (row & 7)) << 4));
                const int u = tw + i * kL1XfWarps;
                const int rg = u / vp_t, piece = p_lo_t + (u - rg * vp_t);
                const int trow = rg * 32 + lane;
                off_tail[i] = (uint32_t)(trow * kRowBytes + ((piece ^ (trow & 7)) << 4));
                ch_tail[i] = u < vp_t * 4 ? gt.ch_base + piece * EPV : -1;
            }
            const bool tail_partial = vp_t != 8;
            const uint32_t smem_base = SmemAddr(smem);
            uint32_t sc[kPairs], sh[kPairs];
            const uint32_t pre_sc_addr = SmemAddr(s_pre_scale), pre_sh_addr = SmemAddr(s_pre_shift);
            auto load_consts = [&](int ch0) {
#pragma unroll
                for (int e = 0; e < kPairs; e += 4) {
                    const uint4 a = LdsV4(pre_sc_addr + (ch0 / 2 + e) * 4);
                    const uint4 b = LdsV4(pre_sh_addr + (ch0 / 2 + e) * 4);
                    sc[e] = a.x; sc[e + 1] = a.y; sc[e + 2] = a.z; sc[e + 3] = a.w;
                    sh[e] = b.x; sh[e + 1] = b.y; sh[e + 2] = b.z; sh[e + 3] = b.w;
                }
            };
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int c = 0; c < p.num_chunks; ++c) {
                    const uint32_t a_base = smem_base + stage * Cfg::kStageBytes;
                    MbarWaitWarp(&raw_full[stage], phase);
                    uint4 v[kMaxU];
                    if (POOL) {
                        // four raw planes -> plane 0 (all chunks are full in POOL mode); two units at a time (8 loads in flight)
                        load_consts(c * CH + tw * EPV);
#pragma unroll
                        for (int i = 0; i < kMaxU; i += 2) {
                            uint4 q[2][4];
#pragma unroll
                            for (int j = 0; j < 2; ++j)
#pragma unroll
                                for (int pl = 0; pl < 4; ++pl) q[j][pl] = LdsV4(a_base + pl * kATileBytes + off_full[i + j]);
#pragma unroll
                            for (int j = 0; j < 2; ++j)
                                v[j] = p.pre_relu ? PoolPiece<MmaT, true>(q[j], sc, sh) : PoolPiece<MmaT, false>(q[j], sc, sh);
#pragma unroll
                            for (int j = 0; j < 2; ++j) StsV4(a_base + off_full[i + j], v[j]);
                        }
                    } else if (!(tail_partial && c == p.num_chunks - 1)) {
                        // all loads first, then the math, then the stores: four independent dependency chains in flight
#pragma unroll
                        for (int i = 0; i < kMaxU; ++i) v[i] = LdsV4(a_base + off_full[i]);
                        load_consts(c * CH + tw * EPV);
#pragma unroll
                        for (int i = 0; i < kMaxU; ++i)
                            v[i] = p.pre_relu ? ProloguePiece<MmaT, true>(v[i], sc, sh) : ProloguePiece<MmaT, false>(v[i], sc, sh);
#pragma unroll
                        for (int i = 0; i < kMaxU; ++i) StsV4(a_base + off_full[i], v[i]);
                    } else {
#pragma unroll
                        for (int i = 0; i < kMaxU; ++i)
                            if (ch_tail[i] >= 0) v[i] = LdsV4(a_base + off_tail[i]);
                        int loaded = -1;
#pragma unroll
                        for (int i = 0; i < kMaxU; ++i) {
                            if (ch_tail[i] < 0) continue;
                            if (ch_tail[i] != loaded) { load_consts(ch_tail[i]); loaded = ch_tail[i]; }
                            v[i] = p.pre_relu ? ProloguePiece<MmaT, true>(v[i], sc, sh) : ProloguePiece<MmaT, false>(v[i], sc, sh);
                        }
#pragma unroll
                        for (int i = 0; i < kMaxU; ++i)
                            if (ch_tail[i] >= 0) StsV4(a_base + off_tail[i], v[i]);
                    }
                    FenceProxyAsync();
                    __syncwarp();
                    if (lane == 0) MbarArrive(&xf_full[stage]);  // one arrive per warp (per-thread arrives serialise on the barrier word)
                    if (++stage == NS) { stage = 0; phase ^= 1u; }
                }
            }
        }
    } else {
        // =========================================================== epilogue
        const int ew = wrole - (2 + kL1XfWarps);
        const int q = warp & 3;             // TMEM lane quarter this warp may access
        const int h = ew >> 2;              // which half of the column groups
        constexpr int kCgs = BN / 32;       // 32-column groups per tile
        constexpr int kCgPerWarp = kCgs >= 2 ? kCgs / 2 : 1;
        constexpr int kOutB = (int)sizeof(OutT);
        constexpr int kPiecesPerCg = 32 * kOutB / 16;  // 16-byte pieces per row of one column group
        const bool leader = (wrole == 2 + kL1XfWarps) && lane == 0;
        const int row = q * 32 + lane;
        if (leader) GridDepWait();
        uint32_t tile_iter = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const int m_tile = tile / p.num_n_tiles, n_tile = tile - m_tile * p.num_n_tiles;
            const uint32_t acc = tile_iter & 1u, acc_phase = (tile_iter >> 1) & 1u;
            uint8_t* stg = s_staging + (Cfg::kStagingBufs == 2 ? acc : 0u) * Cfg::kStagingBytes;
            if (leader) {  // the store that last read this staging buffer is done
                if (Cfg::kStagingBufs == 2) BulkWaitRead<1>();
                else BulkWaitRead<0>();
            }
            // ONE warp waits for the accumulator, the barrier that follows (needed for the staging buffer anyway) releases the other
            // seven: every poll of an mbarrier is a shared-memory access, and this kernel's shared-memory port is its busiest resource
            if (ew == 0) MbarWaitWarp(&tmem_full[acc], acc_phase);
            NamedBarSync(1, kL1EpiWarps * 32);
            TcFenceAfter();
            if (kCgs >= 2 || h == 0) {
#pragma unroll
                for (int ci = 0; ci < kCgPerWarp; ++ci) {
                    const int cg = h * kCgPerWarp + ci;
                    uint32_t r[32];
                    TmemLoad32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * BN + cg * 32, r);
                    TmemLoadWait();
                    constexpr int kWords = 32 * kOutB / 4;
                    uint32_t w[kWords];
                    const uint32_t scp = SmemAddr(s_out_scale) + (n_tile * BN + cg * 32) * 4;
                    const uint32_t bip = SmemAddr(s_bias) + (n_tile * BN + cg * 32) * 4;
                    if (CSTP) {
                        // compile-time column group -> the constants are immediate offsets into the parameter (constant) bank
                        if (h == 0) {
                            if (p.post_relu) EpiloguePack32<OutT, true>(r, cst.s + ci * 32, cst.b + ci * 32, w);
                            else EpiloguePack32<OutT, false>(r, cst.s + ci * 32, cst.b + ci * 32, w);
                        } else {
                            if (p.post_relu) EpiloguePack32<OutT, true>(r, cst.s + (kCgPerWarp + ci) * 32, cst.b + (kCgPerWarp + ci) * 32, w);
                            else EpiloguePack32<OutT, false>(r, cst.s + (kCgPerWarp + ci) * 32, cst.b + (kCgPerWarp + ci) * 32, w);
                        }
                    } else if (p.post_relu) EpiloguePack32Smem<OutT, true>(r, scp, bip, w);
                    else EpiloguePack32Smem<OutT, false>(r, scp, bip, w);
                    // staging: slabs of [128 rows][128 B], SWIZZLE_128B (conflict-free: 8 consecutive rows hit 8 different pieces)
                    const int byte0 = cg * 32 * kOutB;  // byte offset of this column group inside the output row
                    const uint32_t slab = SmemAddr(stg) + (byte0 >> 7) * (kTileM * 128) + row * 128;
                    const int piece0 = (byte0 & 127) >> 4;
#pragma unroll
                    for (int i = 0; i < kPiecesPerCg; ++i)
                        StsV4(slab + (((piece0 + i) ^ (row & 7)) << 4), make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]));
                }
            }
            TcFenceBefore();
            __syncwarp();
            if (lane == 0) MbarArrive(&tmem_empty[acc]);
            FenceProxyAsync();  // staging writes -> visible to the TMA store
            NamedBarSync(2, kL1EpiWarps * 32);
            if (leader) {
#pragma unroll
                for (int s = 0; s < Cfg::kSlabs; ++s)
                    TmaStore2D(&tmap_out, stg + s * (kTileM * 128), p.out_coff + n_tile * BN + s * (128 / kOutB), m_tile * p.tile_rows);
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

template <typename MmaT, typename OutT, int BN, bool RESB, bool POOL = false, bool CSTP = false>
cudaError_t LaunchL1(const CUtensorMap& tw, const CUtensorMap& tin, const CUtensorMap& tout, const L1Params& p, cudaStream_t stream,
                     const L1Consts* cst = nullptr) {
    using Cfg = L1Cfg<BN, (int)sizeof(OutT), RESB, POOL>;
    auto kern = conv1x1_tma_kernel<MmaT, OutT, BN, RESB, POOL, CSTP>;
    static const L1Consts kNoConsts = {};
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
    const int tiles = p.num_m_tiles * p.num_n_tiles;
    const int grid = tiles < sm_count[dev] ? tiles : sm_count[dev];
    cudaError_t le = LaunchPdl(kern, grid, kL1Threads, Cfg::kSmemBytes, stream, tw, tin, tout, p, cst ? *cst : kNoConsts);
    CountLaunch();
    return le;
}

}  // namespace

bool Conv1x1TmaSupported(const ConvArgs& a) {
    const DType it = a.in.dtype, ot = a.out.dtype;
    if (it != ot || (it != DType::BF16 && it != DType::FP8)) return false;
    if (!(a.R == 1 && a.S == 1 && a.stride == 1 && a.pad == 0) || a.stem_nchw) return false;
    const int esz = (int)DTypeSize(it);
    if (a.pool2) {  // transition: whole 128-byte chunks, a prologue, even images, N tile 128, at least one output row per tile
        if (a.Cin % (128 / esz) != 0 || !a.pre_scale || a.in.H != 2 * a.out.H || a.in.W != 2 * a.out.W || a.out.W > 128 || a.Cout % 128 != 0) return false;
    } else if (a.in.H != a.out.H || a.in.W != a.out.W) {
        return false;
    }
    const int step_k = 32 / esz, slab = 128 / esz;
    if (a.Cin % step_k != 0 || a.Cin > kL1MaxCin || a.Cin < step_k) return false;
    if (a.Cout % slab != 0 || a.Cout > kL1MaxCout) return false;  // whole 128-byte slabs per TMA store
    if (a.Cout > 64 && a.Cout % 128 != 0) return false;           // N tile = 64 or 128, same rule as the weight packer
    if ((a.in.pitch * esz) % 16 != 0 || (a.out.pitch * esz) % 16 != 0) return false;
    if ((a.in.c_off * esz) % 16 != 0 || (a.out.c_off * esz) % 16 != 0) return false;
    if (reinterpret_cast<uintptr_t>(a.in.base) % 16 != 0 || reinterpret_cast<uintptr_t>(a.out.base) % 16 != 0) return false;
    // a short single chunk reads a whole 128-byte box: it must stay inside the pixel (values past Cin are never multiplied)
    if (a.Cin < slab && a.in.c_off + slab > a.in.pitch) return false;
    return true;
}

cudaError_t Conv1x1Tma(const ConvArgs& a, const UmmaWeights& w, cudaStream_t stream) {
    if (!Conv1x1TmaSupported(a) || !w.tensor_map) return cudaErrorInvalidValue;
    const DType it = a.in.dtype;
    const int esz = (int)DTypeSize(it);
    const int ch = 128 / esz;
    L1Params p;
    p.pre_scale = a.pre_scale; p.pre_shift = a.pre_shift; p.out_scale = w.out_scale; p.bias = a.bias;
    p.pre_relu = a.pre_relu; p.post_relu = a.post_relu;
    p.in_coff = a.in.c_off; p.out_coff = a.out.c_off;
    p.Cin = a.Cin; p.Cout = a.Cout;
    p.M = a.n * a.out.H * a.out.W;
    if (p.M <= 0) return cudaSuccess;
    const int bn = a.Cout <= 64 ? 64 : 128;
    p.num_n_tiles = a.Cout / bn;
    p.num_chunks = (a.Cin + ch - 1) / ch;
    p.tile_rows = kTileM; p.pool_k = 0; p.out_scale_mul = a.out_mul;
    p.num_m_tiles = (p.M + kTileM - 1) / kTileM;
    TensorMap tin, tout;
    if (a.pool2) {
        const int Wo = a.out.W, Ho = a.out.H;
        p.pool_k = kTileM / Wo;
        if (p.pool_k > 256) p.pool_k = 256;
        p.tile_rows = p.pool_k * Wo;
        p.out_scale_mul = 0.25f;
        p.num_m_tiles = (a.n * Ho + p.pool_k - 1) / p.pool_k;
        // input as [row pair r = img*Ho + oy][dy][ox][dx][c]: one box = {128 B of channels, Wo, k rows} for a fixed (dx, dy)
        const uint64_t px = (uint64_t)a.in.pitch * esz;
        const uint64_t dims[5] = {(uint64_t)a.in.pitch, (uint64_t)Wo, (uint64_t)a.n * Ho, 2, 2};
        const uint64_t strides[4] = {2 * px, 2 * (uint64_t)a.in.W * px, px, (uint64_t)a.in.W * px};
        const uint32_t box[5] = {(uint32_t)ch, (uint32_t)Wo, (uint32_t)p.pool_k, 1, 1};
        if (MakeTensorMap(&tin, a.in.base, esz, 5, dims, strides, box, true) != 0) return cudaErrorInvalidValue;
    } else {
        const uint64_t dims[2] = {(uint64_t)a.in.pitch, (uint64_t)p.M};
        const uint64_t strides[1] = {(uint64_t)a.in.pitch * esz};
        const uint32_t box[2] = {(uint32_t)ch, (uint32_t)kTileM};
        if (MakeTensorMap(&tin, a.in.base, esz, 2, dims, strides, box, true) != 0) return cudaErrorInvalidValue;
    }
    {
        const uint64_t dims[2] = {(uint64_t)a.out.pitch, (uint64_t)p.M};
        const uint64_t strides[1] = {(uint64_t)a.out.pitch * esz};
        const uint32_t box[2] = {(uint32_t)ch, (uint32_t)p.tile_rows};
        if (MakeTensorMap(&tout, a.out.base, esz, 2, dims, strides, box, true) != 0) return cudaErrorInvalidValue;
    }
    const CUtensorMap& tw = *reinterpret_cast<const CUtensorMap*>(w.tensor_map);
    const CUtensorMap& ti = *reinterpret_cast<const CUtensorMap*>(&tin);
    const CUtensorMap& to = *reinterpret_cast<const CUtensorMap*>(&tout);
    if (a.pool2) {
        if (bn != 128) return cudaErrorInvalidValue;
        if (it == DType::BF16) return LaunchL1<__nv_bfloat16, __nv_bfloat16, 128, false, true>(tw, ti, to, p, stream);
        return LaunchL1<__nv_fp8_e4m3, __nv_fp8_e4m3, 128, false, true>(tw, ti, to, p, stream);
    }
    static const bool resb_enabled = [] { const char* e = getenv("B200_ENGINE_RESB"); return !(e && e[0] == '0'); }();
    const bool resb = resb_enabled && p.num_n_tiles == 1 && p.num_chunks <= kL1ResChunks;
    static const bool cstp_enabled = [] { const char* e = getenv("B200_ENGINE_L1CSTP"); return !(e && e[0] == '0'); }();
    // (only where there is throughput to win: at a handful of tiles per launch the 1 KB of extra parameters costs launch latency)
    const bool cstp = cstp_enabled && resb && a.Cout == 128 && w.h_out_scale && (a.h_bias || !a.bias) && p.num_m_tiles >= 148;
    L1Consts cst;
    if (cstp)
        for (int i = 0; i < 128; ++i) {
            cst.s[i] = w.h_out_scale[i] * p.out_scale_mul;   // the same float product the kernel's preamble forms
            cst.b[i] = a.h_bias ? a.h_bias[i] : 0.f;
        }
    if (it == DType::BF16) {
        if (bn == 128 && cstp) return LaunchL1<__nv_bfloat16, __nv_bfloat16, 128, true, false, true>(tw, ti, to, p, stream, &cst);
        if (bn == 128) return resb ? LaunchL1<__nv_bfloat16, __nv_bfloat16, 128, true>(tw, ti, to, p, stream)
                                   : LaunchL1<__nv_bfloat16, __nv_bfloat16, 128, false>(tw, ti, to, p, stream);
        return resb ? LaunchL1<__nv_bfloat16, __nv_bfloat16, 64, true>(tw, ti, to, p, stream)
                    : LaunchL1<__nv_bfloat16, __nv_bfloat16, 64, false>(tw, ti, to, p, stream);
    }
    if (bn != 128) return cudaErrorInvalidValue;
    if (cstp) return LaunchL1<__nv_fp8_e4m3, __nv_fp8_e4m3, 128, true, false, true>(tw, ti, to, p, stream, &cst);
    return resb ? LaunchL1<__nv_fp8_e4m3, __nv_fp8_e4m3, 128, true>(tw, ti, to, p, stream)
                : LaunchL1<__nv_fp8_e4m3, __nv_fp8_e4m3, 128, false>(tw, ti, to, p, stream);
}

}  // namespace kernels
}  // namespace b200
