// kernels_umma.cu — implicit-GEMM convolution on Blackwell 5th-gen tensor cores (sm_100a only).
//
//   D[M = n*Ho*Wo pixels][Cout] = A[M][K = taps * Cin] * W[Cout][K]^T          (fp32 accumulate)
//
// One persistent CTA per SM, 14 warps, warp-specialised:
//   warps 0-3, 4-7  A producers (two groups alternating K chunks): coalesced 128-bit global loads of
//                   NHWC channel slices -> fused prologue (folded BatchNorm scale/shift + ReLU of the
//                   pre-activation DenseNet layer, optional 2x2 average pooling of the transition)
//                   -> st.shared into the 128-byte-swizzled K-major UMMA operand layout.
//   warps 8-11      epilogue: tcgen05.ld the fp32 accumulator out of TMEM, per-channel dequant scale
//                   + bias + ReLU, convert (bf16 / e4m3), 128-bit stores into the output's channel
//                   slice (dense-block concat in place).
//   warp 12         B producer: TMA (cp.async.bulk.tensor.2d, SWIZZLE_128B) of the packed weights.
//   warp 13         MMA issuer: one thread issues tcgen05.mma (M=128, N=BN, K=32 bytes per step),
//                   tcgen05.commit releases smem stages / publishes the accumulator.
// Pipelines: NS smem stages (full/empty mbarriers) and 2 TMEM accumulators (full/empty mbarriers), so
// the epilogue of tile i overlaps the main loop of tile i+1.
//
// Why A is not loaded by TMA: every dense layer applies its OWN BatchNorm+ReLU to the shared concat
// buffer (pre-activation), so the A tile has to pass through registers anyway; loading it there
// directly saves one shared-memory round trip and makes zero padding, the 2x2 pooling and the 7x7
// stem gather trivial.  Weights (B) have no such transform and do use TMA.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp8.h>

#include <cstdlib>

#include "kernels.h"
#include "umma_ptx.cuh"

namespace b200 {
namespace kernels {

namespace {

struct UParams {
    const void* in;
    void* out;
    const float* pre_scale;
    const float* pre_shift;
    const float* out_scale;
    const float* bias;
    int pre_relu, post_relu;
    int H, W, in_pitch, in_coff;
    int Ho, Wo, out_pitch, out_coff, Cout;
    int Cin, R, S, stride, pad;
    int M;
    int num_m_tiles, num_n_tiles;
    int chunks_per_tap, num_chunks;
    int cin_pad, cout_pad;
};

// A-producer flavours
enum : int {
    kModeLinear = 0,     // 1x1, stride 1, no padding: input pixel == output pixel; cp.async, optional in-place prologue
    kModeGather = 1,     // RxS window with zero padding: cp.async (LDGSTS) with zero fill, optional in-place prologue
    kModePool2 = 3,      // 1x1 on the 2x2 average of the prologue-transformed input (transition layers)
    kModeStem = 4        // 7x7/s2 on a 3(+1 pad)-channel image: 8-byte cp.async, 8 pixels x 4 channels per filter row
};

constexpr int kMaxCin = 1536;       // prologue vectors staged in shared memory
constexpr int kMaxCoutPad = 1024;   // epilogue vectors staged in shared memory
constexpr int kEpiStageBytes = 4 * 32 * 80;
constexpr int kVecSmemBytes = (2 * kMaxCin + 2 * kMaxCoutPad) * 4 + kEpiStageBytes;

template <int BN> struct TileCfg {
    static constexpr int kStageBytes = kATileBytes + BN * kRowBytes;
    static constexpr int kStages = BN == 128 ? 6 : 8;
    static constexpr int kTmemCols = BN == 128 ? 256 : BN == 64 ? 128 : 64;
    static constexpr int kSmemBytes = kStages * kStageBytes + kVecSmemBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

// Walks the (tile, K-chunk) sequence of this CTA, visiting every second chunk (one producer group's share).
struct ChunkIter {
    int tile, c;
    uint32_t it;
    __device__ __forceinline__ void Init(int group, int num_chunks) {
        tile = blockIdx.x;
        c = 0;
        it = 0;
        if (group) Step(1, num_chunks);
    }
    __device__ __forceinline__ void Step(int n, int num_chunks) {
        c += n;
        it += n;
        while (c >= num_chunks) {
            c -= num_chunks;
            tile += gridDim.x;
        }
    }
};

struct RowInfo {
    int pix[8];  // img * H * W, or -1 for rows past M
    int oyx[8];  // (iy0 << 16) | (ix0 & 0xFFFF): top-left input coordinate of the receptive field
};

template <int MODE>
__device__ __forceinline__ void DecodeRows(const UParams& p, int m_tile, int rbase, RowInfo& ri) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int m = m_tile * kTileM + rbase + 16 * i;
        if (m < p.M) {
            int ox = m % p.Wo, oy = (m / p.Wo) % p.Ho, img = m / (p.Wo * p.Ho);
            int iy0 = (MODE == kModePool2) ? oy * 2 : oy * p.stride - p.pad;
            int ix0 = (MODE == kModePool2) ? ox * 2 : ox * p.stride - p.pad;
            if (MODE == kModeStem) {
                // pix = element offset of the receptive field's top-left pixel (may point before the image);
                // oyx = validity masks: bit 8+r <=> filter row r is inside the image, bit d <=> column ix0+d is
                ri.pix[i] = ((img * p.H + iy0) * p.W + ix0) * p.in_pitch;
                uint32_t ymask = 0, xmask = 0;
#pragma unroll
                for (int r = 0; r < 8; ++r) ymask |= (uint32_t)(r < p.R && iy0 + r >= 0 && iy0 + r < p.H) << r;
#pragma unroll
                for (int d = 0; d < 8; ++d) xmask |= (uint32_t)(d < p.S && ix0 + d >= 0 && ix0 + d < p.W) << d;
                ri.oyx[i] = (int)((ymask << 8) | xmask);
            } else {
                ri.pix[i] = img * p.H * p.W;
                ri.oyx[i] = (int)(((unsigned)iy0 << 16) | ((unsigned)ix0 & 0xFFFFu));
            }
        } else {
            ri.pix[i] = (MODE == kModeStem) ? 0 : -1;
            ri.oyx[i] = 0;
        }
    }
}

// ------------------------------------------------------------------ the kernel
template <typename MmaT, typename OutT, int BN, int MODE>
__global__ void __launch_bounds__(kThreads, 1) conv_umma_kernel(const __grid_constant__ CUtensorMap tmap_w, const UParams p) {
    using ME = MmaElem<MmaT>;
    using Cfg = TileCfg<BN>;
    constexpr int NS = Cfg::kStages;
    constexpr int EPV = ME::kPerVec;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float* s_pre_scale = reinterpret_cast<float*>(smem + NS * Cfg::kStageBytes);
    float* s_pre_shift = s_pre_scale + kMaxCin;
    float* s_out_scale = s_pre_shift + kMaxCin;
    float* s_bias = s_out_scale + kMaxCoutPad;
    uint8_t* s_stage = reinterpret_cast<uint8_t*>(s_bias + kMaxCoutPad);  // 4 epilogue warps x 32 rows x 80 B
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(s_stage + kEpiStageBytes);
    uint64_t* empty_bar = full_bar + NS;
    uint64_t* tmem_full = empty_bar + NS;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int num_tiles = p.num_m_tiles * p.num_n_tiles;

    if (warp == 12 && lane == 0) {
        for (int s = 0; s < NS; ++s) {
            MbarInit(&full_bar[s], 128 + 1);  // 128 A-producer threads + the TMA thread's expect_tx arrive
            MbarInit(&empty_bar[s], 1);       // tcgen05.commit
        }
        for (int a = 0; a < 2; ++a) {
            MbarInit(&tmem_full[a], 1);
            MbarInit(&tmem_empty[a], 128);
        }
        FenceBarrierInit();
        PrefetchTensorMap(&tmap_w);
    }
    if (warp == 13) TmemAlloc(tmem_slot, Cfg::kTmemCols);
    // per-channel vectors -> shared memory (read as broadcasts by the producers / the epilogue)
    if (p.pre_scale) {
        if (MODE == kModePool2) {  // fp32 math (four pixels are summed)
            for (int i = threadIdx.x; i < p.cin_pad; i += kThreads) {
                s_pre_scale[i] = i < p.Cin ? p.pre_scale[i] : 0.f;
                s_pre_shift[i] = i < p.Cin ? p.pre_shift[i] : 0.f;
            }
        } else {  // packed pairs in the prologue's arithmetic type (bf16x2 / f16x2)
            uint32_t* sc = reinterpret_cast<uint32_t*>(s_pre_scale);
            uint32_t* sh = reinterpret_cast<uint32_t*>(s_pre_shift);
            for (int i = threadIdx.x; i < p.cin_pad / 2; i += kThreads) {
                const int c0 = 2 * i, c1 = 2 * i + 1;
                sc[i] = PackPair<MmaT>(c0 < p.Cin ? p.pre_scale[c0] : 0.f, c1 < p.Cin ? p.pre_scale[c1] : 0.f);
                sh[i] = PackPair<MmaT>(c0 < p.Cin ? p.pre_shift[c0] : 0.f, c1 < p.Cin ? p.pre_shift[c1] : 0.f);
            }
        }
    }
    for (int i = threadIdx.x; i < p.cout_pad; i += kThreads) {
        s_out_scale[i] = i < p.Cout ? p.out_scale[i] : 0.f;
        s_bias[i] = (p.bias && i < p.Cout) ? p.bias[i] : 0.f;
    }
    TcFenceBefore();
    __syncthreads();
    TcFenceAfter();
    const uint32_t tmem_base = *tmem_slot;
    GridDepLaunch();
    if (warp < 12) GridDepWait();  // activations: readers (producers) and writers (epilogue); weight TMA runs ahead

    if (warp < 8) {
        // =========================================================== A producers
        const int group = warp >> 2;
        const int pt = threadIdx.x & 127;
        const int sub = pt & 7;     // which 16-byte piece of the 128-byte row
        const int rbase = pt >> 3;  // rows rbase + 16*i
        const MmaT* in = reinterpret_cast<const MmaT*>(p.in);
        const bool has_pre = p.pre_scale != nullptr;
        uint32_t sw_off[8];         // swizzled byte offset of this thread's piece in each of its 8 rows
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int row = rbase + 16 * i;
            sw_off[i] = row * kRowBytes + ((sub ^ (row & 7)) << 4);
        }
        ChunkIter cur;
        cur.Init(group, p.num_chunks);

        if (MODE == kModeLinear || MODE == kModeGather || MODE == kModeStem) {
            // ---- cp.async path: up to kDepth chunks per group in flight with no registers held.  When the layer
            //      has a prologue (folded BN + ReLU) each thread transforms, IN PLACE, exactly the 16-byte pieces
            //      it copied itself once they have landed (so no cross-thread hazard, no extra staging buffer).
            constexpr int kDepth = 3;  // chunks of this group in flight (the other group adds as many)
            RowInfo ri, ri_tail;
            int decoded_tile = -1, decoded_tail = -1;
            ChunkIter tail = cur;
            auto finish = [&](const ChunkIter& q) {  // chunk q has landed: optional in-place prologue, then publish
                const int stage = q.it % NS;
                if (MODE != kModeStem && has_pre) {
                    const uint32_t a_base = SmemAddr(smem + stage * Cfg::kStageBytes);
                    int ch0, fr = 0, fs = 0;
                    if (MODE == kModeLinear) {
                        ch0 = q.c * ME::kChunk + sub * EPV;
                    } else {
                        const int tap = q.c / p.chunks_per_tap, j = q.c - tap * p.chunks_per_tap;
                        fr = tap / p.S;
                        fs = tap - fr * p.S;
                        ch0 = j * ME::kChunk + sub * EPV;
                        if (q.tile != decoded_tail) {
                            DecodeRows<MODE>(p, q.tile / p.num_n_tiles, rbase, ri_tail);
                            decoded_tail = q.tile;
                        }
                    }
                    if (ch0 < p.Cin) {
                        constexpr int kPairs = EPV / 2;  // packed constants per 16-byte piece
                        uint32_t sc[kPairs], sh[kPairs];
#pragma unroll
                        for (int e = 0; e < kPairs; e += 4) {
                            uint4 a = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(s_pre_scale) + ch0 / 2 + e);
                            uint4 b = *reinterpret_cast<const uint4*>(reinterpret_cast<const uint32_t*>(s_pre_shift) + ch0 / 2 + e);
                            sc[e] = a.x; sc[e + 1] = a.y; sc[e + 2] = a.z; sc[e + 3] = a.w;
                            sh[e] = b.x; sh[e + 1] = b.y; sh[e + 2] = b.z; sh[e + 3] = b.w;
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            if (MODE == kModeGather) {  // zero padding must stay zero: skip out-of-image taps
                                int iy = (ri_tail.oyx[i] >> 16) + fr;
                                int ix = (int)(short)(ri_tail.oyx[i] & 0xFFFF) + fs;
                                if (!(ri_tail.pix[i] >= 0 && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W)) continue;
                            }
                            uint4 v;
                            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a_base + sw_off[i]) : "memory");
                            v = p.pre_relu ? ProloguePiece<MmaT, true>(v, sc, sh) : ProloguePiece<MmaT, false>(v, sc, sh);
                            StsV4(a_base + sw_off[i], v);
                        }
                    }
                }
                FenceProxyAsync();  // generic-proxy smem writes -> visible to the tensor core (async proxy)
                MbarArrive(&full_bar[stage]);
            };

            // Adaptive ring: issue while a stage is free and fewer than kDepth chunks of this group are in flight;
            // otherwise retire (publish) the oldest landed chunk.  Publishing never waits on stage availability.
            int in_flight = 0;
            for (;;) {
                if (cur.tile < num_tiles && in_flight < kDepth) {
                    const int stage = cur.it % NS;
                    const uint32_t phase = (cur.it / NS) & 1u;
                    bool free_slot = MbarTest(&empty_bar[stage], phase ^ 1u);
                    if (!free_slot && in_flight == 0) {
                        MbarWait(&empty_bar[stage], phase ^ 1u);
                        free_slot = true;
                    }
                    if (free_slot) {
                        const uint32_t a_base = SmemAddr(smem + stage * Cfg::kStageBytes);
                        if (MODE != kModeLinear && cur.tile != decoded_tile) {
                            DecodeRows<MODE>(p, cur.tile / p.num_n_tiles, rbase, ri);
                            decoded_tile = cur.tile;
                        }
                        if (MODE == kModeStem) {
                            // chunk c = filter rows 2c, 2c+1; a filter row is 8 pixels x 4 channels (64 B); this thread owns
                            // 2 pixels of one filter row: piece = sub & 3, filter row = 2c + (sub >> 2)
                            const int r = 2 * cur.c + (sub >> 2);
                            const int dx = 2 * (sub & 3);
                            const int off = (r * p.W + dx) * p.in_pitch;
#pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const uint32_t mk = (uint32_t)ri.oyx[i];
                                const bool rok = (mk >> (8 + r)) & 1u;
                                const bool ok0 = rok && ((mk >> dx) & 1u), ok1 = rok && ((mk >> (dx + 1)) & 1u);
                                const MmaT* src = in + (ri.pix[i] + off);
                                CpAsync8(a_base + sw_off[i], ok0 ? src : in, ok0);
                                CpAsync8(a_base + sw_off[i] + 8, ok1 ? src + p.in_pitch : in, ok1);
                            }
                        } else if (MODE == kModeLinear) {
                            const int m_tile = cur.tile / p.num_n_tiles;
                            const int ch0 = cur.c * ME::kChunk + sub * EPV;
                            const bool ch_ok = ch0 < p.Cin;
                            const MmaT* src0 = in + (size_t)(m_tile * kTileM + rbase) * p.in_pitch + p.in_coff + ch0;
        #pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                const bool ok = ch_ok && (m_tile * kTileM + rbase + 16 * i) < p.M;
                                CpAsync16(a_base + sw_off[i], ok ? src0 + (size_t)(16 * i) * p.in_pitch : in, ok);
                            }
                        } else {
                            const int tap = cur.c / p.chunks_per_tap, j = cur.c - tap * p.chunks_per_tap;
                            const int fr = tap / p.S, fs = tap - fr * p.S;
                            const int ch0 = j * ME::kChunk + sub * EPV;
                            const bool ch_ok = ch0 < p.Cin;
        #pragma unroll
                            for (int i = 0; i < 8; ++i) {
                                int iy = (ri.oyx[i] >> 16) + fr;
                                int ix = (int)(short)(ri.oyx[i] & 0xFFFF) + fs;
                                bool ok = ch_ok && ri.pix[i] >= 0 && iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                                const MmaT* src = ok ? in + ((size_t)ri.pix[i] + (size_t)iy * p.W + ix) * p.in_pitch + p.in_coff + ch0 : in;
                                CpAsync16(a_base + sw_off[i], src, ok);  // src-size 0 => 16 bytes of zeros (padding)
                            }
                        }
                        CpAsyncCommit();
                        ++in_flight;
                        cur.Step(2, p.num_chunks);
                        continue;
                    }
                }
                if (in_flight == 0) break;
                if (in_flight == 1) CpAsyncWait<0>();
                else if (in_flight == 2) CpAsyncWait<1>();
                else CpAsyncWait<2>();
                finish(tail);
                tail.Step(2, p.num_chunks);
                --in_flight;
            }
        } else {
            // ---- register path: 2x2 average pooling of the prologue-transformed input (transition layers)
            RowInfo ri;
            int decoded_tile = -1;
            for (; cur.tile < num_tiles; cur.Step(2, p.num_chunks)) {
                if (cur.tile != decoded_tile) {
                    DecodeRows<MODE>(p, cur.tile / p.num_n_tiles, rbase, ri);
                    decoded_tile = cur.tile;
                }
                const int stage = cur.it % NS;
                const uint32_t phase = (cur.it / NS) & 1u;
                const uint32_t a_base = SmemAddr(smem + stage * Cfg::kStageBytes);
                const int tap = cur.c / p.chunks_per_tap, j = cur.c - tap * p.chunks_per_tap;
                const int fr = tap / p.S, fs = tap - fr * p.S;
                const int ch0 = j * ME::kChunk + sub * EPV;
                const bool ch_ok = ch0 < p.Cin;
                float sc[EPV], sh[EPV];
#pragma unroll
                for (int e = 0; e < EPV; ++e) {
                    sc[e] = has_pre ? s_pre_scale[ch0 + e] : 1.f;
                    sh[e] = has_pre ? s_pre_shift[ch0 + e] : 0.f;
                }
                {  // kModePool2: A row = mean of the 2x2 input pixels after the prologue
                    MbarWait(&empty_bar[stage], phase ^ 1u);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint4 q[4][4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            int ii = half * 4 + i;
                            bool ok = ch_ok && ri.pix[ii] >= 0;
                            int iy = (ri.oyx[ii] >> 16), ix = (int)(short)(ri.oyx[ii] & 0xFFFF);
#pragma unroll
                            for (int d = 0; d < 4; ++d) {
                                q[i][d] = make_uint4(0u, 0u, 0u, 0u);
                                if (ok) q[i][d] = LdgNc(in + ((size_t)ri.pix[ii] + (size_t)(iy + (d >> 1)) * p.W + ix + (d & 1)) * p.in_pitch + p.in_coff + ch0);
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            int ii = half * 4 + i;
                            float acc[EPV];
#pragma unroll
                            for (int e = 0; e < EPV; ++e) acc[e] = 0.f;
                            if (ch_ok && ri.pix[ii] >= 0) {
#pragma unroll
                                for (int d = 0; d < 4; ++d) {
                                    float f[EPV];
                                    ME::Unpack(q[i][d], f);
#pragma unroll
                                    for (int e = 0; e < EPV; ++e) {
                                        float t = fmaf(f[e], sc[e], sh[e]);
                                        acc[e] += p.pre_relu ? fmaxf(t, 0.f) : t;
                                    }
                                }
#pragma unroll
                                for (int e = 0; e < EPV; ++e) acc[e] *= 0.25f;
                            }
                            StsV4(a_base + sw_off[ii], ME::Pack(acc));
                        }
                    }
                }
                FenceProxyAsync();
                MbarArrive(&full_bar[stage]);
            }
        }
    } else if (warp < 12) {
        // =========================================================== epilogue
        const int e = warp & 3;  // TMEM lane quarter this warp may access
        OutT* out = reinterpret_cast<OutT*>(p.out);
        uint32_t tile_iter = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const int m_tile = tile / p.num_n_tiles, n_tile = tile - m_tile * p.num_n_tiles;
            const uint32_t acc = tile_iter & 1u, acc_phase = (tile_iter >> 1) & 1u;
            MbarWait(&tmem_full[acc], acc_phase);
            TcFenceAfter();
            const int m = m_tile * kTileM + e * 32 + lane;
#pragma unroll 1
            for (int cg = 0; cg < BN / 32; ++cg) {
                uint32_t r[32];
                TmemLoad32(tmem_base + ((uint32_t)(e * 32) << 16) + acc * BN + cg * 32, r);
                TmemLoadWait();
                const int co0 = n_tile * BN + cg * 32;
                if (m < p.M && co0 < p.Cout) {
                    float f[32];
#pragma unroll
                    for (int q = 0; q < 32; q += 4) {
                        float4 s4 = *reinterpret_cast<const float4*>(s_out_scale + co0 + q);  // smem broadcast
                        float4 b4 = *reinterpret_cast<const float4*>(s_bias + co0 + q);
                        f[q] = fmaf(__uint_as_float(r[q]), s4.x, b4.x);
                        f[q + 1] = fmaf(__uint_as_float(r[q + 1]), s4.y, b4.y);
                        f[q + 2] = fmaf(__uint_as_float(r[q + 2]), s4.z, b4.z);
                        f[q + 3] = fmaf(__uint_as_float(r[q + 3]), s4.w, b4.w);
                    }
                    if (p.post_relu) {
#pragma unroll
                        for (int q = 0; q < 32; ++q) f[q] = fmaxf(f[q], 0.f);
                    }
                    // stage this thread's 32 channels (one row) in shared memory ...
                    constexpr int kRowB = 32 * (int)sizeof(OutT);       // bytes per row of this column group
                    constexpr int kPieces = kRowB / 16;                 // 16-byte pieces per row (4 bf16 / 2 e4m3)
                    constexpr int kPitch = kRowB + 16;                  // padded: conflict-free row-wise writes
                    const uint32_t st_base = SmemAddr(s_stage + e * (32 * kPitch));
#pragma unroll
                    for (int q = 0; q < kPieces; ++q) {
                        uint4 w = sizeof(OutT) == 2 ? MmaElem<__nv_bfloat16>::Pack(f + 8 * q) : MmaElem<__nv_fp8_e4m3>::Pack(f + 16 * q);
                        StsV4(st_base + lane * kPitch + q * 16, w);
                    }
                }
                __syncwarp();
                {
                    // ... and write it out with kPieces consecutive lanes per row (full 32/64-byte segments per row
                    // instead of 32 scattered 16-byte stores per instruction)
                    constexpr int kRowB = 32 * (int)sizeof(OutT);
                    constexpr int kPieces = kRowB / 16;
                    constexpr int kPitch = kRowB + 16;
                    const uint32_t st_base = SmemAddr(s_stage + e * (32 * kPitch));
                    const int co0 = n_tile * BN + cg * 32;
#pragma unroll
                    for (int i = 0; i < kPieces; ++i) {
                        const int idx = lane + 32 * i;
                        const int row = idx / kPieces, piece = idx % kPieces;
                        const int mr = m_tile * kTileM + e * 32 + row;
                        const int co = co0 + piece * (16 / (int)sizeof(OutT));
                        if (mr < p.M && co < p.Cout) {
                            uint4 v;
                            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(st_base + row * kPitch + piece * 16) : "memory");
                            *reinterpret_cast<uint4*>(out + (size_t)mr * p.out_pitch + p.out_coff + co) = v;
                        }
                    }
                }
                __syncwarp();
            }
            TcFenceBefore();
            MbarArrive(&tmem_empty[acc]);
        }
    } else if (warp == 12) {
        // =========================================================== B producer (TMA); whole warp converged, one lane issues
        uint32_t it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
            const int n_tile = tile % p.num_n_tiles;
            for (int c = 0; c < p.num_chunks; ++c, ++it) {
                const int stage = it % NS;
                const uint32_t phase = (it / NS) & 1u;
                MbarWait(&empty_bar[stage], phase ^ 1u);
                if (ElectOne()) {
                    MbarArriveExpectTx(&full_bar[stage], BN * kRowBytes);
                    TmaLoad2D(smem + stage * Cfg::kStageBytes + kATileBytes, &tmap_w, &full_bar[stage], c * ME::kChunk, n_tile * BN);
                }
                __syncwarp();
            }
        }
    } else {
        // =========================================================== MMA issuer; whole warp converged, one lane issues
        constexpr uint32_t idesc = MakeInstrDesc(ME::kFmt, BN);
        const uint64_t stage_desc = MakeSmemDesc(SmemAddr(smem));
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        uint32_t it = 0, tile_iter = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++tile_iter) {
            const uint32_t acc = tile_iter & 1u, acc_phase = (tile_iter >> 1) & 1u;
            MbarWait(&tmem_empty[acc], acc_phase ^ 1u);
            TcFenceAfter();
            const uint32_t d_addr = tmem_u + acc * BN;
            int j = 0;  // chunk index inside the current filter tap
            for (int c = 0; c < p.num_chunks; ++c, ++it) {
                const int stage = it % NS;
                const uint32_t phase = (it / NS) & 1u;
                int ksteps;  // K steps of this chunk that carry data
                if (MODE == kModeStem) {
                    ksteps = (p.R - 2 * c >= 2 ? 2 : 1) * (32 / ME::kStepK);
                } else {
                    int valid = p.Cin - j * ME::kChunk;
                    ksteps = valid >= ME::kChunk ? ME::kChunk / ME::kStepK : valid / ME::kStepK;
                    if (++j == p.chunks_per_tap) j = 0;
                }
                const uint64_t a_desc = stage_desc + (uint64_t)((uint32_t)stage * (Cfg::kStageBytes >> 4));
                const uint64_t b_desc = a_desc + (uint64_t)(kATileBytes >> 4);
                MbarWait(&full_bar[stage], phase);
                TcFenceAfter();
                if (ElectOne()) {
                    // +32 bytes per K step inside the 128-byte swizzle row (start-address field is >>4)
                    UmmaSS<ME::kKind>(d_addr, a_desc, b_desc, idesc, c > 0 ? 1u : 0u);
#pragma unroll
                    for (int ks = 1; ks < ME::kChunk / ME::kStepK; ++ks)
                        if (ks < ksteps) UmmaSS<ME::kKind>(d_addr, a_desc + (uint64_t)(2 * ks), b_desc + (uint64_t)(2 * ks), idesc, 1u);
                    UmmaCommit(&empty_bar[stage]);
                    if (c == p.num_chunks - 1) UmmaCommit(&tmem_full[acc]);
                }
                __syncwarp();
            }
        }
    }

    TcFenceBefore();
    __syncthreads();
    if (warp == 13) {
        TcFenceAfter();
        TmemDealloc(tmem_base, Cfg::kTmemCols);
    }
}

// =====================================================================================================
// 3x3 / stride 1 / pad 1 convolution with the input patch staged ONCE in shared memory ("halo" kernel).
//
// The gather kernel above re-reads every input pixel 9 times from L2 (once per filter tap); with only
// Cout = 32 output channels per dense layer that makes the 3x3 convs L2-bandwidth bound.  Here a CTA owns a
// TH x TW patch of output pixels of one image, loads the (TH+2) x 16 input pixels (zero filled outside the
// image) once, and runs the 9 taps as 9 row-SHIFTED views of the same shared-memory tile:
//   output row m' = y*16 + x   reads patch pixel  q = m' + fr*16 + fs        (fr, fs = filter tap)
// so each tap is the same UMMA A operand with its start address advanced by (fr*16 + fs) rows.  Row shifts
// that are not a multiple of 8 are incompatible with the swizzled layouts, so A uses the non-swizzled
// K-major canonical layout with rows 16 bytes apart: one "plane" [pixel][16 B] per 16-byte K piece
// (LBO = plane stride, SBO = 128 B).  Columns x >= TW of a tile are junk rows that are never stored.
// All 9 x (Cin/chunk) weight tiles stay resident in shared memory for the lifetime of the CTA.
constexpr int kHaloPW = 16;                       // padded patch width (TW <= 14)
constexpr int kHaloPatchPixels = 10 * kHaloPW + 8;  // (TH+2 <= 10) rows + slack for the junk rows' overreach
constexpr int kHaloPlaneStride = kHaloPatchPixels * 16 + 16;  // +16 B: planes start in different banks
constexpr int kHaloMaxPieces = 16;                // Cin * esz / 16 <= 16  (128 bf16 or 256 e4m3... capped by Cin<=128)
constexpr int kHaloPatchBytes = ((kHaloMaxPieces * kHaloPlaneStride + 1023) / 1024) * 1024;
constexpr int kHaloWeightBytes = 18 * 32 * kRowBytes;  // 9 taps x <=2 chunks x [32][128 B]
constexpr int kHaloBufs = 3;  // patch ring: two tiles in flight while one is consumed
constexpr int kHaloSmemBytes = 1024 + kHaloWeightBytes + kHaloBufs * kHaloPatchBytes + 2 * 32 * 4 + 256;

struct HParams {
    const void* in;
    void* out;
    const float* out_scale;
    const float* bias;
    int post_relu;
    int H, W, in_pitch, in_coff, out_pitch, out_coff, Cin, Cout;
    int n, TH, TW, tiles_x, tiles_y, num_tiles;
    int chunks_per_tap;  // weight chunks (128 B of K) per tap
};

template <typename MmaT, typename OutT>
__global__ void __launch_bounds__(kThreads, 1) conv3x3_halo_kernel(const __grid_constant__ CUtensorMap tmap_w, const HParams p) {
    using ME = MmaElem<MmaT>;
    constexpr int EPV = ME::kPerVec;
    constexpr int BN = 32;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w = smem;                                  // resident weights, SW128 K-major tiles of 4 KB
    uint8_t* s_patch = smem + kHaloWeightBytes;           // 2 patch buffers
    float* s_out_scale = reinterpret_cast<float*>(s_patch + kHaloBufs * kHaloPatchBytes);
    float* s_bias = s_out_scale + 32;
    uint64_t* w_bar = reinterpret_cast<uint64_t*>(s_bias + 32);
    uint64_t* patch_full = w_bar + 1;
    uint64_t* patch_empty = patch_full + kHaloBufs;
    uint64_t* tmem_full = patch_empty + kHaloBufs;
    uint64_t* tmem_empty = tmem_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int pieces = p.Cin / EPV;               // 16-byte K pieces per pixel
    const int num_wtiles = 9 * p.chunks_per_tap;  // resident weight tiles

    if (warp == 12 && lane == 0) {
        MbarInit(w_bar, 1);
        for (int b = 0; b < kHaloBufs; ++b) {
            MbarInit(&patch_full[b], 256);
            MbarInit(&patch_empty[b], 1);
        }
        for (int b = 0; b < 2; ++b) {
            MbarInit(&tmem_full[b], 1);
            MbarInit(&tmem_empty[b], 128);
        }
        FenceBarrierInit();
        PrefetchTensorMap(&tmap_w);
    }
    if (warp == 13) TmemAlloc(tmem_slot, 64);
    if (threadIdx.x < 32) {
        s_out_scale[threadIdx.x] = threadIdx.x < p.Cout ? p.out_scale[threadIdx.x] : 0.f;
        s_bias[threadIdx.x] = (p.bias && threadIdx.x < p.Cout) ? p.bias[threadIdx.x] : 0.f;
    }
    // the slack pixels past the patch are read by junk rows only, but must never hold NaN-producing garbage
    // for rows that ARE stored: they are not (junk rows only); still, clear both buffers once for hygiene
    for (int i = threadIdx.x; i < kHaloBufs * kHaloPatchBytes / 16; i += kThreads)
        reinterpret_cast<uint4*>(s_patch)[i] = make_uint4(0u, 0u, 0u, 0u);
    FenceProxyAsync();
    TcFenceBefore();
    __syncthreads();
    TcFenceAfter();
    const uint32_t tmem_base = *tmem_slot;
    GridDepLaunch();
    if (warp < 12) GridDepWait();

    if (warp < 8) {
        // =========================================================== patch producers (256 threads, cp.async)
        const MmaT* in = reinterpret_cast<const MmaT*>(p.in);
        const int tid = threadIdx.x;  // 0..255
        uint32_t k = 0;
        constexpr int kFullPieces = 128 / EPV;
        const bool full_c = pieces == kFullPieces;
        // fixed K piece per thread (256 % pieces == 0): no div/mod in the copy loop
        const int kp = full_c ? tid % kFullPieces : tid % pieces;
        const int q0 = full_c ? tid / kFullPieces : tid / pieces;
        const int qstep = full_c ? 256 / kFullPieces : 256 / pieces;
        // Adaptive ring (see the main kernel): issue while a patch buffer is free and < kHaloBufs-1 tiles are in
        // flight, otherwise publish the oldest landed patch.
        int tile = blockIdx.x;
        uint32_t done = 0;  // patches published
        int in_flight = 0;
        for (;;) {
            if (tile < p.num_tiles && in_flight < kHaloBufs - 1) {
                const int buf = k % kHaloBufs;
                const uint32_t par = ((k / kHaloBufs) & 1u) ^ 1u;
                bool free_buf = MbarTest(&patch_empty[buf], par);
                if (!free_buf && in_flight == 0) {
                    MbarWait(&patch_empty[buf], par);
                    free_buf = true;
                }
                if (free_buf) {
                    const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, img = tile / (p.tiles_x * p.tiles_y);
                    const int iy0 = ty * p.TH - 1, ix0 = tx * p.TW - 1;
                    const int npix = (p.TH + 2) * kHaloPW;
                    const uint32_t pbase = SmemAddr(s_patch + buf * kHaloPatchBytes) + kp * kHaloPlaneStride;
                    const MmaT* ibase = in + (size_t)img * p.H * p.W * p.in_pitch + p.in_coff + kp * EPV;
                    for (int q = q0; q < npix; q += qstep) {
                        const int iy = iy0 + (q >> 4), ix = ix0 + (q & 15);
                        const bool ok = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                        CpAsync16(pbase + q * 16, ok ? ibase + ((size_t)iy * p.W + ix) * p.in_pitch : in, ok);
                    }
                    CpAsyncCommit();
                    ++in_flight;
                    ++k;
                    tile += gridDim.x;
                    continue;
                }
            }
            if (in_flight == 0) break;
            if (in_flight == 1) CpAsyncWait<0>();
            else CpAsyncWait<1>();
            FenceProxyAsync();
            MbarArrive(&patch_full[done % kHaloBufs]);
            ++done;
            --in_flight;
        }
    } else if (warp < 12) {
        // =========================================================== epilogue
        const int e = warp & 3;
        OutT* out = reinterpret_cast<OutT*>(p.out);
        uint32_t k = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++k) {
            const uint32_t acc = k & 1u, acc_phase = (k >> 1) & 1u;
            const int tx = tile % p.tiles_x, ty = (tile / p.tiles_x) % p.tiles_y, img = tile / (p.tiles_x * p.tiles_y);
            MbarWait(&tmem_full[acc], acc_phase);
            TcFenceAfter();
            uint32_t r[32];
            TmemLoad32(tmem_base + ((uint32_t)(e * 32) << 16) + acc * BN, r);
            TmemLoadWait();
            TcFenceBefore();
            MbarArrive(&tmem_empty[acc]);  // the accumulator is in registers: release it before the stores
            const int mrow = e * 32 + lane;
            const int y = mrow / kHaloPW, x = mrow - y * kHaloPW;
            const int oy = ty * p.TH + y, ox = tx * p.TW + x;
            if (y < p.TH && x < p.TW && oy < p.H && ox < p.W) {
                float f[32];
#pragma unroll
                for (int q = 0; q < 32; q += 4) {
                    float4 s4 = *reinterpret_cast<const float4*>(s_out_scale + q);
                    float4 b4 = *reinterpret_cast<const float4*>(s_bias + q);
                    f[q] = fmaf(__uint_as_float(r[q]), s4.x, b4.x);
                    f[q + 1] = fmaf(__uint_as_float(r[q + 1]), s4.y, b4.y);
                    f[q + 2] = fmaf(__uint_as_float(r[q + 2]), s4.z, b4.z);
                    f[q + 3] = fmaf(__uint_as_float(r[q + 3]), s4.w, b4.w);
                }
                if (p.post_relu) {
#pragma unroll
                    for (int q = 0; q < 32; ++q) f[q] = fmaxf(f[q], 0.f);
                }
                OutT* orow = out + ((size_t)(img * p.H + oy) * p.W + ox) * p.out_pitch + p.out_coff;
                if (sizeof(OutT) == 2) {
#pragma unroll
                    for (int q = 0; q < 32; q += 8)
                        if (q < p.Cout) *reinterpret_cast<uint4*>(orow + q) = MmaElem<__nv_bfloat16>::Pack(f + q);
                } else {
#pragma unroll
                    for (int q = 0; q < 32; q += 16)
                        if (q < p.Cout) *reinterpret_cast<uint4*>(orow + q) = MmaElem<__nv_fp8_e4m3>::Pack(f + q);
                }
            }
        }
    } else if (warp == 12) {
        // =========================================================== weights: TMA once, resident
        if (ElectOne()) {
            MbarArriveExpectTx(w_bar, (uint32_t)num_wtiles * BN * kRowBytes);
            for (int t = 0; t < num_wtiles; ++t) TmaLoad2D(s_w + t * BN * kRowBytes, &tmap_w, w_bar, t * ME::kChunk, 0);
        }
        __syncwarp();
    } else {
        // =========================================================== MMA issuer; whole warp converged, one lane issues
        constexpr uint32_t idesc = MakeInstrDesc(ME::kFmt, BN);
        const uint64_t a_base = MakeSmemDescNoSwizzle(SmemAddr(s_patch), kHaloPlaneStride, 128);
        const uint64_t b_base = MakeSmemDesc(SmemAddr(s_w));
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        MbarWait(w_bar, 0);
        uint32_t k = 0;
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++k) {
            const uint32_t buf = k % kHaloBufs, ph = (k / kHaloBufs) & 1u;
            const uint32_t acc = k & 1u, acc_ph = (k >> 1) & 1u;
            MbarWait(&tmem_empty[acc], acc_ph ^ 1u);
            MbarWait(&patch_full[buf], ph);
            TcFenceAfter();
            const uint32_t d_addr = tmem_u + acc * BN;
            const uint64_t a0 = a_base + (uint64_t)(buf * (kHaloPatchBytes >> 4));
            constexpr int CPT = 128 / ME::kChunk;  // chunks per tap when Cin == 128
            if (ElectOne()) {
                if (p.Cin == 128) {
                    // fully unrolled: every descriptor is base + compile-time constant
#pragma unroll
                    for (int tap = 0; tap < 9; ++tap) {
#pragma unroll
                        for (int j = 0; j < CPT; ++j) {
#pragma unroll
                            for (int ks = 0; ks < 4; ++ks) {
                                const uint32_t aoff = (uint32_t)(((j * 8 + 2 * ks) * kHaloPlaneStride) / 16 + (tap / 3) * kHaloPW + (tap % 3));
                                const uint32_t boff = (uint32_t)(((tap * CPT + j) * BN * kRowBytes) / 16 + 2 * ks);
                                UmmaSS<ME::kKind>(d_addr, a0 + (uint64_t)aoff, b_base + (uint64_t)boff, idesc, (tap | j | ks) ? 1u : 0u);
                            }
                        }
                    }
                } else {
                    uint32_t first = 1;
                    for (int tap = 0; tap < 9; ++tap) {
                        const uint32_t shift = (uint32_t)((tap / 3) * kHaloPW + (tap % 3));
                        for (int j = 0; j < p.chunks_per_tap; ++j) {
                            int valid = p.Cin - j * ME::kChunk;
                            if (valid > ME::kChunk) valid = ME::kChunk;
                            const int ksteps = valid / ME::kStepK;
                            for (int ks = 0; ks < ksteps; ++ks) {
                                // a K piece is 16 bytes of K (one plane); 8 pieces per 128-byte weight chunk
                                const uint32_t aoff = (uint32_t)((j * 8 + 2 * ks) * (kHaloPlaneStride / 16)) + shift;
                                const uint32_t boff = (uint32_t)((tap * p.chunks_per_tap + j) * (BN * kRowBytes / 16) + 2 * ks);
                                UmmaSS<ME::kKind>(d_addr, a0 + (uint64_t)aoff, b_base + (uint64_t)boff, idesc, first ? 0u : 1u);
                                first = 0;
                            }
                        }
                    }
                }
                UmmaCommit(&patch_empty[buf]);
                UmmaCommit(&tmem_full[acc]);
            }
            __syncwarp();
        }
    }
    TcFenceBefore();
    __syncthreads();
    if (warp == 13) {
        TcFenceAfter();
        TmemDealloc(tmem_base, 64);
    }
}

template <typename MmaT, typename OutT>
cudaError_t LaunchHalo(const CUtensorMap& tm, const HParams& p, cudaStream_t stream) {
    auto kern = conv3x3_halo_kernel<MmaT, OutT>;
    static int sm_count[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!sm_count[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kHaloSmemBytes);
        if (e != cudaSuccess) return e;
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sm_count[dev] = n > 0 ? n : 148;
    }
    int grid = p.num_tiles < sm_count[dev] ? p.num_tiles : sm_count[dev];
    cudaError_t le = LaunchPdl(kern, grid, kThreads, kHaloSmemBytes, stream, tm, p);
    CountLaunch();
    return le;
}

template <typename MmaT, typename OutT, int BN, int MODE>
cudaError_t Launch(const CUtensorMap& tm, const UParams& p, cudaStream_t stream) {
    using Cfg = TileCfg<BN>;
    auto kern = conv_umma_kernel<MmaT, OutT, BN, MODE>;
    static int sm_count[64] = {0};  // per instantiation and device; the smem opt-in is a per-device attribute
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
    int tiles = p.num_m_tiles * p.num_n_tiles;
    int grid = tiles < sm_count[dev] ? tiles : sm_count[dev];
    cudaError_t le = LaunchPdl(kern, grid, kThreads, Cfg::kSmemBytes, stream, tm, p);
    CountLaunch();
    return le;
}

template <typename MmaT, typename OutT, int MODE>
cudaError_t LaunchBN(int bn, const CUtensorMap& tm, const UParams& p, cudaStream_t stream) {
    switch (bn) {
        case 32: return Launch<MmaT, OutT, 32, MODE>(tm, p, stream);
        case 64: return Launch<MmaT, OutT, 64, MODE>(tm, p, stream);
        case 128: return Launch<MmaT, OutT, 128, MODE>(tm, p, stream);
    }
    return cudaErrorInvalidValue;
}

template <typename MmaT, typename OutT>
cudaError_t LaunchMode(int mode, int bn, const CUtensorMap& tm, const UParams& p, cudaStream_t stream) {
    switch (mode) {
        case kModeLinear: return LaunchBN<MmaT, OutT, kModeLinear>(bn, tm, p, stream);
        case kModeGather: return LaunchBN<MmaT, OutT, kModeGather>(bn, tm, p, stream);
        case kModePool2: return LaunchBN<MmaT, OutT, kModePool2>(bn, tm, p, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace

int UmmaKChunkElems(DType d) { return d == DType::FP8 ? 128 : 64; }

int UmmaPaddedCin(int Cin, int, int, DType d) {
    int kc = UmmaKChunkElems(d);
    return (Cin + kc - 1) / kc * kc;
}

bool UmmaSupported(const ConvArgs& a) {
    if (a.stem_nchw) return StemNchwSupported(a);
    const DType it = a.in.dtype, ot = a.out.dtype;
    if (it != DType::BF16 && it != DType::FP8) return false;
    if (ot != DType::BF16 && ot != DType::FP8) return false;
    if (it == DType::FP8 && ot != DType::FP8) return false;
    const int esz_in = (int)DTypeSize(it), esz_out = (int)DTypeSize(ot);
    if (a.Cout % 8 != 0 || (a.out.pitch * esz_out) % 16 != 0 || (a.out.c_off * esz_out) % 16 != 0) return false;
    if (a.Cin < 16) {  // stem: bf16 NHWC4, 7 rows x (8 pixels x 4 channels)
        return it == DType::BF16 && a.in.pitch == 4 && a.in.c_off == 0 && a.S <= 7 && a.R <= 8 && !a.pool2 && !a.pre_scale;
    }
    const int step = it == DType::BF16 ? 16 : 32;
    if (a.Cin % step != 0) return false;
    if ((a.in.pitch * esz_in) % 16 != 0 || (a.in.c_off * esz_in) % 16 != 0) return false;
    if (a.pool2 && !(a.R == 1 && a.S == 1 && a.in.H % 2 == 0 && a.in.W % 2 == 0)) return false;
    if (a.Cin > kMaxCin || (a.Cout + 127) / 128 * 128 > kMaxCoutPad) return false;
    if (it != ot) return false;  // mixed bf16 -> e4m3 only exists for the stem
    return true;
}

cudaError_t ConvUmma(const ConvArgs& a, const UmmaWeights& w, cudaStream_t stream) {
    if (!UmmaSupported(a) || !w.tensor_map) return cudaErrorInvalidValue;
    if (a.stem_nchw) return ConvStemNchw(a, w, stream);
    const DType it = a.in.dtype, ot = a.out.dtype;
    const bool stem = a.Cin < 16;
    const int kc = UmmaKChunkElems(it);
    UParams p;
    p.in = a.in.base; p.out = a.out.base;
    p.pre_scale = a.pre_scale; p.pre_shift = a.pre_shift; p.out_scale = w.out_scale; p.bias = a.bias;
    p.pre_relu = a.pre_relu; p.post_relu = a.post_relu;
    p.H = a.in.H; p.W = a.in.W; p.in_pitch = a.in.pitch; p.in_coff = a.in.c_off;
    p.Ho = a.out.H; p.Wo = a.out.W; p.out_pitch = a.out.pitch; p.out_coff = a.out.c_off; p.Cout = a.Cout;
    p.Cin = a.Cin; p.R = a.R; p.S = a.S; p.stride = a.stride; p.pad = a.pad;
    p.M = a.n * p.Ho * p.Wo;
    if (p.M <= 0) return cudaSuccess;
    const int bn = a.Cout <= 32 ? 32 : a.Cout <= 64 ? 64 : 128;
    p.num_m_tiles = (p.M + kTileM - 1) / kTileM;
    p.num_n_tiles = w.Cout_pad / bn;
    p.chunks_per_tap = stem ? 1 : (a.Cin + kc - 1) / kc;
    p.num_chunks = stem ? w.K_pad / kc : a.R * a.S * p.chunks_per_tap;
    p.cin_pad = stem ? 0 : p.chunks_per_tap * kc;
    p.cout_pad = w.Cout_pad;
    const CUtensorMap& tm = *reinterpret_cast<const CUtensorMap*>(w.tensor_map);
    if (stem) {
        if (ot == DType::BF16) return LaunchBN<__nv_bfloat16, __nv_bfloat16, kModeStem>(bn, tm, p, stream);
        return LaunchBN<__nv_bfloat16, __nv_fp8_e4m3, kModeStem>(bn, tm, p, stream);
    }
    auto env_on = [](const char* name) { const char* e = getenv(name); return !(e && e[0] == '0'); };  // read per launch: tests toggle them
    const bool l1tma_enabled = env_on("B200_ENGINE_L1TMA");
    if (l1tma_enabled && Conv1x1TmaSupported(a)) return Conv1x1Tma(a, w, stream);
    // 3x3/s1/p1 bottleneck conv: TMA-loaded swizzled patch, nine row-shifted descriptors (kernels_conv3x3.cu)
    const bool c3tma_enabled = env_on("B200_ENGINE_C3TMA");
    if (c3tma_enabled && Conv3x3TmaSupported(a)) return Conv3x3Tma(a, w, stream);
    // 3x3/s1/p1 with a narrow output: stage the input patch once in shared memory (9 shifted views)
    const bool halo_enabled = env_on("B200_ENGINE_HALO");
    const int step_k = it == DType::BF16 ? 16 : 32;
    if (halo_enabled && a.R == 3 && a.S == 3 && a.stride == 1 && a.pad == 1 && !a.pre_scale && !a.pool2 && a.Cout <= 32 &&
        a.Cin <= 128 && a.Cin % step_k == 0 && it == ot && a.in.W >= 7) {
        HParams h;
        h.in = a.in.base; h.out = a.out.base; h.out_scale = w.out_scale; h.bias = a.bias; h.post_relu = a.post_relu;
        h.H = a.in.H; h.W = a.in.W; h.in_pitch = a.in.pitch; h.in_coff = a.in.c_off;
        h.out_pitch = a.out.pitch; h.out_coff = a.out.c_off; h.Cin = a.Cin; h.Cout = a.Cout; h.n = a.n;
        h.TW = h.W % 14 == 0 ? 14 : (h.W < 14 ? h.W : (h.W % 13 == 0 ? 13 : (h.W % 12 == 0 ? 12 : 14)));
        h.TH = h.H % 8 == 0 ? 8 : (h.H % 7 == 0 ? 7 : (h.H < 8 ? h.H : 8));
        h.tiles_x = (h.W + h.TW - 1) / h.TW;
        h.tiles_y = (h.H + h.TH - 1) / h.TH;
        h.num_tiles = a.n * h.tiles_x * h.tiles_y;
        h.chunks_per_tap = p.chunks_per_tap;
        if (it == DType::BF16) return LaunchHalo<__nv_bfloat16, __nv_bfloat16>(tm, h, stream);
        return LaunchHalo<__nv_fp8_e4m3, __nv_fp8_e4m3>(tm, h, stream);
    }
    int mode;
    if (a.pool2) mode = kModePool2;
    else if (a.R == 1 && a.S == 1 && a.stride == 1 && a.pad == 0) mode = kModeLinear;
    else mode = kModeGather;
    if (it == DType::BF16 && ot == DType::BF16) return LaunchMode<__nv_bfloat16, __nv_bfloat16>(mode, bn, tm, p, stream);
    if (it == DType::FP8 && ot == DType::FP8) return LaunchMode<__nv_fp8_e4m3, __nv_fp8_e4m3>(mode, bn, tm, p, stream);
    return cudaErrorInvalidValue;
}

}  // namespace kernels
}  // namespace b200
