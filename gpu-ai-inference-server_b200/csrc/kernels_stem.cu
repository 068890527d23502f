// kernels_stem.cu — the 7x7 / stride 2 / pad 3 image stem on tcgen05, straight from the caller's fp32 NCHW batch.
//
// Replaces, for the first convolution of the network, what `Ort::Session::Run` does inside the reference
// (inference_engine/src/model.cpp:1264-1270): NCHW fp32 image -> Conv(7x7,s2) + folded BN + ReLU.
//
// Idea: the A operand of the implicit GEMM is never materialised.  A persistent CTA owns a strip of T output rows
// of one image.  Its producer warps read the 2T+5 input rows (3 fp32 planes, coalesced float4 loads), convert to
// bf16 and interleave them as 8-byte pixels (c0,c1,c2,0) into shared memory, even and odd input rows in two
// separate planes, each row 4 zero pixels + W pixels + 4 zero pixels (row pitch P = (W+8)*8 bytes).  For output
// pixel (oy, ox) and filter row r the eight pixels 2ox-4 .. 2ox+3 of input row 2oy-3+r are 64 contiguous bytes at
//     plane[r&1] + ((oy-oy0) + (r>>1)) * P + 16*ox
// i.e. 16 bytes further for the next ox and exactly P bytes further for the next oy.  With "slot" s = (oy-oy0)*(P/16)
// + ox that is the canonical NON-swizzled K-major UMMA layout (rows 16 bytes apart, SBO = 128 B per 8 rows) whose
// K pieces overlap (LBO = 16 B): one smem descriptor per (filter row, K step) describes a 128-slot x 16-element
// A tile in place.  14 tcgen05.mma (7 filter rows x 2 K steps of 4 pixels x 4 channels) per 128 slots accumulate
// into TMEM; slots with ox >= Wo (4 per output row) are junk rows that are never stored.
// Weights: bf16 [64][256], k = r*32 + (s+1)*4 + c (tap s sits at pixel s+1 of the 8-pixel window), TMA-loaded once
// (SWIZZLE_128B) and resident for the lifetime of the CTA.
//
// Warp roles (448 threads): warps 0-3 input producers, 4-11 epilogue (two per TMEM lane quarter, 32 channels each:
// tcgen05.ld -> packed scale/bias -> cvt(.relu) -> NHWC stores), 12 weight TMA, 13 MMA issuer.  Two input buffers (load of strip k+1 overlaps the MMAs of strip k)
// and four TMEM accumulators of 64 columns.
#include "kernels.h"
#include "umma_ptx.cuh"

namespace b200 {
namespace kernels {

namespace {

constexpr int kStemN = 64;                                  // MMA N (Cout padded to 64)
constexpr int kStemWChunks = 4;                             // 7 x 32 bf16 = 448 B of K -> 4 chunks of 128 B
constexpr int kStemWBytes = kStemWChunks * kStemN * 128;    // 32 KB resident weights
constexpr int kStemProducers = 128;                         // warps 0-3 fill the input planes, warps 4-11 drain accumulators
constexpr int kStemAcc = 4;                                 // TMEM accumulators (64 columns each)
constexpr int kStemSlack = 2304;                            // junk slots of the last tile may read this far past a buffer
constexpr int kStemMaxSmem = 227 * 1024;

struct SParams {
    const uint8_t* in_u8;  // raw [n][H][W][Cin] uint8 pixels (value / 255) when non-null, else:
    const float* in;  // [n][Cin][H][W] fp32
    void* out;        // NHWC
    const float* out_scale;
    const float* bias;
    int post_relu;
    int Cin, H, W, Ho, Wo, out_pitch, out_coff, Cout;
    int T;               // output rows per strip
    int strips_per_img, num_strips;
    int spr;             // slots per output row = Wo + 4
    int row_pitch;       // P
    int plane1_off;      // byte offset of the odd-row plane inside a buffer
    int buf_bytes;
    int tiles_per_strip;
    int lo_off;          // SPLIT (fp32 mode): byte offset of the residual-term planes inside a buffer
};

__device__ __forceinline__ uint32_t PackBf16(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// fp32 -> leading bf16 term (returned packed with its neighbour) and residual term, see kernels_f32x3.cu
__device__ __forceinline__ void SplitPackBf16(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = PackBf16(a, b);
    lo = PackBf16(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xFFFF0000u));
}

// SPLIT = FP32 mode (OutT = float): every pixel is stored twice, as its leading bf16 terms and as the bf16 residuals
// (x = x0 + x1), the weights likewise (w = w0 + w1, two resident tiles), and each tile runs x1*w0 + x0*w1 + x0*w0.
template <typename OutT, bool SPLIT>
__global__ void __launch_bounds__(kThreads, 1) stem_conv7x7_kernel(const __grid_constant__ CUtensorMap tmap_w, const SParams p) {
    constexpr int kWSets = SPLIT ? 2 : 1;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint8_t* s_w = smem;
    uint8_t* s_buf = smem + kWSets * kStemWBytes;
    float* s_out_scale = reinterpret_cast<float*>(s_buf + 2 * p.buf_bytes);
    float* s_bias = s_out_scale + kStemN;
    uint64_t* w_bar = reinterpret_cast<uint64_t*>(s_bias + kStemN);
    uint64_t* in_full = w_bar + 1;
    uint64_t* in_empty = in_full + 2;
    uint64_t* tmem_full = in_empty + 2;
    uint64_t* tmem_empty = tmem_full + kStemAcc;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + kStemAcc);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (warp == 12 && lane == 0) {
        MbarInit(w_bar, 1);
        for (int b = 0; b < 2; ++b) {
            MbarInit(&in_full[b], kStemProducers);
            MbarInit(&in_empty[b], 1);
        }
        for (int a = 0; a < kStemAcc; ++a) {
            MbarInit(&tmem_full[a], 1);
            MbarInit(&tmem_empty[a], 8);  // one arrive per epilogue warp
        }
        FenceBarrierInit();
        PrefetchTensorMap(&tmap_w);
    }
    if (warp == 13) TmemAlloc(tmem_slot, kStemAcc * kStemN);
    if (threadIdx.x < kStemN) {
        s_out_scale[threadIdx.x] = threadIdx.x < p.Cout ? p.out_scale[threadIdx.x] : 0.f;
        s_bias[threadIdx.x] = (p.bias && threadIdx.x < p.Cout) ? p.bias[threadIdx.x] : 0.f;
    }
    // the left/right zero pixels of every row are written here once and never again; the slack keeps junk slots finite
    for (int i = threadIdx.x; i < 2 * p.buf_bytes / 16; i += kThreads) reinterpret_cast<uint4*>(s_buf)[i] = make_uint4(0u, 0u, 0u, 0u);
    FenceProxyAsync();
    TcFenceBefore();
    __syncthreads();
    TcFenceAfter();
    const uint32_t tmem_base = *tmem_slot;
    GridDepLaunch();
    if (warp < 12) GridDepWait();

    if (warp < 4) {
        // =========================================================== input producers: fp32 NCHW -> bf16 (c0,c1,c2,0) pixels
        const int tid = threadIdx.x;
        const int groups = p.W >> 2;  // 4-pixel groups per row
        const int tasks = (2 * p.T + 5) * groups;
        const size_t plane = (size_t)p.H * p.W;
        uint32_t k = 0;
        for (int strip = blockIdx.x; strip < p.num_strips; strip += gridDim.x, ++k) {
            const uint32_t buf = k & 1u, par = ((k >> 1) & 1u) ^ 1u;
            MbarWait(&in_empty[buf], par);
            const int img = strip / p.strips_per_img, sy = strip - img * p.strips_per_img;
            const int iy_base = 2 * sy * p.T - 3;
            const float* src = p.in + (size_t)img * p.Cin * plane;
            const uint32_t b0 = SmemAddr(s_buf + buf * p.buf_bytes);
            for (int t0 = tid; t0 < tasks; t0 += 4 * kStemProducers) {
                float4 c[4][3];
                uint32_t dst[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int t = t0 + u * kStemProducers;
                    const int li = t / groups, g = t - li * groups;
                    const int iy = iy_base + li;
                    const bool ok = t < tasks && iy >= 0 && iy < p.H;
                    const float* q = src + (size_t)(ok ? iy : 0) * p.W + 4 * g;
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch) c[u][ch] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (p.in_u8) {
                        // uint8 ingestion: four HWC pixels = 4*Cin consecutive bytes; value / 255 exactly as the reference client
                        if (ok) {
                            const uint8_t* b = p.in_u8 + (((size_t)img * p.H + iy) * p.W + 4 * g) * p.Cin;
                            float v[4][3];
                            if (p.Cin == 3) {
                                const uint32_t* w = reinterpret_cast<const uint32_t*>(b);  // 12 bytes, 4-byte aligned (W % 4 == 0)
                                const uint32_t w0 = __ldg(w), w1 = __ldg(w + 1), w2 = __ldg(w + 2);
                                const uint32_t by[12] = {w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u, w0 >> 24, w1 & 255u, (w1 >> 8) & 255u,
                                                         (w1 >> 16) & 255u, w1 >> 24, w2 & 255u, (w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24};
#pragma unroll
                                for (int px = 0; px < 4; ++px)
#pragma unroll
                                    for (int ch = 0; ch < 3; ++ch) v[px][ch] = (float)by[3 * px + ch] / 255.0f;
                            } else {
#pragma unroll
                                for (int px = 0; px < 4; ++px)
#pragma unroll
                                    for (int ch = 0; ch < 3; ++ch) v[px][ch] = ch < p.Cin ? (float)b[px * p.Cin + ch] / 255.0f : 0.f;
                            }
#pragma unroll
                            for (int ch = 0; ch < 3; ++ch) c[u][ch] = make_float4(v[0][ch], v[1][ch], v[2][ch], v[3][ch]);
                        }
                    } else {
#pragma unroll
                        for (int ch = 0; ch < 3; ++ch)
                            if (ok && ch < p.Cin) c[u][ch] = __ldg(reinterpret_cast<const float4*>(q + ch * plane));
                    }
                    dst[u] = t < tasks ? b0 + (uint32_t)((li & 1) * p.plane1_off + (li >> 1) * p.row_pitch + 32 + 32 * g) : 0u;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    if (!dst[u]) continue;
                    if (SPLIT) {
                        uint32_t h[8], l[8];
                        SplitPackBf16(c[u][0].x, c[u][1].x, h[0], l[0]); SplitPackBf16(c[u][2].x, 0.f, h[1], l[1]);
                        SplitPackBf16(c[u][0].y, c[u][1].y, h[2], l[2]); SplitPackBf16(c[u][2].y, 0.f, h[3], l[3]);
                        SplitPackBf16(c[u][0].z, c[u][1].z, h[4], l[4]); SplitPackBf16(c[u][2].z, 0.f, h[5], l[5]);
                        SplitPackBf16(c[u][0].w, c[u][1].w, h[6], l[6]); SplitPackBf16(c[u][2].w, 0.f, h[7], l[7]);
                        StsV4(dst[u], make_uint4(h[0], h[1], h[2], h[3]));
                        StsV4(dst[u] + 16, make_uint4(h[4], h[5], h[6], h[7]));
                        StsV4(dst[u] + p.lo_off, make_uint4(l[0], l[1], l[2], l[3]));
                        StsV4(dst[u] + p.lo_off + 16, make_uint4(l[4], l[5], l[6], l[7]));
                        continue;
                    }
                    StsV4(dst[u], make_uint4(PackBf16(c[u][0].x, c[u][1].x), PackBf16(c[u][2].x, 0.f),
                                             PackBf16(c[u][0].y, c[u][1].y), PackBf16(c[u][2].y, 0.f)));
                    StsV4(dst[u] + 16, make_uint4(PackBf16(c[u][0].z, c[u][1].z), PackBf16(c[u][2].z, 0.f),
                                                  PackBf16(c[u][0].w, c[u][1].w), PackBf16(c[u][2].w, 0.f)));
                }
            }
            FenceProxyAsync();
            MbarArrive(&in_full[buf]);
        }
    } else if (warp < 12) {
        // =========================================================== epilogue
        // eight warps: TMEM lane quarter e = warp & 3, column half = (warp - 4) >> 2 (32 of the 64 channels each)
        const int e = warp & 3, half = (warp - 4) >> 2;
        OutT* out = reinterpret_cast<OutT*>(p.out);
        // this warp's 32 channels' scale / bias live in registers for the whole kernel: as 16-byte broadcast loads from shared memory
        // they cost 2 wavefronts each, 256 per tile, on a kernel that is bound by the MMAs' shared-memory operand fetch
        float sc[32], bi[32];
#pragma unroll
        for (int q = 0; q < 32; ++q) {
            sc[q] = s_out_scale[half * 32 + q];
            bi[q] = s_bias[half * 32 + q];
        }
        uint32_t tk = 0;
        for (int strip = blockIdx.x; strip < p.num_strips; strip += gridDim.x) {
            const int img = strip / p.strips_per_img, sy = strip - img * p.strips_per_img;
            for (int tile = 0; tile < p.tiles_per_strip; ++tile, ++tk) {
                const uint32_t acc = tk % kStemAcc, aph = (tk / kStemAcc) & 1u;
                const int s = tile * kTileM + e * 32 + lane;
                const int row = s / p.spr, ox = s - row * p.spr;
                const int oy = sy * p.T + row;
                const bool valid = row < p.T && ox < p.Wo && oy < p.Ho;
                OutT* orow = out + ((size_t)(img * p.Ho + oy) * p.Wo + ox) * p.out_pitch + p.out_coff + half * 32;
                MbarWait(&tmem_full[acc], aph);
                TcFenceAfter();
                uint32_t r[32];
                TmemLoad32(tmem_base + ((uint32_t)(e * 32) << 16) + acc * kStemN + half * 32, r);
                TmemLoadWait();
                TcFenceBefore();  // the accumulator slice is in registers: hand it back before the math and the stores
                __syncwarp();
                if (lane == 0) MbarArrive(&tmem_empty[acc]);
                if (valid) {
                    if constexpr (SPLIT) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            if (half * 32 + q * 4 >= p.Cout) continue;
                            float4 v;
                            v.x = fmaf(__uint_as_float(r[4 * q]), sc[4 * q], bi[4 * q]);
                            v.y = fmaf(__uint_as_float(r[4 * q + 1]), sc[4 * q + 1], bi[4 * q + 1]);
                            v.z = fmaf(__uint_as_float(r[4 * q + 2]), sc[4 * q + 2], bi[4 * q + 2]);
                            v.w = fmaf(__uint_as_float(r[4 * q + 3]), sc[4 * q + 3], bi[4 * q + 3]);
                            if (p.post_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                            *reinterpret_cast<float4*>(orow + q * 4) = v;
                        }
                    } else {
                        constexpr int kWords = 32 * (int)sizeof(OutT) / 4;
                        uint32_t w[kWords];
                        if (p.post_relu) EpiloguePack32<OutT, true>(r, sc, bi, w);
                        else EpiloguePack32<OutT, false>(r, sc, bi, w);
                        constexpr int kPer = 16 / (int)sizeof(OutT);  // channels per 16-byte store
#pragma unroll
                        for (int q = 0; q < kWords / 4; ++q)
                            if (half * 32 + q * kPer < p.Cout) *reinterpret_cast<uint4*>(orow + q * kPer) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
                    }
                }
                __syncwarp();
            }
        }
    } else if (warp == 12) {
        // =========================================================== weights: TMA once, resident
        if (ElectOne()) {
            MbarArriveExpectTx(w_bar, (uint32_t)(kWSets * kStemWBytes));
            for (int c = 0; c < kWSets * kStemWChunks; ++c) TmaLoad2D(s_w + c * kStemN * 128, &tmap_w, w_bar, c * 64, 0);
        }
        __syncwarp();
    } else {
        // =========================================================== MMA issuer
        constexpr uint32_t idesc = MakeInstrDesc(MmaElem<__nv_bfloat16>::kFmt, kStemN);
        const uint64_t b_base = MakeSmemDesc(SmemAddr(s_w));
        const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem_base, 0);
        MbarWait(w_bar, 0);
        uint32_t k = 0, tk = 0;
        for (int strip = blockIdx.x; strip < p.num_strips; strip += gridDim.x, ++k) {
            const uint32_t buf = k & 1u, ph = (k >> 1) & 1u;
            MbarWait(&in_full[buf], ph);
            TcFenceAfter();
            const uint64_t a_buf = MakeSmemDescNoSwizzle(SmemAddr(s_buf + buf * p.buf_bytes), 16, 128);
            for (int tile = 0; tile < p.tiles_per_strip; ++tile, ++tk) {
                const uint32_t acc = tk % kStemAcc, aph = (tk / kStemAcc) & 1u;
                MbarWait(&tmem_empty[acc], aph ^ 1u);
                TcFenceAfter();
                if (ElectOne()) {
                    const uint32_t d_addr = tmem_u + acc * kStemN;
                    const uint64_t a_tile = a_buf + (uint64_t)(tile * kTileM);  // 16 bytes per slot, address field is >>4
#pragma unroll
                    for (int pass = SPLIT ? 0 : 2; pass < 3; ++pass) {  // SPLIT: x1*w0, x0*w1, x0*w0 (smallest terms first)
                        const uint32_t a_set = pass == 0 ? (uint32_t)(p.lo_off >> 4) : 0u;
                        const uint32_t b_set = pass == 1 ? (uint32_t)(kStemWBytes >> 4) : 0u;
#pragma unroll
                        for (int r = 0; r < 7; ++r) {
                            const uint32_t arow = (uint32_t)(((r & 1) * p.plane1_off + (r >> 1) * p.row_pitch) >> 4) + a_set;
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {
                                const uint32_t boff = (uint32_t)((r >> 1) * (kStemN * 128 / 16) + (r & 1) * 4 + ks * 2) + b_set;
                                UmmaSS<0>(d_addr, a_tile + (uint64_t)(arow + 2 * ks), b_base + (uint64_t)boff, idesc,
                                          (pass > (SPLIT ? 0 : 2) || r || ks) ? 1u : 0u);
                            }
                        }
                    }
                    UmmaCommit(&tmem_full[acc]);
                    if (tile == p.tiles_per_strip - 1) UmmaCommit(&in_empty[buf]);
                }
                __syncwarp();
            }
        }
    }
    TcFenceBefore();
    __syncthreads();
    if (warp == 13) {
        TcFenceAfter();
        TmemDealloc(tmem_base, kStemAcc * kStemN);
    }
}

// Strip height and buffer geometry; returns false when even T = 1 does not fit in shared memory.
bool StemGeometry(int H, int W, SParams* p, bool split = false) {
    const int Ho = (H + 2 * 3 - 7) / 2 + 1, Wo = (W + 2 * 3 - 7) / 2 + 1;
    const int P = (W + 8) * 8;
    for (int T = Ho < 16 ? Ho : 16; T >= 1; --T) {
        const int rows0 = T + 3, rows1 = T + 2;  // even / odd plane rows of a (2T+5)-row strip
        const int set_bytes = (rows0 + rows1) * P;  // the residual planes follow the leading planes (same geometry)
        const int buf = ((split ? 2 : 1) * set_bytes + kStemSlack + 127) / 128 * 128;
        const int smem = 1024 + (split ? 2 : 1) * kStemWBytes + 2 * buf + 2 * kStemN * 4 + 256;
        if (smem > kStemMaxSmem) continue;
        p->H = H; p->W = W; p->Ho = Ho; p->Wo = Wo;
        p->T = T;
        p->strips_per_img = (Ho + T - 1) / T;
        p->spr = P / 16;
        p->row_pitch = P;
        p->plane1_off = rows0 * P;
        p->buf_bytes = buf;
        p->lo_off = split ? set_bytes : 0;
        p->tiles_per_strip = (T * p->spr + kTileM - 1) / kTileM;
        return true;
    }
    return false;
}

template <typename OutT, bool SPLIT = false>
cudaError_t LaunchStem(const CUtensorMap& tm, const SParams& p, cudaStream_t stream) {
    auto kern = stem_conv7x7_kernel<OutT, SPLIT>;
    static int sm_count[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 63;
    if (!sm_count[dev]) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemMaxSmem);
        if (e != cudaSuccess) return e;
        int n = 0;
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sm_count[dev] = n > 0 ? n : 148;
    }
    const int smem = 1024 + (SPLIT ? 2 : 1) * kStemWBytes + 2 * p.buf_bytes + 2 * kStemN * 4 + 256;
    const int grid = p.num_strips < sm_count[dev] ? p.num_strips : sm_count[dev];
    cudaError_t le = LaunchPdl(kern, grid, kThreads, smem, stream, tm, p);
    CountLaunch();
    return le;
}

}  // namespace

bool StemNchwSupported(const ConvArgs& a) {
    if (!a.stem_nchw) return false;
    if (!a.in_u8_hwc && a.in.dtype != DType::F32) return false;  // out F32 = FP32 mode (bf16-split operands)
    if (!StemFusable(a.Cin, a.Cout, a.R, a.S, a.stride, a.pad, a.in.H, a.in.W)) return false;
    if (a.pre_scale || a.pool2) return false;
    const int esz = (int)DTypeSize(a.out.dtype);
    if ((a.out.pitch * esz) % 16 != 0 || (a.out.c_off * esz) % 16 != 0) return false;
    if (!a.in_u8_hwc && reinterpret_cast<uintptr_t>(a.in.base) % 16 != 0) return false;
    if (a.in_u8_hwc && reinterpret_cast<uintptr_t>(a.in_u8_hwc) % 4 != 0) return false;
    SParams p;
    return StemGeometry(a.in.H, a.in.W, &p, a.out.dtype == DType::F32);
}

cudaError_t ConvStemNchw(const ConvArgs& a, const UmmaWeights& w, cudaStream_t stream) {
    if (!StemNchwSupported(a) || !w.tensor_map) return cudaErrorInvalidValue;
    if (a.n <= 0) return cudaSuccess;
    SParams p;
    if (!StemGeometry(a.in.H, a.in.W, &p, a.out.dtype == DType::F32)) return cudaErrorInvalidValue;
    if (p.Ho != a.out.H || p.Wo != a.out.W) return cudaErrorInvalidValue;
    p.in = reinterpret_cast<const float*>(a.in.base);
    p.in_u8 = a.in_u8_hwc;
    p.out = a.out.base;
    p.out_scale = w.out_scale;
    p.bias = a.bias;
    p.post_relu = a.post_relu;
    p.Cin = a.Cin;
    p.out_pitch = a.out.pitch; p.out_coff = a.out.c_off; p.Cout = a.Cout;
    p.num_strips = a.n * p.strips_per_img;
    const CUtensorMap& tm = *reinterpret_cast<const CUtensorMap*>(w.tensor_map);
    if (a.out.dtype == DType::F32) return LaunchStem<float, true>(tm, p, stream);
    if (a.out.dtype == DType::BF16) return LaunchStem<__nv_bfloat16>(tm, p, stream);
    return LaunchStem<__nv_fp8_e4m3>(tm, p, stream);
}

}  // namespace kernels
}  // namespace b200
