// kernels_simt.cu — SIMT kernels of the engine (sm_100a):
//   * conv_simt_f32: implicit-GEMM convolution in exact fp32 FFMA — the "FP32 reference mode" of the
//     engine and the executor of every rank-2 Gemm/MatMul.  Fused A-operand prologue (folded
//     BatchNormalization + ReLU, applied to in-bounds taps only) and bias/ReLU epilogue; reads and
//     writes channel slices of wider NHWC pixels so dense-block concats never materialise.
//   * the memory-bound kernels (layout conversion, pooling, BN+ReLU, global average pool, add, copy,
//     softmax), templated on the storage type (f32 / bf16 / e4m3) with 128-bit vector accesses
//     whenever the channel slice is 16-byte aligned.
// These replace the cuDNN/cuBLAS library calls that ONNX Runtime's CUDA EP would make inside
// `Ort::Session::Run` (reference inference_engine/src/model.cpp:1264-1270).
#include <cuda_bf16.h>
#include <cuda_fp8.h>

#include <atomic>
#include <cfloat>

#include "kernels.h"

namespace b200 {
namespace kernels {

static std::atomic<uint64_t> g_launches{0};
void CountLaunch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }
uint64_t LaunchCount() { return g_launches.load(std::memory_order_relaxed); }

namespace {

// ------------------------------------------------------------------ element helpers
template <typename T> struct Elem;
template <> struct Elem<float> {
    static constexpr int V = 4;
    __device__ static float ld(const float* p) { return *p; }
    __device__ static void st(float* p, float v) { *p = v; }
};
template <> struct Elem<__nv_bfloat16> {
    static constexpr int V = 8;
    __device__ static float ld(const __nv_bfloat16* p) { return __bfloat162float(*p); }
    __device__ static void st(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }
};
template <> struct Elem<__nv_fp8_e4m3> {
    static constexpr int V = 16;
    __device__ static float ld(const __nv_fp8_e4m3* p) { return float(*p); }
    __device__ static void st(__nv_fp8_e4m3* p, float v) { *p = __nv_fp8_e4m3(v); }
};

template <typename T>
__device__ __forceinline__ void LoadVec(const T* p, float* f) {
    constexpr int V = Elem<T>::V;
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const T* e = reinterpret_cast<const T*>(&raw);
#pragma unroll
    for (int i = 0; i < V; ++i) f[i] = Elem<T>::ld(e + i);
}
template <typename T>
__device__ __forceinline__ void StoreVec(T* p, const float* f) {
    constexpr int V = Elem<T>::V;
    uint4 raw;
    T* e = reinterpret_cast<T*>(&raw);
#pragma unroll
    for (int i = 0; i < V; ++i) Elem<T>::st(e + i, f[i]);
    *reinterpret_cast<uint4*>(p) = raw;
}

template <typename T>
bool VecOk(const View& v) {
    constexpr int V = Elem<T>::V;
    return v.C % V == 0 && v.pitch % V == 0 && v.c_off % V == 0 && (reinterpret_cast<uintptr_t>(v.base) % 16) == 0;
}

struct DView {  // device-side copy of View with typed access
    void* base;
    int C, H, W, pitch, c_off;
};
DView ToD(const View& v) { return DView{v.base, v.C, v.H, v.W, v.pitch, v.c_off}; }

#define DISPATCH_DTYPE(dt, ...)                                              \
    switch (dt) {                                                            \
        case DType::F32: { using T = float; __VA_ARGS__; break; }            \
        case DType::BF16: { using T = __nv_bfloat16; __VA_ARGS__; break; }   \
        case DType::FP8: { using T = __nv_fp8_e4m3; __VA_ARGS__; break; }    \
        default: return cudaErrorInvalidValue;                               \
    }

inline unsigned Blocks(size_t work, int threads) { return (unsigned)((work + threads - 1) / threads); }

// ------------------------------------------------------------------ layout conversion
template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, DView out, int n) {
    size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t hw = (size_t)out.H * out.W;
    if (pix >= (size_t)n * hw) return;
    size_t img = pix / hw, p = pix % hw;
    T* o = reinterpret_cast<T*>(out.base) + pix * out.pitch + out.c_off;
    for (int c = 0; c < out.C; ++c) Elem<T>::st(o + c, in[(img * out.C + c) * hw + p]);
    // zero the channel padding so padded-K MMAs see exact zeros
    if (out.c_off == 0)
        for (int c = out.C; c < out.pitch; ++c) Elem<T>::st(o + c, 0.f);
}

template <typename T>
__global__ void u8hwc_to_nhwc_kernel(const uint8_t* __restrict__ in, DView out, int n) {
    size_t pix = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= (size_t)n * out.H * out.W) return;
    T* o = reinterpret_cast<T*>(out.base) + pix * out.pitch + out.c_off;
    const uint8_t* s = in + pix * out.C;
    for (int c = 0; c < out.C; ++c) Elem<T>::st(o + c, (float)s[c] / 255.0f);  // same arithmetic as client/test_client.py: img / 255
    if (out.c_off == 0)
        for (int c = out.C; c < out.pitch; ++c) Elem<T>::st(o + c, 0.f);
}

template <typename T>
__global__ void nhwc_to_nchw_kernel(DView in, float* __restrict__ out, int n) {
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t hw = (size_t)in.H * in.W;
    size_t total = (size_t)n * in.C * hw;
    if (idx >= total) return;
    size_t p = idx % hw, c = (idx / hw) % in.C, img = idx / (hw * in.C);
    const T* src = reinterpret_cast<const T*>(in.base) + (img * hw + p) * in.pitch + in.c_off + c;
    out[idx] = Elem<T>::ld(src);
}

// ------------------------------------------------------------------ pooling
template <typename T, bool VEC, bool IS_MAX>
__global__ void pool_kernel(DView in, DView out, int n, int k, int stride, int pad, bool count_include_pad) {
    constexpr int V = VEC ? Elem<T>::V : 1;
    int cv = out.C / V;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)n * out.H * out.W * cv;
    if (idx >= total) return;
    int c = (int)(idx % cv) * V;
    size_t opix = idx / cv;
    int ow = (int)(opix % out.W);
    int oh = (int)((opix / out.W) % out.H);
    size_t img = opix / ((size_t)out.W * out.H);
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = IS_MAX ? -FLT_MAX : 0.f;
    int cnt = 0;
    for (int r = 0; r < k; ++r) {
        int ih = oh * stride - pad + r;
        if (ih < 0 || ih >= in.H) continue;
        for (int s = 0; s < k; ++s) {
            int iw = ow * stride - pad + s;
            if (iw < 0 || iw >= in.W) continue;
            const T* p = reinterpret_cast<const T*>(in.base) + ((img * in.H + ih) * in.W + iw) * in.pitch + in.c_off + c;
            float v[V];
            if (VEC) LoadVec<T>(p, v);
            else v[0] = Elem<T>::ld(p);
#pragma unroll
            for (int i = 0; i < V; ++i) acc[i] = IS_MAX ? fmaxf(acc[i], v[i]) : acc[i] + v[i];
            ++cnt;
        }
    }
    if (!IS_MAX) {
        float inv = 1.f / (float)(count_include_pad ? k * k : (cnt > 0 ? cnt : 1));
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] *= inv;
    }
    T* o = reinterpret_cast<T*>(out.base) + opix * out.pitch + out.c_off + c;
    if (VEC) StoreVec<T>(o, acc);
    else Elem<T>::st(o, acc[0]);
}

// 3x3 / stride 2 / pad 1 max pool on packed data (the stem's pool): the running maximum stays in packed half
// precision (e4m3 -> f16x2 is exact and full rate, bf16x2 has a native max), so a tap costs 2 instructions per four
// e4m3 values instead of the 8 conversions + 4 fmaxf of the generic kernel.  One thread = one 16-byte channel piece of
// one output pixel.
__device__ __forceinline__ uint32_t MaxF16x2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t MaxBf16x2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("max.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
template <typename T>
__global__ void __launch_bounds__(256) maxpool3x3s2_kernel(DView in, DView out, int n) {
    constexpr int V = Elem<T>::V;
    const int cv = out.C / V;
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)n * out.H * out.W * cv;
    if (idx >= total) return;
    const int c = (int)(idx % cv) * V;
    const size_t opix = idx / cv;
    const int ow = (int)(opix % out.W), oh = (int)((opix / out.W) % out.H);
    const size_t img = opix / ((size_t)out.W * out.H);
    const T* base = reinterpret_cast<const T*>(in.base) + img * in.H * in.W * in.pitch + in.c_off + c;
    constexpr bool kFp8 = sizeof(T) == 1;
    constexpr uint32_t kNegInf = kFp8 ? 0xFC00FC00u : 0xFF80FF80u;
    uint32_t acc[kFp8 ? 8 : 4];
#pragma unroll
    for (int i = 0; i < (kFp8 ? 8 : 4); ++i) acc[i] = kNegInf;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int ih = oh * 2 - 1 + r;
        if (ih < 0 || ih >= in.H) continue;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
            const int iw = ow * 2 - 1 + s;
            if (iw < 0 || iw >= in.W) continue;
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(base + ((size_t)ih * in.W + iw) * in.pitch));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (kFp8) {
                    uint32_t lo, hi;
                    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(lo) : "h"((unsigned short)(w[i] & 0xFFFFu)));
                    asm("cvt.rn.f16x2.e4m3x2 %0, %1;" : "=r"(hi) : "h"((unsigned short)(w[i] >> 16)));
                    acc[2 * i] = MaxF16x2(acc[2 * i], lo);
                    acc[2 * i + 1] = MaxF16x2(acc[2 * i + 1], hi);
                } else {
                    acc[i] = MaxBf16x2(acc[i], w[i]);
                }
            }
        }
    }
    uint4 o;
    if (kFp8) {
        uint32_t q[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            unsigned short lo, hi;
            asm("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(lo) : "r"(acc[2 * i]));
            asm("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(hi) : "r"(acc[2 * i + 1]));
            q[i] = (uint32_t)lo | ((uint32_t)hi << 16);
        }
        o = make_uint4(q[0], q[1], q[2], q[3]);
    } else {
        o = make_uint4(acc[0], acc[1], acc[2], acc[3]);
    }
    *reinterpret_cast<uint4*>(reinterpret_cast<T*>(out.base) + opix * out.pitch + out.c_off + c) = o;
}

// ------------------------------------------------------------------ elementwise family
// mode 0: y = relu?(x*scale+shift)   mode 1: y = a + b   mode 2: copy   mode 3: relu
template <typename TI, typename TO, bool VEC, int MODE>
__global__ void elementwise_kernel(DView a, DView b, DView out, size_t pixels, const float* __restrict__ scale,
                                   const float* __restrict__ shift, bool relu) {
    constexpr int V = VEC ? (Elem<TI>::V < Elem<TO>::V ? Elem<TI>::V : Elem<TO>::V) : 1;
    static_assert(!VEC || sizeof(TI) == sizeof(TO), "vector path needs equal element sizes");
    int cv = out.C / V;
    size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= pixels * cv) return;
    int c = (int)(idx % cv) * V;
    size_t pix = idx / cv;
    const TI* pa = reinterpret_cast<const TI*>(a.base) + pix * a.pitch + a.c_off + c;
    float x[V], y[V];
    if (VEC) LoadVec<TI>(pa, x);
    else x[0] = Elem<TI>::ld(pa);
    if (MODE == 1) {
        const TI* pb = reinterpret_cast<const TI*>(b.base) + pix * b.pitch + b.c_off + c;
        if (VEC) LoadVec<TI>(pb, y);
        else y[0] = Elem<TI>::ld(pb);
    }
#pragma unroll
    for (int i = 0; i < V; ++i) {
        float v = x[i];
        if (MODE == 0) {
            v = fmaf(v, scale[c + i], shift[c + i]);
            if (relu) v = fmaxf(v, 0.f);
        } else if (MODE == 1) {
            v += y[i];
        } else if (MODE == 3) {
            v = fmaxf(v, 0.f);
        }
        x[i] = v;
    }
    TO* po = reinterpret_cast<TO*>(out.base) + pix * out.pitch + out.c_off + c;
    if (VEC) StoreVec<TO>(po, x);
    else Elem<TO>::st(po, x[0]);
}

// ------------------------------------------------------------------ global average pool (+BN+ReLU)
// One CTA per (image, up to 64 channel groups): threadIdx.x = channel group (coalesced 16-byte pieces of a pixel),
// threadIdx.y = one of 8 pixel slices, so 8x more loads are in flight than with a thread per channel group; the
// slices are reduced through shared memory.
constexpr int kGapSlices = 8;
template <typename T, bool VEC>
__global__ void gap_kernel(DView in, float* __restrict__ out, int out_pitch, int n, const float* __restrict__ scale,
                           const float* __restrict__ shift, bool relu) {
    constexpr int V = VEC ? Elem<T>::V : 1;
    __shared__ float red[kGapSlices][64][V + 1];
    const int cv = in.C / V;
    const int g = blockIdx.y * 64 + threadIdx.x;
    const size_t img = blockIdx.x;
    const int hw = in.H * in.W;
    const bool live = g < cv;
    const int c = g * V;
    float sc[V], sh[V], acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) {
        sc[i] = (live && scale) ? scale[c + i] : 1.f;
        sh[i] = (live && shift) ? shift[c + i] : 0.f;
        acc[i] = 0.f;
    }
    if (live) {
        const T* p = reinterpret_cast<const T*>(in.base) + img * hw * (size_t)in.pitch + in.c_off + c;
#pragma unroll 4
        for (int q = threadIdx.y; q < hw; q += kGapSlices) {
            float v[V];
            if (VEC) LoadVec<T>(p + (size_t)q * in.pitch, v);
            else v[0] = Elem<T>::ld(p + (size_t)q * in.pitch);
#pragma unroll
            for (int i = 0; i < V; ++i) {
                float t = fmaf(v[i], sc[i], sh[i]);
                acc[i] += relu ? fmaxf(t, 0.f) : t;
            }
        }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) red[threadIdx.y][threadIdx.x][i] = acc[i];
    __syncthreads();
    if (threadIdx.y == 0 && live) {
        const float inv = 1.f / (float)hw;
#pragma unroll
        for (int i = 0; i < V; ++i) {
            float s = 0.f;
#pragma unroll
            for (int y = 0; y < kGapSlices; ++y) s += red[y][threadIdx.x][i];
            out[img * out_pitch + c + i] = s * inv;
        }
    }
}

// ------------------------------------------------------------------ softmax over rows
__global__ void softmax_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
    int row = blockIdx.x;
    if (row >= rows) return;
    const float* x = in + (size_t)row * cols;
    float* y = out + (size_t)row * cols;
    __shared__ float red[32];
    float m = -FLT_MAX;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) m = fmaxf(m, x[c]);
    for (int o = 16; o; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    m = red[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
    __syncthreads();
    float s = 0.f;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) s += expf(x[c] - m);
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    s = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) s += red[w];
    float inv = 1.f / s;
    for (int c = threadIdx.x; c < cols; c += blockDim.x) y[c] = expf(x[c] - m) * inv;
}

__global__ void flush_kernel(uint4* p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = make_uint4((unsigned)i, 0u, 0u, 0u);
}

__global__ void vector_add_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ r, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) r[i] = a[i] + b[i];
}

// ------------------------------------------------------------------ fp32 implicit-GEMM convolution
struct ConvP {
    const float* in;
    float* out;
    const float* w;  // [K][Cout]
    const float* bias;
    const float* pre_scale;
    const float* pre_shift;
    int pre_relu, post_relu;
    int H, W, Cin, in_pitch, in_coff;
    int Ho, Wo, Cout, out_pitch, out_coff;
    int R, S, stride, pad;
    int M, K;
    int vecA, vecB, vecC;
    // split-K (v2 kernel, small M): blockIdx.z = K slice; partial[z][m][n]; the LAST slice of a tile to finish sums them in z order
    int splits, k_per_split;
    float* partial;
    unsigned int* counters;
};

template <int BM, int BN, int BK, int TM, int TN>
__global__ void __launch_bounds__((BM / TM) * (BN / TN)) conv_simt_f32_kernel(ConvP p) {
    constexpr int NT = (BM / TM) * (BN / TN);
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < p.K; k0 += BK) {
        // ---- A tile: BM x BK gathered from NHWC input with prologue ----
        if (p.vecA) {
            constexpr int KQ = BK / 4;
            for (int e = tid; e < BM * KQ; e += NT) {
                int row = e / KQ, kq = (e % KQ) * 4;
                int m = m0 + row, k = k0 + kq;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (m < p.M && k < p.K) {
                    int c = k % p.Cin, rs = k / p.Cin;
                    int s = rs % p.S, r = rs / p.S;
                    int ow = m % p.Wo, oh = (m / p.Wo) % p.Ho, img = m / (p.Wo * p.Ho);
                    int ih = oh * p.stride - p.pad + r, iw = ow * p.stride - p.pad + s;
                    if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
                        const float* src = p.in + ((size_t)(img * p.H + ih) * p.W + iw) * p.in_pitch + p.in_coff + c;
                        v = *reinterpret_cast<const float4*>(src);
                        if (p.pre_scale) {
                            float4 sc = *reinterpret_cast<const float4*>(p.pre_scale + c);
                            float4 sh = *reinterpret_cast<const float4*>(p.pre_shift + c);
                            v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
                            v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                        }
                        if (p.pre_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
                    }
                }
                As[kq + 0][row] = v.x; As[kq + 1][row] = v.y; As[kq + 2][row] = v.z; As[kq + 3][row] = v.w;
            }
        } else {
            for (int e = tid; e < BM * BK; e += NT) {
                int row = e / BK, kk = e % BK;
                int m = m0 + row, k = k0 + kk;
                float v = 0.f;
                if (m < p.M && k < p.K) {
                    int c = k % p.Cin, rs = k / p.Cin;
                    int s = rs % p.S, r = rs / p.S;
                    int ow = m % p.Wo, oh = (m / p.Wo) % p.Ho, img = m / (p.Wo * p.Ho);
                    int ih = oh * p.stride - p.pad + r, iw = ow * p.stride - p.pad + s;
                    if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
                        v = p.in[((size_t)(img * p.H + ih) * p.W + iw) * p.in_pitch + p.in_coff + c];
                        if (p.pre_scale) v = fmaf(v, p.pre_scale[c], p.pre_shift[c]);
                        if (p.pre_relu) v = fmaxf(v, 0.f);
                    }
                }
                As[kk][row] = v;
            }
        }
        // ---- B tile: BK x BN of w[K][Cout] ----
        if (p.vecB) {
            constexpr int NQ = BN / 4;
            for (int e = tid; e < BK * NQ; e += NT) {
                int kk = e / NQ, nq = (e % NQ) * 4;
                int k = k0 + kk, nn = n0 + nq;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (k < p.K && nn < p.Cout) v = *reinterpret_cast<const float4*>(p.w + (size_t)k * p.Cout + nn);
                *reinterpret_cast<float4*>(&Bs[kk][nq]) = v;
            }
        } else {
            for (int e = tid; e < BK * BN; e += NT) {
                int kk = e / BN, nq = e % BN;
                int k = k0 + kk, nn = n0 + nq;
                Bs[kk][nq] = (k < p.K && nn < p.Cout) ? p.w[(size_t)k * p.Cout + nn] : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; i += 4) {
                float4 t = *reinterpret_cast<const float4*>(&As[kk][ty * TM + i]);
                a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
            }
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                float4 t = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN + j]);
                b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

    // ---- epilogue ----
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int m = m0 + ty * TM + i;
        if (m >= p.M) continue;
        float* orow = p.out + (size_t)m * p.out_pitch + p.out_coff;
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
            int nn = n0 + tx * TN + j;
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float t = acc[i][j + q];
                if (p.bias && nn + q < p.Cout) t += p.bias[nn + q];
                if (p.post_relu) t = fmaxf(t, 0.f);
                v[q] = t;
            }
            if (p.vecC && nn + 3 < p.Cout) {
                *reinterpret_cast<float4*>(orow + nn) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (nn + q < p.Cout) orow[nn + q] = v[q];
            }
        }
    }
}

// ------------------------------------------------------------------ fp32 implicit GEMM, second generation
// Same arithmetic as conv_simt_f32_kernel (exact fp32 FFMA, prologue on in-bounds taps only) for the vectorisable case
// (Cin, pitches, offsets multiples of 4).  The per-row pixel decode is hoisted out of the K loop, the next K slab is
// prefetched into registers while the current one is multiplied, and the M tile shrinks (128/64/32) until the grid fills
// the SMs - the old kernel ran batch-1 layers on 25-50 CTAs (13.4 ms per image).
template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) conv_simt_f32_v2_kernel(ConvP p) {
    constexpr int BK = 16;
    static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
    constexpr int kASlots = BM * 4 / 256 > 0 ? BM * 4 / 256 : 1;  // float4 slots of the A tile per thread
    constexpr bool kAHalf = BM * 4 < 256;                          // BM = 32: only the first 128 threads load A
    constexpr bool kBHalf = BN * 4 < 256;                          // BN = 32: only the first 128 threads load B
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN), ty = tid / (BN / TN);
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    // ---- loader state: fixed (row, k quad) per thread
    const int a_kq = (tid & 3) * 4;
    int a_row[kASlots], a_ih0[kASlots], a_iw0[kASlots];
    size_t a_base[kASlots];
    bool a_ok[kASlots];
#pragma unroll
    for (int i = 0; i < kASlots; ++i) {
        a_row[i] = (tid >> 2) + 64 * i;
        const int m = m0 + a_row[i];
        a_ok[i] = m < p.M && (!kAHalf || tid < BM * 4);
        const int mm = a_ok[i] ? m : 0;
        const int ow = mm % p.Wo, oh = (mm / p.Wo) % p.Ho, img = mm / (p.Wo * p.Ho);
        a_ih0[i] = oh * p.stride - p.pad;
        a_iw0[i] = ow * p.stride - p.pad;
        a_base[i] = (size_t)img * p.H * p.W;
    }
    const int b_kk = tid / (BN / 4), b_nq = (tid % (BN / 4)) * 4;
    const bool b_thread = !kBHalf || tid < BK * (BN / 4);
    const int k_begin = p.splits > 1 ? (int)blockIdx.z * p.k_per_split : 0;
    const int k_end = p.splits > 1 ? (k_begin + p.k_per_split < p.K ? k_begin + p.k_per_split : p.K) : p.K;
    float4 ra[kASlots], rb;
    auto fetch = [&](int k0) {
        const int k = k0 + a_kq;
        const int c = k % p.Cin, rs = k / p.Cin;
        const int s = rs % p.S, r = rs / p.S;
        float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
        if (p.pre_scale && k < k_end) {
            sc = __ldg(reinterpret_cast<const float4*>(p.pre_scale + c));
            sh = __ldg(reinterpret_cast<const float4*>(p.pre_shift + c));
        }
#pragma unroll
        for (int i = 0; i < kASlots; ++i) {
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            const int ih = a_ih0[i] + r, iw = a_iw0[i] + s;
            if (a_ok[i] && k < k_end && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
                v = __ldg(reinterpret_cast<const float4*>(p.in + (a_base[i] + (size_t)ih * p.W + iw) * p.in_pitch + p.in_coff + c));
                if (p.pre_scale) {
                    v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y);
                    v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
                }
                if (p.pre_relu) { v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f); }
            }
            ra[i] = v;
        }
        rb = make_float4(0.f, 0.f, 0.f, 0.f);
        const int kb = k0 + b_kk, nn = n0 + b_nq;
        if (b_thread && kb < k_end && nn < p.Cout) rb = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)kb * p.Cout + nn));
    };

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    fetch(k_begin);
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
        for (int i = 0; i < kASlots; ++i) {
            if (!kAHalf || tid < BM * 4) {
                As[a_kq + 0][a_row[i]] = ra[i].x; As[a_kq + 1][a_row[i]] = ra[i].y;
                As[a_kq + 2][a_row[i]] = ra[i].z; As[a_kq + 3][a_row[i]] = ra[i].w;
            }
        }
        if (b_thread) *reinterpret_cast<float4*>(&Bs[b_kk][b_nq]) = rb;
        __syncthreads();
        if (k0 + BK < k_end) fetch(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; j += 4) {
                float4 tt = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN + j]);
                b[j] = tt.x; b[j + 1] = tt.y; b[j + 2] = tt.z; b[j + 3] = tt.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    if (p.splits > 1) {
        // Deterministic split-K: every K slice parks its partial tile, the slice that arrives last adds them up in slice order
        // (so the result does not depend on which one that is) and runs the epilogue.
        const size_t zstride = (size_t)p.M * p.Cout;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int m = m0 + ty * TM + i;
            if (m >= p.M) continue;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int nn = n0 + tx * TN + j;
                if (nn < p.Cout) __stcg(p.partial + (size_t)blockIdx.z * zstride + (size_t)m * p.Cout + nn, acc[i][j]);
            }
        }
        __threadfence();
        __syncthreads();
        __shared__ unsigned int s_last;
        const unsigned int tile = blockIdx.y * gridDim.x + blockIdx.x;
        if (tid == 0) s_last = atomicAdd(&p.counters[tile], 1u) == (unsigned int)(p.splits - 1) ? 1u : 0u;
        __syncthreads();
        if (!s_last) return;
        __threadfence();
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            const int m = m0 + ty * TM + i;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                const int nn = n0 + tx * TN + j;
                float t = 0.f;
                if (m < p.M && nn < p.Cout)
                    for (int z = 0; z < p.splits; ++z) t += __ldcg(p.partial + (size_t)z * zstride + (size_t)m * p.Cout + nn);
                acc[i][j] = t;
            }
        }
        if (tid == 0) p.counters[tile] = 0u;  // ready for the next launch (launches on a stream are serialised)
    }
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + ty * TM + i;
        if (m >= p.M) continue;
        float* orow = p.out + (size_t)m * p.out_pitch + p.out_coff;
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
            const int nn = n0 + tx * TN + j;
            float v[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float tt = acc[i][j + q];
                if (p.bias && nn + q < p.Cout) tt += p.bias[nn + q];
                if (p.post_relu) tt = fmaxf(tt, 0.f);
                v[q] = tt;
            }
            if (p.vecC && nn + 3 < p.Cout) {
                *reinterpret_cast<float4*>(orow + nn) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    if (nn + q < p.Cout) orow[nn + q] = v[q];
            }
        }
    }
}

// ------------------------------------------------------------------ rank-2 GEMM (classifier, MatMul/Gemm graphs)
// out[m][n] = relu?(bias[n] + sum_k A[m][k] * W[k][n]); A rows are in_pitch apart, W is [K][Cout] row-major.
// 32 x 64 output tile per CTA (2 x 4 per thread), K in steps of 32 with register prefetch of the next tiles, so a
// batch-256 x 1000-class classifier spreads over 128 CTAs instead of the 32 the implicit-GEMM tiling would give.
__global__ void __launch_bounds__(256) fc_f32_kernel(ConvP p) {
    constexpr int BM = 32, BN = 64, BK = 32;
    __shared__ float As[BK][BM + 1];
    __shared__ float Bs[BK][BN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads: rows ty*2.., cols tx*4..
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    // A loader: thread -> (row = tid / 8, k quad = tid % 8); B loader: two (k row, n quad) pieces per thread
    const int a_row = tid >> 3, a_k = (tid & 7) * 4;
    const int b_k = tid >> 4, b_n = (tid & 15) * 4;
    float acc[2][4] = {};
    float ra[4], rb[2][4];
    auto fetch = [&](int k0) {
        const int m = m0 + a_row;
        if (p.vecA && m < p.M && k0 + a_k + 3 < p.K) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p.in + (size_t)m * p.in_pitch + p.in_coff + k0 + a_k));
            ra[0] = v.x; ra[1] = v.y; ra[2] = v.z; ra[3] = v.w;
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int k = k0 + a_k + q;
                ra[q] = (m < p.M && k < p.K) ? p.in[(size_t)m * p.in_pitch + p.in_coff + k] : 0.f;
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = k0 + b_k + 16 * h;
            if (p.vecB && k < p.K && n0 + b_n + 3 < p.Cout) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(p.w + (size_t)k * p.Cout + n0 + b_n));
                rb[h][0] = v.x; rb[h][1] = v.y; rb[h][2] = v.z; rb[h][3] = v.w;
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int n = n0 + b_n + q;
                    rb[h][q] = (k < p.K && n < p.Cout) ? p.w[(size_t)k * p.Cout + n] : 0.f;
                }
            }
        }
    };
    fetch(0);
    for (int k0 = 0; k0 < p.K; k0 += BK) {
#pragma unroll
        for (int q = 0; q < 4; ++q) As[a_k + q][a_row] = ra[q];
#pragma unroll
        for (int h = 0; h < 2; ++h)
            *reinterpret_cast<float4*>(&Bs[b_k + 16 * h][b_n]) = make_float4(rb[h][0], rb[h][1], rb[h][2], rb[h][3]);
        __syncthreads();
        if (k0 + BK < p.K) fetch(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float a0 = As[kk][ty * 2], a1 = As[kk][ty * 2 + 1];
            const float4 b = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
            acc[0][0] = fmaf(a0, b.x, acc[0][0]); acc[0][1] = fmaf(a0, b.y, acc[0][1]);
            acc[0][2] = fmaf(a0, b.z, acc[0][2]); acc[0][3] = fmaf(a0, b.w, acc[0][3]);
            acc[1][0] = fmaf(a1, b.x, acc[1][0]); acc[1][1] = fmaf(a1, b.y, acc[1][1]);
            acc[1][2] = fmaf(a1, b.z, acc[1][2]); acc[1][3] = fmaf(a1, b.w, acc[1][3]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + ty * 2 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.Cout) continue;
            float v = acc[i][j] + (p.bias ? p.bias[n] : 0.f);
            if (p.post_relu) v = fmaxf(v, 0.f);
            p.out[(size_t)m * p.out_pitch + p.out_coff + n] = v;
        }
    }
}

// Same product with a three-deep cp.async pipeline (the register-prefetch kernel above keeps ONE 32-step K slab in flight and is
// bound by global-load latency: 41 us for batch 256 x 1000 classes x K 1024).  Every output is still accumulated by one thread
// with k ascending, so the result is bit-identical.  Needs 16-byte aligned rows (vecA, vecB) and K % 32 == 0.
__device__ __forceinline__ void FcCpAsync16(float* smem_dst, const float* src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(src), "r"(valid ? 16 : 0) : "memory");
}
__global__ void __launch_bounds__(256) fc_f32_pipelined_kernel(ConvP p) {
    constexpr int BM = 32, BN = 64, BK = 32, ST = 3;
    __shared__ __align__(16) float As[ST][BM][BK + 4];
    __shared__ __align__(16) float Bs[ST][BK][BN];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int a_row = tid >> 3, a_k = (tid & 7) * 4;   // A: 32 rows x 8 quads
    const int b_k = tid >> 4, b_n = (tid & 15) * 4;    // B: two (k row, n quad) pieces per thread
    const int nk = p.K / BK;
    auto issue = [&](int it) {
        if (it < nk) {
            const int k0 = it * BK, st = it % ST;
            const int m = m0 + a_row;
            const bool ok = m < p.M;
            FcCpAsync16(&As[st][a_row][a_k], ok ? p.in + (size_t)m * p.in_pitch + p.in_coff + k0 + a_k : p.in, ok);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int k = k0 + b_k + 16 * h;
                const bool okb = n0 + b_n + 3 < p.Cout;
                FcCpAsync16(&Bs[st][b_k + 16 * h][b_n], okb ? p.w + (size_t)k * p.Cout + n0 + b_n : p.w, okb);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    float acc[2][4] = {};
#pragma unroll
    for (int s = 0; s < ST - 1; ++s) issue(s);
    for (int it = 0; it < nk; ++it) {
        asm volatile("cp.async.wait_group %0;" ::"n"(ST - 2) : "memory");
        __syncthreads();               // slab `it` has landed for every thread; slab it-1's buffer is free again
        issue(it + ST - 1);
        const int st = it % ST;
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            const float a0 = As[st][ty * 2][kk], a1 = As[st][ty * 2 + 1][kk];
            const float4 b = *reinterpret_cast<const float4*>(&Bs[st][kk][tx * 4]);
            acc[0][0] = fmaf(a0, b.x, acc[0][0]); acc[0][1] = fmaf(a0, b.y, acc[0][1]);
            acc[0][2] = fmaf(a0, b.z, acc[0][2]); acc[0][3] = fmaf(a0, b.w, acc[0][3]);
            acc[1][0] = fmaf(a1, b.x, acc[1][0]); acc[1][1] = fmaf(a1, b.y, acc[1][1]);
            acc[1][2] = fmaf(a1, b.z, acc[1][2]); acc[1][3] = fmaf(a1, b.w, acc[1][3]);
        }
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int m = m0 + ty * 2 + i;
        if (m >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= p.Cout) continue;
            float v = acc[i][j] + (p.bias ? p.bias[n] : 0.f);
            if (p.post_relu) v = fmaxf(v, 0.f);
            p.out[(size_t)m * p.out_pitch + p.out_coff + n] = v;
        }
    }
}

}  // namespace

// ====================================================================== launchers
cudaError_t ConvSimtF32(const ConvArgs& a, const float* w_kn, cudaStream_t stream) {
    if (a.in.dtype != DType::F32 || a.out.dtype != DType::F32 || a.pool2) return cudaErrorInvalidValue;
    ConvP p;
    p.in = (const float*)a.in.base; p.out = (float*)a.out.base; p.w = w_kn; p.bias = a.bias;
    p.pre_scale = a.pre_scale; p.pre_shift = a.pre_shift; p.pre_relu = a.pre_relu; p.post_relu = a.post_relu;
    p.H = a.in.H; p.W = a.in.W; p.Cin = a.Cin; p.in_pitch = a.in.pitch; p.in_coff = a.in.c_off;
    p.Ho = a.out.H; p.Wo = a.out.W; p.Cout = a.Cout; p.out_pitch = a.out.pitch; p.out_coff = a.out.c_off;
    p.R = a.R; p.S = a.S; p.stride = a.stride; p.pad = a.pad;
    p.M = a.n * p.Ho * p.Wo; p.K = a.R * a.S * a.Cin;
    p.splits = 1; p.k_per_split = p.K; p.partial = nullptr; p.counters = nullptr;
    p.vecA = (a.Cin % 4 == 0 && a.in.pitch % 4 == 0 && a.in.c_off % 4 == 0 && ((uintptr_t)p.in % 16) == 0 &&
              (!a.pre_scale || (((uintptr_t)a.pre_scale % 16) == 0 && ((uintptr_t)a.pre_shift % 16) == 0)));
    p.vecB = (a.Cout % 4 == 0 && ((uintptr_t)w_kn % 16) == 0);
    p.vecC = (a.out.pitch % 4 == 0 && a.out.c_off % 4 == 0 && ((uintptr_t)p.out % 16) == 0);
    if (p.M <= 0) return cudaSuccess;
    if (a.R == 1 && a.S == 1 && a.in.H == 1 && a.in.W == 1 && p.Ho == 1 && p.Wo == 1 && !a.pre_scale) {
        dim3 grid((p.M + 31) / 32, (a.Cout + 63) / 64);
        if (p.vecA && p.vecB && p.K % 32 == 0 && p.K >= 128) fc_f32_pipelined_kernel<<<grid, 256, 0, stream>>>(p);
        else fc_f32_kernel<<<grid, 256, 0, stream>>>(p);
    } else if (p.vecA && p.vecB && p.K % 4 == 0) {
        // second-generation kernel: shrink the M tile until the grid covers the SMs
        int sms = 148;
        {
            int dev = 0;
            cudaGetDevice(&dev);
            int nsm = 0;
            if (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && nsm > 0) sms = nsm;
        }
        const bool narrow = a.Cout <= 32;
        const int bn = narrow ? 32 : 64;
        const int ntile = (a.Cout + bn - 1) / bn;
        auto ctas = [&](int bm) { return ((p.M + bm - 1) / bm) * ntile; };
        // Small batches leave most SMs without a tile and every tile with a long serial K loop (batch 1, 3x3 conv of a 7x7 image:
        // ONE CTA walks K = 1152): cut K into slices across blockIdx.z until ~2 CTAs per SM exist.
        const int bm_small = narrow ? 64 : 32;
        unsigned int gz = 1;
        if (a.splitk_scratch && ctas(bm_small) < sms) {
            const int tiles = ctas(bm_small);
            int want = (2 * sms + tiles - 1) / tiles;
            const int max_by_k = p.K / 64;  // at least four 16-step slabs per slice
            if (want > max_by_k) want = max_by_k;
            if (want > 16) want = 16;
            const size_t counters_bytes = 4096 * sizeof(unsigned int);
            while (want > 1 && (size_t)want * p.M * p.Cout * sizeof(float) + counters_bytes > a.splitk_bytes) --want;
            if (want > 1 && tiles <= 4096) {
                p.k_per_split = ((p.K + want - 1) / want + 15) / 16 * 16;
                p.splits = (p.K + p.k_per_split - 1) / p.k_per_split;
                p.counters = reinterpret_cast<unsigned int*>(a.splitk_scratch);
                p.partial = reinterpret_cast<float*>(reinterpret_cast<char*>(a.splitk_scratch) + counters_bytes);
                gz = (unsigned int)p.splits;
                if (gz <= 1) { p.splits = 1; gz = 1; }
            }
        }
        if (narrow) {
            if (ctas(128) >= 2 * sms) conv_simt_f32_v2_kernel<128, 32, 4, 4><<<dim3((p.M + 127) / 128, ntile), 256, 0, stream>>>(p);
            else conv_simt_f32_v2_kernel<64, 32, 2, 4><<<dim3((p.M + 63) / 64, ntile, gz), 256, 0, stream>>>(p);
        } else {
            if (ctas(128) >= 2 * sms) conv_simt_f32_v2_kernel<128, 64, 8, 4><<<dim3((p.M + 127) / 128, ntile), 256, 0, stream>>>(p);
            else if (ctas(64) >= sms) conv_simt_f32_v2_kernel<64, 64, 4, 4><<<dim3((p.M + 63) / 64, ntile), 256, 0, stream>>>(p);
            else conv_simt_f32_v2_kernel<32, 64, 2, 4><<<dim3((p.M + 31) / 32, ntile, gz), 256, 0, stream>>>(p);
        }
    } else if (a.Cout <= 32) {
        constexpr int BM = 128, BN = 32;
        dim3 grid((p.M + BM - 1) / BM, (a.Cout + BN - 1) / BN);
        conv_simt_f32_kernel<BM, BN, 16, 4, 4><<<grid, 256, 0, stream>>>(p);
    } else {
        constexpr int BM = 128, BN = 64;
        dim3 grid((p.M + BM - 1) / BM, (a.Cout + BN - 1) / BN);
        conv_simt_f32_kernel<BM, BN, 16, 8, 4><<<grid, 256, 0, stream>>>(p);
    }
    CountLaunch();
    return cudaGetLastError();
}

cudaError_t NchwToNhwc(const float* in, View out, int n, cudaStream_t stream) {
    size_t pixels = (size_t)n * out.H * out.W;
    if (!pixels) return cudaSuccess;
    DISPATCH_DTYPE(out.dtype, (nchw_to_nhwc_kernel<T><<<Blocks(pixels, 256), 256, 0, stream>>>(in, ToD(out), n)));
    CountLaunch();
    return cudaGetLastError();
}

cudaError_t U8HwcToNhwc(const uint8_t* in, View out, int n, cudaStream_t stream) {
    size_t pixels = (size_t)n * out.H * out.W;
    if (!pixels) return cudaSuccess;
    DISPATCH_DTYPE(out.dtype, (u8hwc_to_nhwc_kernel<T><<<Blocks(pixels, 256), 256, 0, stream>>>(in, ToD(out), n)));
    CountLaunch();
    return cudaGetLastError();
}

cudaError_t NhwcToNchw(View in, float* out, int n, cudaStream_t stream) {
    size_t total = (size_t)n * in.C * in.H * in.W;
    if (!total) return cudaSuccess;
    DISPATCH_DTYPE(in.dtype, (nhwc_to_nchw_kernel<T><<<Blocks(total, 256), 256, 0, stream>>>(ToD(in), out, n)));
    CountLaunch();
    return cudaGetLastError();
}

template <bool IS_MAX>
static cudaError_t PoolImpl(View in, View out, int n, int k, int stride, int pad, bool cip, cudaStream_t stream) {
    if (in.dtype != out.dtype || in.C != out.C) return cudaErrorInvalidValue;
    size_t opix = (size_t)n * out.H * out.W;
    if (!opix) return cudaSuccess;
    if (IS_MAX && k == 3 && stride == 2 && pad == 1 && in.dtype != DType::F32) {
        if (in.dtype == DType::FP8 && VecOk<__nv_fp8_e4m3>(in) && VecOk<__nv_fp8_e4m3>(out)) {
            maxpool3x3s2_kernel<__nv_fp8_e4m3><<<Blocks(opix * (out.C / 16), 256), 256, 0, stream>>>(ToD(in), ToD(out), n);
            CountLaunch();
            return cudaGetLastError();
        }
        if (in.dtype == DType::BF16 && VecOk<__nv_bfloat16>(in) && VecOk<__nv_bfloat16>(out)) {
            maxpool3x3s2_kernel<__nv_bfloat16><<<Blocks(opix * (out.C / 8), 256), 256, 0, stream>>>(ToD(in), ToD(out), n);
            CountLaunch();
            return cudaGetLastError();
        }
    }
    DISPATCH_DTYPE(in.dtype, {
        if (VecOk<T>(in) && VecOk<T>(out))
            pool_kernel<T, true, IS_MAX><<<Blocks(opix * (out.C / Elem<T>::V), 256), 256, 0, stream>>>(ToD(in), ToD(out), n, k, stride, pad, cip);
        else
            pool_kernel<T, false, IS_MAX><<<Blocks(opix * out.C, 256), 256, 0, stream>>>(ToD(in), ToD(out), n, k, stride, pad, cip);
    });
    CountLaunch();
    return cudaGetLastError();
}
cudaError_t MaxPool(View in, View out, int n, int k, int stride, int pad, cudaStream_t stream) {
    return PoolImpl<true>(in, out, n, k, stride, pad, false, stream);
}
cudaError_t AvgPool(View in, View out, int n, int k, int stride, int pad, bool count_include_pad, cudaStream_t stream) {
    return PoolImpl<false>(in, out, n, k, stride, pad, count_include_pad, stream);
}

template <int MODE>
static cudaError_t ElementwiseImpl(View a, View b, View out, int n, const float* scale, const float* shift, bool relu,
                                   cudaStream_t stream) {
    if (a.C != out.C || a.H != out.H || a.W != out.W) return cudaErrorInvalidValue;
    size_t pixels = (size_t)n * out.H * out.W;
    if (!pixels) return cudaSuccess;
    if (a.dtype == out.dtype) {
        DISPATCH_DTYPE(a.dtype, {
            bool vec = VecOk<T>(a) && VecOk<T>(out) && (MODE != 1 || VecOk<T>(b));
            if (vec)
                elementwise_kernel<T, T, true, MODE><<<Blocks(pixels * (out.C / Elem<T>::V), 256), 256, 0, stream>>>(
                    ToD(a), ToD(b), ToD(out), pixels, scale, shift, relu);
            else
                elementwise_kernel<T, T, false, MODE><<<Blocks(pixels * out.C, 256), 256, 0, stream>>>(
                    ToD(a), ToD(b), ToD(out), pixels, scale, shift, relu);
        });
    } else if (out.dtype == DType::F32) {
        DISPATCH_DTYPE(a.dtype, (elementwise_kernel<T, float, false, MODE><<<Blocks(pixels * out.C, 256), 256, 0, stream>>>(
                                    ToD(a), ToD(b), ToD(out), pixels, scale, shift, relu)));
    } else {
        return cudaErrorInvalidValue;
    }
    CountLaunch();
    return cudaGetLastError();
}
cudaError_t BnRelu(View in, View out, int n, const float* scale, const float* shift, bool relu, cudaStream_t stream) {
    return ElementwiseImpl<0>(in, in, out, n, scale, shift, relu, stream);
}
cudaError_t AddTensors(View a, View b, View out, int n, cudaStream_t stream) {
    if (a.dtype != b.dtype) return cudaErrorInvalidValue;
    return ElementwiseImpl<1>(a, b, out, n, nullptr, nullptr, false, stream);
}
cudaError_t CopyChannels(View in, View out, int n, cudaStream_t stream) {
    return ElementwiseImpl<2>(in, in, out, n, nullptr, nullptr, false, stream);
}
cudaError_t ReluTensor(View in, View out, int n, cudaStream_t stream) {
    return ElementwiseImpl<3>(in, in, out, n, nullptr, nullptr, false, stream);
}

cudaError_t GlobalAvgPool(View in, float* out, int out_pitch, int n, const float* scale, const float* shift, bool relu,
                          cudaStream_t stream) {
    if (!n) return cudaSuccess;
    DISPATCH_DTYPE(in.dtype, {
        if (VecOk<T>(in))
            gap_kernel<T, true><<<dim3((unsigned)n, (unsigned)((in.C / Elem<T>::V + 63) / 64)), dim3(64, kGapSlices), 0, stream>>>(ToD(in), out, out_pitch, n, scale, shift, relu);
        else
            gap_kernel<T, false><<<dim3((unsigned)n, (unsigned)((in.C + 63) / 64)), dim3(64, kGapSlices), 0, stream>>>(ToD(in), out, out_pitch, n, scale, shift, relu);
    });
    CountLaunch();
    return cudaGetLastError();
}

cudaError_t SoftmaxRows(const float* in, float* out, int rows, int cols, cudaStream_t stream) {
    if (!rows) return cudaSuccess;
    softmax_kernel<<<rows, 256, 0, stream>>>(in, out, rows, cols);
    CountLaunch();
    return cudaGetLastError();
}

// ---- on-device softmax + top-k: what the Go handler does per request with a full sort of the
// 1000 logits (reference server/main.go:744-786).  One warp per row: softmax statistics (max, sum of exp), then k rounds of a
// warp-wide argmax over the entries ranked below the previous pick (value descending, lowest index first among equals).
__global__ void __launch_bounds__(256) topk_rows_kernel(const float* __restrict__ in, int rows, int cols, int k, int softmax,
                                                        int* __restrict__ idx_out, float* __restrict__ val_out) {
    const int row = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* x = in + (size_t)row * cols;
    float mx = -INFINITY;
    for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, x[c]);
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
    if (softmax) {
        for (int c = lane; c < cols; c += 32) sum += expf(x[c] - mx);
#pragma unroll
        for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    }
    float prev_v = INFINITY;
    int prev_i = -1;
    for (int r = 0; r < k; ++r) {
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int c = lane; c < cols; c += 32) {
            const float v = x[c];
            const bool eligible = v < prev_v || (v == prev_v && c > prev_i);
            if (eligible && (v > bv || (v == bv && c < bi))) { bv = v; bi = c; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) {
            const bool found = bi != 0x7fffffff;
            idx_out[(size_t)row * k + r] = found ? bi : -1;
            val_out[(size_t)row * k + r] = !found ? 0.f : softmax ? expf(bv - mx) / sum : bv;
        }
        prev_v = bv;
        prev_i = bi;
    }
}

cudaError_t TopKRows(const float* in, int rows, int cols, int k, bool softmax, int* idx_out, float* val_out, cudaStream_t stream) {
    if (rows <= 0 || k <= 0) return cudaSuccess;
    if (cols <= 0) return cudaErrorInvalidValue;
    const int warps_per_block = 8;
    topk_rows_kernel<<<(rows + warps_per_block - 1) / warps_per_block, warps_per_block * 32, 0, stream>>>(in, rows, cols, k, softmax ? 1 : 0, idx_out, val_out);
    CountLaunch();
    return cudaGetLastError();
}

cudaError_t FlushL2(void* scratch, size_t bytes, cudaStream_t stream) {
    flush_kernel<<<148 * 8, 256, 0, stream>>>((uint4*)scratch, bytes / 16);
    return cudaGetLastError();
}

cudaError_t VectorAddF32(const float* a, const float* b, float* out, size_t n, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    vector_add_kernel<<<Blocks(n, 256), 256, 0, stream>>>(a, b, out, n);
    CountLaunch();
    return cudaGetLastError();
}

}  // namespace kernels
}  // namespace b200
