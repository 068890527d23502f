// kernels_poolbn.cu — the A operand of a WIDE transition layer, materialised once: y[n,oy,ox,c] = sum over the 2x2 input
// pixels of relu(bn(x)) in the MMA element type (e4m3 or bf16), exactly the arithmetic (packed f16x2 / bf16x2 pairs, the same
// summation order) that conv1x1_tma_kernel<POOL> applies in its transform warps - results are bit-identical.
//
// Why: conv1x1_tma_kernel tiles (M tile, N tile) pairs and redoes the four-plane load + pooled transform of the A tile for
// every N tile.  For DenseNet's transition 2 (Cout 256) and transition 3 (Cout 512) that is 2x / 4x redundant work in the
// most expensive part of the kernel.  With the pooled tensor written once (25.7 MB / 12.8 MB at bs256, L2 resident) the conv
// becomes a plain TMA -> tcgen05 1x1 conv without prologue.  Part of what ONNX Runtime executes as BatchNormalization -> Relu
// -> Conv -> AveragePool nodes inside `Ort::Session::Run` (reference inference_engine/src/model.cpp:1264-1270).
#include "kernels.h"
#include "umma_ptx.cuh"

namespace b200 {
namespace kernels {

namespace {

constexpr int kPbThreads = 256;
constexpr int kPbMaxCin = 2048;

struct PbParams {
    const uint8_t* in;
    uint8_t* out;
    const float* scale;
    const float* shift;
    int relu;
    int H, W, Ho, Wo, Cin;
    int in_pitch_b, in_coff_b, out_pitch_b, out_coff_b;  // bytes
    long long total;                                     // 16-byte pieces to produce
    int ppp;                                             // pieces per pixel
};

template <typename MmaT>
__global__ void __launch_bounds__(kPbThreads) pool_bn_relu_2x2_kernel(const PbParams p) {
    using ME = MmaElem<MmaT>;
    constexpr int EPV = ME::kPerVec;
    constexpr int kPairs = EPV / 2;
    __shared__ __align__(16) uint32_t s_sc[kPbMaxCin / 2], s_sh[kPbMaxCin / 2];
    for (int i = threadIdx.x; i < (p.Cin + 1) / 2; i += kPbThreads) {
        const int c0 = 2 * i, c1 = 2 * i + 1;
        s_sc[i] = PackPair<MmaT>(p.scale[c0], c1 < p.Cin ? p.scale[c1] : 0.f);
        s_sh[i] = PackPair<MmaT>(p.shift[c0], c1 < p.Cin ? p.shift[c1] : 0.f);
    }
    GridDepLaunch();
    __syncthreads();
    GridDepWait();
    const size_t row_b = (size_t)p.W * p.in_pitch_b;
    for (long long idx = (long long)blockIdx.x * kPbThreads + threadIdx.x; idx < p.total; idx += (long long)gridDim.x * kPbThreads) {
        const int piece = (int)(idx % p.ppp);
        const long long pix = idx / p.ppp;
        const int ox = (int)(pix % p.Wo);
        const long long t = pix / p.Wo;
        const int oy = (int)(t % p.Ho);
        const long long img = t / p.Ho;
        const uint8_t* src = p.in + ((size_t)(img * p.H + 2 * oy) * p.W + 2 * ox) * p.in_pitch_b + p.in_coff_b + piece * 16;
        uint4 q[4];
        q[0] = LdgNc(src);
        q[1] = LdgNc(src + p.in_pitch_b);
        q[2] = LdgNc(src + row_b);
        q[3] = LdgNc(src + row_b + p.in_pitch_b);
        uint32_t sc[kPairs], sh[kPairs];
#pragma unroll
        for (int e = 0; e < kPairs; e += 4) {  // 16-byte shared loads (word loads at this stride were 4-way bank conflicted)
            const uint4 a = *reinterpret_cast<const uint4*>(&s_sc[piece * kPairs + e]);
            const uint4 b = *reinterpret_cast<const uint4*>(&s_sh[piece * kPairs + e]);
            sc[e] = a.x; sc[e + 1] = a.y; sc[e + 2] = a.z; sc[e + 3] = a.w;
            sh[e] = b.x; sh[e + 1] = b.y; sh[e + 2] = b.z; sh[e + 3] = b.w;
        }
        const uint4 v = p.relu ? PoolPiece<MmaT, true>(q, sc, sh) : PoolPiece<MmaT, false>(q, sc, sh);
        *reinterpret_cast<uint4*>(p.out + (size_t)pix * p.out_pitch_b + p.out_coff_b + piece * 16) = v;
    }
}

// FP32 mode: the same pass in fp32 (4 channels per 16-byte piece, (p0 + p1) + (p2 + p3) like the packed variants)
__global__ void __launch_bounds__(kPbThreads) pool_bn_relu_2x2_f32_kernel(const PbParams p) {
    __shared__ __align__(16) float s_sc[kPbMaxCin / 2], s_sh[kPbMaxCin / 2];
    for (int i = threadIdx.x; i < p.Cin; i += kPbThreads) {
        s_sc[i] = p.scale[i];
        s_sh[i] = p.shift[i];
    }
    GridDepLaunch();
    __syncthreads();
    GridDepWait();
    const size_t row_b = (size_t)p.W * p.in_pitch_b;
    for (long long idx = (long long)blockIdx.x * kPbThreads + threadIdx.x; idx < p.total; idx += (long long)gridDim.x * kPbThreads) {
        const int piece = (int)(idx % p.ppp);
        const long long pix = idx / p.ppp;
        const int ox = (int)(pix % p.Wo);
        const long long t = pix / p.Wo;
        const int oy = (int)(t % p.Ho);
        const long long img = t / p.Ho;
        const uint8_t* src = p.in + ((size_t)(img * p.H + 2 * oy) * p.W + 2 * ox) * p.in_pitch_b + p.in_coff_b + piece * 16;
        uint4 q[4];
        q[0] = LdgNc(src);
        q[1] = LdgNc(src + p.in_pitch_b);
        q[2] = LdgNc(src + row_b);
        q[3] = LdgNc(src + row_b + p.in_pitch_b);
        const float4 sc = *reinterpret_cast<const float4*>(&s_sc[piece * 4]);
        const float4 sh = *reinterpret_cast<const float4*>(&s_sh[piece * 4]);
        float y[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            y[k][0] = fmaf(__uint_as_float(q[k].x), sc.x, sh.x);
            y[k][1] = fmaf(__uint_as_float(q[k].y), sc.y, sh.y);
            y[k][2] = fmaf(__uint_as_float(q[k].z), sc.z, sh.z);
            y[k][3] = fmaf(__uint_as_float(q[k].w), sc.w, sh.w);
            if (p.relu) {
#pragma unroll
                for (int e = 0; e < 4; ++e) y[k][e] = fmaxf(y[k][e], 0.f);
            }
        }
        float4 v;
        v.x = (y[0][0] + y[1][0]) + (y[2][0] + y[3][0]);
        v.y = (y[0][1] + y[1][1]) + (y[2][1] + y[3][1]);
        v.z = (y[0][2] + y[1][2]) + (y[2][2] + y[3][2]);
        v.w = (y[0][3] + y[1][3]) + (y[2][3] + y[3][3]);
        *reinterpret_cast<float4*>(p.out + (size_t)pix * p.out_pitch_b + p.out_coff_b + piece * 16) = v;
    }
}

}  // namespace

bool PoolBnRelu2x2Supported(View in, View out) {
    if (in.dtype != out.dtype || (in.dtype != DType::FP8 && in.dtype != DType::BF16 && in.dtype != DType::F32)) return false;
    if (in.dtype == DType::F32 && in.C > kPbMaxCin / 2) return false;
    const int esz = (int)DTypeSize(in.dtype);
    if (in.H != 2 * out.H || in.W != 2 * out.W || in.C != out.C || in.C > kPbMaxCin || (in.C * esz) % 16 != 0) return false;
    if ((in.pitch * esz) % 16 || (in.c_off * esz) % 16 || (out.pitch * esz) % 16 || (out.c_off * esz) % 16) return false;
    return reinterpret_cast<uintptr_t>(in.base) % 16 == 0 && reinterpret_cast<uintptr_t>(out.base) % 16 == 0;
}

cudaError_t PoolBnRelu2x2(View in, View out, int n, const float* scale, const float* shift, bool relu, cudaStream_t stream) {
    if (!PoolBnRelu2x2Supported(in, out) || !scale || !shift) return cudaErrorInvalidValue;
    if (n <= 0) return cudaSuccess;
    const int esz = (int)DTypeSize(in.dtype);
    PbParams p;
    p.in = reinterpret_cast<const uint8_t*>(in.base);
    p.out = reinterpret_cast<uint8_t*>(out.base);
    p.scale = scale; p.shift = shift; p.relu = relu ? 1 : 0;
    p.H = in.H; p.W = in.W; p.Ho = out.H; p.Wo = out.W; p.Cin = in.C;
    p.in_pitch_b = in.pitch * esz; p.in_coff_b = in.c_off * esz;
    p.out_pitch_b = out.pitch * esz; p.out_coff_b = out.c_off * esz;
    p.ppp = in.C * esz / 16;
    p.total = (long long)n * out.H * out.W * p.ppp;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    long long blocks = (p.total + kPbThreads - 1) / kPbThreads;
    const long long cap = (long long)sms * 8;  // grid-stride: a multiple of the SM count
    if (blocks > cap) blocks = cap;
    cudaError_t e = in.dtype == DType::FP8    ? LaunchPdl(pool_bn_relu_2x2_kernel<__nv_fp8_e4m3>, (int)blocks, kPbThreads, 0, stream, p)
                    : in.dtype == DType::BF16 ? LaunchPdl(pool_bn_relu_2x2_kernel<__nv_bfloat16>, (int)blocks, kPbThreads, 0, stream, p)
                                              : LaunchPdl(pool_bn_relu_2x2_f32_kernel, (int)blocks, kPbThreads, 0, stream, p);
    CountLaunch();
    return e;
}

}  // namespace kernels
}  // namespace b200
