// kernels.h — host-callable launchers of the engine's CUDA kernels (sm_100a only).
// All pointers are device pointers; every launcher enqueues on `stream` and returns the CUDA
// status of the launch.  Element types are described by b200::DType.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "plan.h"

namespace b200 {
namespace kernels {

// Every launcher bumps this process-wide counter (exported as B200KernelLaunchCount()).
void CountLaunch(int n = 1);
uint64_t LaunchCount();

// A strided NHWC view: element (n,h,w,c) lives at base[((n*H + h)*W + w)*pitch + c_off + c].
struct View {
    void* base = nullptr;
    DType dtype = DType::F32;
    int C = 0, H = 1, W = 1;
    int pitch = 0, c_off = 0;
};

struct ConvArgs {
    View in, out;
    int n = 0;  // batch
    int R = 1, S = 1, stride = 1, pad = 0;
    int Cin = 0, Cout = 0;
    const float* pre_scale = nullptr;  // [Cin]  (x*scale+shift before the conv, only on in-bounds taps)
    const float* pre_shift = nullptr;
    bool pre_relu = false;
    const float* bias = nullptr;  // [Cout]
    const float* h_bias = nullptr;  // host copy of `bias` (optional; kernels that take their constants as kernel parameters)
    bool post_relu = false;
    bool pool2 = false;  // A rows are 2x2 averages of (prologue-transformed) input pixels
    void* splitk_scratch = nullptr;  // fp32 SIMT path: partial sums + tile counters for split-K on small batches (optional)
    size_t splitk_bytes = 0;
    float out_mul = 1.f; // extra factor on the per-channel output scale (0.25 when the caller pooled the A operand itself)
    bool stem_nchw = false;  // `in` is the caller's fp32 NCHW image batch (7x7/s2/p3 stem, see kernels_stem.cu)
    const uint8_t* in_u8_hwc = nullptr;  // stem_nchw only: read raw uint8 [n][H][W][C] pixels instead (value / 255)
};

// A CUtensorMap by value (128 bytes, 64-byte aligned) without pulling <cuda.h> into every translation unit.
struct alignas(64) TensorMap { unsigned char bytes[128]; };
// Tiled TMA descriptor: dims/box innermost first, strides_bytes for dims 1..rank-1.  Returns 0 on success.
int MakeTensorMap(TensorMap* out, const void* base, int elem_bytes, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, bool swizzle128);

// ---- fp32 SIMT path ("FP32 reference mode"; also every rank-2 GEMM) ----
// w_kn: fp32 [K = R*S*Cin][Cout]
cudaError_t ConvSimtF32(const ConvArgs& a, const float* w_kn, cudaStream_t stream);

// ---- tcgen05 path (bf16 / e4m3 operands, fp32 accumulate in TMEM) ----
struct UmmaWeights {
    const void* w = nullptr;        // [Cout_pad][K_pad] in the MMA element type, K-major, zero padded
    const float* out_scale = nullptr;  // [Cout] per-output-channel dequant scale (1.0 for bf16)
    const float* h_out_scale = nullptr;  // host copy of `out_scale` (optional)
    int K_pad = 0;                  // padded K (multiple of the 128-byte K chunk)
    int Cout_pad = 0;
    void* tensor_map = nullptr;     // device-visible CUtensorMap* (host memory, passed by value at launch)
};
cudaError_t ConvUmma(const ConvArgs& a, const UmmaWeights& w, cudaStream_t stream);
// Host-side packing parameters for the tcgen05 path.
int UmmaKChunkElems(DType mma_dtype);                       // elements per 128-byte K chunk
int UmmaPaddedCin(int Cin, int R, int S, DType mma_dtype);  // per-tap channel padding used by the A loader
bool UmmaSupported(const ConvArgs& a);
// 7x7/s2/p3 image stem straight from fp32 NCHW (kernels_stem.cu); `tensor_map` covers bf16 weights packed
// [64][256] with k = r*32 + (s+1)*4 + c.
cudaError_t ConvStemNchw(const ConvArgs& a, const UmmaWeights& w, cudaStream_t stream);
bool StemNchwSupported(const ConvArgs& a);

// ---- FP32 mode on tcgen05: fp32 activations in HBM, operands split into two bf16 terms on the fly, three MMAs per product
// (kernels_f32x3.cu).  1x1/s1 (Cin, Cout multiples of 32) and 3x3/s1/p1 (Cin multiple of 32, Cout <= 32).
// `w.tensor_map` covers bf16 weights packed [F32x3PackedRows()][F32x3PackedK()], every weight as two terms w0 = bf16(w),
// w1 = bf16(w - w0) placed by F32x3WeightPos(); TMA box {64, F32x3TileN()}.
bool ConvF32x3Supported(const ConvArgs& a);
int F32x3TileN(const ConvArgs& a);
int F32x3PackedRows(const ConvArgs& a);
int F32x3PackedK(const ConvArgs& a);
void F32x3WeightPos(const ConvArgs& a, int o, int tap, int c, int term, int* row, int* col);
cudaError_t ConvF32x3(const ConvArgs& a, const UmmaWeights& w, cudaStream_t stream);

// 3x3/s1/p1 conv of a 128-byte-per-pixel bottleneck into <= 32 channels, patches loaded by 4-D TMA (kernels_conv3x3.cu)
bool Conv3x3TmaSupported(const ConvArgs& a);
cudaError_t Conv3x3Tma(const ConvArgs& a, const UmmaWeights& w, cudaStream_t stream);

// 1x1/s1 conv with TMA loads/stores and an in-place BN+ReLU transform of the landed A tile (kernels_conv1x1.cu)
bool Conv1x1TmaSupported(const ConvArgs& a);
cudaError_t Conv1x1Tma(const ConvArgs& a, const UmmaWeights& w, cudaStream_t stream);

// Wide transition layers: sum over 2x2 of relu(bn(x)) written once as the MMA element type (kernels_poolbn.cu); the conv that
// follows runs without prologue and with ConvArgs::out_mul = 0.25.  Bit-identical to the pooled transform of Conv1x1Tma.
bool PoolBnRelu2x2Supported(View in, View out);
cudaError_t PoolBnRelu2x2(View in, View out, int n, const float* scale, const float* shift, bool relu, cudaStream_t stream);

// ---- a whole dense block in one persistent kernel (kernels_dense.cu; e4m3, whole images per CTA) ----
struct DenseLayerDesc {          // one BN-ReLU-Conv1x1(->128)-BN-ReLU-Conv3x3(->32) layer; lives in device memory
    TensorMap w1;                // conv1 weights [128][K_pad] e4m3, box {128, 128}, SWIZZLE_128B
    TensorMap w2;                // conv2 weights [32][9*128] e4m3 seen as {128 B, 32 rows, 9 taps}, box = all of it, SWIZZLE_128B
    const uint32_t* pre_scale;   // folded BN1 scale/shift as packed f16x2 pairs, Cin/2 words each
    const uint32_t* pre_shift;
    const float* s1;             // conv1 epilogue: per-channel dequant scale, bias (BN2 folded) [128]
    const float* b1;
    const float* s2;             // conv2 epilogue [32]
    const float* b2;
    int Cin, c_off_out;          // input channels [0, Cin); the 32 outputs go to channels [c_off_out, c_off_out + 32)
    int pre_relu, relu1, relu2;
    int pad_[3];
};
struct DenseBlockArgs {
    const DenseLayerDesc* layers_dev = nullptr;
    int num_layers = 0;
    void* buf = nullptr;         // block buffer (NHWC e4m3), channel offset 0 of the first layer's input
    int pitch = 0, n = 0, H = 0, W = 0;
};
bool DenseBlockGeometry(int H, int W, int* images_per_cta, int* m_tiles);
cudaError_t DenseBlockFp8(const DenseBlockArgs& a, cudaStream_t stream);
// ---- one dense layer (1x1 -> 128 -> 3x3 -> 32) of the 56x56 / 28x28 blocks as a streaming kernel (kernels_dense_stream.cu; e4m3):
// the bottleneck tensor stays in shared memory, the conv1 A operand goes through tensor memory.  Per-channel constants are HOST
// arrays (they travel as kernel parameters).
struct DenseLayerStreamArgs {
    const void* w1_map = nullptr;   // host CUtensorMap*: conv1 weights [128][K_pad] e4m3, box {128, 128}
    const void* w2_map = nullptr;   // host CUtensorMap*: conv2 weights [32][9*128] e4m3, box {128, 32}
    void* buf = nullptr;            // block buffer (NHWC e4m3); at least one pixel of readable memory in front of it
    int pitch = 0, n = 0, H = 0, W = 0;
    int Cin = 0, c_off_out = 0;
    int pre_relu = 0, relu1 = 0, relu2 = 0;
    const float* pre_scale = nullptr;  // [Cin] folded BN1 (host)
    const float* pre_shift = nullptr;
    const float* s1 = nullptr;         // [128] conv1 epilogue scale / bias (host; bias may be null)
    const float* b1 = nullptr;
    const float* s2 = nullptr;         // [32] conv2 epilogue (host)
    const float* b2 = nullptr;
};
bool DenseLayerStreamSupported(int H, int W, int Cin, int pitch);
cudaError_t DenseLayerStreamFp8(const DenseLayerStreamArgs& a, cudaStream_t stream);
// ---- memory-bound kernels (templated on element type inside) ----
cudaError_t NchwToNhwc(const float* in, View out, int n, cudaStream_t stream);
// uint8 ingestion: raw [n][H][W][C] uint8 pixels -> value / 255 in the internal NHWC layout
cudaError_t U8HwcToNhwc(const uint8_t* in, View out, int n, cudaStream_t stream);
cudaError_t NhwcToNchw(View in, float* out, int n, cudaStream_t stream);
cudaError_t MaxPool(View in, View out, int n, int k, int stride, int pad, cudaStream_t stream);
cudaError_t AvgPool(View in, View out, int n, int k, int stride, int pad, bool count_include_pad, cudaStream_t stream);
cudaError_t BnRelu(View in, View out, int n, const float* scale, const float* shift, bool relu, cudaStream_t stream);
cudaError_t GlobalAvgPool(View in, float* out, int out_pitch, int n, const float* scale, const float* shift, bool relu,
                          cudaStream_t stream);
cudaError_t AddTensors(View a, View b, View out, int n, cudaStream_t stream);
cudaError_t CopyChannels(View in, View out, int n, cudaStream_t stream);
cudaError_t ReluTensor(View in, View out, int n, cudaStream_t stream);
cudaError_t SoftmaxRows(const float* in, float* out, int rows, int cols, cudaStream_t stream);
// top-k of every row (value descending, lowest index first among equals); `softmax`: values are softmax probabilities
cudaError_t TopKRows(const float* in, int rows, int cols, int k, bool softmax, int* idx_out, float* val_out, cudaStream_t stream);
cudaError_t FlushL2(void* scratch, size_t bytes, cudaStream_t stream);
cudaError_t VectorAddF32(const float* a, const float* b, float* out, size_t n, cudaStream_t stream);

}  // namespace kernels
}  // namespace b200
