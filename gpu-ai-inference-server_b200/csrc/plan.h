// plan.h — static launch plan: the lowering of an ONNX graph to a fixed sequence of fused steps
// over a pre-allocated activation arena.  Pure host code (no CUDA), so it is unit-testable on a
// CPU-only box.  This is the engine's replacement for ONNX Runtime's session construction
// (reference inference_engine/src/model.cpp:825-903: Ort::Session ctor + ORT_ENABLE_ALL graph
// optimisation) for the operator set of SURVEY.md §8 a10.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <unordered_map>
#include <vector>

#include "onnx_wire.h"

namespace b200 {

enum class Precision { FP32 = 0, BF16 = 1, FP8 = 2 };
const char* PrecisionName(Precision p);
bool ParsePrecision(const std::string& s, Precision* out);

enum class DType { F32 = 0, BF16 = 1, FP8 = 2, U8 = 3 };
inline size_t DTypeSize(DType d) { return d == DType::F32 ? 4 : d == DType::BF16 ? 2 : 1; }
const char* DTypeName(DType d);

// A planned value.  Rank-4 values live in NHWC order inside `buffer`, possibly as a channel slice
// [c_off, c_off + C) of a wider pixel (pitch >= C): that is how dense-block concats are done in
// place.  Rank-2 values are [N, C] rows (H = W = 1).
struct TensorDesc {
    std::string name;
    int rank = 4;
    int C = 0, H = 1, W = 1;
    int buffer = -1;
    int c_off = 0;
    int pitch = 0;
    DType dtype = DType::F32;
    size_t PixelsPerSample() const { return (size_t)H * W; }
};

struct BufferDesc {
    enum class Role { Arena, Input, Output };
    Role role = Role::Arena;
    int io_index = -1;            // position among graph inputs / outputs
    size_t elems_per_sample = 0;  // H * W * pitch
    DType dtype = DType::F32;
    int first_step = 1 << 30, last_step = -1;
    size_t offset = 0;  // byte offset inside the arena (for max_batch samples)
    size_t BytesPerSample() const { return elems_per_sample * DTypeSize(dtype); }
};

enum class StepKind {
    NchwToNhwc,     // graph input [N,C,H,W] fp32 -> internal NHWC (dtype cast, optional channel padding)
    NhwcToNchw,     // internal NHWC -> fp32 [N,C,H,W] (graph outputs of rank 4, Flatten of HW>1)
    Conv,           // implicit-GEMM convolution / Gemm / MatMul with fused prologue + epilogue
    MaxPool,
    AvgPool,
    BnRelu,         // standalone per-channel scale/shift (+ReLU)
    GlobalAvgPool,  // optional fused BN+ReLU prologue; output fp32 [N,C]
    Add,            // elementwise tensor + tensor
    Relu,
    Softmax,        // rank-2, over C
    CopyChannels    // concat fallback when in-place placement is impossible
};
const char* StepKindName(StepKind k);

struct Step {
    StepKind kind = StepKind::Conv;
    std::string name;
    int in = -1, in2 = -1, out = -1;  // indices into Plan::tensors
    // Conv / pools
    int R = 1, S = 1, stride = 1, pad = 0;
    int Cin = 0, Cout = 0;
    int weight = -1;     // Plan::consts index, layout [Cout][R][S][Cin]
    int bias = -1;       // [Cout]
    int pre_scale = -1;  // [Cin]  A-operand prologue: x*scale+shift (folded BatchNormalization)
    int pre_shift = -1;
    bool pre_relu = false;
    bool post_relu = false;
    bool pool2_fused = false;  // 2x2/s2 AveragePool commuted in front of a 1x1 conv (linear ops commute)
    bool stem_nchw = false;    // 7x7/s2/p3 image stem reading the caller's fp32 NCHW input directly (layout pass fused)
    bool count_include_pad = false;
    bool ceil_mode = false;
    // BnRelu / GlobalAvgPool prologue
    int bn_scale = -1, bn_shift = -1;
    bool relu = false;
    // accounting, per sample, for the ONNX-order (algorithmic) computation
    double flops = 0.0;
    double bytes = 0.0;  // minimal HBM traffic of this step: inputs read once + outputs written once + weights
};

struct ConstBlob {
    std::string name;
    std::vector<int64_t> dims;
    std::vector<float> data;
};

struct Plan {
    Precision precision = Precision::FP32;
    int max_batch = 1;
    std::vector<TensorDesc> tensors;
    std::vector<BufferDesc> buffers;
    std::vector<Step> steps;
    std::vector<ConstBlob> consts;
    std::vector<int> inputs;   // tensor index of each graph input, as the caller passes it (fp32, ONNX layout)
    std::vector<int> outputs;  // tensor index of each graph output (fp32, ONNX layout)
    std::vector<std::string> input_names, output_names;
    std::vector<std::vector<int64_t>> input_dims, output_dims;  // ONNX dims, -1 = batch / unknown
    std::unordered_map<std::string, int> value_to_tensor;
    size_t arena_bytes = 0;
    size_t weight_bytes = 0;
    double flops_per_sample = 0.0;
    double hbm_bytes_per_sample = 0.0;
    int inplace_concats = 0, copied_concats = 0;

    std::string ToJson() const;
};

// True when a convolution is the image stem the fused tcgen05 stem kernel executes straight from fp32 NCHW
// (7x7, stride 2, pad 3, <= 4 input channels, <= 64 output channels).
bool StemFusable(int Cin, int Cout, int R, int S, int stride, int pad, int H, int W);

// Throws std::runtime_error with a descriptive message for unsupported graphs.
Plan BuildPlan(const onnx::Model& model, Precision precision, int max_batch);

}  // namespace b200
