// engine.cu — Replica: executes a Plan on one GPU.  Replaces `Ort::Session::Run`
// (reference inference_engine/src/model.cpp:1264-1270) and the host<->device traffic that lives
// inside ONNX Runtime's CUDA EP in the reference.
#include "engine.h"

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp8.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <sstream>

namespace b200 {

void CudaCheck(cudaError_t e, const char* what) {
    if (e != cudaSuccess) {
        std::string msg = std::string("CUDA error in ") + what + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
        throw CudaError(msg);
    }
}

namespace {

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        CudaCheck(cudaSetDevice(dev), "cudaSetDevice");
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn GetEncodeTiled() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    if (!fn) throw CudaError("cuTensorMapEncodeTiled is not available from this driver");
    return fn;
}

uint16_t F32ToBf16(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) == 0x7F800000u && (u & 0x7FFFFFu)) return (uint16_t)((u >> 16) | 0x40);  // NaN
    u += 0x7FFFu + ((u >> 16) & 1u);  // round to nearest even
    return (uint16_t)(u >> 16);
}

}  // namespace

// --------------------------------------------------------------------------------------------
ComputeChain::ComputeChain(int dev) : device(dev) {
    DeviceGuard g(device);
    CudaCheck(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate(chain)");
}
ComputeChain::~ComputeChain() {
    if (ev) { cudaSetDevice(device); cudaEventDestroy(ev); }
}

PinnedPool& PinnedPool::Get() {
    static PinnedPool* pool = new PinnedPool();  // leaked on purpose: request threads may outlive static destruction
    return *pool;
}
PinnedPool::~PinnedPool() {}
bool PinnedPool::IsPageable(const void* p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
    return a.type == cudaMemoryTypeUnregistered;
}
void* PinnedPool::Take(size_t bytes) {
    // size classes: 1 MiB granules up to 16 MiB, 4 MiB granules above (a power-of-two class pinned 256 MB for a 154 MB request)
    const size_t gran = bytes <= (16u << 20) ? (1u << 20) : (4u << 20);
    const size_t cap = std::max(gran, (bytes + gran - 1) / gran * gran);
    {
        std::lock_guard<std::mutex> lk(mu_);
        if (!budget_) {
            const char* e = getenv("B200_ENGINE_STAGING_MB");
            budget_ = (size_t)std::max(1, e ? atoi(e) : 2048) << 20;
        }
        for (auto& b : bufs_)
            if (!b.used && b.cap == cap) { b.used = true; return b.p; }
        if (total_ + cap > budget_) return nullptr;
        total_ += cap;
    }
    void* p = nullptr;
    if (cudaHostAlloc(&p, cap, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        std::lock_guard<std::mutex> lk(mu_);
        total_ -= cap;
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(mu_);
    bufs_.push_back({p, cap, true});
    return p;
}
void PinnedPool::Give(void* p) {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto& b : bufs_)
        if (b.p == p) { b.used = false; return; }
}
// Releases every idle staging buffer (called when a model is unloaded: the pool is process-wide and would otherwise keep
// its high-water mark pinned for the life of the process).
void PinnedPool::Trim() {
    std::vector<void*> drop;
    {
        std::lock_guard<std::mutex> lk(mu_);
        for (size_t i = 0; i < bufs_.size();) {
            if (bufs_[i].used) { ++i; continue; }
            drop.push_back(bufs_[i].p);
            total_ -= bufs_[i].cap;
            bufs_[i] = bufs_.back();
            bufs_.pop_back();
        }
    }
    for (void* p : drop) cudaFreeHost(p);
}

Replica::Replica(int device, std::shared_ptr<const Plan> plan, bool use_graphs, std::shared_ptr<ComputeChain> chain, const Replica* weights_of)
    : device_(device), plan_(std::move(plan)), use_graphs_(use_graphs), chain_(std::move(chain)), weights_of_(weights_of) {
    if (weights_of_ && (weights_of_->device_ != device_ || weights_of_->plan_.get() != plan_.get()))
        throw CudaError("a replica can only borrow the weights of a replica of the same plan on the same device");
    DeviceGuard g(device_);
    cudaDeviceProp prop;
    CudaCheck(cudaGetDeviceProperties(&prop, device_), "cudaGetDeviceProperties");
    if (prop.major != 10) {
        throw CudaError("device " + std::to_string(device_) + " (" + prop.name + ", sm_" + std::to_string(prop.major) +
                        std::to_string(prop.minor) + ") is not a Blackwell sm_100 GPU; this engine ships sm_100a code only");
    }
    CudaCheck(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "cudaStreamCreate");
    CudaCheck(cudaStreamCreateWithFlags(&copy_stream_, cudaStreamNonBlocking), "cudaStreamCreate");
    copy_events_.resize(32);
    for (auto& e : copy_events_) CudaCheck(cudaEventCreateWithFlags(&e, cudaEventDisableTiming), "cudaEventCreate");
    if (const char* pc = getenv("B200_ENGINE_PIPELINE_CHUNK")) pipeline_chunk_ = atoi(pc);
    if (const char* pc = getenv("B200_ENGINE_CHAIN_MIN_BATCH")) chain_min_batch_ = atoi(pc);
    CudaCheck(cudaEventCreate(&ev0_), "cudaEventCreate");
    CudaCheck(cudaEventCreate(&ev1_), "cudaEventCreate");
    // a guard in front of the first buffer: the streaming dense-layer kernel's TMA boxes begin one pixel before a row
    constexpr size_t kArenaGuard = 4096;
    CudaCheck(cudaMalloc((void**)&arena_alloc_, plan_->arena_bytes + 4096 + kArenaGuard), "cudaMalloc(arena)");
    CudaCheck(cudaMemsetAsync(arena_alloc_, 0, plan_->arena_bytes + 4096 + kArenaGuard, stream_), "cudaMemset(arena)");
    arena_ = arena_alloc_ + kArenaGuard;
    device_bytes_ += plan_->arena_bytes + 4096 + kArenaGuard;

    const Plan& P = *plan_;
    // fp32 device copies of per-channel vectors, uploaded lazily below
    dconst_.assign(P.consts.size(), nullptr);
    auto vec = [&](int idx) -> const float* {
        if (idx < 0) return nullptr;
        if (weights_of_) return weights_of_->dconst_[idx];
        if (!dconst_[idx]) dconst_[idx] = (const float*)Upload(P.consts[idx].data.data(), P.consts[idx].data.size() * 4);
        return dconst_[idx];
    };

    prepared_.resize(P.steps.size());
    size_t f32_pool_need = 0;
    for (size_t i = 0; i < P.steps.size(); ++i) {
        const Step& s = P.steps[i];
        Prepared& pr = prepared_[i];
        auto io_stride = [&](int tensor) -> size_t {
            const BufferDesc& b = P.buffers[P.tensors[tensor].buffer];
            return b.role == BufferDesc::Role::Arena ? 0 : b.BytesPerSample();
        };
        if (s.in >= 0) { pr.in = MakeView(s.in); pr.in_io_stride = io_stride(s.in); }
        if (s.in2 >= 0) { pr.in2 = MakeView(s.in2); pr.in2_io_stride = io_stride(s.in2); }
        if (s.out >= 0) { pr.out = MakeView(s.out); pr.out_io_stride = io_stride(s.out); }
        pr.scale = vec(s.bn_scale);
        pr.shift = vec(s.bn_shift);
        if (s.kind != StepKind::Conv) continue;

        kernels::ConvArgs& a = pr.conv;
        a.in = pr.in;
        a.out = pr.out;
        a.R = s.R; a.S = s.S; a.stride = s.stride; a.pad = s.pad;
        a.Cin = s.Cin; a.Cout = s.Cout;
        a.pre_scale = vec(s.pre_scale);
        a.pre_shift = vec(s.pre_shift);
        a.pre_relu = s.pre_relu;
        a.bias = vec(s.bias);
        a.post_relu = s.post_relu;
        a.pool2 = s.pool2_fused;
        a.stem_nchw = s.stem_nchw;
        const std::vector<float>& w = P.consts[s.weight].data;  // [Cout][R][S][Cin]
        const int K = s.R * s.S * s.Cin;
        pr.use_umma = s.stem_nchw || pr.in.dtype != DType::F32;
        // FP32 mode: tcgen05 with bf16-split operands (kernels_f32x3.cu) wherever the shape allows; B200_ENGINE_FP32_EXACT=1 keeps
        // every convolution on the exact FFMA kernels
        const char* exact_env = getenv("B200_ENGINE_FP32_EXACT");
        const bool f32x3_enabled = !(exact_env && exact_env[0] == '1');
        kernels::ConvArgs a_eff = a;  // the conv the tensor cores actually run (a pooled transition multiplies the pooled operand)
        if (!pr.use_umma && s.pool2_fused) {
            // FP32 transition: sum_2x2 relu(bn(x)) is materialised once (kernels_poolbn.cu), then a plain 1x1 conv with out_mul 0.25
            kernels::View vin = pr.in;
            vin.C = s.Cin;
            kernels::View pooled = vin;
            pooled.base = arena_;
            pooled.H = pr.out.H; pooled.W = pr.out.W; pooled.pitch = s.Cin; pooled.c_off = 0;
            a_eff.in = pooled;
            a_eff.pre_scale = a_eff.pre_shift = nullptr;
            a_eff.pre_relu = false; a_eff.pool2 = false; a_eff.out_mul = 0.25f;
            if (!f32x3_enabled || !a.pre_scale || !kernels::PoolBnRelu2x2Supported(vin, pooled) || !kernels::ConvF32x3Supported(a_eff))
                throw CudaError("transition '" + s.name + "' was planned with its 2x2 average pool in front of the conv, which FP32 mode cannot run for this shape");
            pr.split_pool = true;
            f32_pool_need = std::max(f32_pool_need, (size_t)P.max_batch * pooled.H * pooled.W * s.Cin * 4 + 256);
        }
        pr.use_f32x3 = !pr.use_umma && f32x3_enabled && kernels::ConvF32x3Supported(a_eff);
        if (weights_of_) {  // same plan, same device: the lender's packed weights, scales and tensor map serve this instance too
            const Prepared& lp = weights_of_->prepared_[i];
            pr.w_kn = lp.w_kn;
            pr.umma = lp.umma;
            pr.h_out_scale = lp.h_out_scale;
            continue;
        }
        if (pr.use_f32x3) {
            // [Cout_pad][R*S*Cin*2] bf16: per filter tap and 32 input channels one 128-byte row [w0 x32 | w1 x32], w = w0 + w1
            const int bn = kernels::F32x3TileN(a_eff);
            const int cout_pad = kernels::F32x3PackedRows(a_eff);
            const int K_pad = kernels::F32x3PackedK(a_eff);
            std::vector<uint16_t> packed((size_t)cout_pad * K_pad, 0);
            for (int o = 0; o < s.Cout; ++o)
                for (int tap = 0; tap < s.R * s.S; ++tap)
                    for (int c = 0; c < s.Cin; ++c) {
                        const float v = w[((size_t)o * s.R * s.S + tap) * s.Cin + c];
                        const uint16_t h0 = F32ToBf16(v);
                        const uint32_t u0 = (uint32_t)h0 << 16;
                        float f0;
                        memcpy(&f0, &u0, 4);
                        int row, col;
                        kernels::F32x3WeightPos(a_eff, o, tap, c, 0, &row, &col);
                        packed[(size_t)row * K_pad + col] = h0;
                        kernels::F32x3WeightPos(a_eff, o, tap, c, 1, &row, &col);
                        packed[(size_t)row * K_pad + col] = F32ToBf16(v - f0);
                    }
            std::vector<float> ones(s.Cout, 1.f);
            pr.umma.w = Upload(packed.data(), packed.size() * 2);
            pr.umma.out_scale = (const float*)Upload(ones.data(), ones.size() * 4);
            pr.umma.K_pad = K_pad;
            pr.umma.Cout_pad = cout_pad;
            CUtensorMap* tm = new CUtensorMap;
            cuuint64_t dims[2] = {(cuuint64_t)K_pad, (cuuint64_t)cout_pad};
            cuuint64_t strides[1] = {(cuuint64_t)K_pad * 2};
            cuuint32_t box[2] = {64u, (cuuint32_t)bn};
            cuuint32_t estr[2] = {1, 1};
            CUresult r = GetEncodeTiled()(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(pr.umma.w), dims, strides, box, estr,
                                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            if (r != CUDA_SUCCESS) {
                delete tm;
                throw CudaError("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ") for conv '" + s.name + "'");
            }
            pr.umma.tensor_map = tm;
            continue;
        }
        if (!pr.use_umma) {
            std::vector<float> kn((size_t)K * s.Cout);
            for (int o = 0; o < s.Cout; ++o)
                for (int k = 0; k < K; ++k) kn[(size_t)k * s.Cout + o] = w[(size_t)o * K + k];
            pr.w_kn = (const float*)Upload(kn.data(), kn.size() * 4);
            continue;
        }
        // ---- tcgen05 path: [Cout_pad][K_pad] K-major in the MMA element type ----
        if (!kernels::UmmaSupported(a))
            throw CudaError("conv '" + s.name + "' has a shape the tcgen05 path does not support in " +
                            PrecisionName(P.precision) + " mode (Cin=" + std::to_string(s.Cin) + ", Cout=" + std::to_string(s.Cout) + ")");
        const DType mt = s.stem_nchw ? DType::BF16 : pr.in.dtype;  // the stem always multiplies in bf16
        const int esz = (int)DTypeSize(mt);
        const int kc = kernels::UmmaKChunkElems(mt);
        const int cin_pad = kernels::UmmaPaddedCin(s.Cin, s.R, s.S, mt);
        const bool stem = s.Cin < 16;
        // stem packing: k = r*32 + s*4 + c (8 pixels x 4 channels per filter row, zero padded); the fused NCHW stem
        // starts its 8-pixel window one pixel further left (16-byte aligned rows), so its taps sit at s+1
        // FP32 mode stem: w = w0 + w1, the residual tile follows the leading tile along K (kernels_stem.cu, SPLIT)
        const bool stem_split = s.stem_nchw && pr.out.dtype == DType::F32;
        const int K_stem = (s.R * 32 + kc - 1) / kc * kc;
        const int K_pad = stem ? (stem_split ? 2 : 1) * K_stem : s.R * s.S * cin_pad;
        const int bn = s.stem_nchw ? 64 : s.Cout <= 32 ? 32 : s.Cout <= 64 ? 64 : 128;
        const int cout_pad = (s.Cout + bn - 1) / bn * bn;
        std::vector<float> scale(s.Cout, 1.f);
        if (mt == DType::FP8) {
            for (int o = 0; o < s.Cout; ++o) {
                float amax = 0.f;
                for (int k = 0; k < K; ++k) amax = std::max(amax, std::fabs(w[(size_t)o * K + k]));
                scale[o] = amax > 0.f ? amax / 448.f : 1.f;
            }
        }
        std::vector<uint8_t> packed((size_t)cout_pad * K_pad * esz, 0);
        auto put = [&](int o, int kk, float v) {
            size_t idx = (size_t)o * K_pad + kk;
            if (mt == DType::BF16) {
                uint16_t h = F32ToBf16(v);
                memcpy(&packed[idx * 2], &h, 2);
            } else {
                __nv_fp8_e4m3 q(v / scale[o]);
                memcpy(&packed[idx], &q, 1);
            }
        };
        for (int o = 0; o < s.Cout; ++o)
            for (int r = 0; r < s.R; ++r)
                for (int ss = 0; ss < s.S; ++ss)
                    for (int c = 0; c < s.Cin; ++c) {
                        float v = w[(((size_t)o * s.R + r) * s.S + ss) * s.Cin + c];
                        int kk = stem ? r * 32 + (ss + (s.stem_nchw ? 1 : 0)) * 4 + c : (r * s.S + ss) * cin_pad + c;
                        put(o, kk, v);
                        if (stem_split) {
                            const uint32_t u0 = (uint32_t)F32ToBf16(v) << 16;
                            float f0;
                            memcpy(&f0, &u0, 4);
                            put(o, K_stem + kk, v - f0);
                        }
                    }
        pr.umma.w = Upload(packed.data(), packed.size());
        pr.umma.out_scale = (const float*)Upload(scale.data(), scale.size() * 4);
        pr.h_out_scale = scale;
        pr.umma.K_pad = K_pad;
        pr.umma.Cout_pad = cout_pad;
        // TMA descriptor over the packed weights: dims {K_pad, Cout_pad}, box {kc, bn}, 128-byte swizzle
        CUtensorMap* tm = new CUtensorMap;
        cuuint64_t dims[2] = {(cuuint64_t)K_pad, (cuuint64_t)cout_pad};
        cuuint64_t strides[1] = {(cuuint64_t)K_pad * esz};
        cuuint32_t box[2] = {(cuuint32_t)kc, (cuuint32_t)bn};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = GetEncodeTiled()(tm, mt == DType::BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2,
                                      const_cast<void*>(pr.umma.w), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
            delete tm;
            throw CudaError("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ") for conv '" + s.name + "'");
        }
        pr.umma.tensor_map = tm;
    }
    if (weights_of_) BorrowDenseRuns(*weights_of_);
    else BuildDenseRuns();
    MarkStreamPairs();
    // fp32 reference mode: scratch for the deterministic split-K of the SIMT convolutions at small batch
    {
        bool any_simt = false;
        for (size_t i = 0; i < P.steps.size(); ++i) any_simt = any_simt || (P.steps[i].kind == StepKind::Conv && !prepared_[i].use_umma && !prepared_[i].use_f32x3);
        const char* e = getenv("B200_ENGINE_SPLITK");
        if (any_simt && !(e && e[0] == '0')) {
            splitk_bytes_ = (8u << 20) + 4096 * sizeof(unsigned int);
            CudaCheck(cudaMalloc(&splitk_scratch_, splitk_bytes_), "cudaMalloc(split-K scratch)");
            CudaCheck(cudaMemsetAsync(splitk_scratch_, 0, splitk_bytes_, stream_), "cudaMemset(split-K scratch)");
            allocations_.push_back(splitk_scratch_);
            device_bytes_ += splitk_bytes_;
        }
    }
    // Transition layers: conv1x1_tma<POOL> redoes the pooled transform of the A tile for every 128-column N tile (Cout 256 / 512);
    // materialising sum_2x2 relu(bn(x)) once and running a plain 1x1 conv over it is cheaper (kernels_poolbn.cu).
    {
        const char* e = getenv("B200_ENGINE_SPLIT_TRANSITION");
        const bool enabled = !(e && e[0] == '0');
        // "1": only the wide ones (Cout > 128).  Default: transition 1 (one N tile) as well - the fused kernel's pooled transform
        // reads four planes through 8 warps and streams at 2.8 TB/s; the memory-bound pass + plain conv is 73 -> 60 us.
        const int min_cout = (e && e[0] == '1') ? 128 : 0;
        size_t need = f32_pool_need;
        for (size_t i = 0; enabled && i < P.steps.size(); ++i) {
            const Step& s = P.steps[i];
            Prepared& pr = prepared_[i];
            if (s.kind != StepKind::Conv || !s.pool2_fused || !pr.use_umma || pr.fused_run >= 0) continue;
            // single-N-tile transitions only pay off in e4m3 (bf16 doubles the bytes of the extra round trip: 3.60 -> 3.62 ms)
            if (s.Cout <= (pr.in.dtype == DType::FP8 ? min_cout : 128)) continue;
            kernels::View vin = pr.in;
            vin.C = s.Cin;
            kernels::View pooled = vin;
            pooled.base = arena_;  // any 16-byte aligned address: only the geometry is checked here
            pooled.H = pr.out.H; pooled.W = pr.out.W; pooled.pitch = s.Cin; pooled.c_off = 0;
            kernels::ConvArgs a2 = pr.conv;
            a2.in = pooled;
            a2.pre_scale = a2.pre_shift = nullptr;
            a2.pre_relu = false; a2.pool2 = false; a2.out_mul = 0.25f;
            if (!pr.conv.pre_scale || !kernels::PoolBnRelu2x2Supported(vin, pooled) || !kernels::Conv1x1TmaSupported(a2)) continue;
            pr.split_pool = true;
            need = std::max(need, (size_t)P.max_batch * pooled.H * pooled.W * s.Cin * DTypeSize(vin.dtype) + 256);
        }
        if (need) {
            CudaCheck(cudaMalloc(&pool_scratch_, need), "cudaMalloc(pooled transition operand)");
            allocations_.push_back(pool_scratch_);
            device_bytes_ += need;
        }
    }
    flush_bytes_ = 256u << 20;
    CudaCheck(cudaStreamSynchronize(stream_), "replica init");
}

// The dense-layer tables hold only weights (tensor maps, packed BN constants, epilogue vectors): an execution instance that
// borrows another's weight set reuses them and points the runs at its OWN arena.
void Replica::BorrowDenseRuns(const Replica& lender) {
    dense_runs_ = lender.dense_runs_;
    for (size_t i = 0; i < prepared_.size(); ++i) prepared_[i].fused_run = lender.prepared_[i].fused_run;
    for (DenseRun& run : dense_runs_) run.args.buf = BufferPtr(plan_->tensors[plan_->steps[run.first_step].in].buffer);
}

// Finds maximal runs of dense layers (conv1x1 -> 128 channels -> conv3x3 -> 32 channels appended to the buffer the 1x1
// read) whose images are small enough for a CTA to own whole images, and builds the device-side layer tables of the
// dense-block kernel.  e4m3 mode only; everything else keeps the layer-per-kernel schedule.
void Replica::BuildDenseRuns() {
    const Plan& P = *plan_;
    if (P.precision != Precision::FP8) return;
    if (const char* e = getenv("B200_ENGINE_DENSEFUSE")) if (e[0] == '0') return;
    // (A fused dense-layer kernel over 14x14 tiles for the 56x56 / 28x28 blocks existed in round 1: bit-identical but slower than
    // the conv1x1 + conv3x3 pair - block 1: 941 us against 735 us - and was removed; DESIGN.md section 10 keeps the analysis.)
    auto readers = [&](int tensor) {
        int c = 0;
        for (const Step& s : P.steps) c += (s.in == tensor) + (s.in2 == tensor);
        for (int o : P.outputs) c += (o == tensor);
        return c;
    };
    auto pair_ok = [&](size_t i) -> bool {
        if (i + 1 >= P.steps.size()) return false;
        const Step& a = P.steps[i];
        const Step& b = P.steps[i + 1];
        if (a.kind != StepKind::Conv || b.kind != StepKind::Conv) return false;
        const Prepared& pa = prepared_[i];
        const Prepared& pb = prepared_[i + 1];
        if (!pa.use_umma || !pb.use_umma || !pa.umma.tensor_map || !pb.umma.tensor_map) return false;
        if (a.R != 1 || a.S != 1 || a.stride != 1 || a.pad != 0 || a.pool2_fused || a.stem_nchw || a.pre_scale < 0) return false;
        if (b.R != 3 || b.S != 3 || b.stride != 1 || b.pad != 1 || b.pool2_fused || b.pre_scale >= 0) return false;
        if (a.Cout != 128 || b.Cin != 128 || b.Cout != 32 || a.Cin % 32 != 0 || a.Cin < 32 || a.Cin > 2048) return false;
        if (b.in != a.out || readers(a.out) != 1) return false;
        const TensorDesc& ain = P.tensors[a.in];
        const TensorDesc& bout = P.tensors[b.out];
        if (ain.dtype != DType::FP8 || bout.dtype != DType::FP8 || P.tensors[a.out].dtype != DType::FP8) return false;
        if (ain.buffer != bout.buffer || ain.pitch != bout.pitch || ain.c_off != 0 || bout.c_off != a.Cin) return false;
        if (ain.H != bout.H || ain.W != bout.W || ain.pitch % 16 != 0) return false;
        int ipc = 0, mt = 0;
        return kernels::DenseBlockGeometry(ain.H, ain.W, &ipc, &mt);
    };
    for (size_t i = 0; i < P.steps.size();) {
        if (!pair_ok(i)) { ++i; continue; }
        const TensorDesc& first_in = P.tensors[P.steps[i].in];
        size_t j = i;
        std::vector<kernels::DenseLayerDesc> table;
        while (pair_ok(j)) {
            const Step& a = P.steps[j];
            const Step& b = P.steps[j + 1];
            const TensorDesc& ain = P.tensors[a.in];
            if (ain.buffer != first_in.buffer || ain.H != first_in.H || ain.W != first_in.W || ain.pitch != first_in.pitch) break;
            // dense connectivity: every layer reads exactly its predecessor's inputs plus the 32 channels it appended (the
            // dense-block kernel forwards those through shared memory)
            if (!table.empty() && a.Cin != table.back().Cin + 32) break;
            kernels::DenseLayerDesc d;
            memset(&d, 0, sizeof(d));
            memcpy(&d.w1, prepared_[j].umma.tensor_map, sizeof(CUtensorMap));
            {   // the nine 3x3 taps as ONE box: dims {128 B of K inside a tap, 32 output channels, 9 taps} -> lands as [tap][row][128 B]
                const kernels::UmmaWeights& uw = prepared_[j + 1].umma;
                const uint64_t wdims[3] = {128, 32, 9};
                const uint64_t wstrides[2] = {(uint64_t)uw.K_pad, 128};
                const uint32_t wbox[3] = {128u, 32u, 9u};
                if (uw.K_pad != 9 * 128 || uw.Cout_pad != 32 ||
                    kernels::MakeTensorMap(&d.w2, uw.w, 1, 3, wdims, wstrides, wbox, true) != 0)
                    throw CudaError("dense block: cannot describe the 3x3 weights of '" + b.name + "' as one TMA box");
            }
            // folded BN1 as packed f16x2 pairs (the transform warps' arithmetic type in e4m3 mode)
            const std::vector<float>& sc = P.consts[a.pre_scale].data;
            const std::vector<float>& sh = P.consts[a.pre_shift].data;
            std::vector<uint32_t> psc((a.Cin + 1) / 2 + 8, 0), psh((a.Cin + 1) / 2 + 8, 0);
            for (int c = 0; c < a.Cin; ++c) {
                const uint32_t hs = __half_as_ushort(__float2half_rn(sc[c])), ht = __half_as_ushort(__float2half_rn(sh[c]));
                psc[c / 2] |= hs << (16 * (c & 1));
                psh[c / 2] |= ht << (16 * (c & 1));
            }
            d.pre_scale = (const uint32_t*)Upload(psc.data(), psc.size() * 4);
            d.pre_shift = (const uint32_t*)Upload(psh.data(), psh.size() * 4);
            d.s1 = prepared_[j].umma.out_scale;
            d.b1 = prepared_[j].conv.bias;
            d.s2 = prepared_[j + 1].umma.out_scale;
            d.b2 = prepared_[j + 1].conv.bias;
            d.Cin = a.Cin;
            d.c_off_out = P.tensors[b.out].c_off;
            d.pre_relu = a.pre_relu; d.relu1 = a.post_relu; d.relu2 = b.post_relu;
            table.push_back(d);
            j += 2;
        }
        if (table.empty()) { ++i; continue; }
        DenseRun run;
        run.first_step = i;
        run.num_layers = (int)table.size();
        run.args.layers_dev = (const kernels::DenseLayerDesc*)Upload(table.data(), table.size() * sizeof(kernels::DenseLayerDesc));
        run.args.num_layers = run.num_layers;
        run.args.buf = BufferPtr(first_in.buffer);
        run.args.pitch = first_in.pitch;
        run.args.H = first_in.H; run.args.W = first_in.W;
        prepared_[i].fused_run = (int)dense_runs_.size();
        dense_runs_.push_back(run);
        i = j;
    }
}

// Dense layers of the large-image blocks (56x56, 28x28; e4m3): the (1x1 conv, 3x3 conv) step pair can run as ONE streaming kernel
// that keeps the 128-channel bottleneck in shared memory (kernels_dense_stream.cu).  Bit-identical to the kernel pair and half the
// HBM traffic.  Measured per layer at bs256 (tools/perlayer_fuse.py): the streaming kernel wins where the layer has ONE K chunk
// (56x56, Cin <= 128: 90-93 us against 97-102 us) and loses 0.55 us per tile for every further chunk (its row-owning transform into
// tensor memory is a latency chain per chunk), so the default ("auto") fuses exactly those layers; B200_ENGINE_LAYERFUSE=1 fuses
// every layer of both blocks (forward 2.20 ms against 1.98 ms), =0 none.
void Replica::MarkStreamPairs() {
    const Plan& P = *plan_;
    if (P.precision != Precision::FP8) return;
    const char* lf = getenv("B200_ENGINE_LAYERFUSE");
    const int mode = !lf ? 2 : lf[0] == '0' ? 0 : lf[0] == '1' ? 1 : 2;   // 0 off, 1 all, 2 auto
    if (mode == 0) return;
    auto readers = [&](int tensor) {
        int c = 0;
        for (const Step& s : P.steps) c += (s.in == tensor) + (s.in2 == tensor);
        for (int o : P.outputs) c += (o == tensor);
        return c;
    };
    for (size_t i = 0; i + 1 < P.steps.size(); ++i) {
        const Step& a = P.steps[i];
        const Step& b = P.steps[i + 1];
        Prepared& pa = prepared_[i];
        const Prepared& pb = prepared_[i + 1];
        if (pa.fused_run >= 0 || a.kind != StepKind::Conv || b.kind != StepKind::Conv) continue;
        if (!pa.use_umma || !pb.use_umma || !pa.umma.tensor_map || !pb.umma.tensor_map) continue;
        if (a.R != 1 || a.S != 1 || a.stride != 1 || a.pad != 0 || a.pool2_fused || a.stem_nchw || a.pre_scale < 0 || a.pre_shift < 0) continue;
        if (b.R != 3 || b.S != 3 || b.stride != 1 || b.pad != 1 || b.pool2_fused || b.pre_scale >= 0) continue;
        if (a.Cout != 128 || b.Cin != 128 || b.Cout != 32) continue;
        if (b.in != a.out || readers(a.out) != 1) continue;
        const TensorDesc& ain = P.tensors[a.in];
        const TensorDesc& bout = P.tensors[b.out];
        if (ain.dtype != DType::FP8 || bout.dtype != DType::FP8 || P.tensors[a.out].dtype != DType::FP8) continue;
        if (ain.buffer != bout.buffer || ain.pitch != bout.pitch || ain.c_off != 0 || bout.c_off != a.Cin) continue;
        if (ain.H != bout.H || ain.W != bout.W) continue;
        if (P.buffers[ain.buffer].role != BufferDesc::Role::Arena) continue;  // graph I/O buffers are indexed by the sub-batch offset
        if (pa.h_out_scale.size() != 128 || pb.h_out_scale.size() != 32) continue;
        if ((int)P.consts[a.pre_scale].data.size() < a.Cin || (int)P.consts[a.pre_shift].data.size() < a.Cin) continue;
        if (!kernels::DenseLayerStreamSupported(ain.H, ain.W, a.Cin, ain.pitch)) continue;
        if (mode == 2 && !(ain.W == 56 && a.Cin <= 128)) continue;
        pa.stream_pair = true;
        pa.stream_min_batch = mode == 2 ? 32 : 0;
        ++i;  // the 3x3 conv is consumed by the pair
    }
}

Replica::~Replica() {
    cudaSetDevice(device_);
    if (stream_) cudaStreamSynchronize(stream_);
    if (copy_stream_) { cudaStreamSynchronize(copy_stream_); cudaStreamDestroy(copy_stream_); }
    for (auto& e : copy_events_) if (e) cudaEventDestroy(e);
    for (auto& kv : graphs_) cudaGraphExecDestroy(kv.second);
    if (!weights_of_)
        for (auto& pr : prepared_)
            if (pr.umma.tensor_map) delete (CUtensorMap*)pr.umma.tensor_map;
    for (void* p : allocations_) cudaFree(p);
    if (flush_buf_) cudaFree(flush_buf_);
    if (arena_alloc_) cudaFree(arena_alloc_);
    if (ev0_) cudaEventDestroy(ev0_);
    if (ev1_) cudaEventDestroy(ev1_);
    if (stream_) cudaStreamDestroy(stream_);
}

void* Replica::Upload(const void* host, size_t bytes) {
    void* d = nullptr;
    size_t padded = (bytes + 255) / 256 * 256 + 256;
    CudaCheck(cudaMalloc(&d, padded), "cudaMalloc(weights)");
    allocations_.push_back(d);
    device_bytes_ += padded;
    CudaCheck(cudaMemsetAsync(d, 0, padded, stream_), "cudaMemset(weights)");
    CudaCheck(cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, stream_), "cudaMemcpy(weights)");
    CudaCheck(cudaStreamSynchronize(stream_), "upload sync");
    return d;
}

void* Replica::BufferPtr(int buffer) const { return arena_ + plan_->buffers[buffer].offset; }

kernels::View Replica::MakeView(int tensor) const {
    const TensorDesc& t = plan_->tensors[tensor];
    kernels::View v;
    v.base = BufferPtr(t.buffer);
    v.dtype = t.dtype;
    v.C = t.C; v.H = t.H; v.W = t.W;
    v.pitch = t.pitch; v.c_off = t.c_off;
    return v;
}

const uint8_t* Replica::U8Source(int tensor, int off, unsigned u8_mask) {
    if (!u8_mask || tensor < 0) return nullptr;
    const TensorDesc& t = plan_->tensors[tensor];
    const BufferDesc& b = plan_->buffers[t.buffer];
    if (b.role != BufferDesc::Role::Input || b.io_index < 0 || !((u8_mask >> b.io_index) & 1u)) return nullptr;
    if ((size_t)b.io_index >= u8_stage_.size() || !u8_stage_[b.io_index]) throw CudaError("uint8 input was not staged");
    return u8_stage_[b.io_index] + (size_t)off * t.C * t.H * t.W;
}

void Replica::EnqueueStep(size_t i, int n, int off, unsigned u8_mask) {
    const Step& s = plan_->steps[i];
    const Prepared& pr0 = prepared_[i];
    const uint8_t* u8 = U8Source(s.in, off, u8_mask);
    if (u8 && !(s.kind == StepKind::NchwToNhwc || (s.kind == StepKind::Conv && s.stem_nchw && pr0.use_umma)))
        throw CudaError("step '" + s.name + "' cannot ingest a uint8 input (only image inputs feeding the layout pass or the stem can)");
    // views of graph inputs/outputs are indexed by the sub-batch offset; everything else is scratch
    kernels::View vin = pr0.in, vin2 = pr0.in2, vout = pr0.out;
    if (off) {
        if (vin.base) vin.base = (char*)vin.base + (size_t)off * pr0.in_io_stride;
        if (vin2.base) vin2.base = (char*)vin2.base + (size_t)off * pr0.in2_io_stride;
        if (vout.base) vout.base = (char*)vout.base + (size_t)off * pr0.out_io_stride;
    }
    cudaError_t e = cudaSuccess;
    switch (s.kind) {
        case StepKind::NchwToNhwc:
            e = u8 ? kernels::U8HwcToNhwc(u8, vout, n, stream_) : kernels::NchwToNhwc((const float*)vin.base, vout, n, stream_);
            break;
        case StepKind::NhwcToNchw: e = kernels::NhwcToNchw(vin, (float*)vout.base, n, stream_); break;
        case StepKind::Conv: {
            kernels::ConvArgs a = pr0.conv;
            a.in = vin;
            a.out = vout;
            a.n = n;
            a.in_u8_hwc = u8;
            a.splitk_scratch = splitk_scratch_;
            a.splitk_bytes = splitk_bytes_;
            a.h_bias = s.bias >= 0 ? plan_->consts[s.bias].data.data() : nullptr;
            kernels::UmmaWeights uw = pr0.umma;
            uw.h_out_scale = pr0.h_out_scale.empty() ? nullptr : pr0.h_out_scale.data();
            if (pr0.split_pool) {
                kernels::View src = vin;
                src.C = s.Cin;
                kernels::View pooled = src;
                pooled.base = pool_scratch_;
                pooled.H = vout.H; pooled.W = vout.W; pooled.pitch = s.Cin; pooled.c_off = 0;
                e = kernels::PoolBnRelu2x2(src, pooled, n, a.pre_scale, a.pre_shift, a.pre_relu, stream_);
                if (e != cudaSuccess) break;
                a.in = pooled;
                a.pre_scale = a.pre_shift = nullptr;
                a.pre_relu = false; a.pool2 = false; a.out_mul = 0.25f;
                e = pr0.use_f32x3 ? kernels::ConvF32x3(a, pr0.umma, stream_) : kernels::Conv1x1Tma(a, uw, stream_);
                break;
            }
            e = pr0.use_umma ? kernels::ConvUmma(a, uw, stream_)
                : pr0.use_f32x3 ? kernels::ConvF32x3(a, pr0.umma, stream_)
                                : kernels::ConvSimtF32(a, pr0.w_kn, stream_);
            break;
        }
        case StepKind::MaxPool: e = kernels::MaxPool(vin, vout, n, s.R, s.stride, s.pad, stream_); break;
        case StepKind::AvgPool: e = kernels::AvgPool(vin, vout, n, s.R, s.stride, s.pad, s.count_include_pad, stream_); break;
        case StepKind::BnRelu: e = kernels::BnRelu(vin, vout, n, pr0.scale, pr0.shift, s.relu, stream_); break;
        case StepKind::GlobalAvgPool:
            e = kernels::GlobalAvgPool(vin, (float*)vout.base + vout.c_off, vout.pitch, n, pr0.scale, pr0.shift, s.relu, stream_);
            break;
        case StepKind::Add: e = kernels::AddTensors(vin, vin2, vout, n, stream_); break;
        case StepKind::Relu: e = kernels::ReluTensor(vin, vout, n, stream_); break;
        case StepKind::Softmax: e = kernels::SoftmaxRows((const float*)vin.base, (float*)vout.base, n, vin.C, stream_); break;
        case StepKind::CopyChannels: e = kernels::CopyChannels(vin, vout, n, stream_); break;
    }
    if (e != cudaSuccess) CudaCheck(e, ("step '" + s.name + "' (" + StepKindName(s.kind) + ")").c_str());
}

size_t Replica::EnqueueAt(size_t i, int n, int off, unsigned u8_mask) {
    const int r = prepared_[i].fused_run;
    if (r < 0 && prepared_[i].stream_pair && n >= prepared_[i].stream_min_batch) {
        const Plan& P = *plan_;
        const Step& a = P.steps[i];
        const Step& b = P.steps[i + 1];
        const Prepared& pa = prepared_[i];
        const Prepared& pb = prepared_[i + 1];
        kernels::DenseLayerStreamArgs d;
        d.w1_map = pa.umma.tensor_map;
        d.w2_map = pb.umma.tensor_map;
        d.buf = pa.in.base;
        d.pitch = pa.in.pitch; d.n = n; d.H = pa.in.H; d.W = pa.in.W;
        d.Cin = a.Cin; d.c_off_out = pb.out.c_off;
        d.pre_relu = a.pre_relu; d.relu1 = a.post_relu; d.relu2 = b.post_relu;
        d.pre_scale = P.consts[a.pre_scale].data.data();
        d.pre_shift = P.consts[a.pre_shift].data.data();
        d.s1 = pa.h_out_scale.data();
        d.b1 = a.bias >= 0 ? P.consts[a.bias].data.data() : nullptr;
        d.s2 = pb.h_out_scale.data();
        d.b2 = b.bias >= 0 ? P.consts[b.bias].data.data() : nullptr;
        cudaError_t e = kernels::DenseLayerStreamFp8(d, stream_);
        if (e != cudaSuccess) CudaCheck(e, ("dense layer (streaming) starting at step '" + a.name + "'").c_str());
        return 2;
    }
    if (r < 0) {
        EnqueueStep(i, n, off, u8_mask);
        return 1;
    }
    DenseRun& run = dense_runs_[r];
    kernels::DenseBlockArgs a = run.args;
    a.n = n;
    cudaError_t e = cudaSuccess;
    e = kernels::DenseBlockFp8(a, stream_);
    if (e != cudaSuccess) CudaCheck(e, ("dense block starting at step '" + plan_->steps[i].name + "'").c_str());
    return (size_t)run.num_layers * 2;
}

void Replica::Enqueue(int n, int off, unsigned u8_mask) {
    if (n <= 0 || off < 0 || off + n > plan_->max_batch) throw CudaError("batch " + std::to_string(off + n) + " exceeds the planned maximum");
    if (!use_graphs_) {
        for (size_t i = 0; i < plan_->steps.size();) i += EnqueueAt(i, n, off, u8_mask);
        return;
    }
    const int64_t key = ((int64_t)(u8_mask & 0xFFu) << 44) | ((int64_t)off << 20) | (int64_t)n;
    auto it = graphs_.find(key);
    if (it == graphs_.end()) {
        cudaGraph_t graph = nullptr;
        uint64_t before = kernels::LaunchCount();
        CudaCheck(cudaStreamBeginCapture(stream_, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture");
        try {
            for (size_t i = 0; i < plan_->steps.size();) i += EnqueueAt(i, n, off, u8_mask);
        } catch (...) {
            cudaStreamEndCapture(stream_, &graph);
            if (graph) cudaGraphDestroy(graph);
            throw;
        }
        CudaCheck(cudaStreamEndCapture(stream_, &graph), "cudaStreamEndCapture");
        launches_per_forward_ = (int)(kernels::LaunchCount() - before);
        kernels::CountLaunch(-launches_per_forward_);  // capture does not execute anything
        cudaGraphExec_t exec = nullptr;
        cudaError_t e = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        CudaCheck(e, "cudaGraphInstantiate");
        if (graphs_.size() >= 64) {  // bound the cache
            cudaGraphExecDestroy(graphs_.begin()->second);
            graph_launches_.erase(graphs_.begin()->first);
            graphs_.erase(graphs_.begin());
        }
        it = graphs_.emplace(key, exec).first;
        graph_launches_[key] = launches_per_forward_;
    }
    CudaCheck(cudaGraphLaunch(it->second, stream_), "cudaGraphLaunch");
    kernels::CountLaunch(graph_launches_[key]);
}

void Replica::Run(int n, const std::vector<const void*>& host_inputs, const std::vector<void*>& host_outputs,
                  const std::vector<size_t>& out_capacity_bytes, unsigned u8_mask, bool alone, const TopK* topk) {
    std::lock_guard<std::mutex> lk(mu_);
    DeviceGuard g(device_);
    const Plan& P = *plan_;
    // destination and bytes per sample of every graph input (fp32 NCHW into the arena, or raw uint8 into its staging buffer)
    if (u8_mask) u8_stage_.resize(P.inputs.size(), nullptr);
    auto in_dst = [&](size_t i) -> char* {
        const TensorDesc& t = P.tensors[P.inputs[i]];
        if (!((u8_mask >> i) & 1u)) return (char*)BufferPtr(t.buffer);
        if (!u8_stage_[i]) {
            const size_t bytes = (size_t)P.max_batch * t.C * t.H * t.W + 256;
            CudaCheck(cudaMalloc((void**)&u8_stage_[i], bytes), "cudaMalloc(uint8 staging)");
            allocations_.push_back(u8_stage_[i]);
            device_bytes_ += bytes;
        }
        return (char*)u8_stage_[i];
    };
    auto in_stride = [&](size_t i) -> size_t {
        const TensorDesc& t = P.tensors[P.inputs[i]];
        return (size_t)t.C * t.H * t.W * (((u8_mask >> i) & 1u) ? 1 : 4);
    };
    // Large batches are pipelined in sub-batches: the H2D copy of sub-batch k+1 (copy stream) overlaps the forward
    // of sub-batch k (compute stream).  The fp32 NCHW input is 602 KB per image, so at bs256 the PCIe transfer is
    // as long as the whole forward; without overlap the two add up.
    const int chunk = pipeline_chunk_ > 0 && alone ? pipeline_chunk_ : n;
    const int pieces = std::min<int>((n + chunk - 1) / chunk, (int)copy_events_.size());
    std::unique_lock<std::mutex> chain_lk;
    const bool chained = chain_ && n >= chain_min_batch_;
    auto chain_begin = [&] {
        if (!chained) return;
        chain_lk = std::unique_lock<std::mutex>(chain_->mu);
        CudaCheck(cudaStreamWaitEvent(stream_, chain_->ev, 0), "chain wait");
    };
    auto chain_end = [&] {
        if (!chained) return;
        CudaCheck(cudaEventRecord(chain_->ev, stream_), "chain record");
        chain_lk.unlock();
    };
    if (pieces <= 1) {
        for (size_t i = 0; i < P.inputs.size(); ++i)
            CudaCheck(cudaMemcpyAsync(in_dst(i), host_inputs[i], (size_t)n * in_stride(i), cudaMemcpyHostToDevice, stream_), "H2D input");
        chain_begin();
        Enqueue(n, 0, u8_mask);
        chain_end();
    } else {
        const int per = (n + pieces - 1) / pieces;
        for (int k = 0; k < pieces; ++k) {
            const int off = k * per, cnt = std::min(per, n - off);
            if (cnt <= 0) break;
            for (size_t i = 0; i < P.inputs.size(); ++i) {
                const size_t stride = in_stride(i);
                CudaCheck(cudaMemcpyAsync(in_dst(i) + off * stride, (const char*)host_inputs[i] + off * stride,
                                          cnt * stride, cudaMemcpyHostToDevice, copy_stream_), "H2D input (pipelined)");
            }
            CudaCheck(cudaEventRecord(copy_events_[k], copy_stream_), "event record");
        }
        chain_begin();
        for (int k = 0; k < pieces; ++k) {
            const int off = k * per, cnt = std::min(per, n - off);
            if (cnt <= 0) break;
            CudaCheck(cudaStreamWaitEvent(stream_, copy_events_[k], 0), "stream wait event");
            Enqueue(cnt, off, u8_mask);
        }
        chain_end();
    }
    for (size_t i = 0; i < P.outputs.size() && i < host_outputs.size(); ++i) {
        const TensorDesc& t = P.tensors[P.outputs[i]];
        size_t bytes = std::min((size_t)n * t.C * t.H * t.W * 4, out_capacity_bytes[i]);
        if (host_outputs[i] && bytes)
            CudaCheck(cudaMemcpyAsync(host_outputs[i], BufferPtr(t.buffer), bytes, cudaMemcpyDeviceToHost, stream_), "D2H output");
    }
    if (topk && topk->k > 0 && !P.outputs.empty()) {
        if (topk->k > kMaxTopK || !topk->idx || !topk->val) throw CudaError("top-k: k must be 1.." + std::to_string(kMaxTopK) + " with both result arrays");
        const TensorDesc& t = P.tensors[P.outputs[0]];
        const size_t half = (size_t)P.max_batch * kMaxTopK * 4;
        if (!topk_dev_) {
            CudaCheck(cudaMalloc((void**)&topk_dev_, 2 * half), "cudaMalloc(top-k)");
            allocations_.push_back(topk_dev_);
            device_bytes_ += 2 * half;
        }
        int* d_idx = reinterpret_cast<int*>(topk_dev_);
        float* d_val = reinterpret_cast<float*>(topk_dev_ + half);
        CudaCheck(kernels::TopKRows((const float*)BufferPtr(t.buffer), n, t.C * t.H * t.W, topk->k, topk->softmax, d_idx, d_val, stream_), "top-k");
        CudaCheck(cudaMemcpyAsync(topk->idx, d_idx, (size_t)n * topk->k * 4, cudaMemcpyDeviceToHost, stream_), "D2H top-k classes");
        CudaCheck(cudaMemcpyAsync(topk->val, d_val, (size_t)n * topk->k * 4, cudaMemcpyDeviceToHost, stream_), "D2H top-k scores");
    }
    CudaCheck(cudaStreamSynchronize(stream_), "forward");
}

void Replica::RunSegments(const std::vector<Segment>& segs, unsigned u8_mask) {
    std::lock_guard<std::mutex> lk(mu_);
    DeviceGuard g(device_);
    const Plan& P = *plan_;
    int total = 0;
    for (const auto& s : segs) total += s.n;
    if (total <= 0) return;
    if (total > P.max_batch) throw CudaError("coalesced batch exceeds the planned maximum");
    if (u8_mask) u8_stage_.resize(P.inputs.size(), nullptr);
    int off = 0;
    for (const auto& s : segs) {
        for (size_t i = 0; i < P.inputs.size(); ++i) {
            const TensorDesc& t = P.tensors[P.inputs[i]];
            const bool u8 = (u8_mask >> i) & 1u;
            const size_t stride = (size_t)t.C * t.H * t.W * (u8 ? 1 : 4);
            char* dst;
            if (u8) {
                if (!u8_stage_[i]) {
                    const size_t bytes = (size_t)P.max_batch * t.C * t.H * t.W + 256;
                    CudaCheck(cudaMalloc((void**)&u8_stage_[i], bytes), "cudaMalloc(uint8 staging)");
                    allocations_.push_back(u8_stage_[i]);
                    device_bytes_ += bytes;
                }
                dst = (char*)u8_stage_[i];
            } else {
                dst = (char*)BufferPtr(t.buffer);
            }
            CudaCheck(cudaMemcpyAsync(dst + (size_t)off * stride, s.in[i], (size_t)s.n * stride, cudaMemcpyHostToDevice, stream_), "H2D input (coalesced)");
        }
        off += s.n;
    }
    {
        std::unique_lock<std::mutex> chain_lk;
        const bool chained = chain_ && total >= chain_min_batch_;
        if (chained) {
            chain_lk = std::unique_lock<std::mutex>(chain_->mu);
            CudaCheck(cudaStreamWaitEvent(stream_, chain_->ev, 0), "chain wait");
        }
        Enqueue(total, 0, u8_mask);
        if (chained) CudaCheck(cudaEventRecord(chain_->ev, stream_), "chain record");
    }
    off = 0;
    for (const auto& s : segs) {
        for (size_t i = 0; i < P.outputs.size() && i < s.out.size(); ++i) {
            const TensorDesc& t = P.tensors[P.outputs[i]];
            const size_t stride = (size_t)t.C * t.H * t.W * 4;
            const size_t bytes = std::min((size_t)s.n * stride, s.cap[i]);
            if (s.out[i] && bytes)
                CudaCheck(cudaMemcpyAsync(s.out[i], (char*)BufferPtr(t.buffer) + (size_t)off * stride, bytes, cudaMemcpyDeviceToHost, stream_), "D2H output (coalesced)");
        }
        off += s.n;
    }
    CudaCheck(cudaStreamSynchronize(stream_), "forward (coalesced)");
}

void Replica::StageInput(int input_index, const void* host, int n) {
    std::lock_guard<std::mutex> lk(mu_);
    DeviceGuard g(device_);
    const TensorDesc& t = plan_->tensors[plan_->inputs.at(input_index)];
    size_t bytes = (size_t)n * t.C * t.H * t.W * 4;
    CudaCheck(cudaMemcpyAsync(BufferPtr(t.buffer), host, bytes, cudaMemcpyHostToDevice, stream_), "H2D stage");
    CudaCheck(cudaStreamSynchronize(stream_), "stage sync");
}

float Replica::ForwardTimed(int n, bool flush_l2) {
    std::lock_guard<std::mutex> lk(mu_);
    DeviceGuard g(device_);
    if (flush_l2) {
        if (!flush_buf_) CudaCheck(cudaMalloc(&flush_buf_, flush_bytes_), "cudaMalloc(flush)");
        CudaCheck(kernels::FlushL2(flush_buf_, flush_bytes_, stream_), "flush L2");
    }
    CudaCheck(cudaEventRecord(ev0_, stream_), "event record");
    Enqueue(n);
    CudaCheck(cudaEventRecord(ev1_, stream_), "event record");
    CudaCheck(cudaEventSynchronize(ev1_), "event sync");
    float ms = 0.f;
    CudaCheck(cudaEventElapsedTime(&ms, ev0_, ev1_), "event elapsed");
    return ms;
}

void Replica::ReadOutput(int output_index, void* host, size_t bytes) {
    std::lock_guard<std::mutex> lk(mu_);
    DeviceGuard g(device_);
    const TensorDesc& t = plan_->tensors[plan_->outputs.at(output_index)];
    bytes = std::min(bytes, (size_t)plan_->max_batch * t.C * t.H * t.W * 4);  // never read past the output buffer
    CudaCheck(cudaMemcpyAsync(host, BufferPtr(t.buffer), bytes, cudaMemcpyDeviceToHost, stream_), "D2H read");
    CudaCheck(cudaStreamSynchronize(stream_), "read sync");
}

std::string Replica::ProfileSteps(int n, int repeats) {
    std::lock_guard<std::mutex> lk(mu_);
    DeviceGuard g(device_);
    const Plan& P = *plan_;
    size_t ns = P.steps.size();
    std::vector<cudaEvent_t> ev(ns + 1);
    for (auto& e : ev) CudaCheck(cudaEventCreate(&e), "cudaEventCreate");
    std::vector<double> ms(ns, 0.0);
    repeats = std::max(1, repeats);
    for (int r = 0; r < repeats + 1; ++r) {  // first pass is warm-up
        CudaCheck(cudaEventRecord(ev[0], stream_), "event");
        for (size_t i = 0; i < ns;) {  // a fused run is timed as one step (its remaining steps read 0)
            const size_t used = EnqueueAt(i, n, 0);
            for (size_t u = 0; u < used; ++u) CudaCheck(cudaEventRecord(ev[i + u + 1], stream_), "event");
            i += used;
        }
        CudaCheck(cudaStreamSynchronize(stream_), "profile sync");
        if (r == 0) continue;
        for (size_t i = 0; i < ns; ++i) {
            float t = 0.f;
            cudaEventElapsedTime(&t, ev[i], ev[i + 1]);
            ms[i] += t;
        }
    }
    for (auto& e : ev) cudaEventDestroy(e);
    std::ostringstream os;
    os.precision(9);
    os << "[";
    for (size_t i = 0; i < ns; ++i) {
        const Step& s = P.steps[i];
        if (i) os << ",";
        os << "{\"step\":" << i << ",\"kind\":\"" << StepKindName(s.kind) << "\",\"name\":\"";
        for (char c : s.name) os << ((c == '"' || c == '\\') ? '_' : c);
        os << "\",\"ms\":" << ms[i] / repeats << ",\"flops\":" << s.flops * n << ",\"bytes\":" << s.bytes * n
           << ",\"umma\":" << ((prepared_[i].use_umma || prepared_[i].use_f32x3) ? "true" : "false");
        if (s.kind == StepKind::Conv)
            os << ",\"Cin\":" << s.Cin << ",\"Cout\":" << s.Cout << ",\"R\":" << s.R << ",\"H\":" << P.tensors[s.out].H;
        os << "}";
    }
    os << "]";
    return os.str();
}

int64_t Replica::ReadValue(const std::string& value_name, float* out, size_t capacity, int n) {
    std::lock_guard<std::mutex> lk(mu_);
    DeviceGuard g(device_);
    auto it = plan_->value_to_tensor.find(value_name);
    if (it == plan_->value_to_tensor.end()) return -1;
    const TensorDesc& t = plan_->tensors[it->second];
    size_t elems = (size_t)n * t.C * t.H * t.W;
    if (elems > capacity) return -1;
    float* d = nullptr;
    CudaCheck(cudaMalloc((void**)&d, elems * 4 + 16), "cudaMalloc(read value)");
    cudaError_t e = kernels::NhwcToNchw(MakeView(it->second), d, n, stream_);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d, elems * 4, cudaMemcpyDeviceToHost, stream_);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream_);
    cudaFree(d);
    CudaCheck(e, "read value");
    return (int64_t)elems;
}

}  // namespace b200
