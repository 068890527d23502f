// model.cpp — inference::Tensor, inference::Model and ModelImpl.
//
// Mirrors the control flow of reference `inference_engine/src/model.cpp`:
//   Load   : file checks + per-type dispatch (:503-548)  -> here: ONNX import + plan + per-GPU replicas
//   Infer  : loaded check, ValidateInputs (:734-794), dispatch, wall-clock stats (:557-613)
//   InferONNX (:1158-1328): inputs looked up BY NAME, outputs returned in graph order, fp32 only
// What differs on purpose (INTEGRATION.md "deviations"): I/O names come from the ONNX graph when the
// configured names do not occur in it (fixes SURVEY.md 0.7), any leading batch dimension is accepted,
// stats are atomics, and nothing executes on the CPU.
#include "model.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <future>
#include <sstream>
#include <sys/stat.h>

#include "cuda_utils.h"
#include "json_lite.h"
#include "model_impl.h"
#include "onnx_wire.h"

namespace inference {

// ============================================================================ Tensor
class Tensor::TensorImpl {
public:
    std::string name;
    DataType dtype = DataType::FLOAT32;
    Shape shape;
    std::vector<uint8_t> host;
    void* dev = nullptr;
    int device_id = 0;

    static size_t ElemSize(DataType t) {
        switch (t) {
            case DataType::FLOAT32: case DataType::INT32: return 4;
            case DataType::INT64: return 8;
            case DataType::FP16: return 2;
            case DataType::UINT8: case DataType::INT8: case DataType::BOOL: return 1;
            default: return 0;  // STRING / UNKNOWN carry no fixed-size payload
        }
    }
    size_t Bytes() const { return shape.NumElements() * ElemSize(dtype); }
    void DropDevice() {
        if (dev) {
            cudaSetDevice(device_id);
            cudaFree(dev);
            dev = nullptr;
        }
    }
    ~TensorImpl() { DropDevice(); }
};

Tensor::Tensor() : impl_(new TensorImpl()) {}
Tensor::Tensor(const std::string& name, DataType dtype, const Shape& shape) : impl_(new TensorImpl()) {
    impl_->name = name;
    impl_->dtype = dtype;
    impl_->shape = shape;
    impl_->host.resize(impl_->Bytes());
}
Tensor::Tensor(const Tensor& o) : impl_(new TensorImpl()) {
    // deep copy of the host payload; a device mirror is never shared (the reference aliased it, model.cpp:205)
    impl_->name = o.impl_->name;
    impl_->dtype = o.impl_->dtype;
    impl_->shape = o.impl_->shape;
    impl_->host = o.impl_->host;
}
Tensor::Tensor(Tensor&& o) noexcept : impl_(std::move(o.impl_)) { o.impl_.reset(new TensorImpl()); }
Tensor& Tensor::operator=(const Tensor& o) {
    if (this != &o) {
        impl_->DropDevice();
        impl_->name = o.impl_->name;
        impl_->dtype = o.impl_->dtype;
        impl_->shape = o.impl_->shape;
        impl_->host = o.impl_->host;
    }
    return *this;
}
Tensor::~Tensor() = default;

const std::string& Tensor::GetName() const { return impl_->name; }
DataType Tensor::GetDataType() const { return impl_->dtype; }
const Shape& Tensor::GetShape() const { return impl_->shape; }
const void* Tensor::RawData() const { return impl_->host.data(); }
void* Tensor::MutableRawData() { return impl_->host.data(); }
size_t Tensor::ByteSize() const { return impl_->host.size(); }

bool Tensor::Reshape(const Shape& new_shape) {
    size_t before = impl_->shape.NumElements();
    impl_->shape = new_shape;  // (the reference forgot this when the element count is unchanged, model.cpp:278-305)
    if (impl_->shape.NumElements() != before) {
        impl_->host.resize(impl_->Bytes());
        impl_->DropDevice();
    }
    return true;
}

bool Tensor::toGPU(int device_id) {
    if (!cuda::IsCudaAvailable()) return false;
    if (impl_->dev) return true;
    if (cudaSetDevice(device_id) != cudaSuccess) return false;
    impl_->device_id = device_id;
    if (cudaMalloc(&impl_->dev, std::max<size_t>(impl_->Bytes(), 1)) != cudaSuccess) { impl_->dev = nullptr; return false; }
    if (cudaMemcpy(impl_->dev, impl_->host.data(), impl_->Bytes(), cudaMemcpyHostToDevice) != cudaSuccess) {
        impl_->DropDevice();
        return false;
    }
    return true;
}
bool Tensor::toCPU() {
    if (!impl_->dev) return true;
    if (cudaSetDevice(impl_->device_id) != cudaSuccess) return false;
    if (cudaMemcpy(impl_->host.data(), impl_->dev, impl_->Bytes(), cudaMemcpyDeviceToHost) != cudaSuccess) return false;
    impl_->DropDevice();
    return true;
}

namespace {
template <typename T>
bool SetTyped(std::vector<uint8_t>& host, const Shape& shape, DataType have, DataType want, const std::vector<T>& data) {
    if (have != want || data.size() != shape.NumElements()) return false;  // reference model.cpp:96-106
    host.resize(data.size() * sizeof(T));
    if (!data.empty()) memcpy(host.data(), data.data(), host.size());
    return true;
}
template <typename T>
bool GetTyped(const std::vector<uint8_t>& host, const Shape& shape, DataType have, DataType want, std::vector<T>& data) {
    if (have != want) return false;
    data.resize(shape.NumElements());
    if (!data.empty()) memcpy(data.data(), host.data(), std::min(host.size(), data.size() * sizeof(T)));
    return true;
}
}  // namespace

template <> bool Tensor::SetData(const std::vector<float>& d) { return SetTyped(impl_->host, impl_->shape, impl_->dtype, DataType::FLOAT32, d); }
template <> bool Tensor::GetData(std::vector<float>& d) const { return GetTyped(impl_->host, impl_->shape, impl_->dtype, DataType::FLOAT32, d); }
template <> bool Tensor::SetData(const std::vector<int>& d) { return SetTyped(impl_->host, impl_->shape, impl_->dtype, DataType::INT32, d); }
template <> bool Tensor::GetData(std::vector<int>& d) const { return GetTyped(impl_->host, impl_->shape, impl_->dtype, DataType::INT32, d); }
template <> bool Tensor::SetData(const std::vector<long>& d) { return SetTyped(impl_->host, impl_->shape, impl_->dtype, DataType::INT64, d); }
template <> bool Tensor::GetData(std::vector<long>& d) const { return GetTyped(impl_->host, impl_->shape, impl_->dtype, DataType::INT64, d); }
template <> bool Tensor::SetData(const std::vector<uint8_t>& d) { return SetTyped(impl_->host, impl_->shape, impl_->dtype, DataType::UINT8, d); }
template <> bool Tensor::GetData(std::vector<uint8_t>& d) const { return GetTyped(impl_->host, impl_->shape, impl_->dtype, DataType::UINT8, d); }

// ============================================================================ ModelImpl
namespace {

bool PathExists(const std::string& p) {
    struct stat st;
    return stat(p.c_str(), &st) == 0;
}

std::string EnvOr(const char* name, const std::string& dflt) {
    const char* v = getenv(name);
    return (v && *v) ? std::string(v) : dflt;
}

std::vector<int> ParseDeviceList(const std::string& spec, int device_count, int default_device) {
    std::vector<int> out;
    if (spec.empty() || spec == "all") {
        for (int i = 0; i < device_count; ++i) out.push_back(i);
        return out;
    }
    if (spec == "default") return {default_device};
    std::stringstream ss(spec);
    std::string tok;
    while (std::getline(ss, tok, ',')) {
        if (tok.empty()) continue;
        int d = atoi(tok.c_str());
        if (d < 0 || d >= device_count) throw std::runtime_error("B200_ENGINE_DEVICES names device " + tok + " but only " + std::to_string(device_count) + " are visible");
        out.push_back(d);
    }
    if (out.empty()) out.push_back(default_device);
    return out;
}

}  // namespace

ModelImpl::ModelImpl(const std::string& model_path, ModelType type, const ModelConfig& config, DeviceType device, int device_id)
    : model_path_(model_path), type_(type), config_(config), device_type_(device), device_id_(device_id) {
    metadata_.name = config_.name;
    metadata_.version = config_.version;
    metadata_.type = type_;
    metadata_.inputs = config_.input_names;
    metadata_.outputs = config_.output_names;
    metadata_.load_time_ns = 0;
}

ModelImpl::~ModelImpl() { Unload(); }

std::string ModelImpl::GetLastError() const {
    std::lock_guard<std::mutex> lk(err_mu_);
    return last_error_;
}
void ModelImpl::SetLastError(const std::string& e) const {
    std::lock_guard<std::mutex> lk(err_mu_);
    last_error_ = e;
}
ModelMetadata ModelImpl::GetMetadata() const {
    std::lock_guard<std::mutex> lk(state_mu_);
    return metadata_;
}
Model::Stats ModelImpl::GetStats() const {
    Model::Stats s;
    s.inference_count = inference_count_.load();
    s.total_inference_time_ns = total_ns_.load();
    s.last_inference_time_ns = last_ns_.load();
    s.memory_usage_bytes = memory_bytes_.load();
    return s;
}
std::shared_ptr<ModelImpl::Loaded> ModelImpl::Pin() const {
    std::lock_guard<std::mutex> lk(state_mu_);
    return state_;
}

bool ModelImpl::Load() {
    auto t0 = std::chrono::high_resolution_clock::now();
    if (IsLoaded()) return true;
    if (!PathExists(model_path_)) {
        SetLastError("Model file not found: " + model_path_);
        return false;
    }
    switch (type_) {
        case ModelType::ONNX: break;
        case ModelType::TENSORFLOW: SetLastError("TensorFlow model loading not implemented"); return false;
        case ModelType::TENSORRT: SetLastError("TensorRT model loading not implemented"); return false;
        case ModelType::PYTORCH: SetLastError("PyTorch model loading not implemented"); return false;
        case ModelType::CUSTOM: SetLastError("Custom model loading not implemented"); return false;
        default: SetLastError("Unsupported model type"); return false;
    }
    const std::string onnx_path = model_path_ + "/model.onnx";
    if (!PathExists(onnx_path)) {
        SetLastError("ONNX model file not found: " + onnx_path);
        return false;
    }
    try {
        if (device_type_ != DeviceType::GPU)
            throw std::runtime_error("this engine executes on NVIDIA B200 (sm_100a) only; DeviceType::CPU is not available");
        int ndev = cuda::GetDeviceCount();
        if (ndev <= 0) throw std::runtime_error("no CUDA device is visible; this engine has no CPU execution path");

        // optional per-model settings (config.json next to model.onnx), environment wins
        std::string precision_s = "fp32";
        int max_batch = config_.max_batch_size > 0 ? config_.max_batch_size : 256;
        {
            std::ifstream cf(model_path_ + "/config.json");
            if (cf) {
                std::stringstream ss;
                ss << cf.rdbuf();
                try {
                    b200::json::Value v = b200::json::ParseString(ss.str());
                    if (auto* p = v.Get("precision")) if (p->kind == b200::json::Value::String) precision_s = p->str;
                    if (auto* p = v.Get("max_batch_size")) if (p->kind == b200::json::Value::Number && p->num >= 1) max_batch = (int)p->num;
                    if (auto* p = v.Get("dynamic_batching")) if (p->kind == b200::json::Value::Bool && p->b) config_.dynamic_batching = true;
                } catch (...) { /* a malformed config.json never blocks loading (the reference ignores it) */ }
            }
        }
        // request coalescing window: the reference's dead `dynamic_batching` flag switches it on (200 us), the environment wins
        coalesce_us_ = atoi(EnvOr("B200_ENGINE_COALESCE_US", config_.dynamic_batching ? "200" : "0").c_str());
        coalesce_small_ = std::max(1, atoi(EnvOr("B200_ENGINE_COALESCE_MAX_REQUEST", "8").c_str()));
        stage_pageable_ = EnvOr("B200_ENGINE_STAGE_PAGEABLE", "1") != "0";
        // a request is split over GPUs only in shards of at least this many samples (bs256 on 8 GPUs -> 32 each);
        // smaller requests go whole to one GPU, round-robin, which is what keeps mixed concurrent traffic efficient
        min_shard_ = std::max(1, atoi(EnvOr("B200_ENGINE_MIN_SHARD", "32").c_str()));
        precision_s = EnvOr("B200_ENGINE_PRECISION", precision_s);
        max_batch = atoi(EnvOr("B200_ENGINE_MAX_BATCH", std::to_string(max_batch)).c_str());
        if (max_batch < 1) max_batch = 1;
        b200::Precision precision;
        if (!b200::ParsePrecision(precision_s, &precision)) throw std::runtime_error("unknown precision '" + precision_s + "' (use fp32, bf16 or fp8)");

        b200::onnx::Model om = b200::onnx::ParseFile(onnx_path);
        auto plan = std::make_shared<b200::Plan>(b200::BuildPlan(om, precision, max_batch));

        std::vector<int> devices = ParseDeviceList(EnvOr("B200_ENGINE_DEVICES", "all"), ndev, device_id_);
        bool graphs = EnvOr("B200_ENGINE_GRAPHS", "1") != "0";
        auto st = std::make_shared<Loaded>();
        st->plan = plan;
        size_t mem = 0;
        const bool chain_on = EnvOr("B200_ENGINE_CHAIN", "1") != "0";
        std::vector<std::shared_ptr<b200::ComputeChain>> chains;
        for (int d : devices) {
            chains.push_back(chain_on ? std::make_shared<b200::ComputeChain>(d) : nullptr);
            st->replicas.emplace_back(new b200::Replica(d, plan, graphs, chains.back()));
            mem += st->replicas.back()->DeviceBytes();
        }
        // execution instances per GPU: config.json / ModelConfig `instance_count`, default 4 (measured on the mixed-size replay: 2 -> 38.7 k img/s, 4 -> 55.6 k), the environment wins
        st->instances = std::min(8, std::max(1, atoi(EnvOr("B200_ENGINE_INSTANCES", std::to_string(std::max(4, config_.instance_count))).c_str())));
        for (size_t di = 0; di < devices.size(); ++di)
            for (int j = 1; j < st->instances; ++j) {
                const int d = devices[di];
                st->extra.emplace_back(new b200::Replica(d, plan, graphs, chains[di], st->replicas[di].get()));
                mem += st->extra.back()->DeviceBytes();
            }
        st->faulted = std::vector<std::atomic<int>>(st->replicas.size());
        for (auto& f : st->faulted) f.store(0);
        if (st->replicas.size() > 1) st->workers.reset(new GpuWorkers((int)st->replicas.size(), st->instances));
        st->inject_fault = atoi(EnvOr("B200_ENGINE_FAULT_REPLICA", "-1").c_str());
        st->busy.assign(st->replicas.size() * st->instances, 0);
        st->next_ticket.assign(st->replicas.size(), 0);
        st->serving.assign(st->replicas.size(), 0);
        memory_bytes_.store(mem);

        {
            std::lock_guard<std::mutex> lk(state_mu_);
            // Names: keep the configured ones when they exist in the graph, else adopt the graph's
            // (reference defaults to "input"/"output", model_repository.cpp:143-144, which rejects DenseNet).
            auto all_in = [&](const std::vector<std::string>& want, const std::vector<std::string>& have) {
                if (want.size() != have.size()) return false;
                for (auto& w : want) if (std::find(have.begin(), have.end(), w) == have.end()) return false;
                return true;
            };
            if (!all_in(config_.input_names, plan->input_names)) config_.input_names = plan->input_names;
            if (!all_in(config_.output_names, plan->output_names)) config_.output_names = plan->output_names;
            for (size_t i = 0; i < plan->input_names.size(); ++i) {
                Shape s;
                s.dims = plan->input_dims[i];  // -1 batch wildcard
                config_.input_shapes[plan->input_names[i]] = s;
                config_.input_types[plan->input_names[i]] = DataType::FLOAT32;
            }
            for (size_t i = 0; i < plan->output_names.size(); ++i) {
                Shape s;
                s.dims = plan->output_dims[i];
                config_.output_shapes[plan->output_names[i]] = s;
                config_.output_types[plan->output_names[i]] = DataType::FLOAT32;
            }
            metadata_.inputs = config_.input_names;
            metadata_.outputs = config_.output_names;
            metadata_.description = std::string("b200-engine ") + b200::PrecisionName(precision) + " plan: " +
                                    std::to_string(plan->steps.size()) + " steps, " + std::to_string(devices.size()) + " GPU replica(s)";
            state_ = st;
        }
        loaded_.store(true, std::memory_order_release);
    } catch (const std::exception& e) {
        SetLastError(std::string("ONNX model loading error: ") + e.what());
        return false;
    }
    auto t1 = std::chrono::high_resolution_clock::now();
    {
        std::lock_guard<std::mutex> lk(state_mu_);
        metadata_.load_time_ns = std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
    }
    return true;
}

void ModelImpl::Unload() {
    std::shared_ptr<Loaded> old;
    {
        std::lock_guard<std::mutex> lk(state_mu_);
        old.swap(state_);
    }
    loaded_.store(false, std::memory_order_release);
    memory_bytes_.store(0);
    if (old) b200::PinnedPool::Get().Trim();
    // `old` (replicas, arenas) is released here unless an in-flight Infer still pins it.
}

bool ModelImpl::ValidateInputs(const std::vector<IoDesc>& ins) const {
    // reference model.cpp:734-794, same messages
    if (ins.size() != config_.input_names.size()) {
        SetLastError("Expected " + std::to_string(config_.input_names.size()) + " inputs, got " + std::to_string(ins.size()));
        return false;
    }
    for (const auto& in : ins) {
        if (std::find(config_.input_names.begin(), config_.input_names.end(), in.name) == config_.input_names.end()) {
            SetLastError("Unexpected input name: " + in.name);
            return false;
        }
        auto ti = config_.input_types.find(in.name);
        auto si = config_.input_shapes.find(in.name);
        // Extension: an image input declared FLOAT32 [N,C,H,W] with C <= 4 also accepts the raw
        // UINT8 [N,H,W,C] pixels; the GPU applies value / 255 and the layout change (client/test_client.py:186-194).
        if (in.dtype == DataType::UINT8 && ti != config_.input_types.end() && ti->second == DataType::FLOAT32 &&
            si != config_.input_shapes.end() && si->second.dims.size() == 4 && in.dims.size() == 4) {
            const auto& want = si->second.dims;
            if (want[1] >= 1 && want[1] <= 4 && in.dims[3] == want[1] && (want[2] == -1 || in.dims[1] == want[2]) &&
                (want[3] == -1 || in.dims[2] == want[3]) && (want[0] == -1 || want[0] == in.dims[0]))
                continue;
            SetLastError("Input shape mismatch for " + in.name + ": a UINT8 image must be [N,H,W,C]");
            return false;
        }
        if (ti != config_.input_types.end() && ti->second != in.dtype) {
            SetLastError("Input data type mismatch for " + in.name);
            return false;
        }
        if (si != config_.input_shapes.end()) {
            const auto& want = si->second.dims;
            if (want.size() != in.dims.size()) {
                SetLastError("Input shape mismatch for " + in.name + ": expected " + std::to_string(want.size()) +
                             " dimensions, got " + std::to_string(in.dims.size()));
                return false;
            }
            for (size_t i = 0; i < want.size(); ++i)
                if (want[i] != -1 && want[i] != in.dims[i]) {
                    SetLastError("Input shape mismatch for " + in.name + " at dimension " + std::to_string(i));
                    return false;
                }
        }
    }
    return true;
}

bool ModelImpl::Infer(const std::vector<Tensor>& inputs, std::vector<Tensor>& outputs) {
    std::vector<IoDesc> ins(inputs.size());
    for (size_t i = 0; i < inputs.size(); ++i) {
        ins[i].name = inputs[i].GetName();
        ins[i].dtype = inputs[i].GetDataType();
        ins[i].dims = inputs[i].GetShape().dims;
        ins[i].data = inputs[i].RawData();
        ins[i].bytes = inputs[i].ByteSize();
    }
    auto st = Pin();
    if (!IsLoaded() || !st) {
        SetLastError("Model not loaded");
        return false;
    }
    // Output tensors are created by the engine in graph order (reference model.cpp:1273-1314 clears and refills).
    int64_t n = (!ins.empty() && !ins[0].dims.empty()) ? ins[0].dims[0] : 0;
    outputs.clear();
    std::vector<OutDesc> outs(st->plan->outputs.size());
    if (n > 0) {
        for (size_t i = 0; i < outs.size(); ++i) {
            Shape s;
            s.dims = st->plan->output_dims[i];
            s.dims[0] = n;
            outputs.emplace_back(st->plan->output_names[i], DataType::FLOAT32, s);
        }
        for (size_t i = 0; i < outs.size(); ++i) {
            outs[i].data = outputs[i].MutableRawData();
            outs[i].capacity = outputs[i].ByteSize();
        }
    }
    bool ok = InferBorrowed(ins, outs);
    if (!ok) outputs.clear();
    return ok;
}

bool ModelImpl::InferBorrowed(const std::vector<IoDesc>& ins, std::vector<OutDesc>& outs) {
    auto st = Pin();
    if (!IsLoaded() || !st) {
        SetLastError("Model not loaded");
        return false;
    }
    if (!ValidateInputs(ins)) return false;
    auto t0 = std::chrono::high_resolution_clock::now();
    bool ok = false;
    try {
        const b200::Plan& P = *st->plan;
        // by-name binding of graph inputs (reference model.cpp:1173-1222)
        std::vector<const void*> ptrs(P.input_names.size(), nullptr);
        unsigned u8_mask = 0;
        int64_t n = -1;
        for (size_t gi = 0; gi < P.input_names.size(); ++gi) {
            const IoDesc* found = nullptr;
            for (const auto& in : ins) if (in.name == P.input_names[gi]) { found = &in; break; }
            if (!found) throw std::runtime_error("Required input tensor not provided: " + P.input_names[gi]);
            const auto& gd = P.input_dims[gi];
            if (found->dtype == DataType::UINT8 && gd.size() == 4 && found->dims.size() == 4 && gi < 8) {
                // raw [N,H,W,C] pixels for a [N,C,H,W] graph input
                if (found->dims[0] < 1 || found->dims[1] != gd[2] || found->dims[2] != gd[3] || found->dims[3] != gd[1])
                    throw std::runtime_error("Input shape mismatch for " + found->name + ": a UINT8 image must be [N,H,W,C]");
                if (n < 0) n = found->dims[0];
                else if (n != found->dims[0]) throw std::runtime_error("Inputs disagree on the batch dimension");
                if (!found->data || found->bytes < (size_t)n * gd[1] * gd[2] * gd[3]) throw std::runtime_error("Invalid UINT8 data for input: " + found->name);
                ptrs[gi] = found->data;
                u8_mask |= 1u << gi;
                continue;
            }
            if (found->dtype != DataType::FLOAT32) throw std::runtime_error("Unsupported data type for input: " + found->name);
            if (found->dims.size() != gd.size() || found->dims.empty() || found->dims[0] < 1)
                throw std::runtime_error("Invalid FLOAT32 data for input: " + found->name);
            size_t per = 1;
            for (size_t k = 1; k < gd.size(); ++k) {
                if (found->dims[k] != gd[k]) throw std::runtime_error("Input shape mismatch for " + found->name + " at dimension " + std::to_string(k));
                per *= (size_t)gd[k];
            }
            if (n < 0) n = found->dims[0];
            else if (n != found->dims[0]) throw std::runtime_error("Inputs disagree on the batch dimension");
            if (!found->data || found->bytes < (size_t)n * per * 4) throw std::runtime_error("Invalid FLOAT32 data for input: " + found->name);
            ptrs[gi] = found->data;
        }
        // pageable request buffers are staged into page-locked memory here, on the caller's thread (see PinnedPool)
        struct Staged {
            std::vector<void*> bufs;
            ~Staged() { for (void* b : bufs) b200::PinnedPool::Get().Give(b); }
        } staged;
        if (stage_pageable_) {
            // the pointer query below must not create a primary context on a device this model does not use
            cudaSetDevice(st->replicas[0]->device());
            for (size_t gi = 0; gi < ptrs.size(); ++gi) {
                const auto& gd = P.input_dims[gi];
                size_t bytes = (size_t)n * (((u8_mask >> gi) & 1u) ? 1 : 4);
                for (size_t k = 1; k < gd.size(); ++k) bytes *= (size_t)gd[k];
                if (bytes < (64u << 10) || !b200::PinnedPool::IsPageable(ptrs[gi])) continue;
                void* pin = b200::PinnedPool::Get().Take(bytes);
                if (!pin) continue;  // budget exhausted: the driver's pageable path still works
                staged.bufs.push_back(pin);
                const size_t kPar = 16u << 20;  // big buffers: four copy threads
                if (bytes >= 2 * kPar) {
                    const size_t q = (bytes / 4 + 63) & ~(size_t)63;
                    const char* src = (const char*)ptrs[gi];
                    std::vector<std::future<void>> f;
                    for (int t = 1; t < 4; ++t)
                        f.push_back(std::async(std::launch::async, [pin, src, q, bytes, t] {
                            const size_t o = t * q;
                            if (o < bytes) memcpy((char*)pin + o, src + o, std::min(q, bytes - o));
                        }));
                    memcpy(pin, ptrs[gi], std::min(q, bytes));
                    for (auto& x : f) x.get();
                } else {
                    memcpy(pin, ptrs[gi], bytes);
                }
                ptrs[gi] = pin;
            }
        }
        const bool wants_topk = !outs.empty() && outs[0].topk > 0;  // top-k requests carry their own result arrays: not coalesced
        if (coalesce_us_ > 0 && n <= coalesce_small_ && n < P.max_batch && !wants_topk) ok = Coalesce(st, (int)n, ptrs, outs, u8_mask);
        else ok = Execute(*st, (int)n, ptrs, outs, u8_mask);
    } catch (const std::exception& e) {
        SetLastError(std::string("ONNX inference error: ") + e.what());
        ok = false;
    }
    auto t1 = std::chrono::high_resolution_clock::now();
    int64_t ns = std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
    inference_count_.fetch_add(1);
    total_ns_.fetch_add(ns);
    last_ns_.store(ns);
    return ok;
}

// ---------------------------------------------------------------------------------------------------------------------
// Request coalescer.  gin serves every REST request on its own goroutine / OS thread and the reference's handler sends
// batch 1, so concurrent callers arrive here one image at a time.  The first caller of a quiet period
// becomes the leader: it waits up to `coalesce_us_` (or until the arena is full) while followers append their requests,
// then runs everything that accumulated as ONE batch - each request's input is copied from its own buffer to its sample
// offset, each result goes straight back to its own output buffer - and wakes the followers.  Requests left over when
// the leader closes its batch elect (promote) their own leader immediately.
bool ModelImpl::Coalesce(const std::shared_ptr<Loaded>& st, int n, const std::vector<const void*>& ptrs, std::vector<OutDesc>& outs,
                         unsigned u8_mask) {
    Pending me;
    me.n = n; me.u8_mask = u8_mask; me.ptrs = ptrs; me.outs = &outs;
    const int cap = st->plan->max_batch;
    std::unique_lock<std::mutex> lk(co_mu_);
    co_queue_.push_back(&me);
    bool lead = !co_leader_;
    bool waited = false;
    for (;;) {
        if (!lead) {
            co_cv_.notify_one();  // the leader re-checks whether its batch is full
            me.cv.wait(lk, [&] { return me.done || me.promoted; });
            if (me.done) {
                if (!me.ok) SetLastError(me.err);
                return me.ok;
            }
            lead = true;     // promoted: the requests that did not fit the previous batch already waited their window
            waited = true;
        }
        co_leader_ = true;
        auto queued = [&] { int s = 0; for (auto* p : co_queue_) s += p->n; return s; };
        if (!waited) co_cv_.wait_for(lk, std::chrono::microseconds(coalesce_us_), [&] { return queued() >= cap; });
        // batch while busy: as long as every replica is already executing a batch, keep collecting - under load the batch
        // grows to whatever arrives during one forward, when idle a request only ever waits its window
        const int G = 2 * (int)st->replicas.size();  // per GPU: one batch computing, one copying
        while (co_inflight_ >= G && queued() < cap) co_cv_.wait_for(lk, std::chrono::microseconds(100));
        // close the batch: queue order, same input kind as the first request, total <= cap.  The leader's own request is always
        // in it: a leader is either the front of the queue (fresh quiet period) or was promoted as the front of what was left,
        // and a single request never exceeds cap (checked by the caller).
        std::vector<Pending*> batch, rest;
        int total = 0;
        try {
            for (auto* p : co_queue_) {
                if (p->u8_mask == co_queue_.front()->u8_mask && total + p->n <= cap) { batch.push_back(p); total += p->n; }
                else rest.push_back(p);
            }
        } catch (...) {  // bad_alloc while forming the batch: nobody may be left waiting for a leader that is gone
            for (auto* p : co_queue_)
                if (p != &me) { p->ok = false; p->err = "ONNX inference error: out of memory while batching requests"; p->done = true; p->cv.notify_one(); }
            co_queue_.clear();
            co_leader_ = false;
            throw;
        }
        if (std::find(batch.begin(), batch.end(), &me) == batch.end()) {
            // cannot happen (see above); if it ever does, run my request alone rather than lead an empty batch
            batch.clear();
            rest.clear();
            for (auto* p : co_queue_) (p == &me ? batch : rest).push_back(p);
        }
        co_queue_.swap(rest);
        co_leader_ = false;
        if (!co_queue_.empty()) {  // somebody else must lead what is left
            Pending* next = co_queue_.front();
            co_leader_ = true;
            next->promoted = true;
            next->cv.notify_one();
        }
        ++co_inflight_;
        lk.unlock();
        std::string err;
        bool ok = true;
        try {
            RunCoalesced(*st, batch);
        } catch (const std::exception& e) {
            ok = false;
            err = std::string("ONNX inference error: ") + e.what();
        }
        co_batches_.fetch_add(1);
        co_requests_.fetch_add((int64_t)batch.size());
        lk.lock();
        --co_inflight_;
        co_cv_.notify_all();  // a leader that kept collecting because every replica was busy may go now
        for (auto* p : batch) {
            if (p == &me) continue;
            p->ok = ok; p->err = err; p->done = true;
            p->cv.notify_one();
        }
        if (!ok) SetLastError(err);
        return ok;
    }
}

void ModelImpl::RunCoalesced(Loaded& st, const std::vector<Pending*>& batch) {
    const b200::Plan& P = *st.plan;
    if (batch.empty()) return;
    std::vector<b200::Replica::Segment> segs(batch.size());
    for (size_t k = 0; k < batch.size(); ++k) {
        Pending& p = *batch[k];
        b200::Replica::Segment& s = segs[k];
        s.n = p.n;
        s.in = p.ptrs;
        std::vector<OutDesc>& outs = *p.outs;
        s.out.assign(outs.size(), nullptr);
        s.cap.assign(outs.size(), 0);
        for (size_t i = 0; i < outs.size() && i < P.outputs.size(); ++i) {
            const auto& t = P.tensors[P.outputs[i]];
            outs[i].name = P.output_names[i];
            outs[i].dims = P.output_dims[i];
            outs[i].dims[0] = p.n;
            outs[i].produced = (size_t)p.n * t.C * t.H * t.W * 4;
            s.out[i] = outs[i].data;
            s.cap[i] = outs[i].data ? outs[i].capacity : 0;
        }
    }
    for (;;) {
        std::vector<int> healthy = st.Healthy();
        if (healthy.empty()) throw std::runtime_error("every GPU replica of this model has faulted; unload and load the model again");
        const int g = healthy[round_robin_.fetch_add(1) % (unsigned)healthy.size()];
        int slot = 0;
        b200::Replica* r = st.Acquire(g, &slot);
        try {
            r->RunSegments(segs, batch.front()->u8_mask);
            st.Release(slot);
            return;
        } catch (const b200::CudaError& e) {
            st.Release(slot);
            st.faulted[g].store(1);
            fprintf(stderr, "[b200-engine] GPU replica %d dropped from the shard set: %s\n", g, e.what());
        } catch (...) {
            st.Release(slot);
            throw;
        }
    }
}

// std::mutex hands a contended lock to whoever gets there first, which under 32+ request threads starves some callers for
// seconds (measured p99 > 1 s at p50 1 ms); a ticket per GPU makes the wait first come, first served.
b200::Replica* ModelImpl::Loaded::Acquire(int g, int* slot, bool* alone) {
    std::unique_lock<std::mutex> lk(pick_mu);
    const uint64_t my = next_ticket[g]++;
    int free_j = -1;
    pick_cv.wait(lk, [&] {
        if (serving[g] != my) return false;
        for (int j = 0; j < instances; ++j)
            if (!busy[g * instances + j]) { free_j = j; return true; }
        return false;
    });
    ++serving[g];
    if (alone) {
        *alone = next_ticket[g] == serving[g];  // nobody queued behind me ...
        for (int j = 0; j < instances; ++j) *alone = *alone && !busy[g * instances + j];  // ... and no instance at work
    }
    *slot = g * instances + free_j;
    busy[*slot] = 1;
    pick_cv.notify_all();
    return free_j == 0 ? replicas[g].get() : extra[g * (instances - 1) + free_j - 1].get();
}
void ModelImpl::Loaded::Release(int slot) {
    {
        std::lock_guard<std::mutex> lk(pick_mu);
        busy[slot] = 0;
    }
    pick_cv.notify_all();
}

// Shard planner (pure function, unit-tested on CPU through B200PlanShards): contiguous split of `n` samples
// over min(G, n / min_shard) replicas, each shard cut into arena-sized chunks; batches too small to split go
// to the single replica `round_robin` so that concurrent small requests spread over the GPUs.
std::vector<ShardPlan> PlanShards(int n, int G, int max_batch, int min_shard, int round_robin) {
    std::vector<ShardPlan> shards;
    if (n <= 0 || G <= 0) return shards;
    max_batch = std::max(1, max_batch);
    min_shard = std::max(1, min_shard);
    int use = std::min(G, std::max(1, n / min_shard));
    if (use <= 1) {
        int r = ((round_robin % G) + G) % G;
        for (int off = 0; off < n; off += max_batch) shards.push_back({r, off, std::min(max_batch, n - off)});
        return shards;
    }
    int base = n / use, rem = n % use, off = 0;
    for (int g = 0; g < use; ++g) {
        int cnt = base + (g < rem ? 1 : 0);
        for (int o = 0; o < cnt; o += max_batch) shards.push_back({g, off + o, std::min(max_batch, cnt - o)});
        off += cnt;
    }
    return shards;
}

GpuWorkers::GpuWorkers(int gpus, int per_gpu) {
    for (int g = 0; g < gpus; ++g) queues_.emplace_back(new Queue());
    for (int g = 0; g < gpus; ++g)
        for (int j = 0; j < std::max(1, per_gpu); ++j)
            threads_.emplace_back([this, g] {
                Queue& q = *queues_[g];
                for (;;) {
                    std::packaged_task<void()> task;
                    {
                        std::unique_lock<std::mutex> lk(q.mu);
                        q.cv.wait(lk, [&] { return q.stop || !q.tasks.empty(); });
                        if (q.tasks.empty()) return;  // stop requested and drained
                        task = std::move(q.tasks.front());
                        q.tasks.pop_front();
                    }
                    task();
                }
            });
}
GpuWorkers::~GpuWorkers() {
    for (auto& q : queues_) {
        { std::lock_guard<std::mutex> lk(q->mu); q->stop = true; }
        q->cv.notify_all();
    }
    for (auto& t : threads_) t.join();
}
std::future<void> GpuWorkers::Submit(int gpu, std::function<void()> fn) {
    std::packaged_task<void()> task(std::move(fn));
    std::future<void> fut = task.get_future();
    Queue& q = *queues_[gpu];
    { std::lock_guard<std::mutex> lk(q.mu); q.tasks.push_back(std::move(task)); }
    q.cv.notify_one();
    return fut;
}

std::vector<int> ModelImpl::Loaded::Healthy() const {
    std::vector<int> h;
    for (size_t g = 0; g < replicas.size(); ++g)
        if (!faulted[g].load(std::memory_order_relaxed)) h.push_back((int)g);
    return h;
}
int ModelImpl::FaultedReplicas() const {
    auto st = Pin();
    int n = 0;
    if (st) for (auto& f : st->faulted) n += f.load() ? 1 : 0;
    return n;
}

// The multi-GPU batch scheduler: contiguous split of the batch over the healthy replicas, no collective.
// Small batches go whole to one replica chosen round-robin so concurrent callers spread.  Shards of a split request run on the
// persistent per-GPU workers; a shard whose GPU raises a CUDA error marks that replica faulted (it leaves the shard set for good)
// and is re-run on a healthy one.
bool ModelImpl::Execute(Loaded& st, int n, const std::vector<const void*>& in_ptrs, std::vector<OutDesc>& outs, unsigned u8_mask) {
    const b200::Plan& P = *st.plan;
    const int max_b = P.max_batch;
    std::vector<size_t> in_stride(P.inputs.size()), out_stride(P.outputs.size());
    for (size_t i = 0; i < P.inputs.size(); ++i) {
        const auto& t = P.tensors[P.inputs[i]];
        in_stride[i] = (size_t)t.C * t.H * t.W * (((u8_mask >> i) & 1u) ? 1 : 4);
    }
    for (size_t i = 0; i < P.outputs.size(); ++i) {
        const auto& t = P.tensors[P.outputs[i]];
        out_stride[i] = (size_t)t.C * t.H * t.W * 4;
    }
    for (size_t i = 0; i < outs.size(); ++i) {
        if (i >= P.outputs.size()) break;
        outs[i].name = P.output_names[i];
        outs[i].dims = P.output_dims[i];
        outs[i].dims[0] = n;
        outs[i].produced = (size_t)n * out_stride[i];
    }
    using Shard = ShardPlan;
    std::vector<int> healthy = st.Healthy();
    if (healthy.empty()) throw std::runtime_error("every GPU replica of this model has faulted; unload and load the model again");
    const int rr = (int)(round_robin_.fetch_add(1) % (unsigned)healthy.size());
    std::vector<Shard> shards = PlanShards(n, (int)healthy.size(), max_b, min_shard_, rr);
    for (auto& s : shards) s.replica = healthy[s.replica];  // shard-set ordinal -> replica index
    bool single = true;
    for (auto& s : shards) single = single && s.replica == shards[0].replica;

    auto run_on = [&](const Shard& s, int g) {
        std::vector<const void*> ip(P.inputs.size());
        std::vector<void*> op(outs.size(), nullptr);
        std::vector<size_t> cap(outs.size(), 0);
        for (size_t i = 0; i < ip.size(); ++i) ip[i] = (const char*)in_ptrs[i] + (size_t)s.off * in_stride[i];
        for (size_t i = 0; i < outs.size() && i < P.outputs.size(); ++i) {
            size_t begin = (size_t)s.off * out_stride[i];
            if (!outs[i].data || begin >= outs[i].capacity) continue;
            op[i] = (char*)outs[i].data + begin;
            cap[i] = outs[i].capacity - begin;
        }
        b200::Replica::TopK tk;
        if (!outs.empty() && outs[0].topk > 0) {
            tk.k = outs[0].topk;
            tk.softmax = outs[0].topk_softmax;
            tk.idx = outs[0].topk_idx + (size_t)s.off * tk.k;
            tk.val = outs[0].topk_val + (size_t)s.off * tk.k;
        }
        if (g == st.inject_fault) throw b200::CudaError("injected fault (B200_ENGINE_FAULT_REPLICA)");
        int slot = 0;
        bool alone = true;
        b200::Replica* r = st.Acquire(g, &slot, &alone);
        try { r->Run(s.cnt, ip, op, cap, u8_mask, alone, tk.k > 0 ? &tk : nullptr); } catch (...) { st.Release(slot); throw; }
        st.Release(slot);
    };
    // a CUDA error takes the replica out of the shard set and the shard moves to another healthy replica; any other error
    // (bad request) propagates unchanged
    auto run_shard = [&](const Shard& s) {
        int g = s.replica;
        for (;;) {
            try {
                run_on(s, g);
                return;
            } catch (const b200::CudaError& e) {
                st.faulted[g].store(1);
                fprintf(stderr, "[b200-engine] GPU replica %d dropped from the shard set: %s\n", g, e.what());
                std::vector<int> h = st.Healthy();
                if (h.empty()) throw;
                g = h[(size_t)s.off % h.size()];
            }
        }
    };
    if (single || !st.workers) {
        for (auto& s : shards) run_shard(s);
        return true;
    }
    // one task per participating replica on that GPU's persistent worker; shards of the same replica run back to back
    std::vector<int> used;
    for (auto& s : shards)
        if (std::find(used.begin(), used.end(), s.replica) == used.end()) used.push_back(s.replica);
    std::vector<std::future<void>> futs;
    for (size_t u = 1; u < used.size(); ++u) {
        const int g = used[u];
        futs.push_back(st.workers->Submit(g, [&, g] {
            for (auto& s : shards) if (s.replica == g) run_shard(s);
        }));
    }
    std::exception_ptr first;
    try {
        for (auto& s : shards) if (s.replica == used[0]) run_shard(s);
    } catch (...) { first = std::current_exception(); }
    for (auto& f : futs) {
        try { f.get(); } catch (...) { if (!first) first = std::current_exception(); }
    }
    if (first) std::rethrow_exception(first);
    return true;
}

// ============================================================================ Model (thin PIMPL shell)
Model::Model(const std::string& model_path, ModelType type, const ModelConfig& config, DeviceType device, int device_id)
    : impl_(new ModelImpl(model_path, type, config, device, device_id)) {}
Model::~Model() = default;
Model::Model(Model&&) noexcept = default;
Model& Model::operator=(Model&&) noexcept = default;
bool Model::Load() { return impl_->Load(); }
bool Model::Infer(const std::vector<Tensor>& inputs, std::vector<Tensor>& outputs) { return impl_->Infer(inputs, outputs); }
ModelMetadata Model::GetMetadata() const { return impl_->GetMetadata(); }
bool Model::IsLoaded() const { return impl_->IsLoaded(); }
void Model::Unload() { impl_->Unload(); }
std::string Model::GetLastError() const { return impl_->GetLastError(); }
Model::Stats Model::GetStats() const { return impl_->GetStats(); }

}  // namespace inference
