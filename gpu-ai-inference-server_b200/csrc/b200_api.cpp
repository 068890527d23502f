// b200_api.cpp — extension entry points declared in include/b200_engine.h (measurement, plan
// inspection, parity debugging).  Not part of the reference ABI.
#include "b200_engine.h"

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <future>
#include <memory>
#include <string>
#include <vector>

#include "engine.h"
#include "kernels.h"
#include "model.h"
#include "model_impl.h"
#include "onnx_wire.h"
#include "plan.h"

std::shared_ptr<inference::Model> B200ModelFromHandle(ModelHandle h);

namespace {
void SetErr(ErrorMessage* error, const std::string& msg) {
    if (error) *error = strdup(msg.c_str());
}
std::shared_ptr<inference::ModelImpl::Loaded> PinHandle(ModelHandle h, std::shared_ptr<inference::Model>* keep, ErrorMessage* error) {
    *keep = B200ModelFromHandle(h);
    if (!*keep) {
        SetErr(error, "Invalid model handle");
        return nullptr;
    }
    auto st = (*keep)->Impl()->Pin();
    if (!st) SetErr(error, "Model not loaded");
    return st;
}
}  // namespace

extern "C" {

void* B200HostAlloc(size_t bytes) {
    void* p = nullptr;
    if (bytes == 0 || cudaHostAlloc(&p, bytes, cudaHostAllocPortable) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
void B200HostFree(void* ptr) {
    if (ptr) cudaFreeHost(ptr);
}

bool B200MeasureH2D(int gpus, size_t bytes, int iters, int write_combined, double* gbs_single, double* gbs_all, ErrorMessage* error) {
    try {
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) throw std::runtime_error("no CUDA device");
        gpus = std::max(1, std::min(gpus, ndev));
        iters = std::max(1, iters);
        if (bytes < (1u << 20) || !gbs_single || !gbs_all) throw std::runtime_error("Invalid parameters");
        struct Lane { void* host = nullptr; void* dev = nullptr; cudaStream_t stream = nullptr; };
        std::vector<Lane> lanes((size_t)gpus);
        auto cleanup = [&] {
            for (int g = 0; g < gpus; ++g) {
                cudaSetDevice(g);
                if (lanes[g].stream) cudaStreamDestroy(lanes[g].stream);
                if (lanes[g].dev) cudaFree(lanes[g].dev);
                if (lanes[g].host) cudaFreeHost(lanes[g].host);
            }
        };
        try {
            for (int g = 0; g < gpus; ++g) {
                b200::CudaCheck(cudaSetDevice(g), "cudaSetDevice");
                b200::CudaCheck(cudaHostAlloc(&lanes[g].host, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)), "cudaHostAlloc");
                memset(lanes[g].host, 1, bytes);
                b200::CudaCheck(cudaMalloc(&lanes[g].dev, bytes), "cudaMalloc");
                b200::CudaCheck(cudaStreamCreateWithFlags(&lanes[g].stream, cudaStreamNonBlocking), "cudaStreamCreate");
            }
            auto run = [&](int g) {
                cudaSetDevice(g);
                for (int i = 0; i < iters; ++i) cudaMemcpyAsync(lanes[g].dev, lanes[g].host, bytes, cudaMemcpyHostToDevice, lanes[g].stream);
                b200::CudaCheck(cudaStreamSynchronize(lanes[g].stream), "H2D");
            };
            run(0);  // warm-up
            auto t0 = std::chrono::steady_clock::now();
            run(0);
            double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            *gbs_single = (double)bytes * iters / dt / 1e9;
            std::vector<std::future<void>> f;
            t0 = std::chrono::steady_clock::now();
            for (int g = 1; g < gpus; ++g) f.push_back(std::async(std::launch::async, run, g));
            run(0);
            for (auto& x : f) x.get();
            dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            *gbs_all = (double)bytes * iters * gpus / dt / 1e9;
        } catch (...) {
            cleanup();
            throw;
        }
        cleanup();
        return true;
    } catch (const std::exception& e) {
        SetErr(error, e.what());
        return false;
    }
}

const char* B200EngineVersion(void) { return "b200-engine 0.1 sm_100a"; }

char* B200PlanDescribe(const char* model_dir, const char* precision, int max_batch, ErrorMessage* error) {
    try {
        if (!model_dir) throw std::runtime_error("model_dir is null");
        b200::Precision p = b200::Precision::FP32;
        if (precision && !b200::ParsePrecision(precision, &p)) throw std::runtime_error(std::string("unknown precision '") + precision + "'");
        b200::onnx::Model m = b200::onnx::ParseFile(std::string(model_dir) + "/model.onnx");
        b200::Plan plan = b200::BuildPlan(m, p, max_batch > 0 ? max_batch : 256);
        return strdup(plan.ToJson().c_str());
    } catch (const std::exception& e) {
        SetErr(error, e.what());
        return nullptr;
    }
}

int B200PlanShards(int n, int gpus, int max_batch, int min_shard, int round_robin, int* triples, int capacity) {
    std::vector<inference::ShardPlan> v = inference::PlanShards(n, gpus, max_batch, min_shard, round_robin);
    for (size_t i = 0; i < v.size() && (int)i < capacity; ++i) {
        triples[3 * i] = v[i].replica;
        triples[3 * i + 1] = v[i].off;
        triples[3 * i + 2] = v[i].cnt;
    }
    return (int)v.size();
}

uint64_t B200KernelLaunchCount(void) { return b200::kernels::LaunchCount(); }

bool B200ModelStageInput(ModelHandle handle, const TensorData* input, ErrorMessage* error) {
    std::shared_ptr<inference::Model> keep;
    auto st = PinHandle(handle, &keep, error);
    if (!st) return false;
    try {
        if (!input || !input->data || !input->shape.dims || input->shape.num_dims < 1) throw std::runtime_error("Invalid parameters");
        const b200::Plan& P = *st->plan;
        int idx = -1;
        for (size_t i = 0; i < P.input_names.size(); ++i)
            if (input->name && P.input_names[i] == input->name) idx = (int)i;
        if (idx < 0) throw std::runtime_error(std::string("Unexpected input name: ") + (input->name ? input->name : ""));
        int n = (int)input->shape.dims[0];
        const auto& t = P.tensors[P.inputs[idx]];
        if (n < 1 || n > P.max_batch) throw std::runtime_error("batch exceeds B200_ENGINE_MAX_BATCH");
        if (input->data_size < (size_t)n * t.C * t.H * t.W * 4) throw std::runtime_error("Invalid FLOAT32 data for input");
        for (auto& r : st->replicas) r->StageInput(idx, input->data, n);
        keep->Impl()->staged_batch = n;
        return true;
    } catch (const std::exception& e) {
        SetErr(error, e.what());
        return false;
    }
}

bool B200ModelForwardDevice(ModelHandle handle, int batch, int iters, int l2_flush, float* ms_out, ErrorMessage* error) {
    std::shared_ptr<inference::Model> keep;
    auto st = PinHandle(handle, &keep, error);
    if (!st) return false;
    try {
        if (batch < 1 || batch > st->plan->max_batch || iters < 1 || !ms_out) throw std::runtime_error("Invalid parameters");
        for (int it = 0; it < iters; ++it) {
            std::vector<std::future<float>> futs;
            for (size_t g = 1; g < st->replicas.size(); ++g) {
                b200::Replica* r = st->replicas[g].get();
                futs.push_back(std::async(std::launch::async, [r, batch, l2_flush] { return r->ForwardTimed(batch, l2_flush != 0); }));
            }
            float ms = st->replicas[0]->ForwardTimed(batch, l2_flush != 0);
            for (auto& f : futs) ms = std::max(ms, f.get());
            ms_out[it] = ms;
        }
        return true;
    } catch (const std::exception& e) {
        SetErr(error, e.what());
        return false;
    }
}

bool B200ModelReadOutput(ModelHandle handle, float* out, size_t out_elems, ErrorMessage* error) {
    std::shared_ptr<inference::Model> keep;
    auto st = PinHandle(handle, &keep, error);
    if (!st) return false;
    try {
        st->replicas[0]->ReadOutput(0, out, out_elems * 4);
        return true;
    } catch (const std::exception& e) {
        SetErr(error, e.what());
        return false;
    }
}

char* B200ModelProfileSteps(ModelHandle handle, int batch, int repeats, ErrorMessage* error) {
    std::shared_ptr<inference::Model> keep;
    auto st = PinHandle(handle, &keep, error);
    if (!st) return nullptr;
    try {
        return strdup(st->replicas[0]->ProfileSteps(batch, repeats).c_str());
    } catch (const std::exception& e) {
        SetErr(error, e.what());
        return nullptr;
    }
}

int64_t B200ModelReadValue(ModelHandle handle, const char* value_name, float* out, size_t out_elems, ErrorMessage* error) {
    std::shared_ptr<inference::Model> keep;
    auto st = PinHandle(handle, &keep, error);
    if (!st) return -1;
    try {
        int n = keep->Impl()->staged_batch > 0 ? keep->Impl()->staged_batch : 1;
        int64_t r = st->replicas[0]->ReadValue(value_name ? value_name : "", out, out_elems, n);
        if (r < 0) SetErr(error, "unknown value name or insufficient capacity");
        return r;
    } catch (const std::exception& e) {
        SetErr(error, e.what());
        return -1;
    }
}


bool B200ModelInferTopK(ModelHandle handle, const TensorData* inputs, int num_inputs, int k, int apply_softmax, int32_t* classes,
                        float* scores, ErrorMessage* error) {
    std::shared_ptr<inference::Model> keep = B200ModelFromHandle(handle);
    if (!keep) {
        SetErr(error, "Invalid model handle");
        return false;
    }
    if (!keep->IsLoaded()) {
        SetErr(error, "Model not loaded");
        return false;
    }
    if (!inputs || num_inputs <= 0 || k < 1 || k > b200::Replica::kMaxTopK || !classes || !scores) {
        SetErr(error, "Invalid parameters");
        return false;
    }
    try {
        std::vector<inference::IoDesc> ins((size_t)num_inputs);
        for (int i = 0; i < num_inputs; ++i) {
            const TensorData& t = inputs[i];
            ins[i].name = t.name ? t.name : "";
            ins[i].dtype = (inference::DataType)(int)t.data_type;
            if (t.shape.dims && t.shape.num_dims > 0) ins[i].dims.assign(t.shape.dims, t.shape.dims + t.shape.num_dims);
            ins[i].data = t.data;
            ins[i].bytes = t.data_size;
        }
        std::vector<inference::OutDesc> outs(1);
        outs[0].topk = k;
        outs[0].topk_softmax = apply_softmax != 0;
        outs[0].topk_idx = classes;
        outs[0].topk_val = scores;
        if (!keep->Impl()->InferBorrowed(ins, outs)) {
            SetErr(error, keep->GetLastError());
            return false;
        }
        return true;
    } catch (const std::exception& e) {
        SetErr(error, e.what());
        return false;
    }
}

int B200ModelFaultedReplicas(ModelHandle handle) {
    std::shared_ptr<inference::Model> keep = B200ModelFromHandle(handle);
    return keep ? keep->Impl()->FaultedReplicas() : 0;
}

bool B200ModelCoalesceStats(ModelHandle handle, int64_t* batches, int64_t* requests) {
    std::shared_ptr<inference::Model> keep;
    auto st = PinHandle(handle, &keep, nullptr);
    if (!st || !batches || !requests) return false;
    keep->Impl()->CoalesceStats(batches, requests);
    return true;
}
}  // extern "C"
