// inference_bridge.cpp — the extern "C" boundary (include/inference_bridge.h).  Entry points, error
// strings and ownership rules follow reference inference_engine/src/inference_bridge.cpp (ranges
// cited per function).  Differences, all on the safe side: the loaded-model table is mutex-protected,
// handles returned by GetModelHandle hold a weak reference (no dangling pointer after unload), and
// ModelInfer passes caller buffers straight to the engine (no intermediate std::vector copies).
#include "inference_bridge.h"

#include <cstdlib>
#include <cstring>
#include <filesystem>
#include <iostream>
#include <memory>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "cuda_utils.h"
#include "model.h"
#include "model_impl.h"
#include "model_repository.h"

struct InferenceManager_t {
    std::string model_repository_path;
    std::unique_ptr<inference::ModelRepository> repository;
    std::unordered_map<std::string, std::shared_ptr<inference::Model>> models;  // keyed by NAME (reference :320,389,427)
    std::mutex mu;
};

struct Model_t {
    std::shared_ptr<inference::Model> owned;  // ModelCreate handles
    std::weak_ptr<inference::Model> borrowed; // GetModelHandle handles (non-owning)
    bool owning = false;
    std::shared_ptr<inference::Model> Get() const { return owning ? owned : borrowed.lock(); }
};

struct Tensor_t {
    int unused;
};

namespace {

char* DupString(const std::string& s) { return strdup(s.c_str()); }
void SetErr(ErrorMessage* error, const std::string& msg) {
    if (error) *error = DupString(msg);
}

inference::ModelType ToCpp(ModelType t) {
    switch (t) {
        case MODEL_TENSORFLOW: return inference::ModelType::TENSORFLOW;
        case MODEL_TENSORRT: return inference::ModelType::TENSORRT;
        case MODEL_ONNX: return inference::ModelType::ONNX;
        case MODEL_PYTORCH: return inference::ModelType::PYTORCH;
        case MODEL_CUSTOM: return inference::ModelType::CUSTOM;
        default: return inference::ModelType::UNKNOWN;
    }
}
ModelType ToC(inference::ModelType t) {
    switch (t) {
        case inference::ModelType::TENSORFLOW: return MODEL_TENSORFLOW;
        case inference::ModelType::TENSORRT: return MODEL_TENSORRT;
        case inference::ModelType::ONNX: return MODEL_ONNX;
        case inference::ModelType::PYTORCH: return MODEL_PYTORCH;
        case inference::ModelType::CUSTOM: return MODEL_CUSTOM;
        default: return MODEL_UNKNOWN;
    }
}
inference::DataType ToCpp(DataType t) {
    switch (t) {
        case DATATYPE_FLOAT32: return inference::DataType::FLOAT32;
        case DATATYPE_INT32: return inference::DataType::INT32;
        case DATATYPE_INT64: return inference::DataType::INT64;
        case DATATYPE_UINT8: return inference::DataType::UINT8;
        case DATATYPE_INT8: return inference::DataType::INT8;
        case DATATYPE_STRING: return inference::DataType::STRING;
        case DATATYPE_BOOL: return inference::DataType::BOOL;
        case DATATYPE_FP16: return inference::DataType::FP16;
        default: return inference::DataType::UNKNOWN;
    }
}

}  // namespace

extern "C" {

// ---- device queries (reference :198-228) ----
bool IsCudaAvailable() { return inference::cuda::IsCudaAvailable(); }
int GetDeviceCount() { return inference::cuda::GetDeviceCount(); }
const char* GetDeviceInfo(int device_id) { return DupString(inference::cuda::GetDeviceInfo(device_id)); }
CudaMemoryInfo GetMemoryInfo(int device_id) {
    inference::cuda::MemoryInfo m = inference::cuda::GetMemoryInfo(device_id);
    CudaMemoryInfo out;
    out.total = m.total;
    out.free = m.free;
    out.used = m.used;
    return out;
}

// ---- manager (reference :254-515) ----
InferenceManagerHandle InferenceInitialize(const char* model_repository_path) {
    try {
        auto* mgr = new InferenceManager_t();
        mgr->model_repository_path = model_repository_path ? model_repository_path : "";
        mgr->repository = std::make_unique<inference::ModelRepository>(mgr->model_repository_path);
        mgr->repository->ScanRepository();
        return mgr;
    } catch (const std::exception& e) {
        std::cerr << "Exception in InferenceInitialize: " << e.what() << std::endl;
        return nullptr;
    }
}

void InferenceShutdown(InferenceManagerHandle handle) { delete handle; }

bool InferenceLoadModel(InferenceManagerHandle handle, const char* model_name, const char* version, ErrorMessage* error) {
    if (!handle || !model_name) {
        SetErr(error, "Invalid handle or model name");
        return false;
    }
    try {
        std::shared_ptr<inference::Model> model;
        {
            std::lock_guard<std::mutex> lk(handle->mu);
            handle->repository->ScanRepository();
            std::string resolved = version ? version : handle->repository->GetLatestVersion(model_name);
            std::string model_path = handle->repository->GetModelPath(model_name, resolved);
            if (model_path.empty() || !std::filesystem::exists(model_path)) {
                SetErr(error, "Model path not found: " + model_path);
                return false;
            }
            if (handle->models.count(model_name)) {
                SetErr(error, "Model already loaded");
                return false;
            }
            std::string onnx_file = model_path + "/model.onnx";
            if (!std::filesystem::exists(onnx_file)) {
                SetErr(error, "ONNX file not found at: " + onnx_file);
                return false;
            }
            inference::ModelConfig cfg = handle->repository->GetModelConfig(model_name, resolved);
            if (cfg.type == inference::ModelType::UNKNOWN) {
                SetErr(error, "Unable to determine model type");
                return false;
            }
            model = std::make_shared<inference::Model>(model_path, cfg.type, cfg, inference::DeviceType::GPU, 0);
            handle->models[model_name] = model;  // reserve the name; removed again if Load fails
        }
        if (!model->Load()) {
            std::string msg = model->GetLastError();
            std::lock_guard<std::mutex> lk(handle->mu);
            handle->models.erase(model_name);
            SetErr(error, msg);
            return false;
        }
        return true;
    } catch (const std::exception& e) {
        SetErr(error, e.what());
        return false;
    }
}

bool InferenceUnloadModel(InferenceManagerHandle handle, const char* model_name, const char* /*version*/, ErrorMessage* error) {
    if (!handle || !model_name) {
        SetErr(error, "Invalid handle or model name");
        return false;
    }
    try {
        std::shared_ptr<inference::Model> victim;
        {
            std::lock_guard<std::mutex> lk(handle->mu);
            auto it = handle->models.find(model_name);
            if (it == handle->models.end()) {
                SetErr(error, "Model not found");
                return false;
            }
            victim = std::move(it->second);
            handle->models.erase(it);
        }
        victim->Unload();  // frees GPU memory as soon as in-flight inferences drain
        return true;
    } catch (const std::exception& e) {
        SetErr(error, e.what());
        return false;
    }
}

bool InferenceIsModelLoaded(InferenceManagerHandle handle, const char* model_name, const char* /*version*/) {
    if (!handle || !model_name) return false;
    std::lock_guard<std::mutex> lk(handle->mu);
    auto it = handle->models.find(model_name);
    return it != handle->models.end() && it->second->IsLoaded();
}

char** InferenceListModels(InferenceManagerHandle handle, int* num_models) {
    if (!handle || !num_models) return nullptr;
    try {
        std::vector<std::string> names;
        {
            std::lock_guard<std::mutex> lk(handle->mu);
            if (handle->repository) {
                handle->repository->ScanRepository();  // lists AVAILABLE models (reference :456)
                names = handle->repository->GetAvailableModels();
            } else {
                for (auto& kv : handle->models) names.push_back(kv.first);
            }
        }
        *num_models = (int)names.size();
        if (names.empty()) return nullptr;
        char** out = new char*[names.size()];
        for (size_t i = 0; i < names.size(); ++i) out[i] = DupString(names[i]);
        return out;
    } catch (...) {
        *num_models = 0;
        return nullptr;
    }
}

void InferenceFreeModelList(char** models, int num_models) {
    if (!models) return;
    for (int i = 0; i < num_models; ++i) free(models[i]);
    delete[] models;
}

// ---- model (reference :528-971) ----
ModelHandle ModelCreate(const char* model_path, ModelType type, const ModelConfig* config, DeviceType device, int device_id,
                        ErrorMessage* error) {
    if (!model_path || !config) {
        SetErr(error, "Invalid model path or configuration");
        return nullptr;
    }
    try {
        inference::ModelConfig cfg;
        cfg.name = config->name ? config->name : "";
        cfg.version = config->version ? config->version : "1";
        cfg.type = ToCpp(config->type_);
        cfg.max_batch_size = config->max_batch_size;
        cfg.instance_count = config->instance_count;
        cfg.dynamic_batching = config->dynamic_batching;
        for (int i = 0; i < config->num_inputs; ++i)
            if (config->input_names && config->input_names[i]) cfg.input_names.push_back(config->input_names[i]);
        for (int i = 0; i < config->num_outputs; ++i)
            if (config->output_names && config->output_names[i]) cfg.output_names.push_back(config->output_names[i]);
        auto* h = new Model_t();
        h->owning = true;
        h->owned = std::make_shared<inference::Model>(model_path, ToCpp(type), cfg,
                                                      device == DEVICE_GPU ? inference::DeviceType::GPU : inference::DeviceType::CPU,
                                                      device_id);
        return h;
    } catch (const std::exception& e) {
        SetErr(error, e.what());
        return nullptr;
    }
}

void ModelDestroy(ModelHandle handle) { delete handle; }  // a borrowed wrapper never touches the model

bool ModelLoad(ModelHandle handle, ErrorMessage* error) {
    auto m = handle ? handle->Get() : nullptr;
    if (!m) {
        SetErr(error, "Invalid model handle");
        return false;
    }
    try {
        bool ok = m->Load();
        if (!ok) SetErr(error, m->GetLastError());
        return ok;
    } catch (const std::exception& e) {
        SetErr(error, e.what());
        return false;
    }
}

bool ModelUnload(ModelHandle handle, ErrorMessage* error) {
    auto m = handle ? handle->Get() : nullptr;
    if (!m) {
        SetErr(error, "Invalid model handle");
        return false;
    }
    m->Unload();
    return true;
}

bool ModelIsLoaded(ModelHandle handle) {
    auto m = handle ? handle->Get() : nullptr;
    return m && m->IsLoaded();
}

bool ModelInfer(ModelHandle handle, const TensorData* inputs, int num_inputs, TensorData* outputs, int num_outputs,
                ErrorMessage* error) {
    if (!handle) {
        SetErr(error, "Invalid model handle");
        return false;
    }
    auto m = handle->Get();
    if (!m || !m->IsLoaded()) {  // same precedence as the reference (:694-714): "not loaded" wins over bad arguments
        SetErr(error, "Model not loaded");
        return false;
    }
    if (!inputs || num_inputs <= 0 || !outputs || num_outputs <= 0) {
        SetErr(error, "Invalid parameters");
        return false;
    }
    try {
        std::vector<inference::IoDesc> ins((size_t)num_inputs);
        for (int i = 0; i < num_inputs; ++i) {
            const TensorData& t = inputs[i];
            ins[i].name = t.name ? t.name : "";
            ins[i].dtype = ToCpp(t.data_type);
            if (t.shape.dims && t.shape.num_dims > 0) ins[i].dims.assign(t.shape.dims, t.shape.dims + t.shape.num_dims);
            ins[i].data = t.data;   // borrowed for the duration of the call
            ins[i].bytes = t.data_size;
        }
        std::vector<inference::OutDesc> outs((size_t)num_outputs);
        for (int i = 0; i < num_outputs; ++i) {
            if (outputs[i].data_type == DATATYPE_FLOAT32 && outputs[i].data && outputs[i].data_size > 0) {
                outs[i].data = outputs[i].data;
                outs[i].capacity = outputs[i].data_size;
            }
        }
        bool ok = m->Impl()->InferBorrowed(ins, outs);
        if (!ok) {
            SetErr(error, m->GetLastError());
            return false;
        }
        for (int i = 0; i < num_outputs; ++i) {
            if (outs[i].dims.empty()) continue;  // more output slots than graph outputs: left untouched
            int caller_rank = outputs[i].shape.num_dims;
            int rank = (int)outs[i].dims.size();
            int writable = outputs[i].shape.dims ? std::min(rank, caller_rank > 0 ? caller_rank : 0) : 0;
            for (int j = 0; j < writable; ++j) outputs[i].shape.dims[j] = outs[i].dims[j];
            outputs[i].shape.num_dims = outputs[i].shape.dims ? writable : rank;
        }
        return true;
    } catch (const std::exception& e) {
        SetErr(error, e.what());
        return false;
    }
}

ModelMetadata* ModelGetMetadata(ModelHandle handle) {
    auto m = handle ? handle->Get() : nullptr;
    if (!m) return nullptr;
    try {
        inference::ModelMetadata md = m->GetMetadata();
        auto* out = new ModelMetadata();
        out->name = DupString(md.name);
        out->version = DupString(md.version);
        out->model_type = ToC(md.type);
        out->description = DupString(md.description);
        out->load_time_ns = md.load_time_ns;
        out->num_inputs = (int)md.inputs.size();
        out->inputs = nullptr;
        if (out->num_inputs > 0) {
            out->inputs = new const char*[md.inputs.size()];
            for (size_t i = 0; i < md.inputs.size(); ++i) out->inputs[i] = DupString(md.inputs[i]);
        }
        out->num_outputs = (int)md.outputs.size();
        out->outputs = nullptr;
        if (out->num_outputs > 0) {
            out->outputs = new const char*[md.outputs.size()];
            for (size_t i = 0; i < md.outputs.size(); ++i) out->outputs[i] = DupString(md.outputs[i]);
        }
        return out;
    } catch (...) {
        return nullptr;
    }
}

void ModelFreeMetadata(ModelMetadata* md) {
    if (!md) return;
    free((void*)md->name);
    free((void*)md->version);
    free((void*)md->description);
    if (md->inputs) {
        for (int i = 0; i < md->num_inputs; ++i) free((void*)md->inputs[i]);
        delete[] md->inputs;
    }
    if (md->outputs) {
        for (int i = 0; i < md->num_outputs; ++i) free((void*)md->outputs[i]);
        delete[] md->outputs;
    }
    delete md;
}

ModelStats* ModelGetStats(ModelHandle handle) {
    auto m = handle ? handle->Get() : nullptr;
    if (!m) return nullptr;
    inference::Model::Stats s = m->GetStats();
    auto* out = new ModelStats();
    out->inference_count = s.inference_count;
    out->total_inference_time_ns = s.total_inference_time_ns;
    out->last_inference_time_ns = s.last_inference_time_ns;
    out->memory_usage_bytes = s.memory_usage_bytes;
    return out;
}

void ModelFreeStats(ModelStats* stats) { delete stats; }

// ---- utilities (reference :978-1028) ----
void FreeErrorMessage(ErrorMessage error) { free(error); }

ModelHandle GetModelHandle(InferenceManagerHandle handle, const char* model_name, const char* /*version*/, ErrorMessage* error) {
    if (!handle || !model_name) {
        SetErr(error, "Invalid handle or model name");
        return nullptr;
    }
    std::lock_guard<std::mutex> lk(handle->mu);
    auto it = handle->models.find(model_name);
    if (it == handle->models.end()) {
        SetErr(error, "Model not found in loaded models");
        return nullptr;
    }
    auto* h = new Model_t();
    h->owning = false;
    h->borrowed = it->second;
    return h;
}

}  // extern "C"

// Accessor for the extension API (b200_api.cpp).
std::shared_ptr<inference::Model> B200ModelFromHandle(ModelHandle h) { return h ? h->Get() : nullptr; }
