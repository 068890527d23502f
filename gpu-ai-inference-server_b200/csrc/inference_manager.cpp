// inference_manager.cpp — stateful manager: "name:version"-keyed model table with a small worker
// pool for asynchronous load/unload.  Semantics of reference inference_engine/src/inference_manager.cpp:
// state machine UNAVAILABLE -> LOADING -> LOADED | ERROR (:283-384), unload (:452-515), RunInference
// releases the table lock before Model::Infer (:674-707), status JSON keys (:580-628).
// Difference: the (slow) Model::Load runs OUTSIDE the table lock, so inference on other models and
// status queries are not blocked while a model is being lowered and uploaded to the GPUs.
#include "inference_manager.h"

#include <ctime>
#include <iostream>
#include <sstream>

#include "model_repository.h"

namespace inference {

std::string ModelStateToString(ModelState state) {
    switch (state) {
        case ModelState::UNAVAILABLE: return "UNAVAILABLE";
        case ModelState::UNLOADED: return "UNLOADED";
        case ModelState::LOADING: return "LOADING";
        case ModelState::LOADED: return "LOADED";
        case ModelState::UNLOADING: return "UNLOADING";
        case ModelState::ERROR: return "ERROR";
    }
    return "UNKNOWN";
}

namespace {
std::string JsonEscape(const std::string& s) {
    std::ostringstream os;
    for (unsigned char c : s) {
        switch (c) {
            case '"': os << "\\\""; break;
            case '\\': os << "\\\\"; break;
            case '\n': os << "\\n"; break;
            case '\r': os << "\\r"; break;
            case '\t': os << "\\t"; break;
            default:
                if (c < 32) { char buf[8]; snprintf(buf, sizeof buf, "\\u%04x", c); os << buf; }
                else os << (char)c;
        }
    }
    return os.str();
}
}  // namespace

InferenceManager::InferenceManager(const std::string& model_repository_path, int num_worker_threads)
    : model_repository_path_(model_repository_path), shutdown_flag_(false), num_worker_threads_(num_worker_threads) {
    for (int i = 0; i < num_worker_threads_; ++i) worker_threads_.emplace_back(&InferenceManager::WorkerThreadFunc, this);
}

InferenceManager::~InferenceManager() { Shutdown(); }

bool InferenceManager::Initialize() {
    std::lock_guard<std::mutex> lk(models_mutex_);
    try {
        repository_ = std::make_unique<ModelRepository>(model_repository_path_);
        if (!repository_->ScanRepository()) {
            SetError("Failed to scan model repository");
            return false;
        }
        return true;
    } catch (const std::exception& e) {
        SetError(std::string("Initialization error: ") + e.what());
        return false;
    }
}

void InferenceManager::Shutdown() {
    {
        std::lock_guard<std::mutex> lk(queue_mutex_);
        shutdown_flag_ = true;
    }
    queue_condition_.notify_all();
    for (auto& t : worker_threads_)
        if (t.joinable()) t.join();
    worker_threads_.clear();
    std::unordered_map<std::string, ModelInfo> doomed;
    {
        std::lock_guard<std::mutex> lk(models_mutex_);
        doomed.swap(models_);
        repository_.reset();
    }
    // models are destroyed here, outside the lock
}

void InferenceManager::WorkerThreadFunc() {
    for (;;) {
        AsyncTask task;
        {
            std::unique_lock<std::mutex> lk(queue_mutex_);
            queue_condition_.wait(lk, [this] { return shutdown_flag_ || !task_queue_.empty(); });
            if (task_queue_.empty()) {
                if (shutdown_flag_) return;
                continue;
            }
            task = std::move(task_queue_.front());
            task_queue_.pop();
        }
        bool ok = false;
        std::string err;
        try {
            ok = task.type == AsyncTask::TaskType::LOAD ? LoadModelInternal(task.model_name, task.version, task.model_key)
                                                        : UnloadModelInternal(task.model_name, task.version, task.model_key);
            if (!ok) err = GetLastError();
        } catch (const std::exception& e) {
            err = std::string("Exception during model operation: ") + e.what();
            SetError(err);
        }
        if (task.callback) {
            try {
                task.callback(ok, task.model_key, err);
            } catch (const std::exception& e) {
                std::cerr << "Exception in model operation callback: " << e.what() << std::endl;
            }
        }
    }
}

std::string InferenceManager::MakeModelKey(const std::string& name, const std::string& version) const {
    if (!version.empty()) return name + ":" + version;
    if (repository_) {
        std::string latest = repository_->GetLatestVersion(name);
        if (!latest.empty()) return name + ":" + latest;
    }
    return name;
}

void InferenceManager::SetError(const std::string& error) const {
    std::lock_guard<std::mutex> lk(error_mutex_);
    last_error_ = error;
}
std::string InferenceManager::GetLastError() const {
    std::lock_guard<std::mutex> lk(error_mutex_);
    return last_error_;
}

bool InferenceManager::LoadModel(const std::string& model_name, const std::string& version) {
    std::string key;
    {
        std::lock_guard<std::mutex> lk(models_mutex_);
        if (!repository_) { SetError("Inference manager is not initialized"); return false; }
        repository_->ScanRepository();
        if (!repository_->ModelExists(model_name, version)) {
            SetError("Model not found in repository: " + model_name + (version.empty() ? "" : ":" + version));
            return false;
        }
        key = MakeModelKey(model_name, version);
    }
    return LoadModelInternal(model_name, version, key);
}

bool InferenceManager::LoadModelAsync(const std::string& model_name, const std::string& version, ModelOperationCallback callback) {
    AsyncTask task;
    {
        std::lock_guard<std::mutex> lk(models_mutex_);
        if (!repository_) { SetError("Inference manager is not initialized"); return false; }
        repository_->ScanRepository();
        if (!repository_->ModelExists(model_name, version)) {
            SetError("Model not found in repository: " + model_name + (version.empty() ? "" : ":" + version));
            return false;
        }
        task.model_key = MakeModelKey(model_name, version);
    }
    task.type = AsyncTask::TaskType::LOAD;
    task.model_name = model_name;
    task.version = version;
    task.callback = std::move(callback);
    {
        std::lock_guard<std::mutex> lk(queue_mutex_);
        if (shutdown_flag_) { SetError("Inference manager is shut down"); return false; }
        task_queue_.push(std::move(task));
    }
    queue_condition_.notify_one();
    return true;
}

bool InferenceManager::LoadModelInternal(const std::string& model_name, const std::string& version, const std::string& model_key) {
    std::shared_ptr<Model> model;
    try {
        {
            std::lock_guard<std::mutex> lk(models_mutex_);
            if (!repository_) { SetError("Inference manager is not initialized"); return false; }
            auto it = models_.find(model_key);
            if (it != models_.end()) {
                switch (it->second.state) {
                    case ModelState::LOADED: return true;    // idempotent
                    case ModelState::LOADING: return true;   // someone else is on it
                    case ModelState::UNLOADING:
                        SetError("Model is currently being unloaded: " + model_key);
                        return false;
                    default: break;  // UNLOADED / UNAVAILABLE / ERROR: (re)load
                }
            }
            ModelInfo& info = models_[model_key];
            std::string path = repository_->GetModelPath(model_name, version);
            if (path.empty()) {
                info.state = ModelState::UNAVAILABLE;
                info.error_message = "Model not found in repository";
                info.state_changed_time = std::chrono::system_clock::now();
                SetError("Model not found: " + model_name + (version.empty() ? "" : ":" + version));
                return false;
            }
            ModelConfig cfg = repository_->GetModelConfig(model_name, version);
            info.state = ModelState::LOADING;
            info.error_message.clear();
            info.state_changed_time = std::chrono::system_clock::now();
            model = std::make_shared<Model>(path, cfg.type, cfg, DeviceType::GPU, 0);
        }
        bool ok = model->Load();  // slow part, lock released
        std::lock_guard<std::mutex> lk(models_mutex_);
        ModelInfo& info = models_[model_key];
        info.state_changed_time = std::chrono::system_clock::now();
        if (!ok) {
            info.state = ModelState::ERROR;
            info.error_message = model->GetLastError();
            info.model.reset();
            SetError("Failed to load model: " + model->GetLastError());
            return false;
        }
        info.model = model;
        info.state = ModelState::LOADED;
        return true;
    } catch (const std::exception& e) {
        std::lock_guard<std::mutex> lk(models_mutex_);
        ModelInfo& info = models_[model_key];
        info.state = ModelState::ERROR;
        info.error_message = e.what();
        info.state_changed_time = std::chrono::system_clock::now();
        SetError(std::string("Load model error: ") + e.what());
        return false;
    }
}

bool InferenceManager::UnloadModel(const std::string& model_name, const std::string& version) {
    std::string key;
    {
        std::lock_guard<std::mutex> lk(models_mutex_);
        key = MakeModelKey(model_name, version);
    }
    return UnloadModelInternal(model_name, version, key);
}

bool InferenceManager::UnloadModelAsync(const std::string& model_name, const std::string& version, ModelOperationCallback callback) {
    AsyncTask task;
    task.type = AsyncTask::TaskType::UNLOAD;
    {
        std::lock_guard<std::mutex> lk(models_mutex_);
        task.model_key = MakeModelKey(model_name, version);
    }
    task.model_name = model_name;
    task.version = version;
    task.callback = std::move(callback);
    {
        std::lock_guard<std::mutex> lk(queue_mutex_);
        if (shutdown_flag_) { SetError("Inference manager is shut down"); return false; }
        task_queue_.push(std::move(task));
    }
    queue_condition_.notify_one();
    return true;
}

bool InferenceManager::UnloadModelInternal(const std::string&, const std::string&, const std::string& model_key) {
    std::shared_ptr<Model> victim;
    try {
        {
            std::lock_guard<std::mutex> lk(models_mutex_);
            auto it = models_.find(model_key);
            if (it == models_.end()) return true;  // nothing to do
            switch (it->second.state) {
                case ModelState::UNLOADED: case ModelState::UNAVAILABLE: case ModelState::UNLOADING: return true;
                case ModelState::LOADING:
                    SetError("Model is currently being loaded: " + model_key);
                    return false;
                default: break;
            }
            it->second.state = ModelState::UNLOADING;
            it->second.state_changed_time = std::chrono::system_clock::now();
            victim.swap(it->second.model);
        }
        if (victim) victim->Unload();  // in-flight Infer calls keep their own pin on the replicas
        victim.reset();
        std::lock_guard<std::mutex> lk(models_mutex_);
        models_.erase(model_key);
        return true;
    } catch (const std::exception& e) {
        std::lock_guard<std::mutex> lk(models_mutex_);
        auto it = models_.find(model_key);
        if (it != models_.end()) {
            it->second.state = ModelState::ERROR;
            it->second.error_message = e.what();
        }
        SetError(std::string("Unload model error: ") + e.what());
        return false;
    }
}

bool InferenceManager::IsModelLoaded(const std::string& model_name, const std::string& version) {
    std::lock_guard<std::mutex> lk(models_mutex_);
    auto it = models_.find(MakeModelKey(model_name, version));
    return it != models_.end() && it->second.state == ModelState::LOADED;
}

ModelState InferenceManager::GetModelState(const std::string& model_name, const std::string& version) {
    std::lock_guard<std::mutex> lk(models_mutex_);
    auto it = models_.find(MakeModelKey(model_name, version));
    if (it != models_.end()) return it->second.state;
    if (repository_ && repository_->ModelExists(model_name, version)) return ModelState::UNLOADED;
    return ModelState::UNAVAILABLE;
}

std::string InferenceManager::GetModelStatus(const std::string& model_name, const std::string& version) {
    std::lock_guard<std::mutex> lk(models_mutex_);
    std::ostringstream js;
    js << "{\n";
    auto it = models_.find(MakeModelKey(model_name, version));
    std::string resolved = version.empty() && repository_ ? repository_->GetLatestVersion(model_name) : version;
    if (it != models_.end()) {
        const ModelInfo& info = it->second;
        std::time_t tt = std::chrono::system_clock::to_time_t(info.state_changed_time);
        char tbuf[64] = {0};
        struct tm tmv;
        localtime_r(&tt, &tmv);
        strftime(tbuf, sizeof tbuf, "%a %b %e %H:%M:%S %Y", &tmv);
        js << "  \"name\": \"" << JsonEscape(model_name) << "\",\n";
        js << "  \"version\": \"" << JsonEscape(resolved) << "\",\n";
        js << "  \"state\": \"" << ModelStateToString(info.state) << "\",\n";
        js << "  \"state_changed_time\": \"" << tbuf << "\",\n";
        js << "  \"error_message\": \"" << JsonEscape(info.error_message) << "\"";
        if (info.model && info.state == ModelState::LOADED) {
            Model::Stats st = info.model->GetStats();
            js << ",\n  \"type\": " << static_cast<int>(info.model->GetMetadata().type) << ",\n";
            js << "  \"memory_usage_bytes\": " << st.memory_usage_bytes << ",\n";
            js << "  \"inference_count\": " << st.inference_count;
        }
    } else if (repository_ && repository_->ModelExists(model_name, version)) {
        js << "  \"name\": \"" << JsonEscape(model_name) << "\",\n";
        js << "  \"version\": \"" << JsonEscape(resolved) << "\",\n";
        js << "  \"state\": \"UNLOADED\",\n";
        js << "  \"error_message\": \"\"";
    } else {
        js << "  \"name\": \"" << JsonEscape(model_name) << "\",\n";
        js << "  \"version\": \"" << JsonEscape(version) << "\",\n";
        js << "  \"state\": \"UNAVAILABLE\",\n";
        js << "  \"error_message\": \"Model not found in repository\"";
    }
    js << "\n}";
    return js.str();
}

std::vector<std::string> InferenceManager::ListModels() {
    std::lock_guard<std::mutex> lk(models_mutex_);
    if (!repository_) return {};
    repository_->ScanRepository();
    return repository_->GetAvailableModels();
}

std::shared_ptr<Model> InferenceManager::GetModel(const std::string& model_name, const std::string& version) {
    std::lock_guard<std::mutex> lk(models_mutex_);
    auto it = models_.find(MakeModelKey(model_name, version));
    if (it == models_.end() || it->second.state != ModelState::LOADED) return nullptr;
    return it->second.model;
}

bool InferenceManager::RunInference(const std::string& model_name, const std::string& version,
                                    const std::vector<Tensor>& inputs, std::vector<Tensor>& outputs) {
    std::shared_ptr<Model> model;
    {
        std::lock_guard<std::mutex> lk(models_mutex_);
        std::string key = MakeModelKey(model_name, version);
        auto it = models_.find(key);
        if (it == models_.end()) {
            SetError("Model not found: " + model_name + (version.empty() ? "" : ":" + version));
            return false;
        }
        if (it->second.state != ModelState::LOADED) {
            SetError("Model not in loaded state: " + key + " (current state: " + ModelStateToString(it->second.state) + ")");
            return false;
        }
        model = it->second.model;
    }
    bool ok = model->Infer(inputs, outputs);  // batch is sharded across GPU replicas inside
    if (!ok) SetError(model->GetLastError());
    return ok;
}

}  // namespace inference
