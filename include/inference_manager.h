// inference_manager.h — stateful model manager with async load/unload and the multi-GPU batch
// scheduler entry point.  Public API of reference `inference_engine/include/inference_manager.h`
// (:23-191); semantics of `inference_engine/src/inference_manager.cpp` (state machine :283-384,
// RunInference drops the map lock before Model::Infer :674-707, JSON status :580-628).
#ifndef INFERENCE_MANAGER_H
#define INFERENCE_MANAGER_H

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <future>
#include <memory>
#include <mutex>
#include <queue>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "model.h"

namespace inference {

enum class ModelState { UNAVAILABLE, UNLOADED, LOADING, LOADED, UNLOADING, ERROR };
std::string ModelStateToString(ModelState state);

class ModelRepository;

using ModelOperationCallback =
    std::function<void(bool success, const std::string& model_key, const std::string& error_msg)>;

class InferenceManager {
public:
    InferenceManager(const std::string& model_repository_path, int num_worker_threads = 4);
    ~InferenceManager();

    bool Initialize();
    void Shutdown();

    bool LoadModel(const std::string& model_name, const std::string& version = "");
    bool LoadModelAsync(const std::string& model_name, const std::string& version = "",
                        ModelOperationCallback callback = nullptr);
    bool UnloadModel(const std::string& model_name, const std::string& version = "");
    bool UnloadModelAsync(const std::string& model_name, const std::string& version = "",
                          ModelOperationCallback callback = nullptr);

    bool IsModelLoaded(const std::string& model_name, const std::string& version = "");
    ModelState GetModelState(const std::string& model_name, const std::string& version = "");
    std::string GetModelStatus(const std::string& model_name, const std::string& version = "");
    std::vector<std::string> ListModels();
    std::shared_ptr<Model> GetModel(const std::string& model_name, const std::string& version = "");

    bool RunInference(const std::string& model_name, const std::string& version,
                      const std::vector<Tensor>& inputs, std::vector<Tensor>& outputs);

    std::string GetLastError() const;

private:
    struct ModelInfo {
        std::shared_ptr<Model> model;
        ModelState state = ModelState::UNLOADED;
        std::string error_message;
        std::chrono::time_point<std::chrono::system_clock> state_changed_time =
            std::chrono::system_clock::now();
    };
    struct AsyncTask {
        enum class TaskType { LOAD, UNLOAD };
        TaskType type;
        std::string model_key;
        std::string model_name;
        std::string version;
        ModelOperationCallback callback;
    };

    std::string model_repository_path_;
    std::unique_ptr<ModelRepository> repository_;
    std::unordered_map<std::string, ModelInfo> models_;  // key "name:version"

    std::vector<std::thread> worker_threads_;
    std::queue<AsyncTask> task_queue_;
    std::mutex queue_mutex_;
    std::condition_variable queue_condition_;
    std::atomic<bool> shutdown_flag_;
    int num_worker_threads_;

    mutable std::mutex models_mutex_;
    mutable std::string last_error_;
    mutable std::mutex error_mutex_;

    std::string MakeModelKey(const std::string& name, const std::string& version) const;
    void SetError(const std::string& error) const;
    void WorkerThreadFunc();
    bool LoadModelInternal(const std::string& model_name, const std::string& version,
                           const std::string& model_key);
    bool UnloadModelInternal(const std::string& model_name, const std::string& version,
                             const std::string& model_key);
};

}  // namespace inference

#endif  // INFERENCE_MANAGER_H
