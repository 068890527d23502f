// model_repository.h — on-disk model store `<repo>/<model>/<version>/model.onnx`.
// API of reference `inference_engine/include/model_repository.h:19-96`; behaviour of
// `inference_engine/src/model_repository.cpp:10-187` (numeric-descending version order, type
// detection by file name).  Addition: GetModelConfig really parses `config.json` when present.
#ifndef MODEL_REPOSITORY_H
#define MODEL_REPOSITORY_H

#include <filesystem>
#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "model.h"

namespace inference {

class ModelRepository {
public:
    explicit ModelRepository(const std::string& repository_path);

    bool ScanRepository();
    std::vector<std::string> GetAvailableModels() const;
    bool ModelExists(const std::string& model_name, const std::string& version = "") const;
    std::string GetModelPath(const std::string& model_name, const std::string& version = "") const;
    ModelConfig GetModelConfig(const std::string& model_name, const std::string& version = "") const;
    std::string GetLatestVersion(const std::string& model_name) const;
    std::vector<std::string> GetModelVersions(const std::string& model_name) const;

private:
    std::string repository_path_;
    std::unordered_map<std::string, std::vector<std::string>> model_versions_;
    mutable std::mutex mu_;

    bool HasModelConfig(const std::filesystem::path& model_path) const;
    ModelType DetectModelType(const std::filesystem::path& model_path) const;
};

}  // namespace inference

#endif  // MODEL_REPOSITORY_H
