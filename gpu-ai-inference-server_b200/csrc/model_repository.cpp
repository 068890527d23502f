// model_repository.cpp — filesystem model store.  Behavioural contract of reference
// inference_engine/src/model_repository.cpp: layout `<repo>/<model>/<version>/…` (:18-66), versions
// ordered numerically descending with a lexical fallback (:45-53), a version directory counts when it
// holds config.json or a known model file (:189-195), model type from the file name (:161-178).
// Additions: thread-safe rescans, and config.json inputs/outputs are read when present.
#include "model_repository.h"

#include <algorithm>
#include <fstream>
#include <sstream>

#include "json_lite.h"

namespace fs = std::filesystem;

namespace inference {

namespace {
bool VersionNewer(const std::string& a, const std::string& b) {
    char *ea = nullptr, *eb = nullptr;
    long va = strtol(a.c_str(), &ea, 10), vb = strtol(b.c_str(), &eb, 10);
    bool na = ea != a.c_str(), nb = eb != b.c_str();
    if (na && nb && va != vb) return va > vb;
    if (na != nb) return na;  // numeric versions sort before non-numeric names
    return a > b;
}
DataType ParseDataType(const std::string& s) {
    if (s == "FLOAT32" || s == "FP32" || s == "TYPE_FP32") return DataType::FLOAT32;
    if (s == "INT32") return DataType::INT32;
    if (s == "INT64") return DataType::INT64;
    if (s == "UINT8") return DataType::UINT8;
    if (s == "INT8") return DataType::INT8;
    if (s == "BOOL") return DataType::BOOL;
    if (s == "FP16" || s == "FLOAT16") return DataType::FP16;
    if (s == "STRING") return DataType::STRING;
    return DataType::UNKNOWN;
}
}  // namespace

ModelRepository::ModelRepository(const std::string& repository_path) : repository_path_(repository_path) {
    std::error_code ec;
    if (!repository_path_.empty() && !fs::exists(repository_path_, ec)) fs::create_directories(repository_path_, ec);
}

bool ModelRepository::ScanRepository() {
    std::unordered_map<std::string, std::vector<std::string>> found;
    std::error_code ec;
    if (repository_path_.empty() || !fs::exists(repository_path_, ec)) {
        std::lock_guard<std::mutex> lk(mu_);
        model_versions_.clear();
        return false;
    }
    try {
        for (const auto& model_dir : fs::directory_iterator(repository_path_)) {
            if (!model_dir.is_directory()) continue;
            std::vector<std::string> versions;
            for (const auto& vdir : fs::directory_iterator(model_dir.path()))
                if (vdir.is_directory() && HasModelConfig(vdir.path())) versions.push_back(vdir.path().filename().string());
            std::sort(versions.begin(), versions.end(), VersionNewer);
            if (!versions.empty()) found[model_dir.path().filename().string()] = std::move(versions);
        }
    } catch (const std::exception&) {
        return false;
    }
    std::lock_guard<std::mutex> lk(mu_);
    model_versions_.swap(found);
    return true;
}

std::vector<std::string> ModelRepository::GetAvailableModels() const {
    std::lock_guard<std::mutex> lk(mu_);
    std::vector<std::string> names;
    names.reserve(model_versions_.size());
    for (const auto& kv : model_versions_) names.push_back(kv.first);
    std::sort(names.begin(), names.end());
    return names;
}

bool ModelRepository::ModelExists(const std::string& model_name, const std::string& version) const {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = model_versions_.find(model_name);
    if (it == model_versions_.end() || it->second.empty()) return false;
    return version.empty() || std::find(it->second.begin(), it->second.end(), version) != it->second.end();
}

std::string ModelRepository::GetModelPath(const std::string& model_name, const std::string& version) const {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = model_versions_.find(model_name);
    if (it == model_versions_.end() || it->second.empty()) return "";
    std::string v = version.empty() ? it->second.front() : version;
    if (std::find(it->second.begin(), it->second.end(), v) == it->second.end()) return "";
    return (fs::path(repository_path_) / model_name / v).string();
}

std::string ModelRepository::GetLatestVersion(const std::string& model_name) const {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = model_versions_.find(model_name);
    return (it == model_versions_.end() || it->second.empty()) ? std::string() : it->second.front();
}

std::vector<std::string> ModelRepository::GetModelVersions(const std::string& model_name) const {
    std::lock_guard<std::mutex> lk(mu_);
    auto it = model_versions_.find(model_name);
    return it == model_versions_.end() ? std::vector<std::string>() : it->second;
}

ModelConfig ModelRepository::GetModelConfig(const std::string& model_name, const std::string& version) const {
    ModelConfig config;
    std::string model_path = GetModelPath(model_name, version);
    if (model_path.empty()) return config;
    config.name = model_name;
    config.version = version.empty() ? GetLatestVersion(model_name) : version;
    config.type = DetectModelType(model_path);
    // Reference default (model_repository.cpp:143-144); Model::Load replaces these with the graph's
    // real names when they do not occur in the graph.
    config.input_names = {"input"};
    config.output_names = {"output"};
    std::ifstream f(model_path + "/config.json");
    if (!f) return config;
    std::stringstream ss;
    ss << f.rdbuf();
    try {
        b200::json::Value root = b200::json::ParseString(ss.str());
        auto read_io = [&](const char* key, std::vector<std::string>& names, std::unordered_map<std::string, Shape>& shapes,
                           std::unordered_map<std::string, DataType>& types) {
            const b200::json::Value* arr = root.Get(key);
            if (!arr || arr->kind != b200::json::Value::Array || arr->arr.empty()) return;
            std::vector<std::string> got;
            for (const auto& e : arr->arr) {
                const b200::json::Value* nm = e.Get("name");
                if (!nm || nm->kind != b200::json::Value::String) continue;
                got.push_back(nm->str);
                if (const auto* sh = e.Get("shape")) {
                    Shape s;
                    for (const auto& d : sh->arr) s.dims.push_back((int64_t)d.num);
                    if (!s.dims.empty()) {
                        s.dims[0] = -1;  // the leading dimension is the batch: any N >= 1 is accepted
                        shapes[nm->str] = s;
                    }
                }
                if (const auto* dt = e.Get("data_type"))
                    if (dt->kind == b200::json::Value::String) types[nm->str] = ParseDataType(dt->str);
            }
            if (!got.empty()) names = got;
        };
        read_io("inputs", config.input_names, config.input_shapes, config.input_types);
        read_io("outputs", config.output_names, config.output_shapes, config.output_types);
        if (const auto* mb = root.Get("max_batch_size")) config.max_batch_size = (int)mb->num;
        if (const auto* db = root.Get("dynamic_batching")) config.dynamic_batching = db->b;
        if (const auto* ic = root.Get("instance_count")) config.instance_count = (int)ic->num;
    } catch (const std::exception&) {
        // unreadable config.json: keep defaults, exactly like the reference which never parsed it
    }
    return config;
}

ModelType ModelRepository::DetectModelType(const fs::path& model_path) const {
    std::error_code ec;
    if (fs::exists(model_path / "model.onnx", ec)) return ModelType::ONNX;
    if (fs::exists(model_path / "saved_model.pb", ec)) return ModelType::TENSORFLOW;
    if (fs::exists(model_path / "model.plan", ec)) return ModelType::TENSORRT;
    if (fs::exists(model_path / "model.pt", ec)) return ModelType::PYTORCH;
    return ModelType::UNKNOWN;
}

bool ModelRepository::HasModelConfig(const fs::path& p) const {
    std::error_code ec;
    for (const char* f : {"config.json", "model.onnx", "model.pt", "saved_model.pb", "model.plan"})
        if (fs::exists(p / f, ec)) return true;
    return false;
}

}  // namespace inference
