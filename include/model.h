// model.h — C++ API of the engine: Tensor, Model and their plain-data companions.
//
// Source-compatible with reference `inference_engine/include/model.h:10-181` so that callers
// written against the reference (e.g. `test/onnx_test.cpp`) compile unchanged.  What differs is
// everything underneath: Model::Load lowers `<model_dir>/model.onnx` to a static sm_100a launch
// plan and Model::Infer executes it with this library's own CUDA kernels (no ONNX Runtime, no
// cuDNN, no CPU execution path).
#ifndef MODEL_H
#define MODEL_H

#include <cstdint>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

namespace inference {

class ModelImpl;

enum class ModelType { UNKNOWN, TENSORFLOW, TENSORRT, ONNX, PYTORCH, CUSTOM };
enum class DeviceType { CPU, GPU };
enum class DataType { FLOAT32, INT32, INT64, UINT8, INT8, STRING, BOOL, FP16, UNKNOWN };

struct Shape {
    std::vector<int64_t> dims;
    // Product of dims; an empty shape has 0 elements (reference model.h:35-42).
    size_t NumElements() const {
        if (dims.empty()) return 0;
        size_t n = 1;
        for (auto d : dims) n *= d;
        return n;
    }
};

struct ModelConfig {
    std::string name;
    std::string version;
    ModelType type = ModelType::UNKNOWN;
    int max_batch_size = 0;
    std::vector<std::string> input_names;
    std::vector<std::string> output_names;
    std::unordered_map<std::string, Shape> input_shapes;   // -1 = wildcard dimension
    std::unordered_map<std::string, Shape> output_shapes;
    std::unordered_map<std::string, DataType> input_types;
    std::unordered_map<std::string, DataType> output_types;
    int instance_count = 1;
    bool dynamic_batching = false;
    ModelConfig() = default;
};

struct ModelMetadata {
    std::string name;
    std::string version;
    ModelType type;
    std::vector<std::string> inputs;
    std::vector<std::string> outputs;
    std::string description;
    int64_t load_time_ns;
};

// Host-side tensor: name + dtype + shape + contiguous row-major byte buffer
// (reference model.h:93-126, model.cpp:30-436).
class Tensor {
public:
    Tensor();
    Tensor(const Tensor& other);
    Tensor(Tensor&& other) noexcept;
    Tensor& operator=(const Tensor& other);
    Tensor(const std::string& name, DataType dtype, const Shape& shape);
    ~Tensor();

    template <typename T> bool SetData(const std::vector<T>& data);  // copies in
    template <typename T> bool GetData(std::vector<T>& data) const;  // copies out

    const std::string& GetName() const;
    DataType GetDataType() const;
    const Shape& GetShape() const;
    bool Reshape(const Shape& new_shape);
    bool toGPU(int device_id = 0);
    bool toCPU();

    // Engine-side zero-copy accessors (additions; not in the reference).
    const void* RawData() const;
    void* MutableRawData();
    size_t ByteSize() const;

private:
    class TensorImpl;
    std::unique_ptr<TensorImpl> impl_;
};

class Model {
public:
    // `model_path` is the version DIRECTORY holding model.onnx (reference model.cpp:830).
    Model(const std::string& model_path, ModelType type, const ModelConfig& config,
          DeviceType device = DeviceType::GPU, int device_id = 0);
    ~Model();
    Model(const Model&) = delete;
    Model& operator=(const Model&) = delete;
    Model(Model&& other) noexcept;
    Model& operator=(Model&& other) noexcept;

    bool Load();
    bool Infer(const std::vector<Tensor>& inputs, std::vector<Tensor>& outputs);
    ModelMetadata GetMetadata() const;
    bool IsLoaded() const;
    void Unload();
    std::string GetLastError() const;

    struct Stats {
        int64_t inference_count;
        int64_t total_inference_time_ns;
        int64_t last_inference_time_ns;
        size_t memory_usage_bytes;
    };
    Stats GetStats() const;

    // Engine access for the C bridge / extension API (addition).
    ModelImpl* Impl() { return impl_.get(); }

private:
    std::unique_ptr<ModelImpl> impl_;
};

}  // namespace inference

#endif  // MODEL_H
