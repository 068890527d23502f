// onnx_wire.h — dependency-free decoder for the subset of onnx.proto3 the serving path needs.
// Replaces what the reference gets from `Ort::Session` construction
// (reference inference_engine/src/model.cpp:847, graph parse inside ONNX Runtime) and the metadata
// getters of `ExtractModelMetadata` (model.cpp:910-972).
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

namespace b200 {
namespace onnx {

// TensorProto.DataType values we understand.
enum : int { kFloat = 1, kUint8 = 2, kInt8 = 3, kInt32 = 6, kInt64 = 7, kBool = 9, kFloat16 = 10, kDouble = 11 };

struct TensorConst {
    std::string name;
    std::vector<int64_t> dims;
    int dtype = kFloat;
    std::vector<float> f32;    // valid for floating types (converted to fp32)
    std::vector<int64_t> i64;  // valid for integer types
    size_t NumElements() const {
        size_t n = 1;
        for (auto d : dims) n *= (size_t)d;
        return n;
    }
};

struct Attr {
    int type = 0;  // AttributeProto.AttributeType: 1 f, 2 i, 3 s, 4 t, 6 floats, 7 ints
    float f = 0.f;
    int64_t i = 0;
    std::string s;
    std::vector<float> floats;
    std::vector<int64_t> ints;
    TensorConst t;
};

struct Node {
    std::string op_type, name;
    std::vector<std::string> inputs, outputs;
    std::map<std::string, Attr> attrs;
    int64_t GetInt(const std::string& k, int64_t dflt) const;
    float GetFloat(const std::string& k, float dflt) const;
    std::vector<int64_t> GetInts(const std::string& k, std::vector<int64_t> dflt = {}) const;
    std::string GetStr(const std::string& k, const std::string& dflt) const;
};

struct ValueInfo {
    std::string name;
    int elem_type = kFloat;
    std::vector<int64_t> dims;  // -1 for symbolic (dim_param) or unknown
};

struct Graph {
    std::string name;
    std::vector<Node> nodes;
    std::unordered_map<std::string, TensorConst> initializers;
    std::vector<ValueInfo> inputs;   // graph inputs minus initializers (what ORT treats as feeds)
    std::vector<ValueInfo> outputs;
};

struct Model {
    int64_t ir_version = 0;
    int64_t opset = 0;
    std::string producer;
    Graph graph;
};

// Throws std::runtime_error on malformed input.
Model ParseBytes(const uint8_t* data, size_t size);
Model ParseFile(const std::string& path);

}  // namespace onnx
}  // namespace b200
