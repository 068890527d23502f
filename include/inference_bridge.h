/*
 * inference_bridge.h — C-ABI of libinference_engine.so (B200-native engine).
 *
 * Drop-in for the reference's cgo boundary: every type and entry point below has the same
 * name, argument order, struct layout (x86-64: Shape 16 B, TensorData 48 B, ModelConfig 64 B,
 * ModelMetadata 72 B, ModelStats 32 B, CudaMemoryInfo 24 B returned by value) and ownership
 * rules as reference `inference_engine/include/inference_bridge.h:12-133`, which the Go
 * binding includes verbatim (`inference_engine/binding/inference_binding.go:5-7`).
 * Behaviour of each call follows reference `inference_engine/src/inference_bridge.cpp`
 * (line ranges cited per function).  Documented deviations are listed in INTEGRATION.md.
 *
 * Ownership summary
 *   - ErrorMessage: on failure, if `error != NULL`, `*error` is set to a malloc'd string the
 *     caller releases with FreeErrorMessage().  On success `*error` is not written.
 *   - GetDeviceInfo(): malloc'd string, caller calls free() (the Go side uses C.free).
 *   - InferenceListModels(): array + strings released with InferenceFreeModelList().
 *   - ModelGetMetadata()/ModelGetStats(): released with ModelFreeMetadata()/ModelFreeStats().
 *   - GetModelHandle(): returns a NON-owning wrapper; ModelDestroy() on it never touches the
 *     model (safe even after InferenceUnloadModel destroyed the model).
 */
#ifndef INFERENCE_BRIDGE_H
#define INFERENCE_BRIDGE_H

#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- opaque handles (reference inference_bridge.h:12-15) ---- */
typedef struct InferenceManager_t* InferenceManagerHandle;
typedef struct Model_t* ModelHandle;
typedef struct Tensor_t* TensorHandle; /* declared by the reference, used by no entry point */

typedef char* ErrorMessage;

/* ---- enums: 4-byte ints, values fixed by reference inference_bridge.h:21-47 ---- */
typedef enum {
    DATATYPE_FLOAT32 = 0,
    DATATYPE_INT32 = 1,
    DATATYPE_INT64 = 2,
    DATATYPE_UINT8 = 3,
    DATATYPE_INT8 = 4,
    DATATYPE_STRING = 5,
    DATATYPE_BOOL = 6,
    DATATYPE_FP16 = 7,
    DATATYPE_UNKNOWN = 8
} DataType;

typedef enum { DEVICE_CPU = 0, DEVICE_GPU = 1 } DeviceType;

typedef enum {
    MODEL_UNKNOWN = 0,
    MODEL_TENSORFLOW = 1,
    MODEL_TENSORRT = 2,
    MODEL_ONNX = 3,
    MODEL_PYTORCH = 4,
    MODEL_CUSTOM = 5
} ModelType;

/* ---- plain structs (reference inference_bridge.h:50-105) ---- */
typedef struct {
    int64_t* dims;
    int num_dims;
} Shape;

typedef struct {
    const char* name;
    DataType data_type;
    Shape shape;
    void* data;       /* host memory, row-major */
    size_t data_size; /* bytes */
} TensorData;

typedef struct {
    const char* name;
    const char* version;
    ModelType type_;
    int max_batch_size;
    const char** input_names;
    int num_inputs;
    const char** output_names;
    int num_outputs;
    int instance_count;
    bool dynamic_batching;
} ModelConfig;

typedef struct {
    const char* name;
    const char* version;
    ModelType model_type;
    const char** inputs;
    int num_inputs;
    const char** outputs;
    int num_outputs;
    const char* description;
    int64_t load_time_ns;
} ModelMetadata;

typedef struct {
    int64_t inference_count;
    int64_t total_inference_time_ns;
    int64_t last_inference_time_ns;
    size_t memory_usage_bytes;
} ModelStats;

typedef struct {
    size_t total;
    size_t free;
    size_t used;
} CudaMemoryInfo;

/* ---- device queries (reference inference_bridge.cpp:198-228, cuda_utils.cu:17-57,129-177) ---- */
bool IsCudaAvailable();
int GetDeviceCount();
const char* GetDeviceInfo(int device_id); /* "Device <id>: <name> (Compute Capability M.m)" */
CudaMemoryInfo GetMemoryInfo(int device_id);

/* ---- repository-level manager (reference inference_bridge.cpp:254-515) ---- */
InferenceManagerHandle InferenceInitialize(const char* model_repository_path);
void InferenceShutdown(InferenceManagerHandle handle);
bool InferenceLoadModel(InferenceManagerHandle handle, const char* model_name, const char* version,
                        ErrorMessage* error);
bool InferenceUnloadModel(InferenceManagerHandle handle, const char* model_name, const char* version,
                          ErrorMessage* error);
bool InferenceIsModelLoaded(InferenceManagerHandle handle, const char* model_name, const char* version);
char** InferenceListModels(InferenceManagerHandle handle, int* num_models);
void InferenceFreeModelList(char** models, int num_models);

/* ---- per-model calls (reference inference_bridge.cpp:528-971) ---- */
ModelHandle ModelCreate(const char* model_path, ModelType type, const ModelConfig* config,
                        DeviceType device, int device_id, ErrorMessage* error);
void ModelDestroy(ModelHandle handle);
bool ModelIsLoaded(ModelHandle handle);
/* THE HOT PATH (reference inference_bridge.cpp:692-828 -> model.cpp:557-613 -> :1158-1328).
 * Inputs are borrowed for the duration of the call.  Outputs are matched by position; the
 * caller pre-allocates `data` (data_size bytes) and `shape.dims`; the callee writes
 * `shape.num_dims`, up to the caller's original `num_dims` entries of `dims`, and
 * min(data_size, produced bytes) bytes of data. */
bool ModelInfer(ModelHandle handle, const TensorData* inputs, int num_inputs, TensorData* outputs,
                int num_outputs, ErrorMessage* error);
ModelMetadata* ModelGetMetadata(ModelHandle handle);
void ModelFreeMetadata(ModelMetadata* metadata);
ModelStats* ModelGetStats(ModelHandle handle);
void ModelFreeStats(ModelStats* stats);

/* Exported by the reference without a declaration (inference_bridge.cpp:603,636). */
bool ModelLoad(ModelHandle handle, ErrorMessage* error);
bool ModelUnload(ModelHandle handle, ErrorMessage* error);

/* ---- utilities (reference inference_bridge.cpp:978-1028) ---- */
void FreeErrorMessage(ErrorMessage error);
ModelHandle GetModelHandle(InferenceManagerHandle handle, const char* model_name, const char* version,
                           ErrorMessage* error);

#ifdef __cplusplus
}
#endif

#endif /* INFERENCE_BRIDGE_H */
