"""gpu-ai-inference-server_b200 — B200-native (sm_100a) drop-in for the hot path of
Oscar-W-Chen/gpu-ai-inference-server: the batched forward pass behind `Model::Infer`.

The product is the C-ABI shared library `lib/libinference_engine.so` (headers in `/include`).
This Python package is the host-side mirror of the reference's cgo binding
(`inference_engine/binding/inference_binding.go`) over ctypes: same type names, call sequence,
allocation pattern and error behaviour, so tests read like the reference's own client code.

The directory name contains '-', so import it with importlib (see `__graft_entry__.load_package`).
"""
from .binding import (  # noqa: F401
    DataType, DeviceType, ModelType, TensorData, OutputConfig, ModelConfig, ModelMetadata, ModelStats,
    MemoryInfo, InferenceManager, Model, EngineError,
    is_cuda_available, get_device_count, get_device_info, get_memory_info,
    library_path, load_library, plan_describe, plan_shards, kernel_launch_count, engine_version, measure_h2d,
)
