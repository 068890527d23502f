"""ctypes mirror of the reference's Go binding `inference_engine/binding/inference_binding.go`.

Go type / func                          -> here
  binding.IsCUDAAvailable (:134)        -> is_cuda_available()
  binding.GetDeviceInfo (:144, C.free)  -> get_device_info()          (frees with libc free, like Go)
  binding.NewInferenceManager (:177)    -> InferenceManager(path)
  (*InferenceManager).LoadModel (:227)  -> InferenceManager.load_model        (NULL version for "")
  (*InferenceManager).GetModel (:387)   -> InferenceManager.get_model         (non-owning handle)
  (*InferenceManager).RunInference(:433)-> InferenceManager.run_inference
  (*Model).Infer (:521-734)             -> Model.infer: outputs pre-allocated from the OutputConfig
                                           shapes, inputs/outputs handed over as C arrays, returned
                                           shape ignored, data_size bytes copied back — same as Go.
There is NO fallback: if the shared library is missing or has no CUDA device, calls fail loudly.
"""
from __future__ import annotations

import ctypes as C
import enum
import json
import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class EngineError(RuntimeError):
    pass


class DataType(enum.IntEnum):
    FLOAT32 = 0
    INT32 = 1
    INT64 = 2
    UINT8 = 3
    INT8 = 4
    STRING = 5
    BOOL = 6
    FP16 = 7
    UNKNOWN = 8


class DeviceType(enum.IntEnum):
    CPU = 0
    GPU = 1


class ModelType(enum.IntEnum):
    UNKNOWN = 0
    TENSORFLOW = 1
    TENSORRT = 2
    ONNX = 3
    PYTORCH = 4
    CUSTOM = 5


_NP_OF = {DataType.FLOAT32: np.float32, DataType.INT32: np.int32, DataType.INT64: np.int64,
          DataType.UINT8: np.uint8, DataType.INT8: np.int8, DataType.BOOL: np.bool_, DataType.FP16: np.float16}


# ---- C structs (include/inference_bridge.h) ----
class CShape(C.Structure):
    _fields_ = [("dims", C.POINTER(C.c_int64)), ("num_dims", C.c_int)]


class CTensorData(C.Structure):
    _fields_ = [("name", C.c_char_p), ("data_type", C.c_int), ("shape", CShape), ("data", C.c_void_p),
                ("data_size", C.c_size_t)]


class CModelConfig(C.Structure):
    _fields_ = [("name", C.c_char_p), ("version", C.c_char_p), ("type_", C.c_int), ("max_batch_size", C.c_int),
                ("input_names", C.POINTER(C.c_char_p)), ("num_inputs", C.c_int),
                ("output_names", C.POINTER(C.c_char_p)), ("num_outputs", C.c_int),
                ("instance_count", C.c_int), ("dynamic_batching", C.c_bool)]


class CModelMetadata(C.Structure):
    _fields_ = [("name", C.c_char_p), ("version", C.c_char_p), ("model_type", C.c_int),
                ("inputs", C.POINTER(C.c_char_p)), ("num_inputs", C.c_int),
                ("outputs", C.POINTER(C.c_char_p)), ("num_outputs", C.c_int),
                ("description", C.c_char_p), ("load_time_ns", C.c_int64)]


class CModelStats(C.Structure):
    _fields_ = [("inference_count", C.c_int64), ("total_inference_time_ns", C.c_int64),
                ("last_inference_time_ns", C.c_int64), ("memory_usage_bytes", C.c_size_t)]


class CCudaMemoryInfo(C.Structure):
    _fields_ = [("total", C.c_size_t), ("free", C.c_size_t), ("used", C.c_size_t)]


# ---- library loading ----
_LIB: Optional[C.CDLL] = None
_LIBC = C.CDLL(None)
_LIBC.free.argtypes = [C.c_void_p]

EXPORTED_SYMBOLS = [
    "IsCudaAvailable", "GetDeviceCount", "GetDeviceInfo", "GetMemoryInfo", "InferenceInitialize", "InferenceShutdown",
    "InferenceLoadModel", "InferenceUnloadModel", "InferenceIsModelLoaded", "InferenceListModels", "InferenceFreeModelList",
    "ModelCreate", "ModelDestroy", "ModelIsLoaded", "ModelInfer", "ModelGetMetadata", "ModelFreeMetadata", "ModelGetStats",
    "ModelFreeStats", "ModelLoad", "ModelUnload", "FreeErrorMessage", "GetModelHandle",
]
EXTENSION_SYMBOLS = [
    "B200EngineVersion", "B200PlanDescribe", "B200PlanShards", "B200KernelLaunchCount", "B200ModelStageInput", "B200ModelForwardDevice",
    "B200ModelReadOutput", "B200ModelProfileSteps", "B200ModelReadValue", "B200ModelCoalesceStats", "B200HostAlloc", "B200HostFree", "B200ModelInferTopK", "B200ModelFaultedReplicas", "B200MeasureH2D",
]


def measure_h2d(gpus: int, mbytes: int = 256, iters: int = 8, write_combined: bool = False):
    """(GB/s to GPU 0 alone, aggregate GB/s to `gpus` GPUs at once) from page-locked host memory (B200MeasureH2D)."""
    one, allg, e = C.c_double(0), C.c_double(0), C.c_void_p()
    if not load_library().B200MeasureH2D(int(gpus), int(mbytes) << 20, int(iters), 1 if write_combined else 0, C.byref(one), C.byref(allg), C.byref(e)):
        raise EngineError(_take_error(e, "H2D measurement failed"))
    return float(one.value), float(allg.value)


def library_path() -> str:
    return os.environ.get("B200_ENGINE_LIB", os.path.join(_HERE, "lib", "libinference_engine.so"))


def load_library() -> C.CDLL:
    global _LIB
    if _LIB is not None:
        return _LIB
    path = library_path()
    if not os.path.exists(path):
        raise EngineError(f"{path} is missing: build it with `python build_engine.py` (there is no Python/CPU fallback)")
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    vp, cp, i, b, sz = C.c_void_p, C.c_char_p, C.c_int, C.c_bool, C.c_size_t
    err = C.POINTER(C.c_void_p)
    sig = {
        "IsCudaAvailable": (b, []), "GetDeviceCount": (i, []), "GetDeviceInfo": (vp, [i]),
        "GetMemoryInfo": (CCudaMemoryInfo, [i]),
        "InferenceInitialize": (vp, [cp]), "InferenceShutdown": (None, [vp]),
        "InferenceLoadModel": (b, [vp, cp, cp, err]), "InferenceUnloadModel": (b, [vp, cp, cp, err]),
        "InferenceIsModelLoaded": (b, [vp, cp, cp]),
        "InferenceListModels": (C.POINTER(C.c_void_p), [vp, C.POINTER(i)]),
        "InferenceFreeModelList": (None, [C.POINTER(C.c_void_p), i]),
        "ModelCreate": (vp, [cp, i, C.POINTER(CModelConfig), i, i, err]), "ModelDestroy": (None, [vp]),
        "ModelIsLoaded": (b, [vp]),
        "ModelInfer": (b, [vp, C.POINTER(CTensorData), i, C.POINTER(CTensorData), i, err]),
        "ModelGetMetadata": (C.POINTER(CModelMetadata), [vp]), "ModelFreeMetadata": (None, [C.POINTER(CModelMetadata)]),
        "ModelGetStats": (C.POINTER(CModelStats), [vp]), "ModelFreeStats": (None, [C.POINTER(CModelStats)]),
        "ModelLoad": (b, [vp, err]), "ModelUnload": (b, [vp, err]),
        "FreeErrorMessage": (None, [vp]), "GetModelHandle": (vp, [vp, cp, cp, err]),
        "B200EngineVersion": (cp, []), "B200PlanDescribe": (vp, [cp, cp, i, err]),
        "B200KernelLaunchCount": (C.c_uint64, []),
        "B200PlanShards": (i, [i, i, i, i, i, C.POINTER(i), i]),
        "B200ModelStageInput": (b, [vp, C.POINTER(CTensorData), err]),
        "B200ModelForwardDevice": (b, [vp, i, i, i, C.POINTER(C.c_float), err]),
        "B200ModelReadOutput": (b, [vp, C.POINTER(C.c_float), sz, err]),
        "B200ModelProfileSteps": (vp, [vp, i, i, err]),
        "B200ModelReadValue": (C.c_int64, [vp, cp, C.POINTER(C.c_float), sz, err]),
        "B200ModelCoalesceStats": (b, [vp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
        "B200HostAlloc": (vp, [sz]), "B200HostFree": (None, [vp]),
        "B200ModelFaultedReplicas": (i, [vp]),
        "B200MeasureH2D": (b, [i, sz, i, i, C.POINTER(C.c_double), C.POINTER(C.c_double), err]),
        "B200ModelInferTopK": (b, [vp, C.POINTER(CTensorData), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_float), C.POINTER(vp)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here == missing export
        fn.restype = res
        fn.argtypes = args
    _LIB = lib
    return lib


def _take_error(err: C.c_void_p, default: str) -> str:
    if err.value:
        msg = C.string_at(err.value).decode(errors="replace")
        load_library().FreeErrorMessage(err)
        return msg
    return default


def _take_string(ptr) -> str:
    """malloc'd char* -> str, freed with libc free (what the Go side does with C.free)."""
    if not ptr:
        return ""
    s = C.string_at(ptr).decode(errors="replace")
    _LIBC.free(ptr)
    return s


# ---- free functions ----
def is_cuda_available() -> bool:
    return bool(load_library().IsCudaAvailable())


def get_device_count() -> int:
    return int(load_library().GetDeviceCount())


def get_device_info(device_id: int = 0) -> str:
    return _take_string(load_library().GetDeviceInfo(device_id))


@dataclass
class MemoryInfo:
    total: int
    free: int
    used: int


def get_memory_info(device_id: int = 0) -> MemoryInfo:
    m = load_library().GetMemoryInfo(device_id)
    return MemoryInfo(int(m.total), int(m.free), int(m.used))


def engine_version() -> str:
    return load_library().B200EngineVersion().decode()


def kernel_launch_count() -> int:
    return int(load_library().B200KernelLaunchCount())


def plan_shards(n: int, gpus: int, max_batch: int = 256, min_shard: int = 8, round_robin: int = 0):
    """(replica, offset, count) triples of the multi-GPU batch scheduler (pure host logic)."""
    cap = 4096
    buf = (C.c_int * (3 * cap))()
    k = load_library().B200PlanShards(n, gpus, max_batch, min_shard, round_robin, buf, cap)
    return [(buf[3 * j], buf[3 * j + 1], buf[3 * j + 2]) for j in range(min(k, cap))]


def plan_describe(model_dir: str, precision: str = "fp32", max_batch: int = 256) -> dict:
    lib = load_library()
    err = C.c_void_p()
    p = lib.B200PlanDescribe(model_dir.encode(), precision.encode(), max_batch, C.byref(err))
    if not p:
        raise EngineError(_take_error(err, "B200PlanDescribe failed"))
    return json.loads(_take_string(p))


# ---- data classes mirroring the Go structs ----
@dataclass
class TensorData:
    name: str
    data: np.ndarray
    data_type: DataType = DataType.FLOAT32
    shape: Optional[Sequence[int]] = None

    def dims(self) -> List[int]:
        return list(self.shape) if self.shape is not None else list(self.data.shape)


@dataclass
class OutputConfig:
    name: str
    shape: Sequence[int]
    data_type: DataType = DataType.FLOAT32


@dataclass
class ModelConfig:
    name: str = ""
    version: str = "1"
    type: ModelType = ModelType.ONNX
    max_batch_size: int = 0
    input_names: List[str] = field(default_factory=list)
    output_names: List[str] = field(default_factory=list)
    instance_count: int = 1
    dynamic_batching: bool = False


@dataclass
class ModelMetadata:
    name: str
    version: str
    type: ModelType
    inputs: List[str]
    outputs: List[str]
    description: str
    load_time_ns: int


@dataclass
class ModelStats:
    inference_count: int
    total_inference_time_ns: int
    last_inference_time_ns: int
    memory_usage_bytes: int


def _c_tensor(name: bytes, dtype: int, dims: Sequence[int], buf: np.ndarray, keep: list) -> CTensorData:
    t = CTensorData()
    t.name = name
    t.data_type = int(dtype)
    arr = (C.c_int64 * max(1, len(dims)))(*[int(d) for d in dims])
    keep.append(arr)
    t.shape.dims = C.cast(arr, C.POINTER(C.c_int64))
    t.shape.num_dims = len(dims)
    t.data = buf.ctypes.data_as(C.c_void_p)
    t.data_size = buf.nbytes
    return t


class Model:
    """Mirror of Go `binding.Model` (inference_binding.go:106-111, :521-797)."""

    def __init__(self, handle: int, owning: bool):
        self._h = C.c_void_p(handle)
        self._owning = owning

    @classmethod
    def create(cls, model_path: str, config: ModelConfig, device: DeviceType = DeviceType.GPU, device_id: int = 0) -> "Model":
        lib = load_library()
        cfg = CModelConfig()
        cfg.name = config.name.encode()
        cfg.version = config.version.encode()
        cfg.type_ = int(config.type)
        cfg.max_batch_size = config.max_batch_size
        ins = (C.c_char_p * max(1, len(config.input_names)))(*[s.encode() for s in config.input_names])
        outs = (C.c_char_p * max(1, len(config.output_names)))(*[s.encode() for s in config.output_names])
        cfg.input_names, cfg.num_inputs = C.cast(ins, C.POINTER(C.c_char_p)), len(config.input_names)
        cfg.output_names, cfg.num_outputs = C.cast(outs, C.POINTER(C.c_char_p)), len(config.output_names)
        cfg.instance_count = config.instance_count
        cfg.dynamic_batching = config.dynamic_batching
        err = C.c_void_p()
        h = lib.ModelCreate(model_path.encode(), int(config.type), C.byref(cfg), int(device), device_id, C.byref(err))
        if not h:
            raise EngineError(_take_error(err, "failed to create model"))
        return cls(h, True)

    def load(self) -> None:
        err = C.c_void_p()
        if not load_library().ModelLoad(self._h, C.byref(err)):
            raise EngineError(_take_error(err, "failed to load model"))

    def unload(self) -> None:
        err = C.c_void_p()
        if not load_library().ModelUnload(self._h, C.byref(err)):
            raise EngineError(_take_error(err, "failed to unload model"))

    def is_loaded(self) -> bool:
        return bool(self._h) and bool(load_library().ModelIsLoaded(self._h))

    def destroy(self) -> None:
        if self._h:
            load_library().ModelDestroy(self._h)
            self._h = C.c_void_p()

    def infer(self, inputs: Sequence[TensorData], output_configs: Sequence[OutputConfig]) -> List[TensorData]:
        """Same marshalling as Go Model.Infer (:521-734)."""
        if not self._h:
            raise EngineError("model handle is nil")
        lib = load_library()
        keep: list = []
        in_bufs = [np.ascontiguousarray(t.data, dtype=_NP_OF[DataType(t.data_type)]) for t in inputs]
        cin = (CTensorData * len(inputs))(*[
            _c_tensor(t.name.encode(), t.data_type, t.dims(), b, keep) for t, b in zip(inputs, in_bufs)])
        out_bufs = [np.zeros(int(np.prod(oc.shape)), dtype=_NP_OF[DataType(oc.data_type)]) for oc in output_configs]
        cout = (CTensorData * len(output_configs))(*[
            _c_tensor(oc.name.encode(), oc.data_type, list(oc.shape), b, keep) for oc, b in zip(output_configs, out_bufs)])
        err = C.c_void_p()
        ok = lib.ModelInfer(self._h, cin, len(inputs), cout, len(output_configs), C.byref(err))
        if not ok:
            raise EngineError(_take_error(err, "inference failed"))
        # Go ignores the returned shape and hands back the configured one (:717-731)
        self.last_returned_shapes = [[int(cout[i].shape.dims[j]) for j in range(cout[i].shape.num_dims)]
                                     for i in range(len(output_configs))]
        return [TensorData(oc.name, b.reshape(tuple(oc.shape)), DataType(oc.data_type), list(oc.shape))
                for oc, b in zip(output_configs, out_bufs)]

    def get_metadata(self) -> ModelMetadata:
        lib = load_library()
        p = lib.ModelGetMetadata(self._h)
        if not p:
            raise EngineError("failed to get model metadata")
        m = p.contents
        md = ModelMetadata(m.name.decode(), m.version.decode(), ModelType(m.model_type),
                           [m.inputs[i].decode() for i in range(m.num_inputs)],
                           [m.outputs[i].decode() for i in range(m.num_outputs)],
                           (m.description or b"").decode(), int(m.load_time_ns))
        lib.ModelFreeMetadata(p)
        return md

    def get_stats(self) -> ModelStats:
        lib = load_library()
        p = lib.ModelGetStats(self._h)
        if not p:
            raise EngineError("failed to get model stats")
        s = p.contents
        st = ModelStats(int(s.inference_count), int(s.total_inference_time_ns), int(s.last_inference_time_ns),
                        int(s.memory_usage_bytes))
        lib.ModelFreeStats(p)
        return st

    # ---- extension API (include/b200_engine.h) ----
    def infer_topk(self, inputs: Sequence[TensorData], k: int = 5, softmax: bool = True):
        """Forward + softmax + top-k on the GPU (B200ModelInferTopK): returns (classes int32 [N,k], scores float32 [N,k])."""
        if not self._h:
            raise EngineError("model handle is nil")
        keep: list = []
        in_bufs = [np.ascontiguousarray(t.data, dtype=_NP_OF[DataType(t.data_type)]) for t in inputs]
        cin = (CTensorData * len(inputs))(*[
            _c_tensor(t.name.encode(), t.data_type, t.dims(), b, keep) for t, b in zip(inputs, in_bufs)])
        n = int(inputs[0].dims()[0])
        classes = np.zeros((n, k), np.int32)
        scores = np.zeros((n, k), np.float32)
        err = C.c_void_p()
        ok = load_library().B200ModelInferTopK(self._h, cin, len(inputs), int(k), 1 if softmax else 0,
                                               classes.ctypes.data_as(C.POINTER(C.c_int32)),
                                               scores.ctypes.data_as(C.POINTER(C.c_float)), C.byref(err))
        if not ok:
            raise EngineError(_take_error(err, "top-k inference failed"))
        return classes, scores

    def faulted_replicas(self) -> int:
        """GPU replicas dropped from the shard set after a CUDA error."""
        return int(load_library().B200ModelFaultedReplicas(self._h))

    def coalesce_stats(self):
        """(batches executed by the request coalescer, requests they carried)"""
        nb, nr = C.c_int64(0), C.c_int64(0)
        load_library().B200ModelCoalesceStats(self._h, C.byref(nb), C.byref(nr))
        return int(nb.value), int(nr.value)

    def stage_input(self, t: TensorData) -> None:
        keep: list = []
        buf = np.ascontiguousarray(t.data, dtype=np.float32)
        ct = _c_tensor(t.name.encode(), t.data_type, t.dims(), buf, keep)
        err = C.c_void_p()
        if not load_library().B200ModelStageInput(self._h, C.byref(ct), C.byref(err)):
            raise EngineError(_take_error(err, "stage failed"))

    def forward_device(self, batch: int, iters: int = 1, l2_flush: bool = True) -> np.ndarray:
        ms = (C.c_float * iters)()
        err = C.c_void_p()
        if not load_library().B200ModelForwardDevice(self._h, batch, iters, int(l2_flush), ms, C.byref(err)):
            raise EngineError(_take_error(err, "forward failed"))
        return np.array(ms[:], dtype=np.float64)

    def read_output(self, elems: int) -> np.ndarray:
        out = np.empty(elems, dtype=np.float32)
        err = C.c_void_p()
        if not load_library().B200ModelReadOutput(self._h, out.ctypes.data_as(C.POINTER(C.c_float)), elems, C.byref(err)):
            raise EngineError(_take_error(err, "read failed"))
        return out

    def profile_steps(self, batch: int, repeats: int = 3) -> list:
        err = C.c_void_p()
        p = load_library().B200ModelProfileSteps(self._h, batch, repeats, C.byref(err))
        if not p:
            raise EngineError(_take_error(err, "profile failed"))
        return json.loads(_take_string(p))

    def read_value(self, name: str, capacity: int) -> np.ndarray:
        out = np.empty(capacity, dtype=np.float32)
        err = C.c_void_p()
        n = load_library().B200ModelReadValue(self._h, name.encode(), out.ctypes.data_as(C.POINTER(C.c_float)), capacity,
                                              C.byref(err))
        if n < 0:
            raise EngineError(_take_error(err, "read_value failed"))
        return out[:n]


class InferenceManager:
    """Mirror of Go `binding.InferenceManager` (inference_binding.go:98-104, :177-446)."""

    def __init__(self, model_repository_path: str):
        lib = load_library()
        self._h = C.c_void_p(lib.InferenceInitialize(model_repository_path.encode()))
        if not self._h:
            raise EngineError("failed to initialize inference manager")
        self._loaded: Dict[str, Model] = {}  # Go-side map keyed name[:version] (:215-220)

    @staticmethod
    def _key(name: str, version: str) -> str:
        return name if not version else f"{name}:{version}"

    def shutdown(self) -> None:
        if not self._h:
            return
        for m in self._loaded.values():
            m.destroy()
        self._loaded.clear()
        load_library().InferenceShutdown(self._h)
        self._h = C.c_void_p()

    def load_model(self, name: str, version: str = "") -> None:
        lib = load_library()
        err = C.c_void_p()
        v = version.encode() if version else None  # Go passes NULL for "" (:235-239)
        if not lib.InferenceLoadModel(self._h, name.encode(), v, C.byref(err)):
            raise EngineError(_take_error(err, "failed to load model"))
        h = lib.GetModelHandle(self._h, name.encode(), v, C.byref(err))
        if not h:
            raise EngineError(_take_error(err, "model loaded but failed to get handle"))
        self._loaded[self._key(name, version)] = Model(h, False)

    def unload_model(self, name: str, version: str = "") -> None:
        lib = load_library()
        err = C.c_void_p()
        v = version.encode() if version else None
        if not lib.InferenceUnloadModel(self._h, name.encode(), v, C.byref(err)):
            raise EngineError(_take_error(err, "failed to unload model"))
        m = self._loaded.pop(self._key(name, version), None)
        if m is not None:
            m.destroy()  # Go destroys the wrapper AFTER the model is gone (:307-330): must be safe

    def is_model_loaded(self, name: str, version: str = "") -> bool:
        v = version.encode() if version else None
        return bool(load_library().InferenceIsModelLoaded(self._h, name.encode(), v))

    def list_models(self) -> List[str]:
        lib = load_library()
        n = C.c_int(0)
        arr = lib.InferenceListModels(self._h, C.byref(n))
        if not arr or n.value == 0:
            return []
        names = [C.string_at(arr[i]).decode() for i in range(n.value)]
        lib.InferenceFreeModelList(arr, n.value)
        return names

    def get_model(self, name: str, version: str = "") -> Model:
        key = self._key(name, version)
        if key in self._loaded and self.is_model_loaded(name, version):
            return self._loaded[key]
        raise EngineError(f"model {key} is not loaded")

    def run_inference(self, name: str, version: str, inputs: Sequence[TensorData],
                      output_configs: Sequence[OutputConfig]) -> List[TensorData]:
        return self.get_model(name, version).infer(inputs, output_configs)
