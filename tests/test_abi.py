"""CPU tests of the drop-in boundary: exported symbols, struct layouts, ownership/error conventions,
repository behaviour and the planner (no CUDA call is made)."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols(header):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", "", txt)
    return re.findall(r"\b([A-Z][A-Za-z0-9]+)\s*\([^;{]*\)\s*;", txt)


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg.load_library()
    declared = set(_declared_symbols("inference_bridge.h")) | set(_declared_symbols("b200_engine.h"))
    assert len(declared) >= 31, declared
    from gais_b200 import binding
    assert set(binding.EXPORTED_SYMBOLS) <= declared and set(binding.EXTENSION_SYMBOLS) <= declared
    for name in declared:
        assert hasattr(lib, name), f"missing export {name}"
    out = subprocess.run(["nm", "-D", "--defined-only", pkg.library_path()], capture_output=True, text=True).stdout
    for name in declared:
        assert re.search(rf"\bT {name}\b", out), name


def test_struct_layouts_match_the_reference_abi(pkg, tmp_path):
    """Compile the public header with the C compiler and compare sizeof/offsetof with the ctypes mirror
    (reference values: SURVEY.md §8b)."""
    src = tmp_path / "layout.c"
    src.write_text(r'''
#include <stdio.h>
#include <stddef.h>
#include "inference_bridge.h"
int main(void) {
  printf("%zu %zu %zu %zu %zu %zu\n", sizeof(Shape), sizeof(TensorData), sizeof(ModelConfig), sizeof(ModelMetadata),
         sizeof(ModelStats), sizeof(CudaMemoryInfo));
  printf("%zu %zu %zu %zu\n", offsetof(TensorData, data_type), offsetof(TensorData, shape), offsetof(TensorData, data),
         offsetof(TensorData, data_size));
  printf("%zu %zu %zu %zu\n", offsetof(ModelConfig, max_batch_size), offsetof(ModelConfig, input_names),
         offsetof(ModelConfig, instance_count), offsetof(ModelConfig, dynamic_batching));
  printf("%d %d %d %zu\n", (int)DATATYPE_UNKNOWN, (int)DEVICE_GPU, (int)MODEL_ONNX, sizeof(DataType));
  return 0;
}''')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    lines = subprocess.check_output([str(exe)], text=True).split("\n")
    assert lines[0].split() == ["16", "48", "64", "72", "32", "24"]
    assert lines[1].split() == ["8", "16", "32", "40"]
    assert lines[2].split() == ["20", "24", "52", "56"]
    assert lines[3].split() == ["8", "1", "3", "4"]
    from gais_b200 import binding as b
    assert (C.sizeof(b.CShape), C.sizeof(b.CTensorData), C.sizeof(b.CModelConfig), C.sizeof(b.CModelMetadata),
            C.sizeof(b.CModelStats), C.sizeof(b.CCudaMemoryInfo)) == (16, 48, 64, 72, 32, 24)


def test_reference_acceptance_programs_compile_against_our_headers(pkg, tmp_path):
    """`test/onnx_test.cpp` / `test/cuda_test.cpp` of the reference are used UNMODIFIED as acceptance
    scripts: they must compile and link against include/ + the .so."""
    ref = "/root/reference/test"
    if not os.path.isdir(ref):
        pytest.skip("reference mount not present (GPU box)")
    for name in ("onnx_test", "cuda_test"):
        exe = tmp_path / name
        subprocess.check_call(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), os.path.join(ref, name + ".cpp"),
                               "-o", str(exe), "-L", os.path.dirname(pkg.library_path()), "-linference_engine",
                               "-Wl,-rpath," + os.path.dirname(pkg.library_path())])
        assert exe.exists()


def test_error_strings_and_ownership_without_gpu(pkg, repo_dir, tmp_path):
    lib = pkg.load_library()
    mgr = pkg.InferenceManager(repo_dir)
    assert sorted(mgr.list_models()) == ["densenet_onnx", "test_model"]
    assert not mgr.is_model_loaded("test_model")
    with pytest.raises(pkg.EngineError, match="Model path not found: "):
        mgr.load_model("no_such_model")
    with pytest.raises(pkg.EngineError, match="^Model not found$"):
        mgr.unload_model("test_model")
    err = C.c_void_p()
    assert not lib.GetModelHandle(mgr._h, b"test_model", None, C.byref(err))
    assert C.string_at(err.value) == b"Model not found in loaded models"
    lib.FreeErrorMessage(err)
    err = C.c_void_p()
    assert not lib.InferenceLoadModel(None, b"x", None, C.byref(err))
    assert C.string_at(err.value) == b"Invalid handle or model name"
    lib.FreeErrorMessage(err)
    err = C.c_void_p()
    assert not lib.ModelInfer(None, None, 0, None, 0, C.byref(err))
    assert C.string_at(err.value) == b"Invalid model handle"
    lib.FreeErrorMessage(err)
    # error pointer may be NULL; success must not write *error
    assert not lib.InferenceLoadModel(None, b"x", None, None)
    # a version directory without model.onnx
    os.makedirs(tmp_path / "repo" / "broken" / "3")
    (tmp_path / "repo" / "broken" / "3" / "config.json").write_text("{}")
    os.makedirs(tmp_path / "repo" / "broken" / "12")
    (tmp_path / "repo" / "broken" / "12" / "config.json").write_text("{}")
    m2 = pkg.InferenceManager(str(tmp_path / "repo"))
    assert m2.list_models() == ["broken"]
    with pytest.raises(pkg.EngineError, match=r"ONNX file not found at: .*broken/12/model.onnx"):
        m2.load_model("broken")  # numeric-descending version order: 12 beats 3
    m2.shutdown()
    # unloaded model object
    mdl = pkg.Model.create(os.path.join(repo_dir, "test_model", "1"), pkg.ModelConfig(name="test_model", input_names=["input"],
                                                                                   output_names=["output"]))
    assert not mdl.is_loaded()
    import numpy as np
    with pytest.raises(pkg.EngineError, match="Model not loaded"):
        mdl.infer([pkg.TensorData("input", np.zeros((1, 3), np.float32))], [pkg.OutputConfig("output", [1, 2])])
    md = mdl.get_metadata()
    assert md.name == "test_model" and md.inputs == ["input"] and md.outputs == ["output"] and md.type == pkg.ModelType.ONNX
    st = mdl.get_stats()
    assert st.inference_count == 0 and st.memory_usage_bytes == 0
    mdl.destroy()
    mgr.shutdown()
    assert pkg.get_device_info(99) == "Unknown device"


def test_load_without_gpu_fails_loudly(pkg, repo_dir):
    if pkg.is_cuda_available():
        pytest.skip("a GPU is present")
    mgr = pkg.InferenceManager(repo_dir)
    with pytest.raises(pkg.EngineError, match="no CPU execution path"):
        mgr.load_model("test_model")
    assert not mgr.is_model_loaded("test_model")
    mgr.shutdown()


@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp8"])
def test_planner_lowers_densenet_to_a_fused_static_plan(pkg, repo_dir, precision):
    d = pkg.plan_describe(os.path.join(repo_dir, "densenet_onnx", "1"), precision, 256)
    kinds = [s["kind"] for s in d["steps"]]
    assert kinds.count("conv") == 121                       # 120 Conv + the Gemm classifier
    assert kinds.count("bn_relu") == 0 and kinds.count("relu") == 0   # all 62 BN / 121 ReLU folded away
    assert kinds.count("copy_channels") == 0 and d["inplace_concats"] == 58 and d["copied_concats"] == 0
    assert kinds.count("global_avgpool") == 1 and kinds.count("maxpool") == 1
    assert kinds.count("avgpool") == 0                                  # transitions pool in front of the conv
    assert abs(d["flops_per_sample"] - 5.668e9) / 5.668e9 < 1e-3
    convs = [s for s in d["steps"] if s["kind"] == "conv"]
    # every mode runs on tensor cores (fp32 with bf16-split operands): the 7x7 stem reads the caller's fp32 NCHW batch
    # directly (no layout pass, no bf16 copy)
    assert kinds.count("nchw_to_nhwc") == 0
    assert convs[0]["stem_nchw"] and convs[0]["R"] == 7
    assert convs[0]["in"]["dtype"] == "f32" and convs[0]["in"]["C"] == 3
    assert sum(1 for s in convs if s["pre_bn"]) == 58 + 3                  # dense layers + transitions
    assert sum(1 for s in convs if s["pool2_fused"]) == 3
    # dense layers write their 32 channels straight into the block buffer slice
    c33 = [s for s in convs if s["R"] == 3]
    assert len(c33) == 58 and all(s["out"]["pitch"] > s["out"]["C"] == 32 for s in c33)
    assert d["inputs"] == [{"name": "data_0", "dims": [-1, 3, 224, 224]}]
    assert d["outputs"] == [{"name": "fc6_1", "dims": [-1, 1000]}]
    esz = {"fp32": 4, "bf16": 2, "fp8": 1}[precision]
    assert d["arena_bytes"] < 256 * 2.2e6 * esz + 3e8


def test_planner_test_model_and_errors(pkg, repo_dir, tmp_path):
    d = pkg.plan_describe(os.path.join(repo_dir, "test_model", "1"), "fp32", 4)
    assert [s["kind"] for s in d["steps"]] == ["conv", "conv"]      # MatMul+Add+Relu, MatMul+Add
    assert d["steps"][0]["post_relu"] and d["steps"][0]["bias"] and not d["steps"][1]["post_relu"]
    with pytest.raises(pkg.EngineError):
        pkg.plan_describe(str(tmp_path), "fp32", 1)                  # no model.onnx
    with pytest.raises(pkg.EngineError, match="unknown precision"):
        pkg.plan_describe(os.path.join(repo_dir, "test_model", "1"), "int4", 1)
    # unsupported operator is reported by name
    from tools import onnx_lite
    import numpy as np
    g = onnx_lite.Graph()
    g.nodes = [onnx_lite.Node("Erf", ["x"], ["y"])]
    g.inputs = [onnx_lite.ValueInfo("x", onnx_lite.FLOAT, [1, 4])]
    g.outputs = [onnx_lite.ValueInfo("y", onnx_lite.FLOAT, [1, 4])]
    os.makedirs(tmp_path / "m")
    onnx_lite.save(onnx_lite.Model(g), str(tmp_path / "m" / "model.onnx"))
    with pytest.raises(pkg.EngineError, match="unsupported operator 'Erf'"):
        pkg.plan_describe(str(tmp_path / "m"), "fp32", 1)


def test_native_replay_tool_links_against_the_c_abi_only(tmp_path):
    """tools/rest_replay.cpp is the stand-in for the Go REST handler (threads calling ModelInfer): it must compile against
    include/ alone and link against nothing but libinference_engine.so; without a GPU it fails loudly at load, like the server."""
    import subprocess
    import build_engine
    exe = build_engine.build_tools(quiet=True)
    r = subprocess.run([exe, "--repo", str(tmp_path), "--requests", "1", "--threads", "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0
    assert "load failed" in r.stderr or "ModelInfer" in r.stderr or r.returncode == 1


def test_importer_rejects_malformed_models_with_a_message(pkg, tmp_path):
    """Unusual or malformed files must fail Load() with a message instead of undefined behaviour: a graph
    input without a shape, an initializer without a payload, wrong operand counts; DOUBLE tensors in `double_data` are read."""
    import struct
    import numpy as np
    from tools import onnx_lite as ol

    def write(name, graph_bytes):
        d = tmp_path / name
        os.makedirs(d, exist_ok=True)
        blob = ol._w_int(1, 7) + ol._w_bytes(7, graph_bytes) + ol._w_bytes(8, ol._w_str(1, "") + ol._w_int(2, 12))
        with open(d / "model.onnx", "wb") as fh:
            fh.write(blob)
        return str(d)

    vi = lambda name, shape: ol._ser_value_info(ol.ValueInfo(name, ol.FLOAT, shape))  # noqa: E731
    matmul = ol._w_bytes(1, ol._ser_node(ol.Node("MatMul", ["x", "w"], ["y"])))
    w_ok = ol._w_bytes(5, ol._ser_tensor("w", np.ones((3, 2), np.float32)))
    # (1) graph input with no shape at all
    no_shape = ol._w_str(1, "x") + ol._w_bytes(2, ol._w_bytes(1, ol._w_int(1, ol.FLOAT)))
    with pytest.raises(pkg.EngineError, match="has no shape"):
        pkg.plan_describe(write("noshape", matmul + w_ok + ol._w_bytes(11, no_shape) + ol._w_bytes(12, vi("y", ["N", 2]))), "fp32", 1)
    # (2) FLOAT initializer that announces 3x2 elements and carries none
    w_empty = ol._w_bytes(5, ol._w_int(1, 3) + ol._w_int(1, 2) + ol._w_int(2, ol.FLOAT) + ol._w_str(8, "w"))
    with pytest.raises(pkg.EngineError, match="element count mismatch for tensor w"):
        pkg.plan_describe(write("empty", matmul + w_empty + ol._w_bytes(11, vi("x", ["N", 3])) + ol._w_bytes(12, vi("y", ["N", 2]))), "fp32", 1)
    # (3) negative dimension
    w_neg = ol._w_bytes(5, ol._w_int(1, 3) + ol._w_key(1, 0) + ol._w_varint((1 << 64) - 2) + ol._w_int(2, ol.FLOAT) + ol._w_str(8, "w"))
    with pytest.raises(pkg.EngineError, match="negative dimension"):
        pkg.plan_describe(write("neg", matmul + w_neg + ol._w_bytes(11, vi("x", ["N", 3])) + ol._w_bytes(12, vi("y", ["N", 2]))), "fp32", 1)
    # (4) BatchNormalization with two operands
    bn = ol._w_bytes(1, ol._ser_node(ol.Node("BatchNormalization", ["x", "w"], ["y"])))
    with pytest.raises(pkg.EngineError, match="BatchNormalization node .* has 2 inputs"):
        pkg.plan_describe(write("bn2", bn + w_ok + ol._w_bytes(11, vi("x", ["N", 3, 4, 4])) + ol._w_bytes(12, vi("y", ["N", 3, 4, 4]))), "fp32", 1)
    # (5) DOUBLE weights stored in double_data (field 10, packed) are converted, not dropped
    dbl = b"".join(struct.pack("<d", float(v)) for v in range(6))
    w_dbl = ol._w_bytes(5, ol._w_int(1, 3) + ol._w_int(1, 2) + ol._w_int(2, 11) + ol._w_str(8, "w") + ol._w_bytes(10, dbl))
    d = pkg.plan_describe(write("dbl", matmul + w_dbl + ol._w_bytes(11, vi("x", ["N", 3])) + ol._w_bytes(12, vi("y", ["N", 2]))), "fp32", 1)
    assert [s["kind"] for s in d["steps"]] == ["conv"] and d["steps"][0]["Cin"] == 3 and d["steps"][0]["Cout"] == 2
