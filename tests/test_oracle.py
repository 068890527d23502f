"""CPU tests: the oracle against the reference's golden vectors, and the fixture tooling."""
import json
import os

import numpy as np
import pytest

from oracle.onnx_oracle import OnnxOracle, topk_indices
from tools import onnx_lite, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TEST_MODEL = "/root/reference/models/test_model/1/model.onnx"


def _kats():
    with open(os.path.join(ROOT, "tests", "golden", "test_model_kat.json")) as fh:
        return json.load(fh)["vectors"]


def test_oracle_matches_reference_known_answers(repo_dir):
    o = OnnxOracle(os.path.join(repo_dir, "test_model", "1", "model.onnx"))
    for v in _kats():
        x = np.asarray(v["input"], np.float32)
        want = np.asarray(v["output"], np.float32)
        np.testing.assert_allclose(o.run_numpy({"input": x})[0], want, rtol=2e-6, atol=1e-6)
        np.testing.assert_allclose(o.run({"input": x})[0], want, rtol=2e-6, atol=1e-6)
    # the illustrative vector of docs/api.md is NOT what the committed model computes
    assert not np.allclose(o.run_numpy({"input": np.array([[1, 2, 3]], np.float32)})[0], [[4.5, 5.5]], atol=0.5)


@pytest.mark.skipif(not os.path.exists(REF_TEST_MODEL), reason="reference mount not present (GPU box)")
def test_regenerated_test_model_equals_reference_fixture(repo_dir):
    mine = onnx_lite.load(os.path.join(repo_dir, "test_model", "1", "model.onnx"))
    ref = onnx_lite.load(REF_TEST_MODEL)
    assert [n.op_type for n in mine.graph.nodes] == [n.op_type for n in ref.graph.nodes]
    assert [n.inputs for n in mine.graph.nodes] == [n.inputs for n in ref.graph.nodes]
    assert ref.opset == mine.opset == 12
    for k, v in ref.graph.initializers.items():
        assert np.array_equal(v, mine.graph.initializers[k]), k
    # and the oracle gives the known answers on the reference's own file too
    o = OnnxOracle(REF_TEST_MODEL)
    for v in _kats():
        np.testing.assert_allclose(o.run_numpy({"input": np.asarray(v["input"], np.float32)})[0],
                                   np.asarray(v["output"], np.float32), rtol=2e-6, atol=1e-6)


def test_onnx_lite_roundtrip():
    g = onnx_lite.Graph(name="g")
    g.initializers["w"] = np.arange(24, dtype=np.float32).reshape(2, 3, 2, 2)
    g.initializers["idx"] = np.array([1, -2, 3], dtype=np.int64)
    g.nodes.append(onnx_lite.Node("Conv", ["x", "w"], ["y"], {"pads": [1, 1, 1, 1], "strides": [2, 2], "group": 1,
                                                               "epsilon": 0.5, "auto_pad": "NOTSET"}, name="c"))
    g.inputs = [onnx_lite.ValueInfo("x", onnx_lite.FLOAT, ["N", 3, 8, 8])]
    g.outputs = [onnx_lite.ValueInfo("y", onnx_lite.FLOAT, ["N", 2, 4, 4])]
    m2 = onnx_lite.load_bytes(onnx_lite.dump_bytes(onnx_lite.Model(g, ir_version=7, opset=12)))
    assert m2.opset == 12 and m2.ir_version == 7
    n = m2.graph.nodes[0]
    assert n.op_type == "Conv" and n.attrs["pads"] == [1, 1, 1, 1] and n.attrs["strides"] == [2, 2]
    assert n.attrs["epsilon"] == 0.5 and n.attrs["auto_pad"] == "NOTSET"
    assert np.array_equal(m2.graph.initializers["w"], g.initializers["w"])
    assert np.array_equal(m2.graph.initializers["idx"], g.initializers["idx"])
    assert m2.graph.inputs[0].shape == ["N", 3, 8, 8]


def test_synthetic_images_are_deterministic_and_sliceable():
    a = synth.synthetic_images_u8(3, start=5)
    b = synth.synthetic_images_u8(5, start=3)
    assert a.dtype == np.uint8 and a.shape == (3, 224, 224, 3)
    assert np.array_equal(a[0], b[2])
    x = synth.to_model_input(a)
    assert x.shape == (3, 3, 224, 224) and x.dtype == np.float32 and 0.0 <= x.min() and x.max() <= 1.0


def test_densenet_oracle_matches_committed_golden_logits(densenet_path):
    g = np.load(os.path.join(ROOT, "tests", "golden", "densenet_logits.npz"))
    x = synth.to_model_input(synth.synthetic_images_u8(2, start=int(g["start"])))
    y = OnnxOracle(densenet_path).run({"data_0": x})[0]
    # different host CPUs may pick different conv kernels: allow fp32 reassociation noise only
    scale = np.abs(g["logits_fp64"]).max()
    assert np.abs(y - g["logits_fp32"][:2]).max() / scale < 2e-5
    assert np.abs(y - g["logits_fp64"][:2]).max() / scale < 2e-5
    assert np.array_equal(topk_indices(y, 1), topk_indices(g["logits_fp32"][:2], 1))


def test_densenet_oracle_matches_torchvision_eager(densenet_path):
    """Independent check of the interpreter: exported graph + oracle == the eager module it came from."""
    import torch
    from tools import make_densenet_onnx
    m = make_densenet_onnx.build_module()
    x = synth.to_model_input(synth.synthetic_images_u8(2, start=77))
    with torch.no_grad():
        want = m(torch.from_numpy(x)).numpy()
    got = OnnxOracle(densenet_path).run({"data_0": x})[0]
    assert np.abs(got - want).max() / np.abs(want).max() < 2e-5
    assert np.array_equal(got.argmax(1), want.argmax(1))


def test_densenet_graph_operator_census(densenet_path):
    from collections import Counter
    m = onnx_lite.load(densenet_path)
    c = Counter(n.op_type for n in m.graph.nodes)
    assert c["Conv"] == 120 and c["BatchNormalization"] == 62 and c["Relu"] == 121 and c["Concat"] == 62
    assert c["AveragePool"] == 3 and c["MaxPool"] == 1 and c["GlobalAveragePool"] == 1 and c["Gemm"] == 1
    assert [v.name for v in m.graph.inputs] == ["data_0"] and [v.name for v in m.graph.outputs] == ["fc6_1"]


def test_engine_arithmetic_oracle_rounding_primitives():
    """oracle/engine_arith.py: the roundings the e4m3 / bf16 engine is specified to make (known answers)."""
    from oracle import engine_arith as ea
    # e4m3: 3 mantissa bits, ties to even, saturating at 448, subnormals down to 2^-9
    assert ea.e4m3(np.array([0.3, 1.0625, 1.1875, 500.0, -1000.0, 2.0 ** -9, 2.0 ** -11])).tolist() == [0.3125, 1.0, 1.25, 448.0, -448.0, 2.0 ** -9, 0.0]
    assert ea.bf16(np.array([1.00390625, 1.01171875])).tolist() == [1.0, 1.015625]       # ties to even, both directions
    # the fused f16 multiply-add rounds ONCE: 0.1 (f16) * 3 + 2^-12 differs from the two-step result
    x = np.full((1, 1, 1, 1), 3.0)
    fused = ea.prologue(x, np.array([0.1]), np.array([2.0 ** -12]), False)[0, 0, 0, 0]
    s16 = float(np.float16(0.1))
    assert fused == float(np.float16(3.0 * s16 + 2.0 ** -12))
    # per-output-channel weight quantisation: the largest weight of every row maps to 448
    q, sc = ea.quantise_weights(np.array([[[[0.5]], [[-0.25]]], [[[2.0]], [[1.0]]]], dtype=np.float32))
    assert q[:, :, 0, 0].tolist() == [[448.0, -224.0], [448.0, 224.0]] and np.allclose(sc * 448.0, [0.5, 2.0])
    # pooled prologue: (p00 + p01) + (p10 + p11) in f16
    xq = np.arange(4, dtype=np.float64).reshape(1, 1, 2, 2)
    assert ea.pooled_prologue_e4m3(xq, np.array([1.0]), np.array([0.0]), True)[0, 0, 0, 0] == 6.0
    assert ea.e4m3_step(np.array([1.0, 300.0])).tolist() == [0.125, 32.0]


def test_engine_arithmetic_densenet_is_close_to_the_fp32_oracle(densenet_path):
    """The e4m3 restatement of the whole network stays within the format's noise of the fp32 ONNX oracle (same top-1)."""
    from oracle import engine_arith as ea
    from tools import synth
    x = synth.to_model_input(synth.clustered_images_u8(1, start=300))
    q = ea.densenet_e4m3_logits(densenet_path, x)
    f = OnnxOracle(densenet_path).run({"data_0": x})[0]
    assert q.shape == f.shape == (1, 1000)
    assert int(q.argmax()) == int(f.argmax())
    assert np.abs(q - f).max() / np.abs(f).max() < 0.25
