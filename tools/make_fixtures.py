"""Materialise ./models (the model repository the engine serves) — fixture recipes, offline.

  models/test_model/1/{model.onnx,config.json}   re-creation of the reference's tiny graph
      (`scripts/create-test-model.py:20-115`: MatMul->Add->Relu->MatMul->Add, numpy seed 42, opset 12);
      weights equal to the reference's committed fixture (checked in tests when /root/reference is
      present).
  models/densenet_onnx/1/{model.onnx,config.json,densenet_labels.txt}
      synthetic DenseNet-121 (tools/make_densenet_onnx.py); config.json as in the reference.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tools import make_densenet_onnx, onnx_lite  # noqa: E402

MODELS = os.path.join(ROOT, "models")


def make_test_model(path: str) -> None:
    np.random.seed(42)
    w1 = np.random.randn(3, 5).astype(np.float32)
    b1 = np.random.randn(5).astype(np.float32)
    w2 = np.random.randn(5, 2).astype(np.float32)
    b2 = np.random.randn(2).astype(np.float32)
    g = onnx_lite.Graph(name="test-model")
    g.initializers = {"weight1": w1, "bias1": b1, "weight2": w2, "bias2": b2}
    g.nodes = [
        onnx_lite.Node("MatMul", ["input", "weight1"], ["matmul1"], name="matmul1"),
        onnx_lite.Node("Add", ["matmul1", "bias1"], ["hidden"], name="add1"),
        onnx_lite.Node("Relu", ["hidden"], ["relu"], name="relu"),
        onnx_lite.Node("MatMul", ["relu", "weight2"], ["matmul2"], name="matmul2"),
        onnx_lite.Node("Add", ["matmul2", "bias2"], ["output"], name="add2"),
    ]
    g.inputs = [onnx_lite.ValueInfo("input", onnx_lite.FLOAT, [1, 3])]
    g.outputs = [onnx_lite.ValueInfo("output", onnx_lite.FLOAT, [1, 2])]
    onnx_lite.save(onnx_lite.Model(g, ir_version=10, opset=12, producer_name="test-model-creator"), path)


def ensure_all(quiet: bool = False) -> None:
    d = os.path.join(MODELS, "test_model", "1")
    os.makedirs(d, exist_ok=True)
    if not os.path.exists(os.path.join(d, "model.onnx")):
        make_test_model(os.path.join(d, "model.onnx"))
    cfg = os.path.join(d, "config.json")
    if not os.path.exists(cfg):
        with open(cfg, "w") as fh:
            json.dump({"name": "test_model", "version": "1",
                       "inputs": [{"name": "input", "shape": [1, 3], "data_type": "FLOAT32"}],
                       "outputs": [{"name": "output", "shape": [1, 2], "data_type": "FLOAT32"}]}, fh, indent=2)
    d = os.path.join(MODELS, "densenet_onnx", "1")
    os.makedirs(d, exist_ok=True)
    make_densenet_onnx.ensure(os.path.join(d, "model.onnx"), quiet=quiet)
    cfg = os.path.join(d, "config.json")
    if not os.path.exists(cfg):
        with open(cfg, "w") as fh:
            json.dump({"name": "densenet_onnx", "platform": "onnxruntime_onnx", "version": "1",
                       "inputs": [{"name": "data_0", "dims": [3, 224, 224], "shape": [1, 3, 224, 224],
                                   "data_type": "FLOAT32"}],
                       "outputs": [{"name": "fc6_1", "dims": [1000], "shape": [1, 1000, 1, 1], "data_type": "FLOAT32",
                                    "label_filename": "densenet_labels.txt"}]}, fh, indent=2)
    labels = os.path.join(d, "densenet_labels.txt")
    if not os.path.exists(labels):
        with open(labels, "w") as fh:
            fh.write("\n".join(f"class_{i:04d}" for i in range(1000)) + "\n")


if __name__ == "__main__":
    ensure_all()
