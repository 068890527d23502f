"""CPU oracle for the serving hot path.  TEST INFRASTRUCTURE ONLY.

Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` may import this module.  The product (`libinference_engine.so`)
never calls it and has no CPU fallback.

What it restates
----------------
The reference hands the whole forward pass to ONNX Runtime 1.21.0:
`inference_engine/src/model.cpp:1158-1328` (`ModelImpl::InferONNX`), with the
arithmetic in `Ort::Session::Run` at `model.cpp:1264-1270`.  ONNX Runtime is a
third-party dependency that is NOT vendored in the reference (it is downloaded by
`scripts/setup-onnxruntime.sh:11`, pinned to v1.21.0) and cannot be installed here
(no network, no `onnxruntime` / `onnx` wheels).  This file therefore restates the
published ONNX operator semantics (opset 12; operators listed in SURVEY.md §8 a10)
as an independent graph interpreter:

* `run(..., dtype=torch.float32)`  — "ORT-CPU stand-in": fp32 torch-CPU kernels.
* `run(..., dtype=torch.float64)`  — truth used to budget the fp32 tolerance.
* `run_numpy(...)`                 — pure-numpy path for the MatMul/Add/Relu subset
                                     (the reference's `models/test_model`).

Reference-side glue that is mirrored: graph inputs are looked up BY NAME
(`model.cpp:1181-1190`), outputs are returned in graph order (`model.cpp:1276-1314`).

Pinning status
--------------
* `test_model`: pinned by the reference's own known-answer vector
  (`docs/run_server.ipynb` cell 4 stdout: input [[-0.01349723,-1.0577109,0.82254493]]
  -> [[-0.6017066, 1.8522782]], ORT 1.21.0 CPU) — see tests/test_oracle.py.
* DenseNet-121: **parity unpinned** against ORT itself (the reference holds no golden
  logits and `models/densenet_onnx/1/model.onnx` is missing from the mount).  The
  interpreter is instead cross-checked against torchvision's own eager forward of the
  module the `.onnx` file was exported from (tests/test_oracle.py), which exercises
  every operator kind on the DenseNet path.
"""
from __future__ import annotations

import os
import sys
from typing import Dict, List, Optional, Sequence

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from tools import onnx_lite  # noqa: E402


def _pads(attrs, nd=2):
    p = attrs.get("pads", [0] * (2 * nd))
    return list(p)


def _auto_pad_ok(attrs):
    ap = attrs.get("auto_pad", "NOTSET")
    if ap not in ("NOTSET", ""):
        raise NotImplementedError(f"auto_pad={ap}")


class OnnxOracle:
    """Interpreter over a parsed `onnx_lite.Model`."""

    def __init__(self, model_or_path):
        if isinstance(model_or_path, (str, os.PathLike)):
            self.model = onnx_lite.load(str(model_or_path))
        else:
            self.model = model_or_path
        self.graph = self.model.graph
        self.input_names = [vi.name for vi in self.graph.inputs]
        self.output_names = [vi.name for vi in self.graph.outputs]

    # ------------------------------------------------------------ torch path
    def run(self, feeds: Dict[str, np.ndarray], dtype=None, outputs: Optional[Sequence[str]] = None,
            keep: Optional[Sequence[str]] = None) -> List[np.ndarray]:
        """Execute the graph with torch-CPU kernels.  Returns outputs in graph order
        (reference: model.cpp:1276-1314)."""
        import torch
        import torch.nn.functional as F

        dtype = dtype or torch.float32
        env: Dict[str, "torch.Tensor"] = {}
        for name, arr in self.graph.initializers.items():
            t = torch.from_numpy(np.ascontiguousarray(arr))
            env[name] = t.to(dtype) if t.is_floating_point() else t
        for name in self.input_names:  # by-name lookup, model.cpp:1181-1190
            if name not in feeds:
                raise KeyError(f"Required input tensor not provided: {name}")
            t = torch.from_numpy(np.ascontiguousarray(feeds[name]))
            env[name] = t.to(dtype) if t.is_floating_point() else t

        with torch.no_grad():
            for n in self.graph.nodes:
                a = n.attrs
                x = [env[i] if i else None for i in n.inputs]
                op = n.op_type
                if op == "Conv":
                    _auto_pad_ok(a)
                    nd = x[0].dim() - 2
                    pads = _pads(a, nd)
                    strides = a.get("strides", [1] * nd)
                    dil = a.get("dilations", [1] * nd)
                    grp = a.get("group", 1)
                    inp = x[0]
                    if pads[:nd] != pads[nd:]:
                        # asymmetric: explicit zero pad (ONNX pads = [b0,b1,e0,e1])
                        inp = F.pad(inp, [pads[1], pads[3], pads[0], pads[2]])
                        pads = [0] * (2 * nd)
                    y = F.conv2d(inp, x[1], x[2] if len(x) > 2 else None, stride=strides,
                                 padding=pads[:nd], dilation=dil, groups=grp)
                elif op == "BatchNormalization":
                    eps = a.get("epsilon", 1e-5)
                    sc, b, mean, var = x[1], x[2], x[3], x[4]
                    # ONNX inference form: y = (x - mean) / sqrt(var + eps) * scale + B
                    y = F.batch_norm(x[0], mean, var, sc, b, training=False, eps=eps)
                elif op == "Relu":
                    y = torch.relu(x[0])
                elif op == "Identity":
                    y = x[0]
                elif op == "Concat":
                    y = torch.cat(x, dim=a.get("axis", 1))
                elif op == "MaxPool":
                    _auto_pad_ok(a)
                    pads = _pads(a)
                    k = a["kernel_shape"]
                    assert pads[:2] == pads[2:], "asymmetric MaxPool pads"
                    y = F.max_pool2d(x[0], k, stride=a.get("strides", [1, 1]), padding=pads[:2],
                                     ceil_mode=bool(a.get("ceil_mode", 0)))
                elif op == "AveragePool":
                    _auto_pad_ok(a)
                    pads = _pads(a)
                    k = a["kernel_shape"]
                    assert pads[:2] == pads[2:], "asymmetric AveragePool pads"
                    y = F.avg_pool2d(x[0], k, stride=a.get("strides", [1, 1]), padding=pads[:2],
                                     ceil_mode=bool(a.get("ceil_mode", 0)),
                                     count_include_pad=bool(a.get("count_include_pad", 0)))
                elif op == "GlobalAveragePool":
                    y = x[0].mean(dim=tuple(range(2, x[0].dim())), keepdim=True)
                elif op == "Flatten":
                    ax = a.get("axis", 1)
                    lead = int(np.prod(x[0].shape[:ax])) if ax > 0 else 1
                    y = x[0].reshape(lead, -1)
                elif op == "Reshape":
                    shape = [int(v) for v in x[1].tolist()]
                    shape = [x[0].shape[i] if s == 0 else s for i, s in enumerate(shape)]
                    y = x[0].reshape(shape)
                elif op == "Gemm":
                    A = x[0].t() if a.get("transA", 0) else x[0]
                    B = x[1].t() if a.get("transB", 0) else x[1]
                    y = a.get("alpha", 1.0) * (A @ B)
                    if len(x) > 2 and x[2] is not None:
                        y = y + a.get("beta", 1.0) * x[2]
                elif op == "MatMul":
                    y = x[0] @ x[1]
                elif op == "Add":
                    y = x[0] + x[1]
                elif op == "Mul":
                    y = x[0] * x[1]
                elif op == "Softmax":
                    ax = a.get("axis", 1 if self.model.opset < 13 else -1)
                    if self.model.opset < 13:
                        # opset < 13: coerce to 2D at `axis`, softmax over the flattened tail
                        shp = x[0].shape
                        lead = int(np.prod(shp[:ax])) if ax > 0 else 1
                        y = torch.softmax(x[0].reshape(lead, -1), dim=1).reshape(shp)
                    else:
                        y = torch.softmax(x[0], dim=ax)
                elif op == "Dropout":
                    y = x[0]
                else:
                    raise NotImplementedError(f"oracle: operator {op}")
                env[n.outputs[0]] = y
        names = list(outputs) if outputs is not None else self.output_names
        return [env[o].to(torch.float32 if env[o].is_floating_point() and dtype == torch.float32 else env[o].dtype)
                .numpy() for o in names]

    # ------------------------------------------------------------ numpy path
    def run_numpy(self, feeds: Dict[str, np.ndarray]) -> List[np.ndarray]:
        """Pure numpy fp32 evaluation of the MatMul/Add/Relu/Gemm/Identity subset
        (the reference's test_model: scripts/create-test-model.py:46-81)."""
        env: Dict[str, np.ndarray] = dict(self.graph.initializers)
        for name in self.input_names:
            if name not in feeds:
                raise KeyError(f"Required input tensor not provided: {name}")
            env[name] = np.asarray(feeds[name], dtype=np.float32)
        for n in self.graph.nodes:
            x = [env[i] for i in n.inputs]
            if n.op_type == "MatMul":
                y = (x[0].astype(np.float32) @ x[1].astype(np.float32)).astype(np.float32)
            elif n.op_type == "Add":
                y = (x[0] + x[1]).astype(np.float32)
            elif n.op_type == "Relu":
                y = np.maximum(x[0], np.float32(0))
            elif n.op_type == "Identity":
                y = x[0]
            elif n.op_type == "Gemm":
                A = x[0].T if n.attrs.get("transA", 0) else x[0]
                B = x[1].T if n.attrs.get("transB", 0) else x[1]
                y = np.float32(n.attrs.get("alpha", 1.0)) * (A @ B)
                if len(x) > 2:
                    y = y + np.float32(n.attrs.get("beta", 1.0)) * x[2]
                y = y.astype(np.float32)
            else:
                raise NotImplementedError(f"numpy oracle: operator {n.op_type}")
            env[n.outputs[0]] = y
        return [env[o] for o in self.output_names]


def topk_indices(logits: np.ndarray, k: int = 5) -> np.ndarray:
    """Indices of the k largest logits per row, descending (ties: lower index first),
    the ordering `server/main.go:744-786` (findTopClasses) produces with a stable sort."""
    order = np.argsort(-logits, axis=1, kind="stable")
    return order[:, :k]
