"""Minimal ONNX protobuf reader/writer (no `onnx`, no `protoc` in this image).

Only the subset of onnx.proto3 that the serving path needs is handled: ModelProto,
GraphProto, NodeProto, AttributeProto, TensorProto, ValueInfoProto.  Field numbers
follow onnx.proto3 (ONNX 1.17, IR v10) and were cross-checked against the reference's
committed fixture `models/test_model/1/model.onnx` (produced by the reference's
`scripts/create-test-model.py`).

This module is fixture tooling: it writes the synthetic model files and lets the
Python oracle read them.  The product has its own, independent C++ decoder
(`csrc/onnx_wire.cpp`); the two are checked against each other in tests.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Tuple

import numpy as np

# ---------------------------------------------------------------- wire helpers

def _read_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    result = 0
    shift = 0
    while True:
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not (b & 0x80):
            return result, pos
        shift += 7


def _iter_fields(buf: bytes):
    """Yield (field_number, wire_type, value) for every field of a message."""
    pos, end = 0, len(buf)
    while pos < end:
        key, pos = _read_varint(buf, pos)
        fno, wt = key >> 3, key & 7
        if wt == 0:
            val, pos = _read_varint(buf, pos)
        elif wt == 1:
            val = buf[pos:pos + 8]
            pos += 8
        elif wt == 2:
            ln, pos = _read_varint(buf, pos)
            val = buf[pos:pos + ln]
            pos += ln
        elif wt == 5:
            val = buf[pos:pos + 4]
            pos += 4
        else:
            raise ValueError(f"unsupported wire type {wt}")
        yield fno, wt, val


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(val, wt) -> List[int]:
    if wt == 0:
        return [_signed64(val)]
    out, pos = [], 0
    while pos < len(val):
        v, pos = _read_varint(val, pos)
        out.append(_signed64(v))
    return out


def _w_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _w_key(fno: int, wt: int) -> bytes:
    return _w_varint((fno << 3) | wt)


def _w_bytes(fno: int, b: bytes) -> bytes:
    return _w_key(fno, 2) + _w_varint(len(b)) + b


def _w_str(fno: int, s: str) -> bytes:
    return _w_bytes(fno, s.encode())


def _w_int(fno: int, v: int) -> bytes:
    return _w_key(fno, 0) + _w_varint(v)


# ---------------------------------------------------------------- data model

# TensorProto.DataType
FLOAT, UINT8, INT8, INT32, INT64, BOOL, FLOAT16, DOUBLE = 1, 2, 3, 6, 7, 9, 10, 11
_NP_OF = {FLOAT: np.float32, UINT8: np.uint8, INT8: np.int8, INT32: np.int32,
          INT64: np.int64, BOOL: np.bool_, FLOAT16: np.float16, DOUBLE: np.float64}
_DT_OF = {np.dtype(v): k for k, v in _NP_OF.items()}


@dataclass
class ValueInfo:
    name: str
    elem_type: int = FLOAT
    shape: List[Any] = field(default_factory=list)  # ints or str (dim_param)


@dataclass
class Node:
    op_type: str
    inputs: List[str]
    outputs: List[str]
    attrs: Dict[str, Any] = field(default_factory=dict)
    name: str = ""


@dataclass
class Graph:
    nodes: List[Node] = field(default_factory=list)
    initializers: Dict[str, np.ndarray] = field(default_factory=dict)
    inputs: List[ValueInfo] = field(default_factory=list)   # real inputs only
    outputs: List[ValueInfo] = field(default_factory=list)
    name: str = "graph"


@dataclass
class Model:
    graph: Graph
    ir_version: int = 7
    opset: int = 12
    producer_name: str = "b200-engine-fixtures"


# ---------------------------------------------------------------- reader

def _parse_tensor(buf: bytes) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype = FLOAT
    name = ""
    raw = None
    floats: List[float] = []
    i32: List[int] = []
    i64: List[int] = []
    for fno, wt, val in _iter_fields(buf):
        if fno == 1:
            dims += _packed_varints(val, wt)
        elif fno == 2:
            dtype = val
        elif fno == 4:
            if wt == 5:
                floats.append(struct.unpack("<f", val)[0])
            else:
                floats += list(np.frombuffer(val, dtype="<f4"))
        elif fno == 5:
            i32 += _packed_varints(val, wt)
        elif fno == 7:
            i64 += _packed_varints(val, wt)
        elif fno == 8:
            name = val.decode()
        elif fno == 9:
            raw = bytes(val)
    npdt = _NP_OF[dtype]
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np.dtype(npdt).newbyteorder("<")).astype(npdt)
    elif floats:
        arr = np.asarray(floats, dtype=npdt)
    elif i64:
        arr = np.asarray(i64, dtype=npdt)
    elif i32:
        arr = np.asarray(i32, dtype=npdt)
    else:
        arr = np.zeros(0, dtype=npdt)
    return name, arr.reshape(dims) if dims or arr.size == 1 else arr


def _parse_attr(buf: bytes) -> Tuple[str, Any]:
    name, atype = "", 0
    f = i = s = t = None
    floats: List[float] = []
    ints: List[int] = []
    strings: List[bytes] = []
    for fno, wt, val in _iter_fields(buf):
        if fno == 1:
            name = val.decode()
        elif fno == 2:
            f = struct.unpack("<f", val)[0]
        elif fno == 3:
            i = _signed64(val)
        elif fno == 4:
            s = bytes(val)
        elif fno == 5:
            t = _parse_tensor(val)[1]
        elif fno == 7:
            if wt == 5:
                floats.append(struct.unpack("<f", val)[0])
            else:
                floats += [float(x) for x in np.frombuffer(val, dtype="<f4")]
        elif fno == 8:
            ints += _packed_varints(val, wt)
        elif fno == 9:
            strings.append(bytes(val))
        elif fno == 20:
            atype = val
    if atype == 1 or (atype == 0 and f is not None):
        return name, float(f if f is not None else 0.0)
    if atype == 2 or (atype == 0 and i is not None):
        return name, int(i if i is not None else 0)
    if atype == 3 or (atype == 0 and s is not None):
        return name, (s or b"").decode()
    if atype == 4 or (atype == 0 and t is not None):
        return name, t
    if atype == 6:
        return name, floats
    if atype == 7:
        return name, ints
    if atype == 8:
        return name, [x.decode() for x in strings]
    return name, ints or floats or None


def _parse_value_info(buf: bytes) -> ValueInfo:
    vi = ValueInfo("")
    for fno, wt, val in _iter_fields(buf):
        if fno == 1:
            vi.name = val.decode()
        elif fno == 2:  # TypeProto
            for f2, _, v2 in _iter_fields(val):
                if f2 != 1:  # tensor_type
                    continue
                for f3, _, v3 in _iter_fields(v2):
                    if f3 == 1:
                        vi.elem_type = v3
                    elif f3 == 2:  # TensorShapeProto
                        for f4, _, v4 in _iter_fields(v3):
                            if f4 != 1:
                                continue
                            dim: Any = -1
                            for f5, _, v5 in _iter_fields(v4):
                                if f5 == 1:
                                    dim = _signed64(v5)
                                elif f5 == 2:
                                    dim = v5.decode()
                            vi.shape.append(dim)
    return vi


def _parse_node(buf: bytes) -> Node:
    n = Node("", [], [])
    for fno, wt, val in _iter_fields(buf):
        if fno == 1:
            n.inputs.append(val.decode())
        elif fno == 2:
            n.outputs.append(val.decode())
        elif fno == 3:
            n.name = val.decode()
        elif fno == 4:
            n.op_type = val.decode()
        elif fno == 5:
            k, v = _parse_attr(val)
            n.attrs[k] = v
    return n


def _parse_graph(buf: bytes) -> Graph:
    g = Graph()
    raw_inputs: List[ValueInfo] = []
    for fno, wt, val in _iter_fields(buf):
        if fno == 1:
            g.nodes.append(_parse_node(val))
        elif fno == 2:
            g.name = val.decode()
        elif fno == 5:
            name, arr = _parse_tensor(val)
            g.initializers[name] = arr
        elif fno == 11:
            raw_inputs.append(_parse_value_info(val))
        elif fno == 12:
            g.outputs.append(_parse_value_info(val))
    g.inputs = [vi for vi in raw_inputs if vi.name not in g.initializers]
    return g


def load_bytes(buf: bytes) -> Model:
    m = Model(Graph())
    for fno, wt, val in _iter_fields(buf):
        if fno == 1:
            m.ir_version = val
        elif fno == 2:
            m.producer_name = val.decode()
        elif fno == 7:
            m.graph = _parse_graph(val)
        elif fno == 8:
            dom, ver = "", 0
            for f2, _, v2 in _iter_fields(val):
                if f2 == 1:
                    dom = v2.decode()
                elif f2 == 2:
                    ver = v2
            if dom in ("", "ai.onnx"):
                m.opset = ver
    return m


def load(path: str) -> Model:
    with open(path, "rb") as fh:
        return load_bytes(fh.read())


# ---------------------------------------------------------------- writer

def _ser_tensor(name: str, arr: np.ndarray) -> bytes:
    arr = np.ascontiguousarray(arr)
    out = b""
    for d in arr.shape:
        out += _w_int(1, int(d))
    out += _w_int(2, _DT_OF[arr.dtype])
    out += _w_str(8, name)
    out += _w_bytes(9, arr.astype(arr.dtype.newbyteorder("<")).tobytes())
    return out


def _ser_attr(name: str, v: Any) -> bytes:
    out = _w_str(1, name)
    if isinstance(v, bool):
        v = int(v)
    if isinstance(v, float):
        out += _w_key(2, 5) + struct.pack("<f", v) + _w_int(20, 1)
    elif isinstance(v, int):
        out += _w_int(3, v) + _w_int(20, 2)
    elif isinstance(v, str):
        out += _w_bytes(4, v.encode()) + _w_int(20, 3)
    elif isinstance(v, np.ndarray):
        out += _w_bytes(5, _ser_tensor("", v)) + _w_int(20, 4)
    elif isinstance(v, (list, tuple)) and v and isinstance(v[0], float):
        for x in v:
            out += _w_key(7, 5) + struct.pack("<f", x)
        out += _w_int(20, 6)
    elif isinstance(v, (list, tuple)):
        for x in v:
            out += _w_int(8, int(x))
        out += _w_int(20, 7)
    else:
        raise TypeError(f"attribute {name}: unsupported {type(v)}")
    return out


def _ser_value_info(vi: ValueInfo) -> bytes:
    dims = b""
    for d in vi.shape:
        if isinstance(d, str):
            dims += _w_bytes(1, _w_str(2, d))
        else:
            dims += _w_bytes(1, _w_int(1, int(d)))
    tensor_type = _w_int(1, vi.elem_type) + _w_bytes(2, dims)
    return _w_str(1, vi.name) + _w_bytes(2, _w_bytes(1, tensor_type))


def _ser_node(n: Node) -> bytes:
    out = b""
    for s in n.inputs:
        out += _w_str(1, s)
    for s in n.outputs:
        out += _w_str(2, s)
    if n.name:
        out += _w_str(3, n.name)
    out += _w_str(4, n.op_type)
    for k, v in n.attrs.items():
        out += _w_bytes(5, _ser_attr(k, v))
    return out


def dump_bytes(m: Model) -> bytes:
    g = m.graph
    gb = b""
    for n in g.nodes:
        gb += _w_bytes(1, _ser_node(n))
    gb += _w_str(2, g.name)
    for name, arr in g.initializers.items():
        gb += _w_bytes(5, _ser_tensor(name, arr))
    for vi in g.inputs:
        gb += _w_bytes(11, _ser_value_info(vi))
    for vi in g.outputs:
        gb += _w_bytes(12, _ser_value_info(vi))
    out = _w_int(1, m.ir_version) + _w_str(2, m.producer_name)
    out += _w_bytes(7, gb)
    out += _w_bytes(8, _w_str(1, "") + _w_int(2, m.opset))
    return out


def save(m: Model, path: str) -> None:
    with open(path, "wb") as fh:
        fh.write(dump_bytes(m))
