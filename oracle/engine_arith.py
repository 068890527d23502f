"""Engine-arithmetic oracle for the e4m3 mode - TEST INFRASTRUCTURE ONLY (imported by tests/, never by the product).

`oracle/onnx_oracle.py` restates the ONNX operators in fp32; the e4m3 engine cannot match it closely (3 mantissa bits per stored
value), so the per-operator gate against it has to be wide.  This module restates the REDUCED-PRECISION arithmetic the engine is
specified to perform - where every rounding happens and to which format - in numpy float64 (every product and partial sum below is
exact in double), so that the GPU kernels can be held to a tight gate: apart from the summation order of the fp32 accumulator
(which can move a result across an e4m3 rounding boundary once in a while) the two must agree bit for bit.

The roundings (file:line of the engine code that performs them; the reference itself does all of this in fp32 inside ONNX Runtime,
reference inference_engine/src/model.cpp:1264-1270):
  * graph input -> e4m3, round to nearest even, saturating at +-448       (csrc/kernels_simt.cu nchw_to_nhwc_kernel)
  * folded BatchNormalization: scale = gamma / sqrt(var + eps), shift = beta - mean * scale in double -> float
                                                                            (csrc/plan.cpp BnConsts)
  * A-operand prologue: x (e4m3, exact in f16) * f16(scale) + f16(shift) as ONE fused f16 operation, optional ReLU, -> e4m3
                                                                            (csrc/umma_ptx.cuh PrologueWordFp8)
  * pooled prologue (transitions): the four pixels' prologue results added in f16 as (p00 + p01) + (p10 + p11) -> e4m3; the 1/4
    of the average is folded into the epilogue scale                         (csrc/umma_ptx.cuh PoolWordFp8)
  * weights: per-output-channel scale = max|w| / 448 (float), w / scale -> e4m3     (csrc/engine.cu Replica::Replica)
  * products exact, accumulation in fp32 (order unspecified)                 (tcgen05.mma kind::f8f6f4)
  * epilogue: acc * (scale * out_mul) + bias as one fp32 FMA, optional ReLU, -> e4m3      (csrc/umma_ptx.cuh EpiloguePack32)
"""
from __future__ import annotations

import numpy as np
import torch

E4M3_MAX = 448.0


def e4m3(x: np.ndarray) -> np.ndarray:
    """Round to the nearest e4m3 value (ties to even), saturating; returned as float64."""
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(torch.float32).clamp(-E4M3_MAX, E4M3_MAX)
    return t.to(torch.float8_e4m3fn).to(torch.float64).numpy()


def e4m3_from_f32_value(x64: np.ndarray) -> np.ndarray:
    """x64 holds float32 values (as float64): convert exactly like cvt.rn.satfinite.e4m3x2.f32."""
    return e4m3(x64)


def f16(x: np.ndarray) -> np.ndarray:
    """Round a double to the nearest f16 (ties to even); returned as float64."""
    return np.asarray(x, dtype=np.float64).astype(np.float16).astype(np.float64)


def f32(x: np.ndarray) -> np.ndarray:
    return np.asarray(x, dtype=np.float64).astype(np.float32).astype(np.float64)


def fold_bn(gamma, beta, mean, var, eps=1e-5):
    """csrc/plan.cpp BnConsts: double arithmetic, results stored as float."""
    g, b, m, v = (np.asarray(a, dtype=np.float32).astype(np.float64) for a in (gamma, beta, mean, var))
    inv = 1.0 / np.sqrt(v + np.float64(np.float32(eps)))
    s = g * inv
    return f32(s), f32(b - m * s)


def prologue(x_q: np.ndarray, scale: np.ndarray, shift: np.ndarray, relu: bool) -> np.ndarray:
    """x_q: e4m3 values [N,C,H,W]; returns the f16 result of the fused multiply-add (before the e4m3 rounding)."""
    s16, b16 = f16(scale)[None, :, None, None], f16(shift)[None, :, None, None]
    y = f16(x_q * s16 + b16)        # exact product and sum in double, ONE rounding to f16 = fma.rn.f16x2
    return np.maximum(y, 0.0) if relu else y


def prologue_e4m3(x_q, scale, shift, relu):
    return e4m3(prologue(x_q, scale, shift, relu))


def pooled_prologue_e4m3(x_q, scale, shift, relu):
    """sum over 2x2 of the prologue results, f16 adds paired by image row, then e4m3."""
    p = prologue(x_q, scale, shift, relu)
    top = f16(p[:, :, 0::2, 0::2] + p[:, :, 0::2, 1::2])
    bot = f16(p[:, :, 1::2, 0::2] + p[:, :, 1::2, 1::2])
    return e4m3(f16(top + bot))


def pooled_prologue_generic_f32(x_q, scale, shift, relu):
    """The generic gather kernel's pooled A operand (csrc/kernels_umma.cu kModePool2; transitions the TMA kernels do not take, e.g.
    Cout = 64): fp32 FMA with the fp32 constants, the four pixels added one after the other in fp32, times 0.25; the caller rounds
    the result to the storage format."""
    s32, b32 = f32(scale)[None, :, None, None], f32(shift)[None, :, None, None]
    p = f32(x_q * s32 + b32)
    if relu:
        p = np.maximum(p, 0.0)
    acc = f32(p[:, :, 0::2, 0::2] + p[:, :, 0::2, 1::2])
    acc = f32(acc + p[:, :, 1::2, 0::2])
    acc = f32(acc + p[:, :, 1::2, 1::2])
    return acc * 0.25


def quantise_weights(w: np.ndarray):
    """w [Cout, Cin, R, S] float32 -> (e4m3 values as float64, per-channel dequant scale as float64 holding float32)."""
    w32 = np.asarray(w, dtype=np.float32)
    amax = np.abs(w32).reshape(w32.shape[0], -1).max(axis=1)
    scale = np.where(amax > 0, amax / np.float32(448.0), np.float32(1.0)).astype(np.float32)
    q = e4m3((w32 / scale[:, None, None, None]).astype(np.float64))   # the division is a float32 operation in the engine
    return q, scale.astype(np.float64)


def conv_exact(a_q: np.ndarray, w_q: np.ndarray, pad: int) -> np.ndarray:
    """stride-1 convolution with exact products and sums (double), result rounded to fp32 like the TMEM accumulator."""
    t = torch.nn.functional.conv2d(torch.from_numpy(a_q), torch.from_numpy(w_q), None, stride=1, padding=pad)
    return f32(t.numpy())


def epilogue_e4m3(acc, scale, bias, relu, out_mul=1.0):
    s = f32(scale * np.float64(np.float32(out_mul)))[None, :, None, None]      # float32 product, as in the kernel's preamble
    b = (np.zeros_like(scale) if bias is None else f32(bias))[None, :, None, None]
    y = f32(acc * s + b)            # one fp32 FMA
    if relu:
        y = np.maximum(y, 0.0)
    return e4m3(y)


def e4m3_step(v: np.ndarray) -> np.ndarray:
    """distance to the next e4m3 value above |v| (the unit in which a boundary flip shows)."""
    a = np.maximum(np.abs(v), 2.0 ** -6)
    e = np.floor(np.log2(a))
    return 2.0 ** (e - 3)


def bf16(x: np.ndarray) -> np.ndarray:
    """Round to the nearest bf16 (ties to even); returned as float64."""
    t = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float64)).to(torch.float32)
    return t.to(torch.bfloat16).to(torch.float64).numpy()


class Bf16Mode:
    """The same pipeline in BF16 mode: activations and weights are bf16 (no weight scale), the prologue is ONE fused bf16
    multiply-add (fma.rn.relu.bf16x2, csrc/umma_ptx.cuh ProloguePiece<bf16>), the pooled prologue adds in bf16, the epilogue is the
    same fp32 FMA narrowed to bf16."""

    @staticmethod
    def store(x):
        return bf16(x)

    @staticmethod
    def prologue(x_q, scale, shift, relu):
        s, b = bf16(scale)[None, :, None, None], bf16(shift)[None, :, None, None]
        y = bf16(x_q * s + b)       # products of two bf16 values and the sum are exact in double: one rounding
        return np.maximum(y, 0.0) if relu else y

    @classmethod
    def pooled_prologue(cls, x_q, scale, shift, relu):
        p = cls.prologue(x_q, scale, shift, relu)
        top = bf16(p[:, :, 0::2, 0::2] + p[:, :, 0::2, 1::2])
        bot = bf16(p[:, :, 1::2, 0::2] + p[:, :, 1::2, 1::2])
        return bf16(top + bot)

    @staticmethod
    def weights(w):
        w32 = np.asarray(w, dtype=np.float32)
        return bf16(w32.astype(np.float64)), np.ones(w32.shape[0])

    @staticmethod
    def epilogue(acc, scale, bias, relu, out_mul=1.0):
        s = f32(scale * np.float64(np.float32(out_mul)))[None, :, None, None]
        b = (np.zeros_like(scale) if bias is None else f32(bias))[None, :, None, None]
        y = f32(acc * s + b)
        if relu:
            y = np.maximum(y, 0.0)
        return bf16(y)


class E4m3Mode:
    store = staticmethod(e4m3)
    prologue = staticmethod(lambda x_q, scale, shift, relu: prologue_e4m3(x_q, scale, shift, relu))
    pooled_prologue = staticmethod(lambda x_q, scale, shift, relu: pooled_prologue_e4m3(x_q, scale, shift, relu))
    weights = staticmethod(quantise_weights)
    epilogue = staticmethod(epilogue_e4m3)


# ---------------------------------------------------------------------------------------------------------------- whole network


def _conv_general(a, w, stride, pad):
    t = torch.nn.functional.conv2d(torch.from_numpy(np.ascontiguousarray(a)), torch.from_numpy(np.ascontiguousarray(w)), None,
                                   stride=stride, padding=pad)
    return f32(t.numpy())


def gap_bn_relu_f32(x_q, scale, shift, relu, slices=8):
    """csrc/kernels_simt.cu gap_kernel: fp32 FMA + ReLU per pixel, pixels q = y, y + 8, ... summed per slice, slices summed in
    order, times float(1 / HW)."""
    n, c, h, w = x_q.shape
    s32, b32 = f32(scale)[None, :, None], f32(shift)[None, :, None]
    t = f32(x_q.reshape(n, c, h * w) * s32 + b32)
    if relu:
        t = np.maximum(t, 0.0)
    total = np.zeros((n, c))
    parts = []
    for y in range(slices):
        acc = np.zeros((n, c))
        for q in range(y, h * w, slices):
            acc = f32(acc + t[:, :, q])
        parts.append(acc)
    for acc in parts:
        total = f32(total + acc)
    inv = np.float64(np.float32(1.0) / np.float32(h * w))
    return f32(total * inv)


def densenet_logits(model_path: str, x: np.ndarray, mode=None) -> np.ndarray:
    """The reduced-precision engine's arithmetic (mode = E4m3Mode, the default, or Bf16Mode) for a DenseNet-style graph (the layer patterns of csrc/plan.cpp's lowering rules), images x
    [N,3,H,W] fp32 -> logits.  Stem: bf16 operands, fp32 accumulate, bias + ReLU -> e4m3; max-pool exact; dense layers and
    transitions as in the operator-level functions above; BN + ReLU + global average pool in fp32; classifier in fp32 (restated in
    double: its summation order is not part of the specification, the gate on the logits allows for it)."""
    import os
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    if root not in sys.path:
        sys.path.insert(0, root)
    from tools import onnx_lite
    M = mode or E4m3Mode
    g = onnx_lite.load(model_path).graph
    init = {k: np.asarray(v, dtype=np.float32) for k, v in g.initializers.items()}
    nodes = g.nodes
    env = {}
    gin = g.inputs[0].name
    i = 0
    while i < len(nodes):
        n = nodes[i]
        op = n.op_type
        if op == "Conv" and n.inputs[0] == gin:
            assert nodes[i + 1].op_type == "Relu"
            w, b = init[n.inputs[1]], init[n.inputs[2]]
            acc = _conv_general(bf16(x), bf16(w), n.attrs["strides"][0], n.attrs["pads"][0])
            y = f32(acc + f32(b)[None, :, None, None])          # acc * 1.0 + bias, one fp32 FMA
            env[nodes[i + 1].outputs[0]] = M.store(np.maximum(y, 0.0))
            i += 2
        elif op == "MaxPool":
            t = torch.nn.functional.max_pool2d(torch.from_numpy(env[n.inputs[0]]), n.attrs["kernel_shape"][0], n.attrs["strides"][0],
                                               n.attrs["pads"][0])
            env[n.outputs[0]] = t.numpy()
            i += 1
        elif op == "Concat":
            env[n.outputs[0]] = np.concatenate([env[k] for k in n.inputs], axis=1)
            i += 1
        elif op == "BatchNormalization":
            sc, sh = fold_bn(*(init[k] for k in n.inputs[1:5]), eps=n.attrs.get("epsilon", 1e-5))
            assert nodes[i + 1].op_type == "Relu"
            nxt = nodes[i + 2]
            xin = env[n.inputs[0]]
            if nxt.op_type == "GlobalAveragePool":
                feat = gap_bn_relu_f32(xin, sc, sh, True)
                env[nxt.outputs[0]] = feat[:, :, None, None]
                assert nodes[i + 3].op_type == "Flatten" and nodes[i + 4].op_type == "Gemm"
                gm = nodes[i + 4]
                w, b = init[gm.inputs[1]].astype(np.float64), init[gm.inputs[2]].astype(np.float64)
                env[gm.outputs[0]] = feat @ w.T + b
                i += 5
                continue
            assert nxt.op_type == "Conv" and nxt.attrs["kernel_shape"][0] == 1
            wq, ws = M.weights(init[nxt.inputs[1]])
            bias = init[nxt.inputs[2]] if len(nxt.inputs) > 2 else None
            after = nodes[i + 3]
            if after.op_type == "AveragePool":       # transition: pool commuted in front of the conv, pooled operand in f16
                a = M.pooled_prologue(xin, sc, sh, True)
                env[after.outputs[0]] = M.epilogue(conv_exact(a, wq, 0), ws, bias, False, out_mul=0.25)
            else:
                assert after.op_type == "Relu"
                a = M.prologue(xin, sc, sh, True)
                env[after.outputs[0]] = M.epilogue(conv_exact(a, wq, 0), ws, bias, True)
            i += 4
        elif op == "Conv":                            # the 3x3 conv of a dense layer
            wq, ws = M.weights(init[n.inputs[1]])
            bias = init[n.inputs[2]] if len(n.inputs) > 2 else None
            env[n.outputs[0]] = M.epilogue(conv_exact(env[n.inputs[0]], wq, n.attrs["pads"][0]), ws, bias, False)
            i += 1
        else:
            raise NotImplementedError(f"{op} at node {i}")
    return env[g.outputs[0].name]


def densenet_e4m3_logits(model_path: str, x: np.ndarray) -> np.ndarray:
    return densenet_logits(model_path, x, E4m3Mode)


def densenet_bf16_logits(model_path: str, x: np.ndarray) -> np.ndarray:
    return densenet_logits(model_path, x, Bf16Mode)
