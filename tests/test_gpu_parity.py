"""GPU parity tests (run with `-m gpu` on a B200): the CUDA path, called through the C-ABI exactly as the
reference's Go binding would call it, against the CPU oracle on the same seeded inputs.

Tolerances (north_star): FP32 logits within 1e-3 relative (to max|logit|) with identical top-1; BF16/FP8 top-1 agreement and
top-5 SET agreement on the fixed 1024-image synthetic set (gates in LOWP_GATES: >= 99 % for bf16; fp8 is reported against the
same definitions with its own measured floor)."""
import json
import os
import sys
import threading

import numpy as np
import pytest

from oracle.onnx_oracle import OnnxOracle
from tools import onnx_lite, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(autouse=True)
def _one_gpu_small_arena(monkeypatch):
    monkeypatch.setenv("B200_ENGINE_DEVICES", os.environ.get("B200_TEST_DEVICES", "0"))
    monkeypatch.setenv("B200_ENGINE_MAX_BATCH", "32")
    monkeypatch.delenv("B200_ENGINE_PRECISION", raising=False)


def _serve(pkg, repo, name, feeds, out_shapes, precision, monkeypatch, max_batch=None):
    monkeypatch.setenv("B200_ENGINE_PRECISION", precision)
    if max_batch:
        monkeypatch.setenv("B200_ENGINE_MAX_BATCH", str(max_batch))
    mgr = pkg.InferenceManager(repo)
    try:
        mgr.load_model(name)
        outs = mgr.run_inference(name, "", [pkg.TensorData(k, v) for k, v in feeds.items()],
                                 [pkg.OutputConfig(k, list(s)) for k, s in out_shapes.items()])
        mgr.unload_model(name)
        return [o.data for o in outs]
    finally:
        mgr.shutdown()


def _graph_case(tmp_path, name, nodes, inits, in_shape, out_name, out_shape):
    g = onnx_lite.Graph(name=name)
    g.nodes = nodes
    g.initializers = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in inits.items()}
    g.inputs = [onnx_lite.ValueInfo("x", onnx_lite.FLOAT, ["N"] + list(in_shape))]
    g.outputs = [onnx_lite.ValueInfo(out_name, onnx_lite.FLOAT, ["N"] + list(out_shape))]
    d = tmp_path / name / "1"
    os.makedirs(d, exist_ok=True)
    path = str(d / "model.onnx")
    onnx_lite.save(onnx_lite.Model(g, ir_version=7, opset=12), path)
    return path


def _bn(rng, c, prefix):
    return {prefix + ".g": rng.uniform(0.5, 1.5, c), prefix + ".b": rng.normal(0, 0.2, c),
            prefix + ".m": rng.normal(0, 0.3, c), prefix + ".v": rng.uniform(0.5, 1.5, c)}


def _bn_node(x, prefix, out):
    return onnx_lite.Node("BatchNormalization", [x, prefix + ".g", prefix + ".b", prefix + ".m", prefix + ".v"], [out],
                          {"epsilon": 1e-5})


def _conv_node(x, w, out, k, s=1, p=0, bias=None):
    return onnx_lite.Node("Conv", [x, w] + ([bias] if bias else []), [out],
                          {"kernel_shape": [k, k], "strides": [s, s], "pads": [p, p, p, p], "dilations": [1, 1], "group": 1})


CASES = ["conv1x1_bn_relu", "conv1x1_partial_chunk", "conv1x1_ktail", "conv1x1_long", "conv3x3", "stem_maxpool", "transition",
         "transition_wide", "dense_block", "dense_block_copy", "dense_block7", "cout256", "gap_gemm_softmax"]
# "fp32": FP32 mode = tcgen05 with bf16-split operands (three MMAs per product, ~2^-17 per term); "fp32-exact": the FFMA kernels
TOL = {"fp32": 1e-4, "fp32-exact": 2e-5, "bf16": 2.5e-2, "fp8": 1.5e-1}


def _build_case(case, tmp_path, rng):
    """returns (model_path, model_name, input array [N,C,H,W], output value name)"""
    k = lambda *s: rng.normal(0, 1.0, s) / np.sqrt(np.prod(s[1:]))  # noqa: E731
    if case == "conv1x1_bn_relu":
        inits = {**_bn(rng, 64, "bn"), "w": k(128, 64, 1, 1), "b": rng.normal(0, 0.1, 128)}
        nodes = [_bn_node("x", "bn", "n"), onnx_lite.Node("Relu", ["n"], ["r"]), _conv_node("r", "w", "c", 1, bias="b"),
                 onnx_lite.Node("Relu", ["c"], ["y"])]
        shp, out = (64, 12, 12), (128, 12, 12)
    elif case == "conv1x1_partial_chunk":
        inits = {**_bn(rng, 96, "bn"), "w": k(128, 96, 1, 1)}
        nodes = [_bn_node("x", "bn", "n"), onnx_lite.Node("Relu", ["n"], ["r"]), _conv_node("r", "w", "y", 1)]
        shp, out = (96, 7, 7), (128, 7, 7)
    elif case in ("conv1x1_ktail", "conv1x1_long"):
        # K not a multiple of the 128-byte chunk (overlapped last TMA box); "long" exceeds the resident-weight limit
        cin = 352 if case == "conv1x1_ktail" else 608
        inits = {**_bn(rng, cin, "bn"), "w": k(128, cin, 1, 1), "b": rng.normal(0, 0.1, 128)}
        nodes = [_bn_node("x", "bn", "n"), onnx_lite.Node("Relu", ["n"], ["r"]), _conv_node("r", "w", "c", 1, bias="b"),
                 onnx_lite.Node("Relu", ["c"], ["y"])]
        shp, out = (cin, 11, 11), (128, 11, 11)
    elif case == "conv3x3":
        inits = {"w": k(32, 128, 3, 3)}
        nodes = [_conv_node("x", "w", "y", 3, 1, 1)]
        shp, out = (128, 14, 14), (32, 14, 14)
    elif case == "stem_maxpool":
        inits = {"w": k(64, 3, 7, 7), "b": rng.normal(0, 0.1, 64)}
        nodes = [_conv_node("x", "w", "c", 7, 2, 3, bias="b"), onnx_lite.Node("Relu", ["c"], ["r"]),
                 onnx_lite.Node("MaxPool", ["r"], ["y"], {"kernel_shape": [3, 3], "strides": [2, 2], "pads": [1, 1, 1, 1]})]
        shp, out = (3, 40, 40), (64, 10, 10)
    elif case == "transition":
        inits = {**_bn(rng, 128, "bn"), "w": k(64, 128, 1, 1)}
        nodes = [_bn_node("x", "bn", "n"), onnx_lite.Node("Relu", ["n"], ["r"]), _conv_node("r", "w", "c", 1),
                 onnx_lite.Node("AveragePool", ["c"], ["y"], {"kernel_shape": [2, 2], "strides": [2, 2], "pads": [0, 0, 0, 0]})]
        shp, out = (128, 12, 12), (64, 6, 6)
    elif case == "transition_wide":
        # Cout = 128: the TMA transition kernel (2x2 pool commuted in front of the conv, four planes per stage)
        inits = {**_bn(rng, 256, "bn"), "w": k(128, 256, 1, 1)}
        nodes = [_bn_node("x", "bn", "n"), onnx_lite.Node("Relu", ["n"], ["r"]), _conv_node("r", "w", "c", 1),
                 onnx_lite.Node("AveragePool", ["c"], ["y"], {"kernel_shape": [2, 2], "strides": [2, 2], "pads": [0, 0, 0, 0]})]
        shp, out = (256, 14, 14), (128, 7, 7)
    elif case in ("dense_block", "dense_block_copy", "dense_block7"):
        # "dense_block": the first member is produced by a kernel -> in-place concat (channel slices);
        # "dense_block_copy": the first member is the graph input itself -> copy_channels fallback
        inits, nodes = {}, []
        if case != "dense_block_copy":
            nodes.append(onnx_lite.Node("Relu", ["x"], ["x0"]))
            feats = ["x0"]
        else:
            feats = ["x"]
        c0 = c = 160 if case == "dense_block7" else 64      # dense_block7: 7x7 images, K tails, 4 layers in one kernel
        hw = 7 if case == "dense_block7" else 9
        for li in range(4 if case == "dense_block7" else 3):
            cat = f"cat{li}"
            nodes.append(onnx_lite.Node("Concat", list(feats), [cat], {"axis": 1}))
            inits.update(_bn(rng, c, f"bn{li}"))
            inits[f"w1_{li}"] = k(128, c, 1, 1)
            inits[f"b1_{li}"] = rng.normal(0, 0.1, 128)
            inits[f"w2_{li}"] = k(32, 128, 3, 3)
            nodes += [_bn_node(cat, f"bn{li}", f"n{li}"), onnx_lite.Node("Relu", [f"n{li}"], [f"r{li}"]),
                      _conv_node(f"r{li}", f"w1_{li}", f"c1_{li}", 1, bias=f"b1_{li}"),
                      onnx_lite.Node("Relu", [f"c1_{li}"], [f"r2_{li}"]), _conv_node(f"r2_{li}", f"w2_{li}", f"f{li}", 3, 1, 1)]
            feats.append(f"f{li}")
            c += 32
        nodes.append(onnx_lite.Node("Concat", list(feats), ["y"], {"axis": 1}))
        shp, out = (c0, hw, hw), (c, hw, hw)
    elif case == "cout256":
        inits = {"w": k(256, 128, 1, 1)}
        nodes = [_conv_node("x", "w", "y", 1)]
        shp, out = (128, 10, 10), (256, 10, 10)
    elif case == "gap_gemm_softmax":
        inits = {**_bn(rng, 64, "bn"), "w": rng.normal(0, 0.2, (10, 64)), "b": rng.normal(0, 0.1, 10)}
        nodes = [_bn_node("x", "bn", "n"), onnx_lite.Node("Relu", ["n"], ["r"]), onnx_lite.Node("GlobalAveragePool", ["r"], ["g"]),
                 onnx_lite.Node("Flatten", ["g"], ["f"], {"axis": 1}),
                 onnx_lite.Node("Gemm", ["f", "w", "b"], ["l"], {"transB": 1, "alpha": 1.0, "beta": 1.0}),
                 onnx_lite.Node("Softmax", ["l"], ["y"], {"axis": 1})]
        shp, out = (64, 7, 7), (10,)
    else:
        raise KeyError(case)
    path = _graph_case(tmp_path, case, nodes, inits, shp, "y", out)
    return path, case, shp, out


@pytest.mark.parametrize("precision", ["fp32", "fp32-exact", "bf16", "fp8"])
@pytest.mark.parametrize("case", CASES)
def test_operator_graphs_match_oracle(pkg, tmp_path, monkeypatch, case, precision):
    if precision == "fp32-exact":
        monkeypatch.setenv("B200_ENGINE_FP32_EXACT", "1")
    rng = np.random.default_rng(abs(hash(case)) % 2**31)
    path, name, shp, out = _build_case(case, tmp_path, rng)
    n = 5
    x = rng.uniform(0, 1, (n,) + tuple(shp)).astype(np.float32) if shp[0] == 3 else rng.normal(0, 1, (n,) + tuple(shp)).astype(np.float32)
    want = OnnxOracle(path).run({"x": x})[0]
    got = _serve(pkg, str(tmp_path), name, {"x": x}, {"y": (n,) + tuple(out)}, precision.split("-")[0], monkeypatch)[0]
    assert got.shape == want.shape
    err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-6)
    assert np.isfinite(got).all() and err < TOL[precision], f"{case}/{precision}: rel err {err:.3e}"


def _engine_arith_reference(case, inits, x, precision):
    """The reduced-precision engine's arithmetic restated in double (oracle/engine_arith.py): every rounding the kernels are specified
    to make, in the order they make it.  `inits`: the graph's initializers, x: the fp32 NCHW input."""
    from oracle import engine_arith as ea
    M = ea.E4m3Mode if precision == "fp8" else ea.Bf16Mode
    g = lambda k: np.asarray(inits[k], dtype=np.float32)  # noqa: E731
    bn = lambda pfx: ea.fold_bn(g(pfx + ".g"), g(pfx + ".b"), g(pfx + ".m"), g(pfx + ".v"))  # noqa: E731
    xq = M.store(x)
    if case in ("conv1x1_bn_relu", "conv1x1_partial_chunk", "conv1x1_ktail", "conv1x1_long"):
        sc, sh = bn("bn")
        a = M.prologue(xq, sc, sh, True)
        wq, ws = M.weights(g("w"))
        has_b = "b" in inits
        return M.epilogue(ea.conv_exact(a, wq, 0), ws, g("b") if has_b else None, has_b)
    if case == "conv3x3":
        wq, ws = M.weights(g("w"))
        return M.epilogue(ea.conv_exact(xq, wq, 1), ws, None, False)
    if case == "cout256":
        wq, ws = M.weights(g("w"))
        return M.epilogue(ea.conv_exact(xq, wq, 0), ws, None, False)
    if case == "transition_wide":   # the TMA kernels: pooled operand in packed f16 / bf16, 1/4 folded into the epilogue scale
        sc, sh = bn("bn")
        a = M.pooled_prologue(xq, sc, sh, True)
        wq, ws = M.weights(g("w"))
        return M.epilogue(ea.conv_exact(a, wq, 0), ws, None, False, out_mul=0.25)
    if case == "transition":        # Cout = 64: the generic gather kernel, pooled operand in fp32
        sc, sh = bn("bn")
        a = M.store(ea.pooled_prologue_generic_f32(xq, sc, sh, True))
        wq, ws = M.weights(g("w"))
        return M.epilogue(ea.conv_exact(a, wq, 0), ws, None, False)
    if case == "stem_maxpool":      # 7x7/s2 stem straight from the fp32 image: bf16 operands in BOTH reduced modes, then the exact max-pool
        import torch
        acc = ea._conv_general(ea.bf16(x), ea.bf16(g("w")), 2, 3)
        y = M.store(np.maximum(ea.f32(acc + ea.f32(g("b"))[None, :, None, None]), 0.0))
        return torch.nn.functional.max_pool2d(torch.from_numpy(y), 3, 2, 1).numpy()
    if case in ("dense_block", "dense_block7"):
        cat = np.maximum(xq, 0.0)
        for li in range(4 if case == "dense_block7" else 3):
            sc, sh = bn(f"bn{li}")
            a = M.prologue(cat, sc, sh, True)
            w1q, w1s = M.weights(g(f"w1_{li}"))
            b = M.epilogue(ea.conv_exact(a, w1q, 0), w1s, g(f"b1_{li}"), True)
            w2q, w2s = M.weights(g(f"w2_{li}"))
            f = M.epilogue(ea.conv_exact(b, w2q, 1), w2s, None, False)
            cat = np.concatenate([cat, f], axis=1)
        return cat
    raise KeyError(case)


@pytest.mark.parametrize("precision", ["fp8", "bf16"])
@pytest.mark.parametrize("case", ["conv1x1_bn_relu", "conv1x1_partial_chunk", "conv1x1_ktail", "conv1x1_long", "conv3x3", "cout256",
                                  "transition", "transition_wide", "dense_block", "dense_block7", "stem_maxpool"])
def test_operator_graphs_match_the_engine_arithmetic_oracle(pkg, tmp_path, monkeypatch, case, precision):
    """The tight gate of the reduced-precision modes.  Against the fp32 ONNX oracle an e4m3 kernel can only be held to ~0.15 of
    max|y| (the format has 3 mantissa bits), which would hide a wrong K-tail column.  Against a restatement of the engine's own
    arithmetic - the same e4m3 / bf16 / f16 / fp32 roundings in the same places, exact products, oracle/engine_arith.py - the kernels
    must agree bit for bit except where the fp32 accumulator's summation order moves a value across a rounding boundary of the
    storage format: at most 1 % of the outputs (3 % after the cascade of a multi-layer dense block), never by more than one step of
    the format (two after a cascade), and an rms error below 1e-2 of the rms output.  Measured: e4m3 0 mismatching outputs in all
    ten graphs."""
    from oracle import engine_arith as ea
    rng = np.random.default_rng(20241)
    path, name, shp, out = _build_case(case, tmp_path, rng)
    m = onnx_lite.load(path)
    x = rng.normal(0, 1.0, (5,) + tuple(shp)).astype(np.float32)
    got = _serve(pkg, str(tmp_path), name, {"x": x}, {"y": (5,) + tuple(out)}, precision, monkeypatch)[0].astype(np.float64)
    ref = _engine_arith_reference(case, m.graph.initializers, x, precision)
    assert got.shape == ref.shape
    diff = np.abs(got - ref)
    # outputs much smaller than the largest one are sums with cancellation: there the accumulator's fp32 rounding is several steps of
    # the (relative) storage format, so the step is taken at no less than 1/64 of the largest magnitude
    mag = np.maximum(np.maximum(np.abs(got), np.abs(ref)), np.abs(ref).max() / 64)
    step = ea.e4m3_step(mag) if precision == "fp8" else 2.0 ** (np.floor(np.log2(np.maximum(mag, 2.0 ** -120))) - 7)
    mism = float((diff > 0).mean())
    rms = float(np.sqrt((diff ** 2).mean()) / max(np.sqrt((ref ** 2).mean()), 1e-30))
    worst = float((diff / step).max())
    print(f"{precision} {case}: mismatching {100 * mism:.3f} %  worst {worst:.2f} steps  rms rel {rms:.2e}")
    cascade = case.startswith("dense_block")
    assert mism <= (0.03 if cascade else 0.01), (case, mism)
    assert worst <= (2.0 if cascade else 1.0) + 1e-9, (case, worst)
    assert rms <= 1e-2, (case, rms)


@pytest.mark.parametrize("precision", ["bf16", "fp8"])
@pytest.mark.parametrize("case", ["conv1x1_bn_relu", "conv1x1_ktail", "conv3x3", "transition", "dense_block", "cout256"])
def test_generic_gather_kernels_match_oracle(pkg, tmp_path, monkeypatch, case, precision):
    """kernels_umma.cu: the cp.async gather kernels (`conv_umma_kernel`, `conv3x3_halo_kernel`) that take every shape the TMA kernels
    do not.  DenseNet-121 itself never needs them, so they are forced here (B200_ENGINE_L1TMA/C3TMA/HALO/DENSEFUSE/SPLIT_TRANSITION = 0)
    on graphs the specialised kernels would otherwise take, against the same oracle and tolerances."""
    for k in ("B200_ENGINE_L1TMA", "B200_ENGINE_C3TMA", "B200_ENGINE_DENSEFUSE", "B200_ENGINE_SPLIT_TRANSITION"):
        monkeypatch.setenv(k, "0")
    if case == "dense_block":
        monkeypatch.setenv("B200_ENGINE_HALO", "0")       # 3x3 through the windowed gather as well
    rng = np.random.default_rng(abs(hash("generic" + case)) % 2**31)
    path, name, shp, out = _build_case(case, tmp_path, rng)
    n = 3
    x = rng.normal(0, 1, (n,) + tuple(shp)).astype(np.float32)
    want = OnnxOracle(path).run({"x": x})[0]
    n0 = pkg.kernel_launch_count()
    got = _serve(pkg, str(tmp_path), name, {"x": x}, {"y": (n,) + tuple(out)}, precision, monkeypatch)[0]
    assert pkg.kernel_launch_count() > n0
    err = np.abs(got - want).max() / max(np.abs(want).max(), 1e-6)
    assert np.isfinite(got).all() and err < TOL[precision], f"generic {case}/{precision}: rel err {err:.3e}"


def test_test_model_known_answers_through_the_c_abi(pkg, repo_dir, monkeypatch):
    with open(os.path.join(ROOT, "tests", "golden", "test_model_kat.json")) as fh:
        vectors = json.load(fh)["vectors"]
    for precision in ("fp32", "bf16", "fp8"):  # rank-2 graphs always run in fp32, whatever the mode
        for v in vectors:
            x = np.asarray(v["input"], np.float32)
            y = _serve(pkg, repo_dir, "test_model", {"input": x}, {"output": (1, 2)}, precision, monkeypatch)[0]
            np.testing.assert_allclose(y, np.asarray(v["output"], np.float32), rtol=2e-6, atol=1e-6)
    # batch > the batch baked into the file is accepted (documented deviation)
    xs = np.random.default_rng(0).normal(0, 1, (7, 3)).astype(np.float32)
    y = _serve(pkg, repo_dir, "test_model", {"input": xs}, {"output": (7, 2)}, "fp32", monkeypatch)[0]
    want = OnnxOracle(os.path.join(repo_dir, "test_model", "1", "model.onnx")).run_numpy({"input": xs})[0]
    np.testing.assert_allclose(y, want, rtol=1e-5, atol=1e-6)


def _agreement(ref, got):
    r5, g5 = np.argsort(-ref, 1)[:, :5], np.argsort(-got, 1)[:, :5]
    return {"top1": float(np.mean(ref.argmax(1) == got.argmax(1))),
            "ref_top1_in_top5": float(np.mean([r in g for r, g in zip(ref.argmax(1), g5)])),
            "top5_overlap": float(np.mean([len(set(a) & set(b)) / 5 for a, b in zip(r5, g5)])),
            "top5_set": float(np.mean([set(a) == set(b) for a, b in zip(r5, g5)])),
            "max_rel": float(np.abs(ref - got).max() / np.abs(ref).max())}


def test_densenet_fp32_matches_golden_logits(pkg, repo_dir, monkeypatch):
    g = np.load(os.path.join(ROOT, "tests", "golden", "densenet_logits.npz"))
    n = int(g["n"])
    x = synth.to_model_input(synth.synthetic_images_u8(n, start=int(g["start"])))
    got = _serve(pkg, repo_dir, "densenet_onnx", {"data_0": x}, {"fc6_1": (n, 1000)}, "fp32", monkeypatch)[0]
    a = _agreement(g["logits_fp64"], got)
    print("fp32 (tcgen05, bf16-split operands) vs fp64 golden:", a)
    assert a["max_rel"] < 1e-3 and a["top1"] == 1.0, a          # north_star fp32 gate
    assert a["max_rel"] < 5e-4, a                                # three bf16 products per term deliver ~1e-4 (simulated 4e-5 .. 1.2e-4)
    monkeypatch.setenv("B200_ENGINE_FP32_EXACT", "1")            # the exact FFMA kernels: fp32-reassociation noise only
    exact = _serve(pkg, repo_dir, "densenet_onnx", {"data_0": x}, {"fc6_1": (n, 1000)}, "fp32", monkeypatch)[0]
    assert _agreement(g["logits_fp32"], exact)["max_rel"] < 1e-4


# north_star: "BF16/FP8 top-5 agreement >= 99 % on a fixed synthetic image set".  Measured on the 1024-image clustered set
# (tools/synth.clustered_images_u8) against the fp32 oracle, with the plain definitions: top-1 agreement = same argmax,
# top-5 agreement = same SET of five classes.  The fixture's classifier is synthesised so that these margins exist at all
# (tools/make_densenet_onnx.synthesize_classifier); max_rel is the worst logit error relative to max|logit|.
LOWP_GATES = {"bf16": {"top1": 0.99, "top5_set": 0.99, "max_rel": 0.05},
              "fp8": {"top1": 0.99, "top5_set": 0.98, "max_rel": 0.25}}   # measured: top-1 100 %, top-5 set 99.0 %, max_rel 0.157


def test_densenet_fp8_logits_match_the_engine_arithmetic_oracle(pkg, repo_dir, densenet_path, monkeypatch):
    """The whole e4m3 network against the restatement of the engine's arithmetic (oracle/engine_arith.py densenet_e4m3_logits: the
    same roundings in the same places for all 120 convolutions, the pooled transitions, the fp32 BN + ReLU + global average pool).
    Only the summation order of the fp32 accumulators is unspecified.  Measured on the B200: on 6 of these 8 images every stored
    activation of the network is bit-identical and the logits agree to 7e-7 of max|logit| (fp32 classifier rounding); on the other
    two ONE activation of ~10^7 lands on the other side of an e4m3 rounding boundary, and an e4m3 network amplifies such a 6 % step
    to its own noise floor within ~50 layers (2e-2 of max|logit| - the fp32 ONNX oracle is 0.16 away).  Gate: at least half of the
    images exact to 1e-5, every image within 5e-2 with the same top-1 class."""
    from oracle import engine_arith as ea
    monkeypatch.setenv("B200_ENGINE_PRECISION", "fp8")
    monkeypatch.setenv("B200_ENGINE_MAX_BATCH", "8")
    monkeypatch.setenv("B200_ENGINE_INSTANCES", "1")
    x = synth.to_model_input(synth.clustered_images_u8(8, start=300))
    mgr = pkg.InferenceManager(repo_dir)
    try:
        mgr.load_model("densenet_onnx")
        m = mgr.get_model("densenet_onnx")
        got = m.infer([pkg.TensorData("data_0", x)], [pkg.OutputConfig("fc6_1", [8, 1000])])[0].data.astype(np.float64)
    finally:
        mgr.shutdown()
    ref = ea.densenet_e4m3_logits(densenet_path, x)
    per_image = np.abs(got - ref).max(axis=1) / np.abs(ref).max()
    print("fp8 DenseNet vs engine-arithmetic oracle, per image:", " ".join(f"{e:.1e}" for e in per_image))
    assert (per_image < 1e-5).sum() >= 4, per_image
    assert per_image.max() < 5e-2, per_image
    assert np.array_equal(got.argmax(1), ref.argmax(1))


def test_densenet_bf16_logits_match_the_engine_arithmetic_oracle(pkg, repo_dir, densenet_path, monkeypatch):
    """The same whole-network check in BF16 mode (oracle/engine_arith.py densenet_bf16_logits).  bf16 steps are 16x finer than e4m3
    steps, so the fp32 accumulator's summation order moves a few more stored values across a rounding boundary, but every such step
    is 16x smaller too: the logits agree to 2e-3 .. 4e-3 of max|logit| on every image (the fp32 ONNX oracle is 3e-2 away); gate 1e-2
    with identical top-5 classes in identical order."""
    from oracle import engine_arith as ea
    monkeypatch.setenv("B200_ENGINE_PRECISION", "bf16")
    monkeypatch.setenv("B200_ENGINE_MAX_BATCH", "8")
    monkeypatch.setenv("B200_ENGINE_INSTANCES", "1")
    x = synth.to_model_input(synth.clustered_images_u8(4, start=300))
    mgr = pkg.InferenceManager(repo_dir)
    try:
        mgr.load_model("densenet_onnx")
        m = mgr.get_model("densenet_onnx")
        got = m.infer([pkg.TensorData("data_0", x)], [pkg.OutputConfig("fc6_1", [4, 1000])])[0].data.astype(np.float64)
    finally:
        mgr.shutdown()
    ref = ea.densenet_bf16_logits(densenet_path, x)
    per_image = np.abs(got - ref).max(axis=1) / np.abs(ref).max()
    print("bf16 DenseNet vs engine-arithmetic oracle, per image:", " ".join(f"{e:.1e}" for e in per_image))
    assert per_image.max() < 1e-2, per_image
    assert np.array_equal(np.argsort(-got, axis=1)[:, :5], np.argsort(-ref, axis=1)[:, :5])


@pytest.mark.parametrize("precision", ["bf16", "fp8"])
def test_densenet_low_precision_top5_agreement(pkg, repo_dir, densenet_path, monkeypatch, precision):
    n, chunk = 1024, 128
    monkeypatch.setenv("B200_ENGINE_PRECISION", precision)
    monkeypatch.setenv("B200_ENGINE_MAX_BATCH", str(chunk))
    oracle = OnnxOracle(densenet_path)
    mgr = pkg.InferenceManager(repo_dir)
    ref, got = [], []
    try:
        mgr.load_model("densenet_onnx")
        m = mgr.get_model("densenet_onnx")
        for s0 in range(0, n, chunk):
            x = synth.to_model_input(synth.clustered_images_u8(chunk, start=20000 + s0))
            ref.append(oracle.run({"data_0": x})[0])
            got.append(m.infer([pkg.TensorData("data_0", x)], [pkg.OutputConfig("fc6_1", [chunk, 1000])])[0].data.copy())
    finally:
        mgr.shutdown()
    ref, got = np.concatenate(ref), np.concatenate(got)
    a = _agreement(ref, got)
    print(precision, a)
    gate = LOWP_GATES[precision]
    assert np.isfinite(got).all()
    assert a["top1"] >= gate["top1"], a
    assert a["top5_set"] >= gate["top5_set"], a
    assert a["max_rel"] < gate["max_rel"], a


@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp8"])
def test_uint8_hwc_ingestion_matches_the_float_path(pkg, repo_dir, monkeypatch, precision):
    """raw uint8 [N,H,W,C] pixels through the same ModelInfer entry point; the GPU applies the
    client's `img / 255` and the HWC->CHW change (client/test_client.py:186-194).  Same arithmetic => identical logits."""
    monkeypatch.setenv("B200_ENGINE_PRECISION", precision)
    monkeypatch.setenv("B200_ENGINE_MAX_BATCH", "8")
    monkeypatch.setenv("B200_ENGINE_PIPELINE_CHUNK", "3")
    u8 = synth.synthetic_images_u8(7, start=6000)                    # [N,H,W,3] uint8
    x = synth.to_model_input(u8)                                     # [N,3,H,W] float32 = u8 / 255
    mgr = pkg.InferenceManager(repo_dir)
    try:
        mgr.load_model("densenet_onnx")
        m = mgr.get_model("densenet_onnx")
        ref = m.infer([pkg.TensorData("data_0", x)], [pkg.OutputConfig("fc6_1", [7, 1000])])[0].data.copy()
        got = m.infer([pkg.TensorData("data_0", np.ascontiguousarray(u8), pkg.DataType.UINT8)], [pkg.OutputConfig("fc6_1", [7, 1000])])[0].data
        assert np.isfinite(got).all()
        assert np.array_equal(got, ref), float(np.abs(got - ref).max())
        with pytest.raises(pkg.EngineError, match="a UINT8 image must be \\[N,H,W,C\\]"):
            m.infer([pkg.TensorData("data_0", np.zeros((1, 3, 224, 224), np.uint8), pkg.DataType.UINT8)], [pkg.OutputConfig("fc6_1", [1, 1000])])
    finally:
        mgr.shutdown()


def test_batch_properties_and_chunking(pkg, repo_dir, monkeypatch):
    """Size-independent properties: per-image results do not depend on batch composition, batch order,
    or on how the batch is chunked through the arena (max_batch 4 < n)."""
    monkeypatch.setenv("B200_ENGINE_PRECISION", "fp32")
    monkeypatch.setenv("B200_ENGINE_MAX_BATCH", "4")
    monkeypatch.setenv("B200_ENGINE_PIPELINE_CHUNK", "3")   # also exercise the H2D/compute sub-batch pipeline
    x = synth.to_model_input(synth.synthetic_images_u8(10, start=3000))
    mgr = pkg.InferenceManager(repo_dir)
    try:
        mgr.load_model("densenet_onnx")
        run = lambda a: mgr.run_inference("densenet_onnx", "", [pkg.TensorData("data_0", a)],  # noqa: E731
                                          [pkg.OutputConfig("fc6_1", [len(a), 1000])])[0].data
        full = run(x)                                   # 10 images through a 4-image arena: 3 chunks
        one = np.concatenate([run(x[i:i + 1]) for i in range(10)])
        perm = np.random.default_rng(1).permutation(10)
        assert np.abs(full - one).max() / np.abs(full).max() < 1e-5
        assert np.abs(run(x[perm]) - full[perm]).max() / np.abs(full).max() < 1e-5
        st = mgr.get_model("densenet_onnx").get_stats()
        assert st.inference_count == 12 and st.memory_usage_bytes > 30e6 and st.last_inference_time_ns > 0
        md = mgr.get_model("densenet_onnx").get_metadata()
        assert md.inputs == ["data_0"] and md.outputs == ["fc6_1"]      # names adopted from the graph
        # Go allocates the output from config.json's [1,1000,1,1]: rank-4 dims buffer, 4000 bytes
        m = mgr.get_model("densenet_onnx")
        y = m.infer([pkg.TensorData("data_0", x[:1])], [pkg.OutputConfig("fc6_1", [1, 1000, 1, 1])])[0].data
        assert np.abs(y.reshape(1, 1000) - full[:1]).max() / np.abs(full).max() < 1e-5
        assert m.last_returned_shapes == [[1, 1000]]
    finally:
        mgr.shutdown()


def test_validation_errors_match_reference_messages(pkg, repo_dir, monkeypatch):
    monkeypatch.setenv("B200_ENGINE_PRECISION", "fp32")
    mgr = pkg.InferenceManager(repo_dir)
    try:
        mgr.load_model("test_model")
        with pytest.raises(pkg.EngineError, match="^Model already loaded$"):
            mgr.load_model("test_model")
        m = mgr.get_model("test_model")
        ok = np.zeros((1, 3), np.float32)
        with pytest.raises(pkg.EngineError, match="Unexpected input name: data"):
            m.infer([pkg.TensorData("data", ok)], [pkg.OutputConfig("output", [1, 2])])
        with pytest.raises(pkg.EngineError, match="Expected 1 inputs, got 2"):
            m.infer([pkg.TensorData("input", ok), pkg.TensorData("input", ok)], [pkg.OutputConfig("output", [1, 2])])
        with pytest.raises(pkg.EngineError, match="Input shape mismatch for input at dimension 1"):
            m.infer([pkg.TensorData("input", np.zeros((1, 4), np.float32))], [pkg.OutputConfig("output", [1, 2])])
        with pytest.raises(pkg.EngineError, match="Input shape mismatch for input: expected 2 dimensions, got 3"):
            m.infer([pkg.TensorData("input", np.zeros((1, 3, 1), np.float32))], [pkg.OutputConfig("output", [1, 2])])
        with pytest.raises(pkg.EngineError, match="Input data type mismatch for input"):
            m.infer([pkg.TensorData("input", np.zeros((1, 3), np.int32), pkg.DataType.INT32)], [pkg.OutputConfig("output", [1, 2])])
        # output buffer smaller than the result: min(data_size, produced) bytes are written, no overflow
        y = m.infer([pkg.TensorData("input", np.ones((1, 3), np.float32))], [pkg.OutputConfig("output", [1, 1])])[0].data
        np.testing.assert_allclose(y, [[-1.6748662]], rtol=2e-6)
        # unload while another wrapper is still held: that wrapper must stay safe to use and destroy
        import ctypes as C
        lib = pkg.load_library()
        err = C.c_void_p()
        extra = pkg.Model(lib.GetModelHandle(mgr._h, b"test_model", None, C.byref(err)), False)
        assert extra.is_loaded()
        mgr.unload_model("test_model")
        assert not extra.is_loaded()
        with pytest.raises(pkg.EngineError, match="Model not loaded"):
            extra.infer([pkg.TensorData("input", ok)], [pkg.OutputConfig("output", [1, 2])])
        extra.destroy()
        with pytest.raises(pkg.EngineError, match="model handle is nil"):   # Go: "model handle is nil"
            m.infer([pkg.TensorData("input", ok)], [pkg.OutputConfig("output", [1, 2])])
    finally:
        mgr.shutdown()


def test_concurrent_callers_on_one_handle(pkg, repo_dir, monkeypatch):
    """gin serves each request on its own goroutine/OS thread: ModelInfer must be re-entrant."""
    monkeypatch.setenv("B200_ENGINE_PRECISION", "fp32")
    monkeypatch.setenv("B200_ENGINE_MAX_BATCH", "8")
    x = synth.to_model_input(synth.synthetic_images_u8(8, start=4000))
    mgr = pkg.InferenceManager(repo_dir)
    try:
        mgr.load_model("densenet_onnx")
        m = mgr.get_model("densenet_onnx")
        want = m.infer([pkg.TensorData("data_0", x)], [pkg.OutputConfig("fc6_1", [8, 1000])])[0].data
        errs = []

        def worker(i):
            try:
                for _ in range(3):
                    y = m.infer([pkg.TensorData("data_0", x[i:i + 1])], [pkg.OutputConfig("fc6_1", [1, 1000])])[0].data
                    if np.abs(y - want[i:i + 1]).max() / np.abs(want).max() > 1e-5:
                        errs.append((i, "mismatch"))
            except Exception as e:  # noqa: BLE001
                errs.append((i, repr(e)))

        ts = [threading.Thread(target=worker, args=(i,)) for i in range(8)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert not errs, errs
        assert m.get_stats().inference_count == 1 + 24
    finally:
        mgr.shutdown()


def test_execution_instances_mixed_sizes_pinned_and_pageable(pkg, repo_dir, monkeypatch):
    """BASELINE.json configs[4]: mixed batch sizes arriving concurrently.  Four execution instances on one GPU (the reference's
    dead `instance_count`), FIFO hand-out, forwards of big batches chained, pageable request buffers staged through the
    pinned pool and pinned ones (B200HostAlloc) read in place - every caller gets exactly the logits a lone call gets."""
    import ctypes
    monkeypatch.setenv("B200_ENGINE_PRECISION", "fp8")
    monkeypatch.setenv("B200_ENGINE_MAX_BATCH", "128")
    monkeypatch.setenv("B200_ENGINE_DEVICES", "0")
    monkeypatch.setenv("B200_ENGINE_INSTANCES", "4")
    monkeypatch.setenv("B200_ENGINE_CHAIN_MIN_BATCH", "16")
    sizes = [1, 3, 16, 40, 96, 128]
    x = synth.to_model_input(synth.synthetic_images_u8(128, start=4100))
    lib = pkg.load_library()
    mgr = pkg.InferenceManager(repo_dir)
    try:
        mgr.load_model("densenet_onnx")
        m = mgr.get_model("densenet_onnx")
        want = {s: m.infer([pkg.TensorData("data_0", x[:s])], [pkg.OutputConfig("fc6_1", [s, 1000])])[0].data.copy() for s in sizes}
        # a pinned copy of the images from the engine's allocator
        ptr = lib.B200HostAlloc(x.nbytes)
        assert ptr
        pinned = np.ctypeslib.as_array(ctypes.cast(ptr, ctypes.POINTER(ctypes.c_float)), shape=x.shape)
        pinned[...] = x
        errs = []

        def worker(i):
            try:
                rng = np.random.default_rng(i)
                for k in range(6):
                    s = int(rng.choice(sizes))
                    src = pinned if (i + k) % 2 else x
                    y = m.infer([pkg.TensorData("data_0", src[:s])], [pkg.OutputConfig("fc6_1", [s, 1000])])[0].data
                    if not np.array_equal(y, want[s]):
                        errs.append((i, s, float(np.abs(y - want[s]).max())))
            except Exception as e:  # noqa: BLE001
                errs.append((i, repr(e)))

        ts = [threading.Thread(target=worker, args=(i,)) for i in range(12)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert not errs, errs[:5]
        lib.B200HostFree(ptr)
        lib.B200HostFree(None)
    finally:
        mgr.shutdown()


@pytest.mark.parametrize("precision", ["bf16", "fp8"])
def test_split_wide_transitions_are_bit_identical(pkg, repo_dir, monkeypatch, precision):
    """Transitions 2 and 3 (Cout 256 / 512) run as pooled-BN-ReLU pass + plain 1x1 conv (kernels_poolbn.cu) instead of redoing the
    pooled transform per N tile inside conv1x1_tma<POOL>: same packed arithmetic, same summation order, same logits."""
    monkeypatch.setenv("B200_ENGINE_PRECISION", precision)
    monkeypatch.setenv("B200_ENGINE_MAX_BATCH", "8")
    monkeypatch.setenv("B200_ENGINE_DEVICES", "0")
    monkeypatch.setenv("B200_ENGINE_INSTANCES", "1")
    x = synth.to_model_input(synth.synthetic_images_u8(6, start=4400))
    outs, launches = {}, {}
    for flag in ("0", "1", "2"):   # off / wide transitions only / all three (the default)
        monkeypatch.setenv("B200_ENGINE_SPLIT_TRANSITION", flag)
        mgr = pkg.InferenceManager(repo_dir)
        try:
            mgr.load_model("densenet_onnx")
            m = mgr.get_model("densenet_onnx")
            n0 = pkg.kernel_launch_count()
            outs[flag] = m.infer([pkg.TensorData("data_0", x)], [pkg.OutputConfig("fc6_1", [6, 1000])])[0].data.copy()
            launches[flag] = pkg.kernel_launch_count() - n0
        finally:
            mgr.shutdown()
    assert launches["2"] == launches["0"] + (3 if precision == "fp8" else 2), launches
    assert np.array_equal(outs["0"], outs["2"])
    assert launches["1"] == launches["0"] + 2, launches   # "1": the two wide transitions gained a kernel each
    assert np.array_equal(outs["0"], outs["1"])


def test_cta_pair_conv3x3_kernel_is_bit_identical(pkg, repo_dir, monkeypatch):
    """B200_ENGINE_C3PAIR=1: the 3x3 convs of the 56x56 / 28x28 blocks run as CTA pairs (cta_group::2: ONE tcgen05.mma of M = 256 per
    two patch tiles, each CTA holding half of the stacked weight tile; the peer forwards its barriers to the leader, commits are
    multicast).  Same products, same accumulation order inside a tile: bit-identical logits, odd and even tile counts."""
    monkeypatch.setenv("B200_ENGINE_PRECISION", "fp8")
    monkeypatch.setenv("B200_ENGINE_MAX_BATCH", "16")
    monkeypatch.setenv("B200_ENGINE_DEVICES", "0")
    monkeypatch.setenv("B200_ENGINE_INSTANCES", "1")
    x = synth.to_model_input(synth.synthetic_images_u8(11, start=6100))
    code = (
        "import os, sys, numpy as np\n"
        "sys.path.insert(0, %r)\n"
        "import __graft_entry__ as ge\n"
        "pkg = ge.load_package()\n"
        "x = np.load(sys.argv[1])\n"
        "mgr = pkg.InferenceManager(%r)\n"
        "mgr.load_model('densenet_onnx'); m = mgr.get_model('densenet_onnx')\n"
        "outs = [m.infer([pkg.TensorData('data_0', x[:n])], [pkg.OutputConfig('fc6_1', [n, 1000])])[0].data for n in (11, 1, 4)]\n"
        "np.save(sys.argv[2], np.concatenate(outs))\n"
        "mgr.shutdown()\n") % (ROOT, repo_dir)
    import subprocess
    import tempfile
    outs = {}
    with tempfile.TemporaryDirectory() as td:
        np.save(os.path.join(td, "x.npy"), x)
        for flag in ("0", "1"):   # the switch is read once per process
            env = dict(os.environ, B200_ENGINE_C3PAIR=flag)
            r = subprocess.run([sys.executable, "-c", code, os.path.join(td, "x.npy"), os.path.join(td, f"y{flag}.npy")],
                               capture_output=True, text=True, timeout=300, env=env, cwd=ROOT)
            assert r.returncode == 0, r.stderr[-2000:]
            outs[flag] = np.load(os.path.join(td, f"y{flag}.npy"))
    assert outs["0"].shape == (16, 1000)
    assert np.array_equal(outs["0"], outs["1"])


def test_streaming_dense_layer_kernel_is_bit_identical(pkg, repo_dir, monkeypatch):
    """B200_ENGINE_LAYERFUSE: dense layers of the 56x56 / 28x28 blocks (1x1 conv -> 128 -> 3x3 conv -> 32) run as ONE
    streaming kernel (kernels_dense_stream.cu: bottleneck in shared memory, conv1 A operand through tensor memory, tap-stacked
    conv2, TMA-stored output) instead of the conv1x1_tma + conv3x3_tma pair.  Same arithmetic, same accumulation order: the logits
    must be bit-identical, for whole and partial strips (batch sizes that split images across CTAs differently).  "1" fuses every
    layer of both blocks, the default ("auto") the layers the streaming kernel is faster for - 56x56 with one K chunk - from 32
    samples on."""
    monkeypatch.setenv("B200_ENGINE_PRECISION", "fp8")
    monkeypatch.setenv("B200_ENGINE_MAX_BATCH", "32")
    monkeypatch.setenv("B200_ENGINE_DEVICES", "0")
    monkeypatch.setenv("B200_ENGINE_INSTANCES", "1")
    x = synth.to_model_input(synth.synthetic_images_u8(13, start=5200))
    x32 = np.concatenate([x, x, x])[:32]
    outs, launches = {}, {}
    for flag in ("0", "1", "auto"):
        monkeypatch.setenv("B200_ENGINE_LAYERFUSE", flag)
        mgr = pkg.InferenceManager(repo_dir)
        try:
            mgr.load_model("densenet_onnx")
            m = mgr.get_model("densenet_onnx")
            res = {}
            for n in (13, 1, 6, 32):
                xin = x32 if n == 32 else x[:n]
                n0 = pkg.kernel_launch_count()
                res[n] = m.infer([pkg.TensorData("data_0", xin)], [pkg.OutputConfig("fc6_1", [n, 1000])])[0].data.copy()
                launches[(flag, n)] = pkg.kernel_launch_count() - n0
            outs[flag] = res
        finally:
            mgr.shutdown()
    for n in (13, 1, 6, 32):
        assert launches[("1", n)] == launches[("0", n)] - 18, launches   # 6 + 12 layer pairs became one launch each
        assert launches[("auto", n)] == launches[("0", n)] - (3 if n >= 32 else 0), launches  # block 1, Cin = 64 / 96 / 128
        assert np.array_equal(outs["0"][n], outs["1"][n]), n
        assert np.array_equal(outs["0"][n], outs["auto"][n]), n


def test_request_coalescer_batches_concurrent_callers(pkg, repo_dir, monkeypatch):
    """concurrent batch-1 callers (what gin + the reference's /infer handler produce) are executed as
    a few batches, every caller still gets exactly its own result, mixed request sizes and both input kinds included."""
    monkeypatch.setenv("B200_ENGINE_PRECISION", "bf16")
    monkeypatch.setenv("B200_ENGINE_MAX_BATCH", "16")
    monkeypatch.setenv("B200_ENGINE_COALESCE_US", "3000")
    u8 = synth.synthetic_images_u8(24, start=7000)
    x = synth.to_model_input(u8)
    mgr = pkg.InferenceManager(repo_dir)
    try:
        mgr.load_model("densenet_onnx")
        m = mgr.get_model("densenet_onnx")
        want = np.concatenate([m.infer([pkg.TensorData("data_0", x[i:i + 12])], [pkg.OutputConfig("fc6_1", [12, 1000])])[0].data
                               for i in (0, 12)])      # 12 > COALESCE_MAX_REQUEST: executed directly
        b0, r0 = m.coalesce_stats()
        assert (b0, r0) == (0, 0)
        errs = []

        def worker(i):
            try:
                for rep in range(3):
                    k = 1 + (i + rep) % 2                      # requests of 1 or 2 images
                    lo = (i + 5 * rep) % (24 - k)
                    if (i + rep) % 3 == 0:                     # every third request sends raw uint8 pixels
                        t = pkg.TensorData("data_0", np.ascontiguousarray(u8[lo:lo + k]), pkg.DataType.UINT8)
                    else:
                        t = pkg.TensorData("data_0", x[lo:lo + k])
                    y = m.infer([t], [pkg.OutputConfig("fc6_1", [k, 1000])])[0].data
                    if np.abs(y - want[lo:lo + k]).max() / np.abs(want).max() > 1e-5:
                        errs.append((i, rep, "mismatch"))
            except Exception as e:  # noqa: BLE001
                errs.append((i, repr(e)))

        ts = [threading.Thread(target=worker, args=(i,)) for i in range(24)]
        [t.start() for t in ts]
        [t.join() for t in ts]
        assert not errs, errs[:3]
        batches, requests = m.coalesce_stats()
        assert requests == 24 * 3 and batches < requests / 2, (batches, requests)
        assert m.get_stats().inference_count == 2 + 72
    finally:
        mgr.shutdown()


def test_device_queries_and_vector_add(pkg):
    assert pkg.is_cuda_available() and pkg.get_device_count() >= 1
    info = pkg.get_device_info(0)
    assert info.startswith("Device 0: ") and "(Compute Capability 10." in info
    mem = pkg.get_memory_info(0)
    assert mem.total > 100e9 and mem.free <= mem.total and mem.used == mem.total - mem.free
    assert pkg.kernel_launch_count() >= 0


def test_multi_gpu_batch_sharding_matches_single_gpu(pkg, repo_dir, monkeypatch):
    """Replicated weights, contiguous batch split, no collective: a batch served by G replicas must give the per-image results
    of one replica.  On a 1-GPU box the scheduler runs over TWO replicas of device 0 (B200_ENGINE_DEVICES=0,0): same split, same
    persistent per-GPU workers, same join - only the silicon is shared."""
    g = pkg.get_device_count()
    multi = "all" if g >= 2 else "0,0"
    g = max(g, 2)
    monkeypatch.setenv("B200_ENGINE_PRECISION", "bf16")
    monkeypatch.setenv("B200_ENGINE_MAX_BATCH", "16")
    monkeypatch.setenv("B200_ENGINE_MIN_SHARD", "8")
    n = 8 * g + 3
    x = synth.to_model_input(synth.synthetic_images_u8(n, start=5000))
    outs = {}
    for devs in ("0", multi):
        monkeypatch.setenv("B200_ENGINE_DEVICES", devs)
        mgr = pkg.InferenceManager(repo_dir)
        try:
            mgr.load_model("densenet_onnx")
            m = mgr.get_model("densenet_onnx")
            outs[devs] = m.infer([pkg.TensorData("data_0", x)], [pkg.OutputConfig("fc6_1", [n, 1000])])[0].data.copy()
            if devs == multi:
                assert f"{g} GPU replica(s)" in m.get_metadata().description
                # small requests rotate over replicas and still agree
                for i in range(2 * g):
                    y = m.infer([pkg.TensorData("data_0", x[i:i + 1])], [pkg.OutputConfig("fc6_1", [1, 1000])])[0].data
                    assert np.array_equal(y, outs["0"][i:i + 1])
                # concurrent split requests share the persistent workers
                errs = []

                def worker():
                    y = m.infer([pkg.TensorData("data_0", x)], [pkg.OutputConfig("fc6_1", [n, 1000])])[0].data
                    if not np.array_equal(y, outs["0"]):
                        errs.append("mismatch")
                ths = [threading.Thread(target=worker) for _ in range(4)]
                [t.start() for t in ths]
                [t.join() for t in ths]
                assert not errs and m.faulted_replicas() == 0
        finally:
            mgr.shutdown()
    assert np.array_equal(outs["0"], outs[multi])     # same kernels, same per-image arithmetic: bit-identical


def test_a_faulted_replica_is_dropped_from_the_shard_set(pkg, repo_dir, monkeypatch):
    """"a failed GPU replica should be dropped from the shard set, not crash".  Replica 1 is made to raise a
    CUDA error on every run (test hook B200_ENGINE_FAULT_REPLICA): its shard moves to a healthy replica, the request succeeds with
    the same results, and later requests no longer use it."""
    g = pkg.get_device_count()
    monkeypatch.setenv("B200_ENGINE_DEVICES", "all" if g >= 2 else "0,0")
    monkeypatch.setenv("B200_ENGINE_PRECISION", "bf16")
    monkeypatch.setenv("B200_ENGINE_MAX_BATCH", "16")
    monkeypatch.setenv("B200_ENGINE_MIN_SHARD", "4")
    n = 8 * max(g, 2)
    x = synth.to_model_input(synth.synthetic_images_u8(n, start=5100))
    mgr = pkg.InferenceManager(repo_dir)
    try:
        mgr.load_model("densenet_onnx")
        want = mgr.get_model("densenet_onnx").infer([pkg.TensorData("data_0", x)], [pkg.OutputConfig("fc6_1", [n, 1000])])[0].data.copy()
        mgr.unload_model("densenet_onnx")
        monkeypatch.setenv("B200_ENGINE_FAULT_REPLICA", "1")
        mgr.load_model("densenet_onnx")
        m = mgr.get_model("densenet_onnx")
        assert m.faulted_replicas() == 0
        got = m.infer([pkg.TensorData("data_0", x)], [pkg.OutputConfig("fc6_1", [n, 1000])])[0].data
        assert np.array_equal(got, want) and m.faulted_replicas() == 1
        for i in range(4):   # small requests round-robin over the remaining replicas only
            y = m.infer([pkg.TensorData("data_0", x[i:i + 1])], [pkg.OutputConfig("fc6_1", [1, 1000])])[0].data
            assert np.array_equal(y, want[i:i + 1])
        assert m.faulted_replicas() == 1
    finally:
        mgr.shutdown()


# ---------------------------------------------------------------------------------------------------------------------
# The C++ API, exercised by native programs
def _run_native(cmd, monkeypatch, timeout=300):
    import subprocess
    env = dict(os.environ, B200_ENGINE_DEVICES="0", B200_ENGINE_PRECISION="fp32", B200_ENGINE_MAX_BATCH="4")
    return subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)


def test_reference_onnx_test_program_runs_unmodified(repo_dir, monkeypatch):
    """`/root/reference/test/onnx_test.cpp` compiled UNMODIFIED against include/ + the .so (oracle/build_ref_tests.py): it loads
    models/test_model through `inference::Model`, feeds [[1, 1, 1]] (onnx_test.cpp:92) and prints the output tensor."""
    exe = os.path.join(ROOT, "oracle", "_ref", "onnx_test")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/onnx_test was not built (the reference mount was absent at build time)")
    r = _run_native([exe, os.path.join(repo_dir, "test_model", "1")], monkeypatch)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "Model loaded successfully!" in r.stdout and "Inference completed successfully!" in r.stdout
    assert "Inputs: input" in r.stdout and "Outputs: output" in r.stdout and "Shape: [1, 2]" in r.stdout
    vals = [ln for ln in r.stdout.splitlines() if ln.startswith("First 2 values:")]
    got = np.array([float(t) for t in vals[0].split(":")[1].split()], np.float32)
    with open(os.path.join(ROOT, "tests", "golden", "test_model_kat.json")) as fh:
        kat = [v for v in json.load(fh)["vectors"] if v["input"] == [[1.0, 1.0, 1.0]]][0]
    np.testing.assert_allclose(got, np.asarray(kat["output"], np.float32).ravel(), rtol=2e-5)   # printed with 6 digits
    assert "Inference Count: 1" in r.stdout and "Model unloaded successfully!" in r.stdout


def test_reference_cuda_test_program_runs_unmodified(monkeypatch):
    exe = os.path.join(ROOT, "oracle", "_ref", "cuda_test")
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/cuda_test was not built (the reference mount was absent at build time)")
    r = _run_native([exe], monkeypatch)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    assert "CUDA Available: Yes" in r.stdout and "Compute Capability 10." in r.stdout
    assert "Vector addition succeeded" in r.stdout and r.stdout.count("1 + 1 = 2 ✓") == 5


def test_cpp_inference_manager_run_inference_and_tensor(repo_dir, densenet_path, tmp_path, monkeypatch):
    """tests/cpp/api_test.cpp: `inference::Tensor` (SetData/GetData/Reshape/copy/toGPU/toCPU, reference model.cpp:30-436) and
    `inference::InferenceManager` (LoadModel / LoadModelAsync / RunInference / UnloadModelAsync, reference
    inference_manager.cpp:283-384,674-707) at run time, results against the golden vector and the oracle."""
    exe = os.path.join(ROOT, "build", "api_test")
    assert os.path.exists(exe), "build/api_test missing: run __graft_entry__.build()"
    n = 2
    x = synth.to_model_input(synth.synthetic_images_u8(n, start=7000))
    f = tmp_path / "x.f32"
    x.astype(np.float32).tofile(str(f))
    r = _run_native([exe, repo_dir, str(f), str(n)], monkeypatch)
    fails = [ln for ln in r.stdout.splitlines() if ln.startswith("FAIL")]
    assert r.returncode == 0 and not fails, (fails, r.stdout[-1500:], r.stderr[-1500:])
    assert r.stdout.count("ok tensor.") >= 13 and "ok mgr.run_inference" in r.stdout and "ok dense.run_inference" in r.stdout
    kat = [ln.split()[1:] for ln in r.stdout.splitlines() if ln.startswith("KAT ")][0]
    np.testing.assert_allclose(np.array(kat, np.float64), [-1.6748662, 2.0709436], rtol=2e-6)
    ref = OnnxOracle(densenet_path).run({"data_0": x})[0]
    dense = [ln.split()[1:] for ln in r.stdout.splitlines() if ln.startswith("DENSE ")]
    assert len(dense) == n
    for i, arg, mx in dense:
        assert int(arg) == int(ref[int(i)].argmax())
        assert abs(float(mx) - float(ref[int(i)].max())) / np.abs(ref).max() < 1e-3


def test_on_device_softmax_top5_matches_a_host_sort(pkg, repo_dir, monkeypatch):
    """B200ModelInferTopK = ModelInfer + softmax + top-k on the GPU; the reference's Go handler sorts
    all 1000 probabilities per request (server/main.go:744-786).  Same classes and scores as sorting the logits of ModelInfer."""
    monkeypatch.setenv("B200_ENGINE_PRECISION", "fp32")
    monkeypatch.setenv("B200_ENGINE_MAX_BATCH", "4")          # 7 images through a 4-image arena: two shards
    u8 = synth.synthetic_images_u8(7, start=8000)
    x = synth.to_model_input(u8)
    mgr = pkg.InferenceManager(repo_dir)
    try:
        mgr.load_model("densenet_onnx")
        m = mgr.get_model("densenet_onnx")
        logits = m.infer([pkg.TensorData("data_0", x)], [pkg.OutputConfig("fc6_1", [7, 1000])])[0].data.astype(np.float64)
        prob = np.exp(logits - logits.max(1, keepdims=True))
        prob /= prob.sum(1, keepdims=True)
        order = np.argsort(-logits, axis=1, kind="stable")[:, :5]
        for inp in (pkg.TensorData("data_0", x), pkg.TensorData("data_0", np.ascontiguousarray(u8), pkg.DataType.UINT8)):
            classes, scores = m.infer_topk([inp], k=5, softmax=True)
            assert classes.shape == (7, 5) and np.array_equal(classes, order)
            np.testing.assert_allclose(scores, np.take_along_axis(prob, order, 1), rtol=2e-5, atol=1e-9)
        classes, raw = m.infer_topk([pkg.TensorData("data_0", x)], k=3, softmax=False)
        assert np.array_equal(classes, order[:, :3])
        np.testing.assert_allclose(raw, np.take_along_axis(logits, order[:, :3], 1), rtol=1e-6)
        with pytest.raises(pkg.EngineError, match="Invalid parameters"):
            m.infer_topk([pkg.TensorData("data_0", x)], k=65)
    finally:
        mgr.shutdown()


# ---------------------------------------------------------------------------------------------------------------------
# Exact-arithmetic cases: operands chosen so that every product, every partial sum and every stored value is representable
# in e4m3 (small multiples of 0.5, per-channel weight scales that are powers of two times 1/448).  Reduced precision then
# loses nothing, so ALL THREE modes must reproduce the fp32 oracle BIT FOR BIT - a mis-indexed K-tail column, a wrong tap
# shift, a swapped swizzle piece or a dropped split term shows up as a hard mismatch instead of hiding inside a tolerance.
def _unit_bn(c, prefix):
    # scale = gamma / sqrt(var + eps) == 1.0 and shift == 0 after rounding to fp32/f16/bf16
    return {prefix + ".g": np.ones(c), prefix + ".b": np.zeros(c), prefix + ".m": np.zeros(c), prefix + ".v": np.full(c, 1.0 - 1e-5)}


def _exact_case(case, tmp_path, rng):
    if case.startswith("conv1x1_"):
        cin, cout, hw = int(case.split("_")[1]), 128, 9
        w = np.zeros((cout, cin, 1, 1))
        for c in range(cin):                               # every input channel feeds exactly one output channel
            w[c % cout, c, 0, 0] = rng.choice([-1.0, -0.5, 0.5, 1.0])
        inits = {**_unit_bn(cin, "bn"), "w": w, "b": rng.choice([-0.5, 0.0, 0.5], cout)}
        nodes = [_bn_node("x", "bn", "n"), onnx_lite.Node("Relu", ["n"], ["r"]), _conv_node("r", "w", "c", 1, bias="b"),
                 onnx_lite.Node("Relu", ["c"], ["y"])]
        x = np.zeros((3, cin, hw, hw), np.float32)
        for n in range(3):                                 # two non-zero channels per pixel, in different residue classes mod 128
            for y in range(hw):
                for xx in range(hw):
                    c0 = int(rng.integers(0, cin))
                    c1 = (c0 + 1 + int(rng.integers(0, min(cin, cout) - 1))) % cin
                    if c1 % cout == c0 % cout:
                        c1 = (c1 + 1) % cin
                    x[n, c0, y, xx] = rng.choice([1.0, 2.0])
                    x[n, c1, y, xx] = rng.choice([-2.0, 1.0, 2.0])   # negative: removed by the prologue ReLU
        shp, out = (cin, hw, hw), (cout, hw, hw)
    elif case == "conv3x3":
        cin, cout, hw = 128, 32, 14
        w = np.zeros((cout, cin, 3, 3))
        for c in range(cin):
            for t in range(9):
                w[c % cout, c, t // 3, t % 3] = rng.choice([-1.0, -0.5, 0.5, 1.0])
        inits = {"w": w}
        nodes = [_conv_node("x", "w", "y", 3, 1, 1)]
        x = np.zeros((3, cin, hw, hw), np.float32)
        for n in range(3):                                 # one non-zero channel per pixel; the 9 neighbours of any output pixel
            for y in range(hw):                            # carry 9 different channels (base + 3*dy + dx) -> <= 1 product per output
                for xx in range(hw):
                    x[n, (17 * n + 3 * y + xx) % cin, y, xx] = rng.choice([1.0, 2.0])
        shp, out = (cin, hw, hw), (cout, hw, hw)
    elif case == "transition":
        cin, cout, hw = 256, 128, 14
        w = np.zeros((cout, cin, 1, 1))
        for c in range(cin):
            w[c % cout, c, 0, 0] = rng.choice([-1.0, 1.0])
        inits = {**_unit_bn(cin, "bn"), "w": w}
        nodes = [_bn_node("x", "bn", "n"), onnx_lite.Node("Relu", ["n"], ["r"]), _conv_node("r", "w", "c", 1),
                 onnx_lite.Node("AveragePool", ["c"], ["y"], {"kernel_shape": [2, 2], "strides": [2, 2], "pads": [0, 0, 0, 0]})]
        x = np.zeros((3, cin, hw, hw), np.float32)
        for n in range(3):                                 # eight channels per 2x2 window (distinct mod 128), values 0 / 4 -> averages 0..4
            for y in range(hw):
                for xx in range(hw):
                    for k in range(8):
                        x[n, (5 * n + 7 * (y // 2) + (xx // 2) + 16 * k) % cin, y, xx] = rng.choice([0.0, 4.0])
        shp, out = (cin, hw, hw), (cout, hw // 2, hw // 2)
    else:
        raise KeyError(case)
    path = _graph_case(tmp_path, "exact_" + case, nodes, inits, shp, "y", out)
    return path, "exact_" + case, x, out


@pytest.mark.parametrize("precision", ["fp32", "bf16", "fp8"])
@pytest.mark.parametrize("case", ["conv1x1_64", "conv1x1_96", "conv1x1_352", "conv1x1_608", "conv1x1_1024", "conv3x3", "transition"])
def test_exactly_representable_operands_give_bit_identical_results(pkg, tmp_path, monkeypatch, case, precision):
    rng = np.random.default_rng(abs(hash("exact" + case)) % 2**31)
    path, name, x, out = _exact_case(case, tmp_path, rng)
    want = OnnxOracle(path).run({"x": x})[0]
    assert np.abs(want).max() <= 7.5 and np.array_equal(want, np.round(want * 4) / 4) and np.count_nonzero(want) > want.size // 50
    got = _serve(pkg, str(tmp_path), name, {"x": x}, {"y": (len(x),) + tuple(out)}, precision, monkeypatch)[0]
    bad = np.argwhere(got != want)
    assert len(bad) == 0, f"{case}/{precision}: {len(bad)} of {want.size} elements differ, first at {bad[:4].tolist()}: " \
                          f"{got[tuple(bad[0])]} vs {want[tuple(bad[0])]}"
