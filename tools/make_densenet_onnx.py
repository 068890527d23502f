"""Generate models/densenet_onnx/1/model.onnx (the blob is missing from the reference mount,
`/root/reference/.MISSING_LARGE_BLOBS`) — fixture recipe, deterministic, offline.

Topology: torchvision DenseNet-121 (growth 32, bn_size 4, blocks 6/12/24/16), random
weights under torch.manual_seed(0); BatchNorm affine parameters perturbed and running
statistics calibrated on the fixed synthetic image set; classifier rescaled so logits
have a non-degenerate spread.  I/O names follow the reference's
`models/densenet_onnx/1/config.json:7,15` (`data_0` / `fc6_1`).

Two emitters:
  * default: torch's legacy TorchScript ONNX exporter (opset 12) with its `onnx`-dependent
    post-step stubbed — gives a real exporter's graph (Identity nodes,
    Conv+BN folding where the exporter does it, Flatten+Gemm tail);
  * `--emitter direct`: hand-serialised graph via tools/onnx_lite (no exporter involved).
"""
from __future__ import annotations

import argparse
import hashlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tools import onnx_lite, synth  # noqa: E402

DEFAULT_OUT = os.path.join(ROOT, "models", "densenet_onnx", "1", "model.onnx")


def build_module(calib_images: int = 32):
    import torch
    import torchvision

    torch.manual_seed(0)
    m = torchvision.models.densenet121(weights=None)
    g = torch.Generator().manual_seed(1)
    for mod in m.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            with torch.no_grad():
                mod.weight.copy_(torch.empty_like(mod.weight).uniform_(0.6, 1.4, generator=g))
                mod.bias.copy_(torch.empty_like(mod.bias).normal_(0.0, 0.15, generator=g))
            mod.momentum = None  # cumulative average during calibration
            mod.reset_running_stats()
    x = torch.from_numpy(synth.to_model_input(synth.synthetic_images_u8(calib_images, seed=0)))
    m.train()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    with torch.no_grad():
        for i in range(0, calib_images, 16):
            m(x[i:i + 16])
    m.eval()
    synthesize_classifier(m, g)
    return m


CLUSTERS = 50             # clusters of tools/synth.clustered_images_u8 the classifier knows
TOP5_GAINS = (1.0, 0.9, 0.8, 0.7, 0.6)
LOGIT_SCALE = 16.0


def class_permutation():
    return np.random.default_rng(12345).permutation(1000)


def synthesize_classifier(m, g) -> None:
    """An untrained network has no decision structure: its logits are a random projection whose top-1/top-5 margins are far
    below any reduced-precision noise floor, so "top-5 agreement" measures nothing (hard parts).  Training
    is what creates margins; here the classifier is SYNTHESISED to the same effect: with c_k the mean penultimate feature of
    cluster k of the synthetic evaluation set and u_k the ridge-regularised dual basis of the centred cluster means
    (<c_j - mu, u_k> ~ delta_jk), the five classes of cluster k get the rows  LOGIT_SCALE * gain_j * u_k  (+ a small independent
    random component), the remaining 500 classes small random rows.  An image of cluster k then has a clear top-1 (margin
    0.15 * LOGIT_SCALE), an ordered top-5 and a gap of 0.4 * LOGIT_SCALE to the rest - the shape of a trained classifier's
    output - while its features still come out of 120 random convolutions, which is what the kernels are tested on."""
    import torch
    rng = np.random.default_rng(2024)
    feats = []
    with torch.no_grad():
        for k0 in range(0, CLUSTERS, 10):
            imgs = np.concatenate([synth.clustered_images_u8(10, start=k0 + rep * CLUSTERS, clusters=CLUSTERS) for rep in range(3)])
            x = torch.from_numpy(synth.to_model_input(imgs))
            f = torch.nn.functional.adaptive_avg_pool2d(torch.relu(m.features(x)), 1).flatten(1).numpy().astype(np.float64)
            feats.append((f[:10] + f[10:20] + f[20:]) / 3.0)
    C = np.concatenate(feats)                         # [CLUSTERS, 1024] cluster means
    mu = C.mean(0)
    G = C - mu
    lam = 1e-2 * np.trace(G @ G.T) / CLUSTERS
    U = np.linalg.solve(G @ G.T + lam * np.eye(CLUSTERS), G)   # rows u_k
    W = np.zeros((1000, 1024))
    unorm = np.linalg.norm(U, axis=1).mean()
    perm = class_permutation()
    for k in range(CLUSTERS):
        for j, gain in enumerate(TOP5_GAINS):
            r = rng.normal(0, 1, 1024)
            W[perm[5 * k + j]] = LOGIT_SCALE * (gain * U[k] + 0.02 * unorm * r / np.linalg.norm(r))
    for c in range(5 * CLUSTERS, 1000):
        r = rng.normal(0, 1, 1024)
        W[perm[c]] = LOGIT_SCALE * 0.05 * unorm * r / np.linalg.norm(r)
    b = -W @ mu + rng.normal(0, 0.1, 1000)
    with torch.no_grad():
        m.classifier.weight.copy_(torch.from_numpy(W.astype(np.float32)))
        m.classifier.bias.copy_(torch.from_numpy(b.astype(np.float32)))


def export_legacy(m, path: str) -> None:
    import torch
    from torch.onnx._internal.torchscript_exporter import onnx_proto_utils

    onnx_proto_utils._add_onnxscript_fn = lambda model_bytes, custom_opsets: model_bytes
    buf = io.BytesIO()
    dummy = torch.zeros(1, 3, 224, 224)
    torch.onnx.export(m, (dummy,), buf, opset_version=12, input_names=["data_0"],
                      output_names=["fc6_1"], dynamo=False,
                      dynamic_axes={"data_0": {0: "N"}, "fc6_1": {0: "N"}})
    with open(path, "wb") as fh:
        fh.write(buf.getvalue())


def export_direct(m, path: str) -> None:
    """Hand-built graph: un-folded BatchNormalization everywhere (123 BN nodes)."""
    g = onnx_lite.Graph(name="densenet121")
    sd = {k: v.detach().numpy() for k, v in m.state_dict().items()}
    nodes = g.nodes

    def init(name, arr):
        g.initializers[name] = np.ascontiguousarray(arr, dtype=np.float32)
        return name

    def bn(x, prefix, out):
        nodes.append(onnx_lite.Node("BatchNormalization",
                                    [x, init(prefix + ".weight", sd[prefix + ".weight"]),
                                     init(prefix + ".bias", sd[prefix + ".bias"]),
                                     init(prefix + ".running_mean", sd[prefix + ".running_mean"]),
                                     init(prefix + ".running_var", sd[prefix + ".running_var"])],
                                    [out], {"epsilon": 1e-5, "momentum": 0.9}))
        return out

    def relu(x, out):
        nodes.append(onnx_lite.Node("Relu", [x], [out]))
        return out

    def conv(x, prefix, out, k, s, p):
        nodes.append(onnx_lite.Node("Conv", [x, init(prefix + ".weight", sd[prefix + ".weight"])], [out],
                                    {"dilations": [1, 1], "group": 1, "kernel_shape": [k, k],
                                     "pads": [p, p, p, p], "strides": [s, s]}))
        return out

    x = conv("data_0", "features.conv0", "conv0", 7, 2, 3)
    x = relu(bn(x, "features.norm0", "norm0"), "relu0")
    nodes.append(onnx_lite.Node("MaxPool", [x], ["pool0"], {"kernel_shape": [3, 3], "pads": [1, 1, 1, 1],
                                                             "strides": [2, 2], "ceil_mode": 0}))
    x = "pool0"
    for bi, nl in enumerate((6, 12, 24, 16), start=1):
        feats = [x]
        for li in range(1, nl + 1):
            p = f"features.denseblock{bi}.denselayer{li}"
            if len(feats) > 1:
                cat = f"b{bi}l{li}.cat"
                nodes.append(onnx_lite.Node("Concat", list(feats), [cat], {"axis": 1}))
            else:
                cat = feats[0]
            t = relu(bn(cat, p + ".norm1", f"b{bi}l{li}.n1"), f"b{bi}l{li}.r1")
            t = conv(t, p + ".conv1", f"b{bi}l{li}.c1", 1, 1, 0)
            t = relu(bn(t, p + ".norm2", f"b{bi}l{li}.n2"), f"b{bi}l{li}.r2")
            t = conv(t, p + ".conv2", f"b{bi}l{li}.c2", 3, 1, 1)
            feats.append(t)
        x = f"b{bi}.out"
        nodes.append(onnx_lite.Node("Concat", list(feats), [x], {"axis": 1}))
        if bi < 4:
            p = f"features.transition{bi}"
            t = relu(bn(x, p + ".norm", f"t{bi}.n"), f"t{bi}.r")
            t = conv(t, p + ".conv", f"t{bi}.c", 1, 1, 0)
            x = f"t{bi}.pool"
            nodes.append(onnx_lite.Node("AveragePool", [t], [x], {"kernel_shape": [2, 2], "strides": [2, 2],
                                                                  "pads": [0, 0, 0, 0], "ceil_mode": 0}))
    x = relu(bn(x, "features.norm5", "norm5"), "relu5")
    nodes.append(onnx_lite.Node("GlobalAveragePool", [x], ["gap"]))
    nodes.append(onnx_lite.Node("Flatten", ["gap"], ["flat"], {"axis": 1}))
    nodes.append(onnx_lite.Node("Gemm", ["flat", init("classifier.weight", sd["classifier.weight"]),
                                         init("classifier.bias", sd["classifier.bias"])], ["fc6_1"],
                                {"alpha": 1.0, "beta": 1.0, "transB": 1}))
    g.inputs = [onnx_lite.ValueInfo("data_0", onnx_lite.FLOAT, ["N", 3, 224, 224])]
    g.outputs = [onnx_lite.ValueInfo("fc6_1", onnx_lite.FLOAT, ["N", 1000])]
    onnx_lite.save(onnx_lite.Model(g, ir_version=7, opset=12), path)


def ensure(path: str = DEFAULT_OUT, emitter: str = "legacy", quiet: bool = False) -> str:
    if os.path.exists(path) and os.path.getsize(path) > 30_000_000:
        return path
    os.makedirs(os.path.dirname(path), exist_ok=True)
    m = build_module()
    tmp = path + ".tmp"
    if emitter == "legacy":
        try:
            export_legacy(m, tmp)
        except Exception as e:  # exporter internals moved: fall back to the direct emitter
            if not quiet:
                print(f"[make_densenet_onnx] legacy exporter failed ({e!r}); using direct emitter")
            export_direct(m, tmp)
    else:
        export_direct(m, tmp)
    os.replace(tmp, path)
    if not quiet:
        with open(path, "rb") as fh:
            digest = hashlib.sha256(fh.read()).hexdigest()
        print(f"[make_densenet_onnx] wrote {path} ({os.path.getsize(path)} B, sha256 {digest[:16]}…)")
    return path


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=DEFAULT_OUT)
    ap.add_argument("--emitter", choices=["legacy", "direct"], default="legacy")
    ap.add_argument("--force", action="store_true")
    a = ap.parse_args()
    if a.force and os.path.exists(a.out):
        os.remove(a.out)
    ensure(a.out, a.emitter)
