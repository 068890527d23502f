"""Per-step device times of one DenseNet-121 forward, grouped by layer family (run under gpurun).
usage: python tools/prof_steps.py <precision> <batch> [iters]   -> prints ms per forward and a grouped table; JSON to gpurun_out/"""
import collections
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import __graft_entry__ as ge  # noqa: E402

precision, batch = sys.argv[1], int(sys.argv[2])
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 10
os.environ["B200_ENGINE_PRECISION"] = precision
os.environ.setdefault("B200_ENGINE_DEVICES", "0")
os.environ["B200_ENGINE_MAX_BATCH"] = str(batch)
pkg = ge.load_package()
ge.ensure_fixtures()
from tools import synth  # noqa: E402

mgr = pkg.InferenceManager(os.path.join(ROOT, "models"))
mgr.load_model("densenet_onnx")
m = mgr.get_model("densenet_onnx")
nimg = min(batch, 16)
imgs = synth.to_model_input(synth.synthetic_images_u8(nimg, start=500))
x = np.concatenate([imgs] * ((batch + nimg - 1) // nimg))[:batch]
m.stage_input(pkg.TensorData("data_0", x))
m.forward_device(batch, 3, True)
ms = m.forward_device(batch, iters, True)
print(f"{precision} bs{batch}: median {np.median(ms):.3f} ms/forward = {batch / np.median(ms) * 1e3:.0f} img/s (min {ms.min():.3f})")
prof = m.profile_steps(batch, 3)
g = collections.OrderedDict()
for p in prof:
    if p["kind"] == "conv":
        key = f"conv{p['R']}x{p['R']} H{p['H']}" + (" trans" if "trans" in p["name"] else "") + ("" if p.get("umma") else " simt")
    else:
        key = p["kind"]
    a = g.setdefault(key, [0, 0.0, 0.0, 0.0])
    a[0] += 1; a[1] += p["ms"]; a[2] += p["flops"]; a[3] += p["bytes"]
tot = sum(a[1] for a in g.values())
for k, a in g.items():
    tf = a[2] / (a[1] * 1e-3) / 1e12 if a[1] > 0 else 0
    gb = a[3] / (a[1] * 1e-3) / 1e9 if a[1] > 0 else 0
    print(f"  {k:28s} n={a[0]:3d} {a[1]*1e3:9.1f} us {100*a[1]/tot:5.1f}%  {tf:7.1f} TFLOP/s {gb:7.0f} GB/s")
print(f"  sum of steps {tot:.3f} ms")
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
tag = os.environ.get("PROF_TAG", "prof")
with open(os.path.join(ROOT, "gpurun_out", f"{tag}_steps_{precision}_bs{batch}.json"), "w") as fh:
    json.dump(prof, fh)
mgr.shutdown()
