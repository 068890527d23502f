"""Runs exactly `reps` forwards of DenseNet-121 (no L2 flush, CUDA graph off so every kernel is a plain launch)
for profiling under ncu.  usage: python tools/ncu_forward.py <precision> <batch> [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge

precision, batch = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
os.environ["B200_ENGINE_PRECISION"] = precision
os.environ.setdefault("B200_ENGINE_DEVICES", "0")
os.environ["B200_ENGINE_MAX_BATCH"] = str(batch)
os.environ["B200_ENGINE_GRAPHS"] = "0"
pkg = ge.load_package(); ge.ensure_fixtures()
from tools import synth
mgr = pkg.InferenceManager(os.path.join(ROOT, "models"))
mgr.load_model("densenet_onnx")
m = mgr.get_model("densenet_onnx")
nimg = min(batch, 16)
imgs = synth.to_model_input(synth.synthetic_images_u8(nimg, start=500))
x = np.concatenate([imgs] * ((batch + nimg - 1) // nimg))[:batch]
m.stage_input(pkg.TensorData("data_0", x))
ms = m.forward_device(batch, reps, False)
print("forward ms", ms)
mgr.shutdown()
