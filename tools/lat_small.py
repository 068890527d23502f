"""Device-resident p50 latency of one e4m3 forward at batch 1..64 (A/B runs of a kernel-selection switch, e.g. B200_ENGINE_LAYERFUSE).
usage (under gpurun): B200_ENGINE_LAYERFUSE=0 python tools/lat_small.py"""
import os, sys
ROOT='/root/repo'; sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as ge
os.environ["B200_ENGINE_PRECISION"]="fp8"; os.environ["B200_ENGINE_DEVICES"]="0"; os.environ["B200_ENGINE_MAX_BATCH"]="64"
pkg = ge.load_package(); ge.ensure_fixtures()
from tools import synth
mgr = pkg.InferenceManager(os.path.join(ROOT,"models")); mgr.load_model("densenet_onnx"); m = mgr.get_model("densenet_onnx")
x = synth.to_model_input(synth.synthetic_images_u8(16, start=500)); x = np.concatenate([x]*4)
m.stage_input(pkg.TensorData("data_0", x))
for b in (1, 2, 4, 8, 16, 32, 64):
    m.forward_device(b, 5, False)
    ms = m.forward_device(b, 50, False)
    print(f"LAYERFUSE={os.environ.get('B200_ENGINE_LAYERFUSE')} bs{b}: p50 {np.median(ms):.4f} ms")
mgr.shutdown()
