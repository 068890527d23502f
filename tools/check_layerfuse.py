"""Bit-exactness check of the streaming dense-layer kernel (kernels_dense_stream.cu) against the conv1x1 + conv3x3 kernel pair:
runs the DenseNet-121 fixture in e4m3 mode with B200_ENGINE_LAYERFUSE=0 and =1 (one subprocess each) and compares the logits.
usage (under gpurun): python tools/check_layerfuse.py [batch]"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if len(sys.argv) > 2 and sys.argv[1] == "--child":
    import numpy as np
    import __graft_entry__ as ge
    batch = int(sys.argv[2])
    os.environ["B200_ENGINE_PRECISION"] = "fp8"
    os.environ.setdefault("B200_ENGINE_DEVICES", "0")
    os.environ["B200_ENGINE_MAX_BATCH"] = str(max(batch, 8))
    pkg = ge.load_package()
    ge.ensure_fixtures()
    from tools import synth
    mgr = pkg.InferenceManager(os.path.join(ROOT, "models"))
    mgr.load_model("densenet_onnx")
    m = mgr.get_model("densenet_onnx")
    x = synth.to_model_input(synth.synthetic_images_u8(batch, start=900))
    out = m.infer([pkg.TensorData("data_0", x)], [pkg.OutputConfig("fc6_1", [batch, 1000])])
    y = np.asarray(out[0].data, dtype=np.float32).reshape(batch, -1)
    np.save(sys.argv[3], y)
    ms = m.forward_device(batch, 5, True)
    print(f"LAYERFUSE={os.environ.get('B200_ENGINE_LAYERFUSE')} bs{batch}: {np.median(ms):.3f} ms/forward", flush=True)
    mgr.shutdown()
    sys.exit(0)

import numpy as np
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
outs = []
for flag in ("0", "1"):
    path = f"/tmp/layerfuse_{flag}.npy"
    env = dict(os.environ, B200_ENGINE_LAYERFUSE=flag)
    r = subprocess.run([sys.executable, __file__, "--child", str(batch), path], env=env, capture_output=True, text=True, timeout=600)
    print(r.stdout.strip()[-400:])
    if r.returncode != 0:
        print("child failed:", r.stderr[-3000:])
        sys.exit(1)
    outs.append(np.load(path))
a, b = outs
diff = np.abs(a - b)
print(f"max|logit| {np.abs(a).max():.4f}  max abs diff {diff.max():.3e}  mismatching {int((a != b).sum())} of {a.size}  top1 equal {(a.argmax(1) == b.argmax(1)).mean():.3f}")
print("BIT-IDENTICAL" if np.array_equal(a, b) else "DIFFERENT")
