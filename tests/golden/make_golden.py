"""Regenerate tests/golden/densenet_logits.npz from the oracle (run in the build container)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle.onnx_oracle import OnnxOracle  # noqa: E402
from tools import make_fixtures, synth  # noqa: E402

make_fixtures.ensure_all()
path = os.path.join(ROOT, "models", "densenet_onnx", "1", "model.onnx")
start, n = 1000, 8
x = synth.to_model_input(synth.synthetic_images_u8(n, start=start))
o = OnnxOracle(path)
y32 = o.run({"data_0": x})[0]
y64 = o.run({"data_0": x}, dtype=torch.float64)[0].astype(np.float64)
import hashlib
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "densenet_logits.npz"), start=start, n=n, logits_fp32=y32,
                    logits_fp64=y64, model_sha256=hashlib.sha256(open(path, "rb").read()).hexdigest())
print("fp32 vs fp64 max rel:", np.abs(y32 - y64).max() / np.abs(y64).max())
