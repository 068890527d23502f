"""Deterministic synthetic inputs for the DenseNet-121 path.

Images are uint8 HWC, 224x224x3: uniform noise from `numpy.random.default_rng(seed)`
blended with low-frequency structure so activations are not pure noise.  Conversion
to the model's input follows the reference client (`client/test_client.py:186-194`):
`/255`, HWC -> CHW, float32.
"""
from __future__ import annotations

import numpy as np


def synthetic_images_u8(n: int, seed: int = 0, start: int = 0, hw: int = 224) -> np.ndarray:
    """Images [start, start+n) of the fixed synthetic set; each image depends only on
    (seed, index) so any slice of the set can be regenerated independently."""
    out = np.empty((n, hw, hw, 3), dtype=np.uint8)
    yy, xx = np.meshgrid(np.linspace(0, 1, hw, dtype=np.float32),
                         np.linspace(0, 1, hw, dtype=np.float32), indexing="ij")
    for i in range(n):
        rng = np.random.default_rng([seed, start + i])
        noise = rng.integers(0, 256, size=(hw, hw, 3)).astype(np.float32)
        f = rng.uniform(0.5, 6.0, size=(3, 2)).astype(np.float32)
        ph = rng.uniform(0, 2 * np.pi, size=(3, 2)).astype(np.float32)
        amp = rng.uniform(40, 110, size=3).astype(np.float32)
        base = rng.uniform(60, 190, size=3).astype(np.float32)
        low = np.stack([base[c] + amp[c] * np.sin(2 * np.pi * f[c, 0] * yy + ph[c, 0])
                        * np.cos(2 * np.pi * f[c, 1] * xx + ph[c, 1]) for c in range(3)], axis=-1)
        img = 0.65 * low + 0.35 * noise
        out[i] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return out


def clustered_images_u8(n: int, seed: int = 0, start: int = 0, clusters: int = 50, hw: int = 224) -> np.ndarray:
    """The evaluation set of the reduced-precision gates: image i belongs to cluster i % clusters.  A cluster fixes the
    low-frequency structure (what the fixture's classifier keys on, tools/make_densenet_onnx.py); every image adds its own
    pixel noise, brightness and contrast jitter, so members of a cluster differ the way photos of one class do."""
    out = np.empty((n, hw, hw, 3), dtype=np.uint8)
    yy, xx = np.meshgrid(np.linspace(0, 1, hw, dtype=np.float32),
                         np.linspace(0, 1, hw, dtype=np.float32), indexing="ij")
    for i in range(n):
        idx = start + i
        crng = np.random.default_rng([seed, 7_000_000 + idx % clusters])
        f = crng.uniform(0.5, 6.0, size=(3, 2)).astype(np.float32)
        ph = crng.uniform(0, 2 * np.pi, size=(3, 2)).astype(np.float32)
        amp = crng.uniform(40, 110, size=3).astype(np.float32)
        base = crng.uniform(60, 190, size=3).astype(np.float32)
        rng = np.random.default_rng([seed, 8_000_000 + idx])
        noise = rng.integers(0, 256, size=(hw, hw, 3)).astype(np.float32)
        gain = rng.uniform(0.9, 1.1, size=3).astype(np.float32)
        shift = rng.uniform(-8, 8, size=3).astype(np.float32)
        low = np.stack([base[c] + shift[c] + gain[c] * amp[c] * np.sin(2 * np.pi * f[c, 0] * yy + ph[c, 0])
                        * np.cos(2 * np.pi * f[c, 1] * xx + ph[c, 1]) for c in range(3)], axis=-1)
        img = 0.65 * low + 0.35 * noise
        out[i] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
    return out


def to_model_input(u8_hwc: np.ndarray) -> np.ndarray:
    """uint8 [N,H,W,3] -> float32 [N,3,H,W] in [0,1] (client/test_client.py:186-194)."""
    x = u8_hwc.astype(np.float32) / np.float32(255.0)
    return np.ascontiguousarray(x.transpose(0, 3, 1, 2))
