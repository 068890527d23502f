#!/usr/bin/env python3
"""Build libinference_engine.so (sm_100a only) in-tree with nvcc — no cmake/ninja needed.

    python build_engine.py            # incremental build
    python build_engine.py --force    # rebuild everything

Output: gpu-ai-inference-server_b200/lib/libinference_engine.so (git-ignored; travels to the GPU box).
Objects are cached under build/obj/ keyed by source + header mtimes.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "gpu-ai-inference-server_b200")
CSRC = os.path.join(PKG, "csrc")
INC = os.path.join(ROOT, "include")
OBJ = os.path.join(ROOT, "build", "obj")
LIB = os.path.join(PKG, "lib", "libinference_engine.so")

SOURCES = [
    "onnx_wire.cpp", "plan.cpp", "model_repository.cpp", "inference_manager.cpp", "model.cpp",
    "inference_bridge.cpp", "b200_api.cpp", "engine.cu", "cuda_utils.cu", "kernels_simt.cu", "kernels_umma.cu", "kernels_stem.cu", "kernels_conv3x3.cu", "kernels_conv1x1.cu", "kernels_dense.cu", "kernels_dense_stream.cu", "kernels_poolbn.cu", "kernels_f32x3.cu", "tmap.cu",
]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
COMMON = ["-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
          "-Xcompiler", "-fPIC,-fvisibility=default,-Wall,-Wno-unused-function,-pthread",
          "-I", INC, "-I", CSRC, "--expt-relaxed-constexpr", "-diag-suppress", "177"]


def _newest_header() -> float:
    t = 0.0
    for d in (INC, CSRC):
        for f in os.listdir(d):
            if f.endswith((".h", ".hpp", ".cuh")):
                t = max(t, os.path.getmtime(os.path.join(d, f)))
    return t


def _compile(src: str, force: bool, hdr_t: float) -> str:
    obj = os.path.join(OBJ, src + ".o")
    sp = os.path.join(CSRC, src)
    if not force and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(sp), hdr_t):
        return obj
    cmd = [NVCC, *COMMON, *os.environ.get("B200_EXTRA_NVCC", "").split(), "-x", "cu", "-c", sp, "-o", obj]
    if src.endswith(".cu") and os.environ.get("B200_PTXAS_V"):
        cmd[1:1] = ["-Xptxas", "-v"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{r.stdout}\n{r.stderr}")
    if r.stderr.strip():
        sys.stderr.write(r.stderr)
    return obj


def build(force: bool = False, quiet: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    hdr_t = _newest_header()
    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(lambda s: _compile(s, force, hdr_t), SOURCES))
    if force or not os.path.exists(LIB) or any(os.path.getmtime(o) > os.path.getmtime(LIB) for o in objs):
        cmd = [NVCC, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a",
               "-Xcompiler", "-fPIC,-pthread", "-cudart", "static", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
        if not quiet:
            print(f"[build] linked {LIB} ({os.path.getsize(LIB) / 1e6:.1f} MB)")
    elif not quiet:
        print(f"[build] up to date: {LIB}")
    return LIB


def build_tools(quiet: bool = False) -> str:
    """The native load generator (tools/rest_replay.cpp): links against nothing but the C-ABI of the library."""
    out = os.path.join(ROOT, "build", "rest_replay")
    src = os.path.join(ROOT, "tools", "rest_replay.cpp")
    if not (os.path.exists(out) and os.path.getmtime(out) > max(os.path.getmtime(src), os.path.getmtime(LIB))):
        cmd = ["g++", "-O2", "-std=c++17", "-pthread", "-I", INC, src, "-o", out, "-L", os.path.dirname(LIB), "-linference_engine",
               "-Wl,-rpath,$ORIGIN/../gpu-ai-inference-server_b200/lib"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"rest_replay build failed:\n{r.stdout}\n{r.stderr}")
        if not quiet:
            print(f"[build] {out}")
    # run-time test program of the C++ API (inference::Tensor, inference::InferenceManager)
    api = os.path.join(ROOT, "build", "api_test")
    api_src = os.path.join(ROOT, "tests", "cpp", "api_test.cpp")
    if not (os.path.exists(api) and os.path.getmtime(api) > max(os.path.getmtime(api_src), os.path.getmtime(LIB))):
        cmd = ["g++", "-O1", "-std=c++17", "-pthread", "-I", INC, api_src, "-o", api, "-L", os.path.dirname(LIB), "-linference_engine",
               "-Wl,-rpath,$ORIGIN/../gpu-ai-inference-server_b200/lib"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"api_test build failed:\n{r.stdout}\n{r.stderr}")
        if not quiet:
            print(f"[build] {api}")
    return out


if __name__ == "__main__":
    build(force="--force" in sys.argv)
    build_tools()
