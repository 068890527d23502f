"""Builds the reference's own acceptance programs, UNMODIFIED, against this repo's headers and shared library:

    /root/reference/test/onnx_test.cpp  ->  oracle/_ref/onnx_test     (Model::Load / Infer on models/test_model, input [[1,1,1]])
    /root/reference/test/cuda_test.cpp  ->  oracle/_ref/cuda_test     (cuda_utils: device info + VectorAdd)

TEST INFRASTRUCTURE ONLY.  The sources are compiled where they lie (never copied into the repo); the binaries land in
oracle/_ref/ (git-ignored, but they travel to the GPU box) and are run by tests/test_gpu_parity.py.  Called from
__graft_entry__.build() whenever /root/reference is present; on the GPU box only the prebuilt binaries exist."""
from __future__ import annotations

import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TEST = "/root/reference/test"
OUT = os.path.join(ROOT, "oracle", "_ref")
LIBDIR = os.path.join(ROOT, "gpu-ai-inference-server_b200", "lib")


def build(quiet: bool = True) -> list[str]:
    if not os.path.isdir(REF_TEST):
        return []
    os.makedirs(OUT, exist_ok=True)
    built = []
    for name in ("onnx_test", "cuda_test"):
        src, exe = os.path.join(REF_TEST, name + ".cpp"), os.path.join(OUT, name)
        lib = os.path.join(LIBDIR, "libinference_engine.so")
        if os.path.exists(exe) and os.path.getmtime(exe) > max(os.path.getmtime(src), os.path.getmtime(lib)):
            built.append(exe)
            continue
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"), src, "-o", exe, "-L", LIBDIR,
                               "-linference_engine", "-Wl,-rpath,$ORIGIN/../../gpu-ai-inference-server_b200/lib"])
        built.append(exe)
        if not quiet:
            print(f"[build] {exe}")
    return built


if __name__ == "__main__":
    build(quiet=False)
