"""Builds par_raytracer_b200/librt_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

-fmad=false is part of the arithmetic contract: every hit-deciding float operation
must round exactly like the reference's x86-64 SSE2 build (no fused multiply-add).
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "rt_api.cu")
OUT = os.path.join(HERE, "librt_b200.so")
DEPS = [os.path.join(HERE, "csrc", f) for f in ("rt_api.cu", "rt_build.cuh", "rt_common.cuh", "rt_rng.cuh", "rt_raygen.cuh", "rt_groups.cuh", "rt_preprocess.cuh", "rt_shade.cuh", "rt_trace.cuh")]
DEPS.append(os.path.join(HERE, "..", "include", "rt_b200.h"))

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fno-strict-aliasing",
    "-shared",
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    cmd = [nvcc] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-o", OUT, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building librt_b200.so")
    if verbose:
        sys.stderr.write(r.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
