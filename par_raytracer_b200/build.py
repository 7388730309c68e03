"""Builds par_raytracer_b200/librt_b200.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

-fmad=false is part of the arithmetic contract: every hit-deciding float operation
must round exactly like the reference's x86-64 SSE2 build (no fused multiply-add).

The library is four translation units (csrc/rt_scene.cu, rt_render.cu, rt_loadtime.cu, rt_comm.cu), compiled in
parallel and linked into one shared object. NCCL is NOT a link-time dependency: rt_comm.cu binds libnccl.so.2 at run
time (the copy already loaded by the host process if there is one), so the library loads on a box without NCCL.
"""
from __future__ import annotations

import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
UNITS = ["rt_scene", "rt_render", "rt_loadtime", "rt_comm"]
OUT = os.path.join(HERE, "librt_b200.so")
OBJ_DIR = os.path.join(HERE, "build")


def _deps():
    return glob.glob(os.path.join(CSRC, "*.cu")) + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        [os.path.join(HERE, "..", "include", "rt_b200.h")]


NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false", "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-fno-strict-aliasing",
]


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = os.environ.get("NVCC", "nvcc")
    os.makedirs(OBJ_DIR, exist_ok=True)
    extra = os.environ.get("RT_B200_NVCC_EXTRA", "").split()

    def compile_unit(u):
        obj = os.path.join(OBJ_DIR, u + ".o")
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, os.path.join(CSRC, u + ".cu")]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return u, obj, r

    with ThreadPoolExecutor(max_workers=len(UNITS)) as ex:
        results = list(ex.map(compile_unit, UNITS))
    objs = []
    for u, obj, r in results:
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed compiling {u}.cu")
        if verbose:
            sys.stderr.write(r.stderr)
        objs.append(obj)
    r = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", OUT] + objs + ["-ldl"], capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking librt_b200.so")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
