"""The C-ABI library loads and exports every symbol include/rt_b200.h declares; struct layouts match the
header; compute entry points fail loudly without a GPU (there is no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import HAS_GPU, ROOT
from par_raytracer_b200 import api, build, cabi, scenes, types


@pytest.fixture(scope="module")
def lib():
    build.build()          # nvcc cross-compiles sm_100a without a GPU
    return api.load_library()


def header_functions():
    src = open(os.path.join(ROOT, "include", "rt_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z_0-9]+)\s*\(", src)))


def test_exports_every_declared_symbol(lib):
    names = header_functions()
    assert set(names) == set(api.EXPORTS), (names, api.EXPORTS)
    for n in names:
        assert hasattr(lib, n), f"librt_b200.so does not export {n}"
    assert lib.rt_abi_version() == 1


def test_struct_layouts_match_header():
    # sizes the reference structs have (measured with g++ 13): Camera 64, Ray 24,
    # BoundingSphere 24, LightSource 48, Texture 24
    assert types.CAMERA.itemsize == 64 and types.RAY.itemsize == 24 and types.BSPHERE.itemsize == 24
    assert types.LIGHT.itemsize == 48 and C.sizeof(cabi.RtTexture) == 24
    assert types.COUNTERS.itemsize == 24 and types.HIT.itemsize == 52 and types.PARAMS.itemsize == 48
    assert types.PARAMS.fields["base_seed"][1] == 40
    assert C.sizeof(cabi.RtMaterial) == types.MATERIAL.itemsize == 96


def test_library_is_the_cuda_build(lib):
    """The product must be device code for sm_100a, not a host shim."""
    import subprocess
    out = subprocess.run(["cuobjdump", "-lelf", api.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_error_paths_without_compute(lib):
    h = C.c_void_p()
    assert lib.rt_scene_create(None, 0, C.byref(h)) == -1           # RT_ERR_ARG
    assert b"null" in lib.rt_last_error()
    sd = scenes.spheres_plane_scene(grid=1, nu=8, nv=4)
    desc, keep = cabi.make_scene_desc(sd)
    bad = sd.idx_positions.copy(); bad[5] = 10 ** 6
    desc.idx_positions = bad.ctypes.data
    rc = lib.rt_scene_create(C.addressof(desc), 0, C.byref(h))
    assert rc in (-1, -2)                                            # bad index (GPU box) or no device (here)
    if not HAS_GPU:
        with pytest.raises(api.RtError, match="no CPU fallback"):
            api.Scene(sd)
        with pytest.raises(api.RtError):
            api.rng_kat(1, 4)


def test_multi_gpu_entry_points_fail_loudly_without_compute(lib):
    """The combine layer's argument checks are host code; without a device every constructor must refuse (no CPU fallback)."""
    h = C.c_void_p()
    assert lib.rt_device_count() >= 0
    assert lib.rt_comm_create(0, 0, None, 0, C.byref(h)) == -1                      # RT_ERR_ARG: 0 ranks
    assert lib.rt_comm_create(2, 5, (C.c_uint8 * 128)(), 0, C.byref(h)) == -1      # rank out of range
    assert lib.rt_render_combined(None, None, None, None, 4, 4, 0, 32, 0, 0, None, None, None, None) == -1
    assert lib.rt_render_multi(None, None, 0, None, None, 4, 4, 0, 32, 0, None, None, None, None) == -1
    assert lib.rt_comm_rank(None) == -1 and lib.rt_comm_size(None) == 0
    n = C.c_uint32(0)
    assert lib.rt_partition_tiles(64, 64, 32, 0, 2, None, C.byref(n)) == 0 and n.value == 2 * 32 * 32      # host-only: works anywhere
    assert lib.rt_partition_tiles(64, 64, 0, 0, 2, None, C.byref(n)) == -1
    if not HAS_GPU:
        assert lib.rt_device_count() == 0
        assert lib.rt_comm_create(1, 0, None, 0, C.byref(h)) == -2                  # RT_ERR_CUDA: no device
        assert b"no CPU fallback" in lib.rt_last_error()
        hs = (C.c_void_p * 1)()
        assert lib.rt_comm_create_local(1, None, hs) == -2
    # draw-counter bound of rt_params (host check): the defaults pass, a recursion tree that could take > 65535 draws is refused up front
    p = types.default_params(spp=1)
    p["bounce_depth"] = 12; p["reflection_samples"] = 4; p["spec_samples"] = 4
    sd = scenes.spheres_plane_scene(grid=1, nu=8, nv=4)
    if HAS_GPU:
        S = api.Scene(sd)
        with pytest.raises(api.RtError, match="65535"):
            S.render(types.make_camera(60.0, 4, 4, (0, 3, 8), (0, -0.3, -1)), p, 4, 4)
        S.close()


def test_product_does_not_import_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may touch oracle/."""
    pkg = os.path.join(ROOT, "par_raytracer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the oracle harness", "").replace("Oracle", "").lower() or f in ("scenes.py", "types.py"), \
                    f"{f} mentions the oracle"


def test_quantisation_grid_placement(lib):
    """rt_quant_grid is host code: the 15-bit grid of the default node format must cover the scene bounds for any finite
    bounds -- tiny, huge, degenerate (zero extent) and far from the origin, where `mid` rounds coarser than the ideal step --
    with the planes the kernel will decode from the two floats per axis; non-finite bounds must be refused (ok = 0)."""
    import ctypes as C
    f3 = C.c_float * 3

    def grid(lo, hi):
        step, mid, ok = f3(), f3(), C.c_int(-1)
        assert lib.rt_quant_grid(f3(*lo), f3(*hi), step, mid, C.byref(ok)) == 0
        return np.array(step[:], np.float32), np.array(mid[:], np.float32), ok.value

    rng = np.random.default_rng(7)
    cases = [((-1, -1, -1), (1, 1, 1)), ((0, 0, 0), (0, 0, 0)), ((5, 5, 5), (5, 6, 5)), ((-1e-30, 0, 0), (1e-30, 1e-38, 0)),
             ((-3e37, -1, 0), (3e37, 1, 1e-20)), ((1e7, -1e7, 3e6), (1e7 + 1, -1e7 + 2, 3e6 + 0.5)), ((16777216, 0, 0), (16777218, 1, 1))]
    for _ in range(2000):
        c = rng.normal(size=3) * 10.0 ** rng.uniform(-3, 8)
        e = np.abs(rng.normal(size=3)) * 10.0 ** rng.uniform(-6, 6, size=3)
        cases.append((tuple(np.float32(c - e)), tuple(np.float32(c + e))))
    for lo, hi in cases:
        lo32, hi32 = np.array(lo, np.float32), np.array(hi, np.float32)
        if not np.all(np.isfinite(lo32) & np.isfinite(hi32)):
            continue
        step, mid, ok = grid(lo32, hi32)
        assert ok == 1, (lo, hi)
        assert np.all(step > 0) and np.all(np.isfinite(step)) and np.all(np.isfinite(mid))
        base = mid.astype(np.float64) + 32768.0 * step.astype(np.float64)            # plane 0, exactly as the build computes it
        top = base + 32767.0 * step.astype(np.float64)                                # plane 32767
        assert np.all(base <= lo32.astype(np.float64)) and np.all(top >= hi32.astype(np.float64)), (lo, hi, step, mid)
        ext = hi32.astype(np.float64) - lo32.astype(np.float64)
        near = np.abs(lo32.astype(np.float64)) < 10.0 * np.maximum(ext, 1e-30)        # not far from the origin (float ulp of `mid` << step): the grid must be tight
        assert np.all(step.astype(np.float64)[near] <= np.maximum(ext[near] / 32764.0, 1e-30) * 1.0001)
    for lo, hi in [((0, 0, 0), (np.inf, 1, 1)), ((np.nan, 0, 0), (1, 1, 1)), ((-np.inf, 0, 0), (np.inf, 1, 1))]:
        assert grid(lo, hi)[2] == 0
    assert lib.rt_quant_grid(None, None, None, None, None) != 0


def test_header_is_plain_c_and_cxx11(tmp_path):
    """The boundary is a C ABI: include/rt_b200.h must compile on its own as C99 and as C++11 (the reference's dialect, build.sh:6),
    with the struct sizes the reference's own headers have."""
    import shutil
    import subprocess
    inc = os.path.join(ROOT, "include")
    body = ('#include "rt_b200.h"\n'
            'typedef char camera_is_64[sizeof(rt_camera) == 64 ? 1 : -1];\n'
            'typedef char ray_is_24[sizeof(rt_ray) == 24 ? 1 : -1];\n'
            'typedef char bsphere_is_24[sizeof(rt_bsphere) == 24 ? 1 : -1];\n'
            'typedef char hit_is_52[sizeof(rt_hit) == 52 ? 1 : -1];\n'
            'int main(void) { int (*f)(void) = rt_abi_version; (void)f; return 0; }\n')
    for cc, flags, name in (("gcc", ["-std=c99", "-pedantic-errors"], "t.c"), ("g++", ["-std=c++11"], "t.cpp")):
        if not shutil.which(cc):
            pytest.skip(f"{cc} not installed")
        src = tmp_path / name
        src.write_text(body)
        r = subprocess.run([cc] + flags + ["-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(src)], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
