"""The drop-in on the UNMODIFIED reference: oracle/_ref/libref_dropin.so is the reference's main.cpp + the product's
host shim (par_raytracer_b200/host/rt_render_shim.hpp) + librt_b200.so. It parses the OBJ, builds tangents and the
hierarchy with the reference's own code and only replaces Render() by RenderB200(). The frame must match the
reference's CPU render of the same scene (oracle/_ref/libref_harness.so, same per-(pixel, sample) seeding)."""
import ctypes as C
import os
import tempfile

import numpy as np
import pytest

from conftest import ROOT
from oracle import ref_harness
from par_raytracer_b200 import scenes, types

pytestmark = pytest.mark.gpu
DROPIN = os.path.join(ROOT, "oracle", "_ref", "libref_dropin.so")


@pytest.mark.skipif(not (os.path.exists(DROPIN) and ref_harness.available()), reason="oracle/_ref not built (needs /root/reference)")
def test_reference_main_with_render_replaced():
    sd = scenes.spheres_plane_scene(grid=2, nu=20, nv=10, textured=True)
    d = tempfile.mkdtemp(prefix="dropin_")
    scenes.write_obj(sd, d)
    W, H, spp = 96, 64, 4
    hint = sd.camera_hint
    pos = np.asarray(hint["position"], np.float32); fac = np.asarray(hint["facing"], np.float32)
    seed = types.DEFAULT_BASE_SEED
    lib = C.CDLL(DROPIN)
    out = np.zeros((W * H, 4), np.float32)
    rays = C.c_ulonglong(0)
    rc = lib.dropin_render(d.encode(), C.c_uint32(W), C.c_uint32(H), C.c_float(hint["fov"]), pos.ctypes.data_as(C.c_void_p),
                           fac.ctypes.data_as(C.c_void_p), C.c_uint32(spp), C.c_uint64(seed), out.ctypes.data_as(C.c_void_p), C.byref(rays))
    assert rc == 0
    R = ref_harness.get()
    R.load_scene(d)
    cam = R.make_camera(hint["fov"], W, H, hint["position"], hint["facing"])
    ref, _, cnt, _ = R.render_seeded(cam, W, H, None, 0, W * H, 0, spp, spp, seed, threads=4)
    assert rays.value == int(cnt["ray_count"])
    assert np.allclose(out, ref, rtol=1e-5, atol=1e-6)
    assert np.all(out[:, 3] == 1.0)
