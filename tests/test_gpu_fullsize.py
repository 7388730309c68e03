"""BASELINE.json config 2 at FULL size (63,490 triangles, 1920x1080, 64 spp) through size-independent
properties, plus an oracle comparison on a pixel subset at the full resolution and sample count."""
import numpy as np
import pytest

from conftest import bits
from oracle import oracle
from par_raytracer_b200 import api, dist, scenes, types

pytestmark = pytest.mark.gpu


def test_config2_full_size_properties():
    sd = scenes.spheres_plane_scene()
    S = api.Scene(sd)
    info = S.hierarchy_info()
    assert info["triangles"] == 63490 and info["depth"] <= 62
    W, H, spp = 1920, 1080, 64
    h = sd.camera_hint
    cam = types.make_camera(h["fov"], W, H, h["position"], h["facing"])
    p = types.default_params(spp=spp)
    img, cnt = S.render(cam, p, W, H)
    flat = img.reshape(-1, 4)
    assert np.all(np.isfinite(flat)) and np.all(flat[:, :3] >= 0) and np.all(flat[:, 3] == 1.0)   # raytracer.cpp:555-558
    n_primary = W * H * spp
    assert n_primary <= cnt["ray_count"] <= 16 * n_primary
    # (1) oracle on every 211th pixel, full spp: identical ray count, colours within tolerance
    ids = np.arange(0, W * H, 211, dtype=np.uint32)
    O = oracle.OracleScene(sd)
    ref, _, cnt_o, _ = O.render(cam, p, W, H, pixel_ids=ids, threads=16)
    sub, cnt_s = S.render_task(cam, p, W, H, pixel_ids=ids)
    assert cnt_s["ray_count"] == cnt_o["ray_count"]
    assert np.allclose(sub, ref, rtol=1e-5, atol=1e-6)
    # (2) the subset render equals the same pixels of the full-frame render bit for bit (partition invariance)
    assert np.array_equal(bits(sub), bits(flat[ids]))
    # (3) 4-way interleaved tile partition covers the frame and reproduces it exactly
    acc = np.zeros_like(flat)
    total_rays = 0
    for r in range(4):
        tids = dist.tile_partition(W, H, r, 4)
        part, c = S.render_task(cam, p, W, H, pixel_ids=tids)
        acc[tids] += part
        total_rays += int(c["ray_count"])
    assert np.array_equal(bits(acc), bits(flat)) and total_rays == int(cnt["ray_count"])
    # (4) hierarchy pruning is conservative: hierarchy == brute force on the frame's own primary rays (subset)
    rays, hits = S.trace_primary(cam, p, W, H, pixel_ids=ids[:4000], sample_count=2)
    hb, _ = S.trace_rays(p, rays, api.RT_TRACE_BRUTE)
    assert hits.tobytes() == hb.tobytes()
    S.close()
