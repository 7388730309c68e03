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


def test_config3_one_million_textured_triangles():
    """BASELINE config 3 at full size: 1,048,352-triangle height field in 529 OBJ-style groups, 8 materials with diffuse /
    ambient maps, bump maps and an alpha mask, 1920x1080, 128 spp. The group hierarchy is the REFERENCE's (BuildHierarchy run on the GPU
    by rt_build_group_hierarchy, bit-identical to the host build), so the equal-t tie-break order is the one the unmodified reference uses."""
    import os
    import tempfile
    from oracle import ref_harness
    sd = scenes.heightfield_scene(724, 724, block=32, size=400.0, amp=20.0, textured=True, tex_size=512, hierarchy="defer")
    assert sd.n_triangles == 2 * 724 * 724 and sd.tangents is not None
    scenes.use_reference_hierarchy(sd)
    assert len(sd.spheres) == 2 * sd.n_groups - 1
    S = api.Scene(sd)
    W, H, spp = 1920, 1080, 128
    h = sd.camera_hint
    cam = types.make_camera(h["fov"], W, H, h["position"], h["facing"])
    p = types.default_params(spp=spp)
    img, cnt = S.render(cam, p, W, H)
    flat = img.reshape(-1, 4)
    assert np.all(np.isfinite(flat)) and np.all(flat[:, :3] >= 0) and np.all(flat[:, 3] == 1.0)
    assert W * H * spp <= cnt["ray_count"] <= 16 * W * H * spp
    # (1) the C port on 2,048 pixels spread over the frame at the full 128 spp: identical ray count, colours within tolerance
    ids = (np.arange(2048, dtype=np.uint64) * np.uint64(W * H) // np.uint64(2048) + np.uint64(977)).astype(np.uint32)
    O = oracle.OracleScene(sd)
    ref, _, cnt_o, _ = O.render(cam, p, W, H, pixel_ids=ids, threads=os.cpu_count() or 16)
    sub, cnt_s = S.render_task(cam, p, W, H, pixel_ids=ids)
    assert cnt_s["ray_count"] == cnt_o["ray_count"]
    assert np.allclose(sub, ref, rtol=1e-5, atol=1e-6)
    assert np.array_equal(bits(sub), bits(flat[ids]))
    # (2) the COMPILED REFERENCE (oracle/_ref: OBJ -> ParseOBJ, CalculateTangents, BuildHierarchy, TraceRayColor) on 512 of those pixels at
    #     16 spp. Its hierarchy must be the one the GPU built, and -- scene exported by the reference -- primary hits must agree bit for bit.
    if ref_harness.available():
        d = tempfile.mkdtemp(prefix="config3_")
        scenes.write_obj(sd, d)
        R = ref_harness.get()
        rs = R.load_scene(d)
        assert np.array_equal(rs.spheres.view(np.uint8), sd.spheres.view(np.uint8)) and np.array_equal(rs.sphere_group, sd.sphere_group)
        R.set_params(p); R.set_lights(sd.lights)
        ids2 = ids[::4]
        ref2, _, cnt_r, _ = R.render_seeded(cam, W, H, ids2, 0, len(ids2), 0, 16, 16, int(p["base_seed"]), threads=os.cpu_count() or 16)
        S2 = api.Scene(rs)                                   # the reference's own arrays (its tangents, its normal maps)
        p16 = types.default_params(spp=16)
        sub2, cnt2 = S2.render_task(cam, p16, W, H, pixel_ids=ids2)
        assert cnt2["ray_count"] == cnt_r["ray_count"]
        assert np.allclose(sub2, ref2, rtol=1e-5, atol=1e-6)
        rr, hr = R.trace_primary(cam, W, H, ids2, 0, len(ids2), 0, 2, int(p["base_seed"]))
        rg, hg = S2.trace_primary(cam, p16, W, H, pixel_ids=ids2, sample_count=2)
        assert rr.tobytes() == rg.tobytes() and hr.tobytes() == hg.tobytes()
        S2.close()
    # alpha-masked / bump-mapped / translucent-free materials all occur among the primary hits
    rays, hits = S.trace_primary(cam, p, W, H, pixel_ids=ids, sample_count=4)
    mats = sd.group_material[sd.sphere_group[hits["object"][hits["hit"] == 1]]]
    assert len(np.unique(mats)) >= 6
    S.close()


def test_config4_ten_million_triangles_tile_of_a_4k_frame():
    """BASELINE config 4 at full size: 9,999,392 triangles (4,900 groups x 2,048), 3840x2160, 256 spp. One rank's share of an
    8-way interleaved tile partition is rendered at the full sample count; pruning is checked against brute force."""
    sd = scenes.heightfield_scene(2236, 2236, block=32, size=400.0, amp=20.0, textured=False, hierarchy="defer")
    assert 9_900_000 < sd.n_triangles < 10_100_000 and sd.n_groups == 4900
    scenes.use_reference_hierarchy(sd)          # the reference's BuildHierarchy over 4,900 groups (minutes on the host, < 1 s here)
    S = api.Scene(sd)
    info = S.hierarchy_info()
    assert info["triangles"] == sd.n_triangles and info["depth"] <= 62
    W, H, spp = 3840, 2160, 256
    h = sd.camera_hint
    cam = types.make_camera(h["fov"], W, H, h["position"], h["facing"])
    p = types.default_params(spp=spp)
    tiles = dist.tile_partition(W, H, 3, 8, 32)
    assert len(tiles) in range(W * H // 8 - 32 * 32 * 8, W * H // 8 + 32 * 32 * 8)
    full = tiles[: 1005 * 1024].reshape(1005, 1024)          # rank 3's 1005 complete 32x32 tiles (tile rows 0..66)
    part = full[::25][:40].reshape(-1).copy()                # 40 of them, spread over the frame: 10.5 M samples at 256 spp
    img, cnt = S.render_task(cam, p, W, H, pixel_ids=part)
    assert np.all(np.isfinite(img)) and np.all(img[:, :3] >= 0) and np.all(img[:, 3] == 1.0)
    assert len(part) * spp <= cnt["ray_count"] <= 16 * len(part) * spp
    again, cnt2 = S.render_task(cam, p, W, H, pixel_ids=part[::-1].copy())
    assert np.array_equal(bits(again[::-1]), bits(img)) and cnt2["ray_count"] == cnt["ray_count"]   # order of the list is irrelevant
    # 256 pixels spread over those tiles against the oracle at 32 spp (the reference scans whole 2,048-triangle groups per leaf), and 16 of
    # them at the full 256 spp
    import os
    ids = part[:: len(part) // 256][:256]
    O = oracle.OracleScene(sd)
    p32 = types.default_params(spp=32)
    ref, _, cnt_o, _ = O.render(cam, p32, W, H, pixel_ids=ids, threads=os.cpu_count() or 16)
    sub, cnt_s = S.render_task(cam, p32, W, H, pixel_ids=ids)
    assert cnt_s["ray_count"] == cnt_o["ray_count"] and np.allclose(sub, ref, rtol=1e-5, atol=1e-6)
    ids_f = ids[::16]
    ref, _, cnt_o, _ = O.render(cam, p, W, H, pixel_ids=ids_f, threads=os.cpu_count() or 16)
    sub, cnt_s = S.render_task(cam, p, W, H, pixel_ids=ids_f)
    assert cnt_s["ray_count"] == cnt_o["ray_count"] and np.allclose(sub, ref, rtol=1e-5, atol=1e-6)
    # hierarchy == brute force over all 10 M triangles for 1,500 primary rays
    rays, hits = S.trace_primary(cam, p, W, H, pixel_ids=part[::27][:1500], sample_count=1)
    hb, _ = S.trace_rays(p, rays, api.RT_TRACE_BRUTE)
    assert hits.tobytes() == hb.tobytes() and hits["hit"].mean() > 0.5
    S.close()
