"""Host-side logic: scene generators, the per-(pixel, sample) seed contract, partitions, OBJ export."""
import os
import tempfile

import numpy as np
import pytest

from oracle import oracle, ref_harness
from par_raytracer_b200 import dist, scenes, types


def test_seed_contract_matches_oracle_primary_rays():
    sd = scenes.spheres_plane_scene(grid=1, nu=8, nv=4)
    O = oracle.OracleScene(sd)
    cam = types.make_camera(60.0, 16, 8, (0, 3, 8), (0, -0.3, -1))
    p = types.default_params(spp=1)
    rays, _ = O.trace_primary(cam, p, 16, 8, None, 0, 128, 3, 2, want_hits=False)
    k, s = 37, 1
    seed = types.sample_seed(int(p["base_seed"]), k, 3 + s)
    j = oracle.rng_float(seed, 2, True)                      # jy first, then jx
    x, y = k % 16, k // 16
    want = oracle.camera_rays(cam, np.array([[np.float32(x) + j[1] * np.float32(0.5), np.float32(y) + j[0] * np.float32(0.5)]], np.float32))
    assert rays[k * 2 + s].tobytes() == want[0].tobytes()


def test_scene_generators_are_valid_and_one_sided():
    for sd in (scenes.spheres_plane_scene(grid=2, nu=12, nv=6, textured=True), scenes.heightfield_scene(16, 12, block=4)):
        sd.validate()
        a = sd.positions[sd.idx_positions[0::3]]; b = sd.positions[sd.idx_positions[1::3]]; c = sd.positions[sd.idx_positions[2::3]]
        n = np.cross(b - a, c - a)
        assert np.all(np.linalg.norm(n, axis=1) > 0), "degenerate triangle emitted"
        vn = sd.normals[sd.idx_normals[0::3]]
        assert np.all(np.einsum("ij,ij->i", n, vn) > 0), "winding disagrees with the shading normal (one-sided test!)"
    sd = scenes.spheres_plane_scene()
    assert sd.n_triangles == 16 * 3968 + 2 and sd.n_groups == 17


def test_group_hierarchy_is_a_valid_bounding_hierarchy():
    sd = scenes.heightfield_scene(32, 32, block=4)
    s = sd.spheres
    assert len(s) == 2 * sd.n_groups - 1
    for i in range(len(s)):
        if s[i]["c0"] and s[i]["c1"]:
            for c in (s[i]["c0"], s[i]["c1"]):
                d = np.linalg.norm(s[c]["center"].astype(np.float64) - s[i]["center"].astype(np.float64))
                assert d + s[c]["radius"] <= s[i]["radius"] * (1 + 1e-6)          # bsphere.cpp:359-362
        else:
            g = sd.sphere_group[i]
            p = sd.positions[sd.idx_positions[sd.group_first[g]:sd.group_first[g + 1]]].astype(np.float64)
            assert np.all(np.linalg.norm(p - s[i]["center"], axis=1) <= s[i]["radius"])   # bsphere.cpp:371-375


def test_partitions_cover_the_frame_exactly_once():
    W, H = 100, 37
    for world in (1, 2, 3, 8):
        ids = np.concatenate([dist.tile_partition(W, H, r, world, tile=16) for r in range(world)])
        assert np.array_equal(np.sort(ids), np.arange(W * H, dtype=np.uint32))
        spans = [dist.range_partition(W, H, r, world) for r in range(world)]
        assert sum(c for _, c in spans) == W * H and spans[0][0] == 0
        ss = [dist.sample_partition(67, r, world) for r in range(world)]
        assert sum(c for _, c in ss) == 67 and all(ss[i][0] + ss[i][1] == ss[i + 1][0] for i in range(world - 1))
    # the reference's own split (main.cpp:313-317)
    assert dist.range_partition(720, 480, 5, 64) == (5 * 5400, 5400)


def test_c_tile_partition_matches_numpy_statement():
    """rt_partition_tiles (host-only C, the partition rt_render_combined uses) against a plain numpy statement of the rule:
    tile t = ty * tiles_x + tx goes to rank t % world, pixels inside a tile row by row; ragged right / bottom tiles clipped."""
    from par_raytracer_b200 import api

    def numpy_tiles(width, height, rank, world, tile):
        tiles_x = (width + tile - 1) // tile
        tiles_y = (height + tile - 1) // tile
        ids = []
        for t in range(rank, tiles_x * tiles_y, world):
            ty, tx = divmod(t, tiles_x)
            xs = np.arange(tx * tile, min(tx * tile + tile, width), dtype=np.uint32)
            ys = np.arange(ty * tile, min(ty * tile + tile, height), dtype=np.uint32)
            ids.append((ys[:, None] * np.uint32(width) + xs[None, :]).reshape(-1))
        return np.concatenate(ids).astype(np.uint32) if ids else np.zeros(0, np.uint32)

    for (W, H, tile, world) in [(100, 37, 16, 3), (1920, 1080, 32, 8), (7, 5, 32, 2), (64, 64, 8, 1), (33, 1, 4, 5)]:
        for r in range(world):
            assert np.array_equal(api.partition_tiles(W, H, r, world, tile), numpy_tiles(W, H, r, world, tile)), (W, H, tile, world, r)
    with pytest.raises(api.RtError):
        api.partition_tiles(10, 10, 2, 2, 4)            # rank out of range
    with pytest.raises(api.RtError):
        api.partition_tiles(10, 10, 0, 1, 0)            # tile 0


def test_png_writer_roundtrip():
    PIL = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    for shape in ((9, 7), (9, 7, 3), (5, 4, 4)):
        img = rng.integers(0, 256, shape).astype(np.uint8)
        path = os.path.join(tempfile.mkdtemp(), "t.png")
        scenes.write_png(path, img)
        assert np.array_equal(np.asarray(PIL.open(path)), img)


@pytest.mark.skipif(not ref_harness.available(), reason="oracle/_ref not built (needs /root/reference)")
def test_obj_export_is_what_the_reference_parses():
    """write_obj -> the reference's ParseOBJ/ParseMTL/CalculateTangents give back the same arrays."""
    sd = scenes.spheres_plane_scene(grid=2, nu=10, nv=5, textured=True)
    d = tempfile.mkdtemp()
    scenes.write_obj(sd, d)
    rs = ref_harness.get().load_scene(d)
    for f in ("positions", "texcoords", "normals", "group_first", "idx_positions", "idx_texcoords", "idx_normals", "group_material"):
        assert np.array_equal(getattr(rs, f), getattr(sd, f)), f
    for k in ("specular_intensity", "index_of_refraction", "alpha", "diffuse_color", "specular_color"):
        assert np.array_equal(rs.materials[k], sd.materials[k]), k
    assert np.abs(rs.tangents - sd.tangents).max() < 1e-5           # numpy restatement of CalculateTangents
    bump_ref = rs.textures[int(rs.materials[2]["bump_texture"])]
    # numpy powf vs glibc powf may disagree by one code value at truncation boundaries
    assert np.abs(bump_ref.texels.astype(int) - sd.textures[1].texels.astype(int)).max() <= 1


def test_bench_roofline_arithmetic_and_inputs():
    """bench.py's roofline numerator is SURVEY 8(d)'s algorithmic bytes per ray; its workloads are BASELINE's configurations; the
    committed traffic file it cites exists and names the dominant kernel."""
    import importlib.util
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec); spec.loader.exec_module(bench)
    assert bench.b_ray(63490) == 776 and bench.b_ray(1048352) == 904 and bench.b_ray(9999392) == 1032
    assert bench.DEFAULT_WORKLOAD == "config3" and bench.WORKLOADS["config3"][:3] == (1920, 1080, 128)
    assert bench.WORKLOADS["config4"][:3] == (3840, 2160, 256) and bench.WORKLOADS["config2"][:3] == (1920, 1080, 64)
    traffic, src = bench.measured_traffic("config3", "k_trace_wave")
    assert src == "profiles/r2_traffic_config3.json" and 1e8 < traffic < 5e9
    assert bench.measured_traffic("config4", "k_trace_wave") == (None, None)
    sd = bench.make_scene("config2")
    cam, params = bench.make_camera_params("config2", sd)
    assert int(params["min_samples"]) == int(params["max_samples"]) == 64 and sd.n_triangles == 63490
    cam1, p1 = bench.make_camera_params("config1", scenes.sponza_standin_scene(detail=0.15))
    assert (int(p1["min_samples"]), int(p1["max_samples"])) == (10, 50)                 # main.cpp:308-309
    assert np.allclose(cam1["position"], (475.0, 250.0, 0.0))                           # main.cpp:426-431


def test_sponza_standin_is_a_valid_one_sided_obj_scene():
    sd = scenes.sponza_standin_scene()               # full detail: coarser tessellations facet the fluted columns against their radial normals
    sd.validate()
    assert 250_000 <= sd.n_triangles <= 320_000
    assert sd.n_groups >= 200 and sd.tangents is not None
    a = sd.positions[sd.idx_positions[0::3]]; b = sd.positions[sd.idx_positions[1::3]]; c = sd.positions[sd.idx_positions[2::3]]
    n = np.cross(b - a, c - a)
    assert np.all(np.linalg.norm(n, axis=1) > 0)
    assert np.all(np.einsum("ij,ij->i", n, sd.normals[sd.idx_normals[0::3]]) > 0)       # winding agrees with the shading normals
    kinds = {k for m in sd.materials for k in ("diffuse_texture", "bump_texture", "alpha_texture") if m[k] >= 0}
    assert kinds == {"diffuse_texture", "bump_texture", "alpha_texture"} and (sd.materials["alpha"] < 1).any()


def test_bench_stdout_carries_only_the_json_line():
    """bench.py's contract is ONE JSON line on stdout; native libraries (NCCL's "NCCL version ..." banner under NCCL_DEBUG) and child processes
    write to file descriptor 1 too, so bench.py points it at stderr and keeps the original descriptor for the line."""
    import subprocess, sys
    from conftest import ROOT
    code = ("import importlib.util, os; spec = importlib.util.spec_from_file_location('bench', %r); b = importlib.util.module_from_spec(spec); "
            "spec.loader.exec_module(b); b._json_only_stdout(); os.write(1, b'NCCL version x\\n'); os.system('echo child'); print('{\"ok\": 1}')"
            % os.path.join(ROOT, "bench.py"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    assert r.stdout == '{"ok": 1}\n'
    assert "NCCL version x" in r.stderr and "child" in r.stderr
