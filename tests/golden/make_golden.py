"""Generates tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/libref_harness.so).

Run in the authoring container only (needs /root/reference to build oracle/_ref):
    make -C oracle ref && python tests/golden/make_golden.py
The reference ships no tests or golden vectors, so every vector here is an output
of the compiled reference itself; the C oracle (oracle/rt_oracle.c) and the CUDA path are both checked
against these files. Everything is seeded; re-running reproduces the files bit for bit.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import ref_harness  # noqa: E402
from par_raytracer_b200 import scenes, types  # noqa: E402
from par_raytracer_b200.types import LIGHT, RAY, TextureData  # noqa: E402


def random_rays(rng, n, lo, hi, lift=1.0):
    rays = np.zeros(n, RAY)
    rays["origin"] = (lo + (hi - lo) * rng.random((n, 3)) + np.array([0, lift, 0])).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["direction"] = d.astype(np.float32)
    return rays


def function_vectors(R):
    rng = np.random.default_rng(20170218)
    out = {}
    seeds = np.array([0, 1, 0x835FDD9143716FE3, 0x766B5859CFF0B8AF, 0x201701260526, 0xFFFFFFFFFFFFFFFF], dtype=np.uint64)
    out["rng_seeds"] = seeds
    out["rng_next"] = np.stack([R.rng_next(int(s), 48) for s in seeds])
    out["rng_f01"] = np.stack([R.rng_float(int(s), 48, False) for s in seeds])
    out["rng_f11"] = np.stack([R.rng_float(int(s), 48, True) for s in seeds])
    out["rng_table"] = np.array([R.rng_table(r, 0) for r in range(16)], dtype=np.uint64)

    cam = R.make_camera(60.0, 720, 480, (475.0, 250.0, 0.0), (1.25, -0.5, 1.25))   # main.cpp:426-434 defaults
    out["cam"] = np.frombuffer(cam.tobytes(), np.uint8)
    xy = (rng.random((512, 2)) * np.array([720, 480])).astype(np.float32)
    out["cam_xy"] = xy
    out["cam_rays"] = R.camera_rays(cam, xy).view(np.uint8)

    # triangles: random, plus rays aimed exactly at vertices / edge midpoints / far away
    n = 4096
    tris = (rng.normal(size=(n, 9)) * 3).astype(np.float32)
    rays = np.zeros(n, RAY)
    rays["origin"] = (rng.normal(size=(n, 3)) * 6).astype(np.float32)
    w = rng.dirichlet((1, 1, 1), size=n)
    kind = rng.integers(0, 5, size=n)
    w[kind == 1] = np.array([1.0, 0.0, 0.0])
    w[kind == 2] = np.array([0.5, 0.5, 0.0])
    w[kind == 3] = np.array([0.0, 0.5, 0.5])
    target = (tris.reshape(n, 3, 3) * w[:, :, None]).sum(axis=1)
    target[kind == 4] += rng.normal(size=(int((kind == 4).sum()), 3)) * 2
    d = target - rays["origin"]
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    rays["direction"] = d.astype(np.float32)
    best = np.where(rng.random(n) < 0.5, np.float32(3.4028235e38), (rng.random(n) * 12).astype(np.float32)).astype(np.float32)
    hit, o10 = R.intersect_triangle(rays, tris, best)
    out["tri_rays"] = rays.view(np.uint8); out["tri_tris"] = tris; out["tri_best"] = best
    out["tri_hit"] = hit; out["tri_out"] = o10

    sph = np.concatenate([(rng.normal(size=(n, 3)) * 5), np.abs(rng.normal(size=(n, 1))) * 3 + 0.01], axis=1).astype(np.float32)
    shit, st = R.intersect_sphere(rays, sph)
    out["sph_spheres"] = sph; out["sph_hit"] = shit; out["sph_t"] = st

    idx = np.arange(1024, dtype=np.uint32)
    out["hamm_1024"] = R.hammersley(idx, np.full(1024, 1024, np.uint32))
    ii = rng.integers(0, 64, 256).astype(np.uint32); nn = rng.integers(1, 65, 256).astype(np.uint32)
    out["hamm_i"] = ii; out["hamm_n"] = nn; out["hamm_misc"] = R.hammersley(ii, nn)

    m = 2048
    normal = rng.normal(size=(m, 3)); normal /= np.linalg.norm(normal, axis=1, keepdims=True)
    normal[:8] = np.array([[0, 0, 1], [0, 0, -1], [1, 0, 0], [0, 1, 0], [0, 0.01, 0.99995], [0, -1, 0], [-1, 0, 0], [0.6, 0.0, 0.8]])
    normal = normal.astype(np.float32)
    origin = (rng.normal(size=(m, 3)) * 4).astype(np.float32)
    xi = R.hammersley(rng.integers(0, 1024, m).astype(np.uint32), np.full(m, 1024, np.uint32))
    spec = rng.choice(np.array([0.0, 1.0, 10.0, 40.0, 200.0, 1000.0], np.float32), m).astype(np.float32)
    out["smp_origin"] = origin; out["smp_normal"] = normal; out["smp_xi"] = xi; out["smp_spec"] = spec
    out["smp_diffuse"] = R.diffuse_rays(origin, normal, xi).view(np.uint8)
    out["smp_specular"] = R.specular_rays(origin, normal, spec, xi).view(np.uint8)
    xi0 = np.zeros((m, 2), np.float32)      # Hammersley(0, 1): the spec_samples == 1 case
    out["smp_specular_xi0"] = R.specular_rays(origin, normal, spec, xi0).view(np.uint8)

    inc = rng.normal(size=(m, 3)); inc /= np.linalg.norm(inc, axis=1, keepdims=True)
    ior_a = rng.choice(np.array([1.0, 1.5, 1.33, 2.4], np.float32), m); ior_b = rng.choice(np.array([1.0, 1.5, 1.33, 0.5], np.float32), m)
    out["fr_exit"] = ior_a; out["fr_enter"] = ior_b; out["fr_incident"] = inc.astype(np.float32)
    out["fresnel"] = R.fresnel(ior_a, ior_b, normal, inc.astype(np.float32))

    out["srgb_lut"] = R.srgb_lut()
    uv = (rng.random((1024, 2)) * 6 - 3).astype(np.float32)
    uv[:6] = np.array([[0, 0], [1, 1], [0.999999, 0.5], [-0.25, 2.75], [1e-8, -1e-8], [5.0, -5.0]], np.float32)
    out["tex_uv"] = uv
    for ch in (1, 3, 4):
        tex = TextureData(32, 16, ch, rng.integers(0, 256, 32 * 16 * ch).astype(np.uint8))
        out[f"tex{ch}_texels"] = tex.texels
        out[f"tex{ch}_samples"] = R.texture_sample_raw(tex, uv)
    hm = scenes.noise_height(32, seed=9)
    out["height_map"] = hm
    out["normal_map"] = R.height_to_normal(hm)
    np.savez_compressed(os.path.join(HERE, "functions.npz"), **out)
    print("functions.npz:", {k: v.shape for k, v in list(out.items())[:6]}, "...")


def scene_vectors(R, name, sd, W, H, params, lights=None, n_random=4000, n_color=1500, render_spp=4, adaptive=(3, 8)):
    d = tempfile.mkdtemp(prefix="golden_" + name)
    scenes.write_obj(sd, d)
    rs = R.load_scene(d, name)
    if lights is not None:
        R.set_lights(lights)
        rs.lights = np.ascontiguousarray(lights, LIGHT)
    R.set_params(params)
    hint = sd.camera_hint
    cam = R.make_camera(hint["fov"], W, H, hint["position"], hint["facing"])
    seed = int(params["base_seed"])
    assert all(R.check_jitter_order(cam, W, H, x, y, 99 + 7 * x + y) for x in range(0, W, 5) for y in range(0, H, 7)), \
        "jitter draw order assumption violated"
    out = rs.to_npz_dict()
    out["cam"] = np.frombuffer(cam.tobytes(), np.uint8)
    out["params"] = np.frombuffer(np.asarray(params).tobytes(), np.uint8)
    out["wh"] = np.array([W, H], np.uint32)
    rays, hits = R.trace_primary(cam, W, H, None, 0, W * H, 0, 2, seed)
    out["primary_rays"] = rays.view(np.uint8); out["primary_hits"] = hits.view(np.uint8)
    rng = np.random.default_rng(42)
    lo, hi = rs.positions.min(0), rs.positions.max(0)
    rr = random_rays(rng, n_random, lo, hi)
    rh, cnt = R.trace_rays(rr)
    out["random_rays"] = rr.view(np.uint8); out["random_hits"] = rh.view(np.uint8)
    out["random_counters"] = np.frombuffer(cnt.tobytes(), np.uint64)
    cr = np.concatenate([rays[:: max(1, len(rays) // n_color)][:n_color], rr[: n_color // 2]])
    cseeds = (np.arange(len(cr), dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) ^ np.uint64(seed)
    col, ccnt = R.trace_color(cr, cseeds)
    out["color_rays"] = cr.view(np.uint8); out["color_seeds"] = cseeds; out["color_rgba"] = col
    out["color_counters"] = np.frombuffer(ccnt.tobytes(), np.uint64)
    img, ns, rcnt, _ = R.render_seeded(cam, W, H, None, 0, W * H, 0, render_spp, render_spp, seed, threads=8)
    out["render_spp"] = np.array([render_spp], np.uint32)
    out["render_rgba"] = img; out["render_counters"] = np.frombuffer(rcnt.tobytes(), np.uint64)
    # sample-range split: samples [2, 2 + render_spp) as raw sums
    img2, _, _, _ = R.render_seeded(cam, W, H, None, 0, W * H, 2, render_spp, render_spp, seed, sum_only=True, threads=8)
    out["render_sum_from2"] = img2
    tm, tl = R.tonemap(img.reshape(H, W, 4))                     # the reference's WriteFramebufferImage on its own frame
    out["tonemap_rgba8"] = tm; out["tonemap_luma"] = np.array([tl], np.float32)
    a0, a1 = adaptive
    imga, nsa, acnt, _ = R.render_seeded(cam, W, H, None, 0, W * H, 0, a0, a1, seed, threads=8)
    out["adaptive_minmax"] = np.array([a0, a1], np.uint32)
    out["adaptive_rgba"] = imga; out["adaptive_nsamples"] = nsa
    out["adaptive_counters"] = np.frombuffer(acnt.tobytes(), np.uint64)
    path = os.path.join(HERE, f"scene_{name}.npz")
    np.savez_compressed(path, **out)
    print(f"scene_{name}.npz: {rs.n_triangles} tris, {rs.n_groups} groups, {len(rs.spheres)} spheres, "
          f"{os.path.getsize(path) / 1024:.0f} KiB, hit rate {hits['hit'].mean():.2f}, rays/render {rcnt['ray_count']}")


def main():
    R = ref_harness.get()
    function_vectors(R)
    # A: textured spheres + plane (diffuse/ambient maps, bump map, alpha mask, translucency), reference defaults
    sdA = scenes.spheres_plane_scene(grid=2, nu=16, nv=8, textured=True, name="spheres")
    scene_vectors(R, "spheres", sdA, 48, 32, types.default_params(spp=4))
    # B: shared-vertex height field, 36 groups (deep reference hierarchy), 2 diffuse + 2 specular samples,
    #    bounce depth 3, a directional AND a point light (inverted visibility quirk, raytracer.cpp:395-396)
    sdB = scenes.heightfield_scene(24, 24, block=4, textured=True, tex_size=32, name="heightfield")
    pB = types.default_params(spp=3)
    pB["reflection_samples"] = 2; pB["spec_samples"] = 2; pB["bounce_depth"] = 3
    pB["base_seed"] = 0x766B5859CFF0B8AF
    lights = np.zeros(2, LIGHT)
    lights[0] = types.default_lights()[0]
    lights[1]["type"] = types.LIGHT_POINT
    lights[1]["color"] = (3.0, 2.5, 2.0, 1.0)
    lights[1]["position"] = (5.0, 25.0, -10.0)
    lights[1]["falloff"] = 30.0
    scene_vectors(R, "heightfield", sdB, 40, 30, pB, lights=lights, n_random=3000, n_color=800, render_spp=3, adaptive=(2, 5))


if __name__ == "__main__":
    main()
