"""Ad-hoc GPU bring-up check: CUDA path vs the C oracle on small scenes (run under gpurun)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from par_raytracer_b200 import api, scenes, types
from oracle import oracle

def cmp_hits(a, b, name):
    ok = True
    for f in ("hit", "object", "vertex0"):
        eq = np.array_equal(a[f], b[f]); ok &= eq
        if not eq: print(f"  {name}.{f}: {np.sum(a[f]!=b[f])} / {len(a)} differ")
    m = a["hit"] == 1
    for f in ("t", "bw", "position", "normal"):
        eq = np.array_equal(a[f][m].view(np.uint32), b[f][m].view(np.uint32)); ok &= eq
        if not eq: print(f"  {name}.{f}: {np.sum(a[f][m].view(np.uint32)!=b[f][m].view(np.uint32))} words differ")
    print(f"{name}: {'BIT-EXACT' if ok else 'MISMATCH'} ({len(a)} rays, {m.mean()*100:.1f}% hit)")
    return ok

def main():
    print("rng kat:", np.array_equal(api.rng_kat(0x835fdd9143716fe3, 40), oracle.rng_next(0x835fdd9143716fe3, 40)),
          np.array_equal(api.rng_kat(0, 70), oracle.rng_next(0, 70)))
    for name, sd in (("spheres_tex", scenes.spheres_plane_scene(grid=2, nu=24, nv=12, textured=True)),
                     ("heightfield", scenes.heightfield_scene(48, 48, block=8, textured=True, tex_size=64))):
        t = time.time(); S = api.Scene(sd); print(name, "scene create", time.time()-t, S.hierarchy_info())
        O = oracle.OracleScene(sd)
        W, H = 160, 120
        h = sd.camera_hint
        cam = types.make_camera(h["fov"], W, H, h["position"], h["facing"])
        p = types.default_params(spp=4)
        rays_g, hits_g = S.trace_primary(cam, p, W, H, sample_count=2)
        rays_o, hits_o = O.trace_primary(cam, p, W, H, None, 0, W*H, 0, 2)
        print("primary rays bit-exact:", rays_g.tobytes() == rays_o.tobytes())
        cmp_hits(hits_g, hits_o, "primary hits")
        rng = np.random.default_rng(1)
        n = 200000
        rays = np.zeros(n, types.RAY)
        lo, hi = sd.positions.min(0), sd.positions.max(0)
        rays["origin"] = (lo + (hi-lo)*rng.random((n,3)) + np.array([0, 2.0, 0])).astype(np.float32)
        d = rng.normal(size=(n,3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        rays["direction"] = d.astype(np.float32)
        hg, cg = S.trace_rays(p, rays)
        hb, _ = S.trace_rays(p, rays, api.RT_TRACE_BRUTE)
        ha, _ = S.trace_rays(p, rays, api.RT_TRACE_ANY)
        ho, co = O.trace_rays(p, rays)
        cmp_hits(hg, ho, "random rays vs oracle")
        cmp_hits(hb, ho, "brute vs oracle")
        print("any-hit flag equal:", np.array_equal(ha["hit"], ho["hit"]), "sphere checks/ray gpu", cg["sphere_check_count"]/n, "cluster", cg["mesh_check_count"]/n,
              "| ref", co["sphere_check_count"]/n, co["mesh_check_count"]/n)
        # render
        t = time.time(); img_g, cnt_g = S.render(cam, p, W, H); tg = time.time()-t
        img_o, ns, cnt_o, sec = O.render(cam, p, W, H, threads=os.cpu_count())
        img_o = img_o.reshape(H, W, 4)
        rel = np.abs(img_g - img_o) / np.maximum(1e-3, np.abs(img_o))
        print("render: ray_count gpu/oracle", cnt_g["ray_count"], cnt_o["ray_count"], "max rel", rel.max(), "mean rel", rel.mean(),
              "n>1e-4:", int((rel > 1e-4).sum()), "gpu s", tg, "cpu s", sec, S.stats())
        mse = np.mean((img_g[..., :3] - img_o[..., :3])**2); print("  PSNR dB:", 10*np.log10(img_o[..., :3].max()**2 / max(mse, 1e-30)))
        # trace_color
        cg2, _ = S.trace_color(p, rays[:50000], np.arange(50000, dtype=np.uint64) + 77)
        co2, _ = O.trace_colors(p, rays[:50000], np.arange(50000, dtype=np.uint64) + 77)
        rel = np.abs(cg2[:, :3] - co2[:, :3]) / np.maximum(1e-3, np.abs(co2[:, :3]))
        print("trace_color max rel", rel.max(), "n>1e-4", int((rel > 1e-4).sum()))
    # throughput probe
    sd = scenes.spheres_plane_scene()
    S = api.Scene(sd); print("config2 scene", S.hierarchy_info())
    W, H = 1920, 1080
    h = sd.camera_hint
    cam = types.make_camera(h["fov"], W, H, h["position"], h["facing"])
    for spp in (1, 4, 8):
        p = types.default_params(spp=spp)
        img, cnt = S.render(cam, p, W, H)
        st = S.stats()
        print(f"1080p spp={spp}: rays {cnt['ray_count']} gpu_ms {st['gpu_ms']:.2f} -> {cnt['ray_count']/st['gpu_ms']/1e3:.1f} Mrays/s waves {st['waves']} launches {st['kernel_launches']}")
    p = types.default_params(spp=2)
    img, cnt = S.render_task(cam, p, W, H, flags=api.RT_FLAG_COUNTERS)
    print("checks/ray:", cnt["sphere_check_count"]/cnt["ray_count"], cnt["mesh_check_count"]/cnt["ray_count"])

if __name__ == "__main__":
    main()
