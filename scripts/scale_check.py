"""Scale check: config-3/4 style height-field scenes (1M / 10M triangles): build time, memory, throughput, parity subset."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from par_raytracer_b200 import api, scenes, types
from oracle import oracle
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 724
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 16
W, H = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (1920, 1080)
t = time.time(); sd = scenes.heightfield_scene(cells, cells, block=32, size=400.0, amp=20.0, textured=True, tex_size=512); print(f"generate {sd.n_triangles} tris, {sd.n_groups} groups: {time.time()-t:.1f}s", flush=True)
t = time.time(); S = api.Scene(sd); print(f"scene create {time.time()-t:.2f}s", S.hierarchy_info(), flush=True)
h = sd.camera_hint
cam = types.make_camera(h["fov"], W, H, h["position"], h["facing"])
p = types.default_params(spp=spp)
for rep in range(2):
    img, cnt = S.render_task(cam, p, W, H, flags=api.RT_FLAG_TIME_KERNELS)
    st = S.stats()
    print(f"{W}x{H} spp={spp} rays={int(cnt['ray_count'])} gpu_ms={float(st['gpu_ms']):.1f} Mrays/s={int(cnt['ray_count'])/float(st['gpu_ms'])/1e3:.0f} trace={float(st['trace_ms']):.1f} logic={float(st['logic_ms']):.1f} waves={int(st['waves'])}", flush=True)
img2, cnt2 = S.render_task(cam, types.default_params(spp=2), W, H, flags=api.RT_FLAG_COUNTERS)
print("checks/ray", cnt2["sphere_check_count"]/cnt2["ray_count"], cnt2["mesh_check_count"]/cnt2["ray_count"], "hit frac", float((img2[:, :3].sum(1) > 0).mean()))
# parity on a pixel subset against the oracle (the python-built group hierarchy is the reference-format input of both)
ids = np.arange(0, W * H, max(1, W * H // 1500), dtype=np.uint32)
ps = types.default_params(spp=4)
O = oracle.OracleScene(sd)
t = time.time(); ref, _, cnt_o, sec = O.render(cam, ps, W, H, pixel_ids=ids, threads=os.cpu_count()); print(f"oracle subset {len(ids)} px x4 spp: {sec:.1f}s {cnt_o['ray_count']/sec/1e6:.3f} Mrays/s on {os.cpu_count()} threads")
sub, cnt_s = S.render_task(cam, ps, W, H, pixel_ids=ids)
print("ray_count equal:", int(cnt_s["ray_count"]) == int(cnt_o["ray_count"]), "allclose:", np.allclose(sub, ref, rtol=1e-5, atol=1e-6), "max rel", float((np.abs(sub-ref)/np.maximum(1e-3, np.abs(ref))).max()))
