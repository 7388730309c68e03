"""The reference's DEFAULT mode (main.cpp:308-309: adaptive 10..50 spp, 720x480) on the config-2 scene: time and per-pixel sample counts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from par_raytracer_b200 import api, scenes, types
sd = scenes.spheres_plane_scene()
W, H = 720, 480
h = sd.camera_hint
cam = types.make_camera(h["fov"], W, H, h["position"], h["facing"])
p = types.default_params(spp=10); p["min_samples"], p["max_samples"] = 10, 50
S = api.Scene(sd)
for rep in range(3):
    img, cnt = S.render_task(cam, p, W, H, flags=api.RT_FLAG_ADAPTIVE | api.RT_FLAG_TIME_KERNELS)
    st = S.stats()
    ns = S.sample_counts(W * H)
    print(f"adaptive 10..50 spp {W}x{H}: rays={int(cnt['ray_count'])} gpu_ms={float(st['gpu_ms']):.2f} Mrays/s={int(cnt['ray_count'])/float(st['gpu_ms'])/1e3:.0f} "
          f"trace={float(st['trace_ms']):.2f} logic={float(st['logic_ms']):.2f} waves={int(st['waves'])} launches={int(st['kernel_launches'])} "
          f"mean spp={ns.mean():.1f} pixels at max={(ns == 50).mean():.2f}", flush=True)
p2 = types.default_params(spp=50)
img, cnt = S.render_task(cam, p2, W, H, flags=api.RT_FLAG_TIME_KERNELS)
st = S.stats()
print(f"fixed 50 spp: rays={int(cnt['ray_count'])} gpu_ms={float(st['gpu_ms']):.2f} Mrays/s={int(cnt['ray_count'])/float(st['gpu_ms'])/1e3:.0f}")
