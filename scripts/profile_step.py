"""One Render() of the bench frame at reduced spp -- a short, representative launch sequence for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from par_raytracer_b200 import api, scenes, types
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sd = scenes.spheres_plane_scene()
W, H = 1920, 1080
h = sd.camera_hint
cam = types.make_camera(h["fov"], W, H, h["position"], h["facing"])
p = types.default_params(spp=spp)
S = api.Scene(sd)
for _ in range(reps):
    img, cnt = S.render_task(cam, p, W, H, flags=api.RT_FLAG_TIME_KERNELS)
    st = S.stats()
    print(f"spp={spp} rays={int(cnt['ray_count'])} gpu_ms={float(st['gpu_ms']):.2f} Mrays/s={int(cnt['ray_count'])/float(st['gpu_ms'])/1e3:.0f} "
          f"trace={float(st['trace_ms']):.2f} shadow={float(st['shadow_ms']):.2f} logic={float(st['logic_ms']):.2f} waves={int(st['waves'])} launches={int(st['kernel_launches'])}")
