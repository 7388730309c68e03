"""BASELINE config 3 scene (1,048,352 textured triangles) at 1080p and low spp: a short launch sequence for ncu (two frames; profile the second)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from par_raytracer_b200 import api, scenes, types
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
sd = scenes.heightfield_scene(724, 724, block=32, size=400.0, amp=20.0, textured=True, tex_size=512, hierarchy="defer")
scenes.use_reference_hierarchy(sd)
S = api.Scene(sd)
W, H = 1920, 1080
h = sd.camera_hint
cam = types.make_camera(h["fov"], W, H, h["position"], h["facing"])
p = types.default_params(spp=spp)
for _ in range(reps):
    img, cnt = S.render_task(cam, p, W, H, flags=api.RT_FLAG_TIME_KERNELS)
    st = S.stats()
    print(f"tris={sd.n_triangles} spp={spp} rays={int(cnt['ray_count'])} gpu_ms={float(st['gpu_ms']):.2f} Mrays/s={int(cnt['ray_count'])/float(st['gpu_ms'])/1e3:.0f} "
          f"trace={float(st['trace_ms']):.2f} logic={float(st['logic_ms']):.2f} waves={int(st['waves'])} launches={int(st['kernel_launches'])}")
