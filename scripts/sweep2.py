"""Knob sweep in one process per scene: RT_B200_LEAF_WAIT x RT_B200_FETCH_MIN (read per launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from par_raytracer_b200 import api, scenes, types
which = sys.argv[1]
W, H = 1920, 1080
if which == "c2":
    sd = scenes.spheres_plane_scene(); spp = 64
else:
    cells = int(which); sd = scenes.heightfield_scene(cells, cells, block=32, size=400.0, amp=20.0, textured=False); spp = 16
h = sd.camera_hint
cam = types.make_camera(h["fov"], W, H, h["position"], h["facing"])
p = types.default_params(spp=spp)
S = api.Scene(sd)
S.render_task(cam, p, W, H, flags=api.RT_FLAG_TIME_KERNELS)
combos = [(int(a), int(b)) for a, b in (c.split(":") for c in sys.argv[2].split(","))]
for lw, fm in combos:
    os.environ["RT_B200_LEAF_WAIT"] = str(lw)
    for k in ("RT_B200_FETCH_MIN", "RT_B200_FETCH_PRIMARY", "RT_B200_FETCH_SHADOW"):
        os.environ[k] = str(fm)
    best = None
    for _ in range(2):
        img, cnt = S.render_task(cam, p, W, H, flags=api.RT_FLAG_TIME_KERNELS)
        st = S.stats()
        r = (float(st['gpu_ms']), float(st['trace_ms']), float(st['logic_ms']))
        best = r if best is None or r[0] < best[0] else best
    print(f"{which} lib={os.environ.get('RT_B200_LIB','default')[-20:]} leaf_wait={lw} fetch={fm} rays={int(cnt['ray_count'])} gpu_ms={best[0]:.2f} trace={best[1]:.2f} logic={best[2]:.2f} Mrays/s={int(cnt['ray_count'])/best[0]/1e3:.0f}", flush=True)
