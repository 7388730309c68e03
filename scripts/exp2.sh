#!/bin/bash
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from par_raytracer_b200 import api, scenes, types
for cells in (707, 2236):
    sd = scenes.heightfield_scene(cells, cells, block=32, size=400.0, amp=20.0, textured=False)
    W, H = 1920, 1080
    h = sd.camera_hint
    cam = types.make_camera(h["fov"], W, H, h["position"], h["facing"])
    p = types.default_params(spp=4)
    for b in ("box", "sphere"):
        os.environ["RT_B200_BOUNDS"] = b
        S = api.Scene(sd)
        img, cnt = S.render_task(cam, p, W, H, flags=api.RT_FLAG_COUNTERS)
        r = int(cnt['ray_count'])
        print(cells, b, "rays", r, "node tests/ray", int(cnt['sphere_check_count'])/r, "clusters/ray", int(cnt['mesh_check_count'])/r, S.hierarchy_info())
        S.close()
PY
ncu --set full --import-source on --clock-control none -k regex:k_trace_wave -c 3 -o gpurun_out/r1_box_big python scripts/profile_big.py 2236 4 > gpurun_out/ncu_box_big.log 2>&1
