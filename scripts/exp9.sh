#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dropin.py -x -q -m gpu 2>&1 | tail -3
RT_B200_BOUNDS=box timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -2
for b in qbox box; do
  for s in c2 707 2236; do RT_B200_BOUNDS=$b python scripts/sweep2.py $s 12:16,16:16; done
done
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
    for b in ("qbox", "box"):
        os.environ["RT_B200_BOUNDS"] = b
        S = api.Scene(sd)
        img, cnt = S.render_task(cam, p, W, H, flags=api.RT_FLAG_COUNTERS)
        r = int(cnt['ray_count'])
        print(cells, b, "rays", r, "node tests/ray", int(cnt['sphere_check_count'])/r, "clusters/ray", int(cnt['mesh_check_count'])/r, S.hierarchy_info())
        S.close()
PY
