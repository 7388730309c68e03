"""How much would ray reordering buy? Diffuse bounce rays off the primary hits of the 10 M-triangle height field, traced through
rt_trace_rays in three orders; run under `ncu --metrics gpu__time_duration.sum -k regex:k_trace_wave` to read kernel times."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from par_raytracer_b200 import api, scenes, types
cells = int(sys.argv[1]) if len(sys.argv) > 1 else 2236
sd = scenes.heightfield_scene(cells, cells, block=32, size=400.0, amp=20.0, textured=False)
S = api.Scene(sd)
W, H = 1920, 1080
h = sd.camera_hint
cam = types.make_camera(h["fov"], W, H, h["position"], h["facing"])
p = types.default_params(spp=2)
rays, hits = S.trace_primary(cam, p, W, H, sample_count=2)
m = hits["hit"] == 1
pos = hits["position"][m]; nrm = hits["normal"][m]
n = len(pos)
rng = np.random.default_rng(7)
u1 = rng.random(n, dtype=np.float32); u2 = rng.random(n, dtype=np.float32)
r = np.sqrt(u1); phi = 2 * np.pi * u2
lx, ly, lz = r * np.cos(phi), r * np.sin(phi), np.sqrt(np.maximum(0, 1 - u1))
up = np.where(np.abs(nrm[:, 2:3]) < 0.9999, np.array([[0, 0, 1.0]], np.float32), np.array([[1.0, 0, 0]], np.float32))
tg = np.cross(up, nrm); tg /= np.linalg.norm(tg, axis=1, keepdims=True)
bt = np.cross(nrm, tg)
d = (tg * lx[:, None] + bt * ly[:, None] + nrm * lz[:, None]).astype(np.float32)
d /= np.linalg.norm(d, axis=1, keepdims=True)
R = np.zeros(n, types.RAY); R["origin"] = pos + nrm * 1e-3; R["direction"] = d
print("bounce rays", n, flush=True)

def morton(q):  # q: (n,3) ints < 1024
    def spread(v):
        v = v.astype(np.uint64) & 0x3FF
        v = (v | (v << 16)) & 0x30000FF; v = (v | (v << 8)) & 0x300F00F; v = (v | (v << 4)) & 0x30C30C3; v = (v | (v << 2)) & 0x9249249
        return v
    return (spread(q[:, 0]) << 2) | (spread(q[:, 1]) << 1) | spread(q[:, 2])
octant = ((d[:, 0] < 0).astype(np.uint64) << 2) | ((d[:, 1] < 0).astype(np.uint64) << 1) | (d[:, 2] < 0).astype(np.uint64)
lo, hi = pos.min(0), pos.max(0)
q = np.minimum(1023, ((pos - lo) / (hi - lo + 1e-9) * 1024).astype(np.int64))
mo = morton(q)
dq = np.minimum(7, ((d * 0.5 + 0.5) * 8).astype(np.int64))
dkey = (dq[:, 0] << 6) | (dq[:, 1] << 3) | dq[:, 2]
orders = {
    "pixel order": np.arange(n),
    "octant (stable)": np.argsort(octant, kind="stable"),
    "octant | morton30": np.argsort((octant << 30) | mo, kind="stable"),
    "morton15hi | dir9 | morton15lo": np.argsort(((mo >> 15) << 24) | (dkey.astype(np.uint64) << 15) | (mo & 0x7FFF), kind="stable"),
    "dir9 | morton30": np.argsort((dkey.astype(np.uint64) << 30) | mo, kind="stable"),
    "random": rng.permutation(n),
}
ref = None
for name, o in orders.items():
    out, cnt = S.trace_rays(p, R[o])
    inv = np.empty(n, np.int64); inv[o] = np.arange(n)
    t = out["t"][inv]
    if ref is None: ref = t
    print(f"{name}: hits {int((out['hit'] == 1).sum())} same_t {bool((t == ref).all())} node tests/ray {int(cnt['sphere_check_count']) / n:.1f}", flush=True)
