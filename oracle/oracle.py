"""ctypes wrapper of oracle/liboracle.so (the C restatement, oracle/rt_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs -- as the checker, never as the thing measured or shipped.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from par_raytracer_b200.cabi import RtTexture, make_scene_desc  # noqa: E402
from par_raytracer_b200.types import CAMERA, COUNTERS, HIT, PARAMS, RAY, SceneData, TextureData  # noqa: E402

LIB_PATH = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "rt_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "rt_b200.h")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_scene_create.restype = C.c_void_p
        L.orc_scene_create.argtypes = [C.c_void_p]
        L.orc_scene_destroy.argtypes = [C.c_void_p]
        L.orc_render.restype = C.c_double
        L.orc_tonemap.restype = C.c_float
        _LIB = L
    return _LIB


class OracleScene:
    def __init__(self, scene: SceneData):
        self.scene = scene
        self.desc, self._keep = make_scene_desc(scene)
        self.h = C.c_void_p(lib().orc_scene_create(C.addressof(self.desc)))

    def __del__(self):
        try:
            if self.h:
                lib().orc_scene_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def trace_rays(self, params: np.ndarray, rays: np.ndarray):
        rays = np.ascontiguousarray(rays, RAY); params = np.asarray(params, PARAMS).reshape(1)
        out = np.zeros(len(rays), HIT); cnt = np.zeros(1, COUNTERS)
        lib().orc_trace_rays(self.h, _p(params), _p(rays), C.c_uint64(len(rays)), _p(out), _p(cnt))
        return out, cnt[0]

    def trace_colors(self, params, rays, seeds):
        rays = np.ascontiguousarray(rays, RAY); seeds = np.ascontiguousarray(seeds, np.uint64)
        params = np.asarray(params, PARAMS).reshape(1)
        out = np.zeros((len(rays), 4), np.float32); cnt = np.zeros(1, COUNTERS)
        lib().orc_trace_colors(self.h, _p(params), _p(rays), _p(seeds), C.c_uint64(len(rays)), _p(out), _p(cnt))
        return out, cnt[0]

    def trace_primary(self, cam, params, width, height, pixel_ids, pixel_begin, pixel_count, sample_begin, sample_count,
                      want_hits=True):
        cam = np.asarray(cam, CAMERA).reshape(1); params = np.asarray(params, PARAMS).reshape(1)
        ids = None if pixel_ids is None else np.ascontiguousarray(pixel_ids, np.uint32)
        n = pixel_count * sample_count
        rays = np.zeros(n, RAY); hits = np.zeros(n, HIT) if want_hits else None
        lib().orc_trace_primary(self.h, _p(cam), _p(params), C.c_uint32(width), C.c_uint32(height), _p(ids),
                                C.c_uint32(pixel_begin), C.c_uint32(pixel_count), C.c_uint32(sample_begin),
                                C.c_uint32(sample_count), _p(rays), _p(hits))
        return rays, hits

    def render(self, cam, params, width, height, pixel_ids=None, pixel_begin=0, pixel_count=None, sample_begin=0,
               sum_only=False, threads=1):
        cam = np.asarray(cam, CAMERA).reshape(1); params = np.asarray(params, PARAMS).reshape(1)
        ids = None if pixel_ids is None else np.ascontiguousarray(pixel_ids, np.uint32)
        if pixel_count is None:
            pixel_count = len(ids) if ids is not None else width * height - pixel_begin
        out = np.zeros((pixel_count, 4), np.float32); ns = np.zeros(pixel_count, np.uint32); cnt = np.zeros(1, COUNTERS)
        sec = lib().orc_render(self.h, _p(cam), _p(params), C.c_uint32(width), C.c_uint32(height), _p(ids),
                               C.c_uint32(pixel_begin), C.c_uint32(pixel_count), C.c_uint32(sample_begin),
                               C.c_int(1 if sum_only else 0), C.c_uint32(threads), _p(out), _p(ns), _p(cnt))
        return out, ns, cnt[0], float(sec)


# ---- function-level probes (no scene) ----------------------------------------------------------
def rng_next(seed: int, n: int) -> np.ndarray:
    out = np.zeros(n, np.uint64)
    lib().orc_rng_next_n(C.c_uint64(seed), C.c_uint32(n), _p(out))
    return out


def rng_float(seed: int, n: int, signed: bool) -> np.ndarray:
    out = np.zeros(n, np.float32)
    lib().orc_rng_float_n(C.c_uint64(seed), C.c_uint32(n), C.c_int(1 if signed else 0), _p(out))
    return out


def camera_rays(cam, xy) -> np.ndarray:
    cam = np.asarray(cam, CAMERA).reshape(1); xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
    out = np.zeros(len(xy), RAY)
    lib().orc_camera_rays(_p(cam), C.c_uint32(len(xy)), _p(xy), _p(out))
    return out


def intersect_triangle(rays, tris, best_t):
    rays = np.ascontiguousarray(rays, RAY); tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
    best_t = np.ascontiguousarray(best_t, np.float32)
    hit = np.zeros(len(rays), np.uint32); out = np.zeros((len(rays), 10), np.float32)
    lib().orc_intersect_triangle_n(C.c_uint32(len(rays)), _p(rays), _p(tris), _p(best_t), _p(hit), _p(out))
    return hit, out


def intersect_sphere(rays, spheres):
    rays = np.ascontiguousarray(rays, RAY); spheres = np.ascontiguousarray(spheres, np.float32).reshape(-1, 4)
    hit = np.zeros(len(rays), np.uint32); t = np.zeros(len(rays), np.float32)
    lib().orc_intersect_sphere_n(C.c_uint32(len(rays)), _p(rays), _p(spheres), _p(hit), _p(t))
    return hit, t


def hammersley(i, n) -> np.ndarray:
    i = np.ascontiguousarray(i, np.uint32); n = np.ascontiguousarray(n, np.uint32)
    out = np.zeros((len(i), 2), np.float32)
    lib().orc_hammersley_n(C.c_uint32(len(i)), _p(i), _p(n), _p(out))
    return out


def diffuse_rays(origin, normal, xi) -> np.ndarray:
    origin = np.ascontiguousarray(origin, np.float32); normal = np.ascontiguousarray(normal, np.float32)
    xi = np.ascontiguousarray(xi, np.float32)
    out = np.zeros(len(origin), RAY)
    lib().orc_diffuse_rays(C.c_uint32(len(origin)), _p(origin), _p(normal), _p(xi), _p(out))
    return out


def specular_rays(origin, normal, spec, xi) -> np.ndarray:
    origin = np.ascontiguousarray(origin, np.float32); normal = np.ascontiguousarray(normal, np.float32)
    spec = np.ascontiguousarray(spec, np.float32); xi = np.ascontiguousarray(xi, np.float32)
    out = np.zeros(len(origin), RAY)
    lib().orc_specular_rays(C.c_uint32(len(origin)), _p(origin), _p(normal), _p(spec), _p(xi), _p(out))
    return out


def fresnel(ior_exit, ior_enter, normal, incident) -> np.ndarray:
    a = np.ascontiguousarray(ior_exit, np.float32); b = np.ascontiguousarray(ior_enter, np.float32)
    n = np.ascontiguousarray(normal, np.float32); i = np.ascontiguousarray(incident, np.float32)
    out = np.zeros(len(a), np.float32)
    lib().orc_fresnel_n(C.c_uint32(len(a)), _p(a), _p(b), _p(n), _p(i), _p(out))
    return out


def texture_sample(tex: TextureData, uv) -> np.ndarray:
    uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
    t = RtTexture(tex.size_x, tex.size_y, tex.channels, tex.texels.ctypes.data)
    out = np.zeros((len(uv), 4), np.float32)
    lib().orc_texture_sample_n(C.byref(t), C.c_uint32(len(uv)), _p(uv), _p(out))
    return out


def srgb_lut() -> np.ndarray:
    out = np.zeros(256, np.float32)
    lib().orc_srgb_lut(_p(out))
    return out


def tonemap(frame: np.ndarray):
    """(H, W, 4) float32 -> ((H, W, 4) uint8, scene_luma): LogAverageLuma + tone map + Color_Pack (main.cpp:78-127)."""
    frame = np.ascontiguousarray(frame, np.float32)
    h, w = frame.shape[:2]
    out = np.zeros((h, w, 4), np.uint8)
    luma = lib().orc_tonemap(_p(frame), C.c_uint32(w), C.c_uint32(h), _p(out))
    return out, float(luma)


def build_hierarchy(scene: SceneData):
    """BuildHierarchy (bsphere.cpp:379-444) restated: returns (spheres BSPHERE array, sphere_group) for the scene's groups."""
    from par_raytracer_b200.types import BSPHERE
    G = scene.n_groups
    spheres = np.zeros(max(1, 2 * G - 1), BSPHERE); sg = np.zeros(max(1, 2 * G - 1), np.int32)
    lib().orc_build_hierarchy.restype = C.c_uint32
    n = lib().orc_build_hierarchy(_p(scene.positions), C.c_uint32(G), _p(scene.group_first), _p(scene.idx_positions), _p(spheres), _p(sg))
    return spheres[:n], sg[:n]


def group_has_bump(scene: SceneData) -> np.ndarray:
    """mesh.h:70: a group takes part in CalculateTangents iff it has a material with a bump texture."""
    gm = scene.group_material
    return np.array([1 if (m >= 0 and scene.materials[m]["bump_texture"] >= 0) else 0 for m in gm], np.uint8)


def calculate_tangents(scene: SceneData) -> np.ndarray:
    out = np.zeros((len(scene.normals), 3), np.float32)
    hb = group_has_bump(scene)
    lib().orc_calculate_tangents(_p(scene.positions), _p(scene.texcoords), C.c_uint32(len(scene.normals)), C.c_uint32(scene.n_groups),
                                 _p(scene.group_first), _p(scene.idx_positions), _p(scene.idx_texcoords), _p(scene.idx_normals), _p(hb), _p(out))
    return out


def height_to_normal_map(height: np.ndarray) -> np.ndarray:
    height = np.ascontiguousarray(height, np.uint8)
    h, w = height.shape
    out = np.zeros((h, w, 3), np.uint8)
    lib().orc_height_to_normal_map(C.c_uint32(w), C.c_uint32(h), _p(height), _p(out))
    return out
