"""ctypes wrapper of oracle/_ref/libref_harness.so -- the UNMODIFIED reference compiled from
/root/reference (oracle/Makefile). TEST INFRASTRUCTURE ONLY: imported by tests/, by
tests/golden/make_golden.py and by bench.py's reference arm / cpu_baseline leg; never by the product.
"""
from __future__ import annotations

import ctypes as C
import os
import sys
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from par_raytracer_b200.types import (BSPHERE, CAMERA, COUNTERS, HIT, LIGHT, MATERIAL, PARAMS, RAY, SceneData,  # noqa: E402
                                      TextureData)

LIB_PATH = os.path.join(_HERE, "_ref", "libref_harness.so")


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class RefHarness:
    """One process-wide instance: the reference keeps its scene in globals (gParams etc.)."""

    def __init__(self):
        if not available():
            raise RuntimeError(f"{LIB_PATH} not built (make -C oracle ref, needs /root/reference)")
        self.lib = C.CDLL(LIB_PATH)
        L = self.lib
        L.ref_load_scene.restype = C.c_int
        L.ref_load_scene.argtypes = [C.c_char_p]
        L.ref_rng_table.restype = C.c_uint64
        L.ref_rng_table.argtypes = [C.c_uint32, C.c_uint32]
        L.ref_render_seeded.restype = C.c_double
        L.ref_render_ranks.restype = C.c_double
        L.ref_check_jitter_order.restype = C.c_int
        L.ref_tonemap_png.restype = C.c_float
        self.loaded = False

    # ---- scene ---------------------------------------------------------------------------
    def load_scene(self, directory: str, name: str = "ref_scene") -> SceneData:
        rc = self.lib.ref_load_scene(directory.encode())
        if rc != 0:
            raise RuntimeError(f"reference ParseOBJ failed for {directory}")
        self.loaded = True
        return self.export_scene(name)

    def export_scene(self, name: str = "ref_scene") -> SceneData:
        sz = np.zeros(16, dtype=np.uint64)
        self.lib.ref_export_sizes(_p(sz))
        nP, nT, nN, nG, nI, nS, nM, nX, nL, hasT = (int(v) for v in sz[:10])
        positions = np.zeros((nP, 3), np.float32)
        texcoords = np.zeros((nT, 2), np.float32)
        normals = np.zeros((nN, 3), np.float32)
        tangents = np.zeros((nN, 3), np.float32) if hasT else None
        group_first = np.zeros(nG + 1, np.uint32)
        ip = np.zeros(nI, np.uint32); it = np.zeros(nI, np.uint32); inn = np.zeros(nI, np.uint32)
        gm = np.zeros(nG, np.int32)
        spheres = np.zeros(nS, BSPHERE)
        sg = np.zeros(nS, np.int32)
        mats = np.zeros(max(nM, 1), MATERIAL)
        dmat = np.zeros(1, MATERIAL)
        lights = np.zeros(max(nL, 1), LIGHT)
        self.lib.ref_export_fill(_p(positions), _p(texcoords), _p(normals), _p(tangents), _p(group_first), _p(ip), _p(it),
                                 _p(inn), _p(gm), _p(spheres), _p(sg), _p(mats), _p(dmat), _p(lights))
        textures = []
        for i in range(nX):
            info = np.zeros(3, np.uint32)
            self.lib.ref_texture_info(C.c_uint32(i), _p(info))
            buf = np.zeros(int(info[0]) * int(info[1]) * int(info[2]), np.uint8)
            self.lib.ref_texture_copy(C.c_uint32(i), _p(buf))
            textures.append(TextureData(int(info[0]), int(info[1]), int(info[2]), buf))
        # the bump path only reads tangents of bump-mapped groups; keep the full array as the reference has it
        sd = SceneData(positions=positions, texcoords=texcoords, normals=normals, tangents=tangents,
                       group_first=group_first, idx_positions=ip, idx_texcoords=it, idx_normals=inn, group_material=gm,
                       spheres=spheres, sphere_group=sg, materials=mats[:nM], default_material=dmat[0], textures=textures,
                       lights=lights[:nL], name=name)
        sd.validate()
        return sd

    def set_params(self, params: np.ndarray) -> None:
        bg = np.ascontiguousarray(params["background_color"], dtype=np.float32)
        self.lib.ref_set_params(C.c_float(float(params["ray_bias"])), C.c_uint32(int(params["reflection_samples"])),
                                C.c_uint32(int(params["spec_samples"])), C.c_uint32(int(params["bounce_depth"])), _p(bg))

    def get_params(self) -> np.ndarray:
        p = np.zeros(1, PARAMS)
        self.lib.ref_get_params(_p(p))
        return p[0]

    def set_lights(self, lights: np.ndarray) -> None:
        lights = np.ascontiguousarray(lights, dtype=LIGHT)
        self.lib.ref_set_lights(C.c_uint32(len(lights)), _p(lights))

    # ---- probes --------------------------------------------------------------------------
    def rng_next(self, seed: int, n: int) -> np.ndarray:
        out = np.zeros(n, np.uint64)
        self.lib.ref_rng_next(C.c_uint64(seed), C.c_uint32(n), _p(out))
        return out

    def rng_float(self, seed: int, n: int, signed: bool) -> np.ndarray:
        out = np.zeros(n, np.float32)
        self.lib.ref_rng_float(C.c_uint64(seed), C.c_uint32(n), C.c_int(1 if signed else 0), _p(out))
        return out

    def rng_table(self, process_id: int, thread_id: int = 0) -> int:
        return int(self.lib.ref_rng_table(process_id, thread_id))

    def make_camera(self, fov: float, w: int, h: int, position, facing) -> np.ndarray:
        cam = np.zeros(1, CAMERA)
        pos = np.asarray(position, np.float32); fac = np.asarray(facing, np.float32)
        self.lib.ref_make_camera(C.c_float(fov), C.c_uint32(w), C.c_uint32(h), _p(pos), _p(fac), _p(cam))
        return cam[0]

    def camera_rays(self, cam: np.ndarray, xy: np.ndarray) -> np.ndarray:
        xy = np.ascontiguousarray(xy, np.float32).reshape(-1, 2)
        out = np.zeros(len(xy), RAY)
        cam = np.asarray(cam, CAMERA).reshape(1)
        self.lib.ref_camera_rays(_p(cam), C.c_uint32(len(xy)), _p(xy), _p(out))
        return out

    def intersect_triangle(self, rays: np.ndarray, tris: np.ndarray, best_t: np.ndarray):
        rays = np.ascontiguousarray(rays, RAY); tris = np.ascontiguousarray(tris, np.float32).reshape(-1, 9)
        best_t = np.ascontiguousarray(best_t, np.float32)
        hit = np.zeros(len(rays), np.uint32); out = np.zeros((len(rays), 10), np.float32)
        self.lib.ref_intersect_triangle(C.c_uint32(len(rays)), _p(rays), _p(tris), _p(best_t), _p(hit), _p(out))
        return hit, out

    def intersect_sphere(self, rays: np.ndarray, spheres: np.ndarray):
        rays = np.ascontiguousarray(rays, RAY); spheres = np.ascontiguousarray(spheres, np.float32).reshape(-1, 4)
        hit = np.zeros(len(rays), np.uint32); t = np.zeros(len(rays), np.float32)
        self.lib.ref_intersect_sphere(C.c_uint32(len(rays)), _p(rays), _p(spheres), _p(hit), _p(t))
        return hit, t

    def hammersley(self, i: np.ndarray, n: np.ndarray) -> np.ndarray:
        i = np.ascontiguousarray(i, np.uint32); n = np.ascontiguousarray(n, np.uint32)
        out = np.zeros((len(i), 2), np.float32)
        self.lib.ref_hammersley(C.c_uint32(len(i)), _p(i), _p(n), _p(out))
        return out

    def diffuse_rays(self, origin, normal, xi) -> np.ndarray:
        origin = np.ascontiguousarray(origin, np.float32); normal = np.ascontiguousarray(normal, np.float32)
        xi = np.ascontiguousarray(xi, np.float32)
        out = np.zeros(len(origin), RAY)
        self.lib.ref_diffuse_rays(C.c_uint32(len(origin)), _p(origin), _p(normal), _p(xi), _p(out))
        return out

    def specular_rays(self, origin, normal, spec, xi) -> np.ndarray:
        origin = np.ascontiguousarray(origin, np.float32); normal = np.ascontiguousarray(normal, np.float32)
        spec = np.ascontiguousarray(spec, np.float32); xi = np.ascontiguousarray(xi, np.float32)
        out = np.zeros(len(origin), RAY)
        self.lib.ref_specular_rays(C.c_uint32(len(origin)), _p(origin), _p(normal), _p(spec), _p(xi), _p(out))
        return out

    def fresnel(self, ior_exit, ior_enter, normal, incident) -> np.ndarray:
        a = np.ascontiguousarray(ior_exit, np.float32); b = np.ascontiguousarray(ior_enter, np.float32)
        n = np.ascontiguousarray(normal, np.float32); i = np.ascontiguousarray(incident, np.float32)
        out = np.zeros(len(a), np.float32)
        self.lib.ref_fresnel(C.c_uint32(len(a)), _p(a), _p(b), _p(n), _p(i), _p(out))
        return out

    def texture_sample_raw(self, tex: TextureData, uv: np.ndarray) -> np.ndarray:
        uv = np.ascontiguousarray(uv, np.float32).reshape(-1, 2)
        out = np.zeros((len(uv), 4), np.float32)
        self.lib.ref_texture_sample_raw(C.c_uint32(tex.size_x), C.c_uint32(tex.size_y), C.c_uint32(tex.channels),
                                        _p(tex.texels), C.c_uint32(len(uv)), _p(uv), _p(out))
        return out

    def height_to_normal(self, height: np.ndarray) -> np.ndarray:
        height = np.ascontiguousarray(height, np.uint8)
        h, w = height.shape
        out = np.zeros((h, w, 3), np.uint8)
        self.lib.ref_height_to_normal(C.c_uint32(w), C.c_uint32(h), _p(height), _p(out))
        return out

    def srgb_lut(self) -> np.ndarray:
        out = np.zeros(256, np.float32)
        self.lib.ref_srgb_lut(_p(out))
        return out

    # ---- TraceRay / TraceRayColor / render -------------------------------------------------
    def trace_rays(self, rays: np.ndarray):
        rays = np.ascontiguousarray(rays, RAY)
        out = np.zeros(len(rays), HIT); cnt = np.zeros(1, COUNTERS)
        self.lib.ref_trace_rays(C.c_uint64(len(rays)), _p(rays), _p(out), _p(cnt))
        return out, cnt[0]

    def trace_color(self, rays: np.ndarray, seeds: np.ndarray):
        rays = np.ascontiguousarray(rays, RAY); seeds = np.ascontiguousarray(seeds, np.uint64)
        out = np.zeros((len(rays), 4), np.float32); cnt = np.zeros(1, COUNTERS)
        self.lib.ref_trace_color(C.c_uint64(len(rays)), _p(rays), _p(seeds), _p(out), _p(cnt))
        return out, cnt[0]

    def trace_primary(self, cam, width, height, pixel_ids, pixel_begin, pixel_count, sample_begin, sample_count, base_seed,
                      want_hits=True):
        cam = np.asarray(cam, CAMERA).reshape(1)
        ids = None if pixel_ids is None else np.ascontiguousarray(pixel_ids, np.uint32)
        n = pixel_count * sample_count
        rays = np.zeros(n, RAY); hits = np.zeros(n, HIT) if want_hits else None
        self.lib.ref_trace_primary(_p(cam), C.c_uint32(width), C.c_uint32(height), _p(ids), C.c_uint32(pixel_begin),
                                   C.c_uint32(pixel_count), C.c_uint32(sample_begin), C.c_uint32(sample_count),
                                   C.c_uint64(base_seed), _p(rays), _p(hits))
        return rays, hits

    def render_seeded(self, cam, width, height, pixel_ids, pixel_begin, pixel_count, sample_begin, min_samples, max_samples,
                      base_seed, sum_only=False, threads=1):
        cam = np.asarray(cam, CAMERA).reshape(1)
        ids = None if pixel_ids is None else np.ascontiguousarray(pixel_ids, np.uint32)
        out = np.zeros((pixel_count, 4), np.float32); ns = np.zeros(pixel_count, np.uint32); cnt = np.zeros(1, COUNTERS)
        sec = self.lib.ref_render_seeded(_p(cam), C.c_uint32(width), C.c_uint32(height), _p(ids), C.c_uint32(pixel_begin),
                                         C.c_uint32(pixel_count), C.c_uint32(sample_begin), C.c_uint32(min_samples),
                                         C.c_uint32(max_samples), C.c_uint64(base_seed), C.c_int(1 if sum_only else 0),
                                         C.c_uint32(threads), _p(out), _p(ns), _p(cnt))
        return out, ns, cnt[0], float(sec)

    def render_ranks(self, cam, width, height, pixel_begin, pixel_count, min_samples, max_samples, threads):
        cam = np.asarray(cam, CAMERA).reshape(1)
        out = np.zeros((pixel_count, 4), np.float32); cnt = np.zeros(1, COUNTERS)
        sec = self.lib.ref_render_ranks(_p(cam), C.c_uint32(width), C.c_uint32(height), C.c_uint32(pixel_begin),
                                        C.c_uint32(pixel_count), C.c_uint32(min_samples), C.c_uint32(max_samples),
                                        C.c_uint32(threads), _p(out), _p(cnt))
        return out, cnt[0], float(sec)

    def tonemap(self, frame: np.ndarray):
        """The reference's WriteFramebufferImage on `frame` ((H, W, 4) float32): returns (RGBA8 image, scene_luma)."""
        import tempfile
        frame = np.ascontiguousarray(frame, np.float32)
        h, w = frame.shape[:2]
        out = np.zeros((h, w, 4), np.uint8)
        path = os.path.join(tempfile.mkdtemp(prefix="ref_png_"), "out.png")
        luma = self.lib.ref_tonemap_png(C.c_uint32(w), C.c_uint32(h), _p(frame), path.encode(), _p(out))
        return out, float(luma)

    def check_jitter_order(self, cam, width, height, x, y, seed) -> bool:
        cam = np.asarray(cam, CAMERA).reshape(1)
        return bool(self.lib.ref_check_jitter_order(_p(cam), C.c_uint32(width), C.c_uint32(height), C.c_uint32(x),
                                                    C.c_uint32(y), C.c_uint64(seed)))


_INSTANCE: Optional[RefHarness] = None


def get() -> RefHarness:
    global _INSTANCE
    if _INSTANCE is None:
        _INSTANCE = RefHarness()
    return _INSTANCE
