"""Host-side mirror of the reference's render entry points over the C ABI (include/rt_b200.h).

    Scene(scene_data)                          <-> main.cpp:546-599 (scene set-up; arrays stay the caller's)
    Scene.render(cam, params, w, h)            <-> Render(Camera*, Scene*, u32, u32) -> Framebuffer   main.cpp:301-358
    Scene.render_task(cam, params, w, h, a, b) <-> RenderTask(RenderJob*) over pixel range [a, b)     main.cpp:267-283
    Scene.trace_rays(params, rays)             <-> TraceRay                                           raytracer.cpp:159-232
    Scene.trace_color(params, rays, seeds)     <-> TraceRayColor                                      raytracer.cpp:413-577

The shared library is the product; this module only marshals numpy arrays into it. If the library is
missing or no GPU is present every compute call raises -- there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from .cabi import RtSceneDesc, make_scene_desc
from .types import CAMERA, COUNTERS, HIT, PARAMS, RAY, STATS, SceneData

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RT_B200_LIB") or os.path.join(_HERE, "librt_b200.so")

RT_OUT_MEAN, RT_OUT_SUM, RT_OUT_FULLFRAME, RT_FLAG_COUNTERS, RT_FLAG_TIME_KERNELS, RT_FLAG_ADAPTIVE, RT_FLAG_PIN_HOST = 0, 1, 2, 4, 8, 16, 32
RT_TRACE_CLOSEST, RT_TRACE_ANY, RT_TRACE_BRUTE = 0, 1, 2

EXPORTS = ["rt_scene_create", "rt_scene_destroy", "rt_last_error", "rt_abi_version", "rt_render", "rt_render_device",
           "rt_trace_rays", "rt_trace_primary", "rt_trace_color", "rt_get_stats", "rt_get_hierarchy_info", "rt_rng_kat",
           "rt_get_sample_counts", "rt_tonemap_device", "rt_tonemap", "rt_build_group_hierarchy",
           "rt_calculate_tangents", "rt_height_to_normal_map", "rt_quant_grid",
           "rt_comm_unique_id", "rt_comm_create", "rt_comm_create_local", "rt_comm_destroy", "rt_comm_rank", "rt_comm_size",
           "rt_partition_tiles", "rt_device_count", "rt_render_combined", "rt_render_multi", "rt_comm_frame", "rt_comm_get_stats"]

RT_PART_TILES, RT_PART_RANGES, RT_PART_SAMPLES = 0, 1, 2
PARTITIONS = {"tiles": RT_PART_TILES, "ranges": RT_PART_RANGES, "samples": RT_PART_SAMPLES}
RT_COMM_ID_BYTES = 128


class RtError(RuntimeError):
    pass


_LIB = None


def load_library():
    """Loads librt_b200.so; raises if it has not been built (no fallback)."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise RtError(f"{LIB_PATH} is missing: run `python -m par_raytracer_b200.build` (nvcc, sm_100a). "
                          "There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        L.rt_last_error.restype = C.c_char_p
        L.rt_scene_create.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.rt_scene_destroy.argtypes = [C.c_void_p]
        L.rt_scene_destroy.restype = None
        L.rt_comm_destroy.argtypes = [C.c_void_p]
        L.rt_comm_destroy.restype = None
        L.rt_comm_create.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_void_p)]
        L.rt_comm_create_local.argtypes = [C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        L.rt_comm_frame.argtypes = [C.c_void_p]
        L.rt_comm_frame.restype = C.c_void_p
        L.rt_comm_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.rt_comm_rank.argtypes = [C.c_void_p]
        L.rt_comm_size.argtypes = [C.c_void_p]
        _LIB = L
    return _LIB


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _check(rc: int, what: str):
    if rc != 0:
        raise RtError(f"{what} failed ({rc}): {load_library().rt_last_error().decode()}")


def rng_kat(seed: int, n: int, device: int = 0) -> np.ndarray:
    out = np.zeros(n, np.uint64)
    _check(load_library().rt_rng_kat(C.c_int(device), C.c_uint64(seed), C.c_uint32(n), _p(out)), "rt_rng_kat")
    return out


def build_group_hierarchy(scene: SceneData, device: int = 0):
    """BuildHierarchy (bsphere.cpp:379-444) on the GPU: returns (spheres, sphere_group) bit-identical to the reference's."""
    from .types import BSPHERE
    G = scene.n_groups
    spheres = np.zeros(max(1, 2 * G - 1), BSPHERE); sg = np.zeros(max(1, 2 * G - 1), np.int32)
    n = C.c_uint32(0)
    _check(load_library().rt_build_group_hierarchy(C.c_int(device), _p(scene.positions), C.c_uint32(len(scene.positions)), C.c_uint32(G),
                                                   _p(scene.group_first), _p(scene.idx_positions), _p(spheres), _p(sg), C.byref(n)),
           "rt_build_group_hierarchy")
    return spheres[:n.value], sg[:n.value]


def calculate_tangents(scene: SceneData, device: int = 0) -> np.ndarray:
    """CalculateTangents (mesh.h:59-129) on the GPU, bit-identical to the reference: (n_normals, 3) float32."""
    hb = np.array([1 if (m >= 0 and scene.materials[m]["bump_texture"] >= 0) else 0 for m in scene.group_material], np.uint8)
    out = np.zeros((len(scene.normals), 3), np.float32)
    _check(load_library().rt_calculate_tangents(C.c_int(device), _p(scene.positions), C.c_uint32(len(scene.positions)), _p(scene.texcoords),
                                                C.c_uint32(len(scene.texcoords)), C.c_uint32(len(scene.normals)), C.c_uint32(scene.n_groups),
                                                _p(scene.group_first), _p(scene.idx_positions), _p(scene.idx_texcoords), _p(scene.idx_normals),
                                                _p(hb), _p(out)), "rt_calculate_tangents")
    return out


def height_to_normal_map(height: np.ndarray, device: int = 0) -> np.ndarray:
    """ConvertHeightMapToNormalMap (texture.cpp:102-144) on the GPU: (H, W) uint8 -> (H, W, 3) uint8."""
    height = np.ascontiguousarray(height, np.uint8)
    h, w = height.shape
    out = np.zeros((h, w, 3), np.uint8)
    _check(load_library().rt_height_to_normal_map(C.c_int(device), C.c_uint32(w), C.c_uint32(h), _p(height), _p(out)), "rt_height_to_normal_map")
    return out


def tonemap(frame: np.ndarray, device: int = 0):
    """WriteFramebufferImage's tone map + Color_Pack (main.cpp:101-127) on the GPU: (H, W, 4) float32 -> (H, W, 4) uint8, scene_luma."""
    frame = np.ascontiguousarray(frame, np.float32)
    h, w = frame.shape[:2]
    out = np.zeros((h, w, 4), np.uint8)
    luma = C.c_float(0)
    _check(load_library().rt_tonemap(C.c_int(device), _p(frame), C.c_uint32(w), C.c_uint32(h), _p(out), C.byref(luma)), "rt_tonemap")
    return out, float(luma.value)


def tonemap_device(frame_ptr: int, width: int, height: int, out_ptr: int, device: int = 0, stream: int = 0) -> float:
    """Same on device buffers (e.g. torch tensors' data_ptr()): float4 frame in, RGBA8 out. Returns scene_luma."""
    luma = C.c_float(0)
    _check(load_library().rt_tonemap_device(C.c_int(device), C.c_void_p(frame_ptr), C.c_uint32(width), C.c_uint32(height),
                                            C.c_void_p(out_ptr), C.byref(luma), C.c_void_p(stream)), "rt_tonemap_device")
    return float(luma.value)


class Scene:
    """Device-resident scene + GPU-built cluster hierarchy (rt_scene)."""

    def __init__(self, scene: SceneData, device: int = 0):
        self.lib = load_library()
        self.data = scene
        desc, keep = make_scene_desc(scene)
        h = C.c_void_p()
        _check(self.lib.rt_scene_create(C.addressof(desc), device, C.byref(h)), "rt_scene_create")
        self.h = h
        self.device = device
        del keep

    def close(self):
        if getattr(self, "h", None):
            self.lib.rt_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- introspection ----------------------------------------------------------------------
    def stats(self) -> np.ndarray:
        s = np.zeros(1, STATS)
        _check(self.lib.rt_get_stats(self.h, _p(s)), "rt_get_stats")
        return s[0]

    def hierarchy_info(self) -> dict:
        o = np.zeros(8, np.uint64)
        _check(self.lib.rt_get_hierarchy_info(self.h, _p(o)), "rt_get_hierarchy_info")
        return dict(triangles=int(o[0]), nodes=int(o[2]), depth=int(o[3]), node_bytes=int(o[4]), triangle_bytes=int(o[5]),
                    build_us=int(o[6]), build_iterations=int(o[7]))

    def sample_counts(self, n: int) -> np.ndarray:
        """Per-pixel sample counts of the last adaptive render (RenderPixel's final `samp`)."""
        out = np.zeros(n, np.uint32)
        _check(self.lib.rt_get_sample_counts(self.h, _p(out), C.c_uint32(n)), "rt_get_sample_counts")
        return out

    # ---- Render / RenderTask ----------------------------------------------------------------
    def render_task(self, cam, params, width, height, pixel_begin=0, pixel_count=None, pixel_ids=None, sample_begin=0,
                    sample_count=None, flags=RT_OUT_MEAN, out=None):
        cam = np.asarray(cam, CAMERA).reshape(1); params = np.asarray(params, PARAMS).reshape(1)
        ids = None if pixel_ids is None else np.ascontiguousarray(pixel_ids, np.uint32)
        if pixel_count is None:
            pixel_count = len(ids) if ids is not None else width * height - pixel_begin
        if sample_count is None:
            sample_count = int(params[0]["min_samples"])
        if out is None:
            out = np.zeros((pixel_count, 4), np.float32)
        assert out.dtype == np.float32 and out.size == pixel_count * 4 and out.flags["C_CONTIGUOUS"]
        cnt = np.zeros(1, COUNTERS)
        _check(self.lib.rt_render(self.h, _p(cam), _p(params), C.c_uint32(width), C.c_uint32(height), _p(ids),
                                  C.c_uint32(pixel_begin), C.c_uint32(pixel_count), C.c_uint32(sample_begin),
                                  C.c_uint32(sample_count), C.c_uint32(flags), _p(out), _p(cnt)), "rt_render")
        return out, cnt[0]

    def render(self, cam, params, width, height, flags=RT_OUT_MEAN):
        """Render(cam, scene, w, h): the whole frame, (h, w, 4) float32 linear HDR, w = 1."""
        out, cnt = self.render_task(cam, params, width, height, 0, width * height, flags=flags)
        return out.reshape(height, width, 4), cnt

    def render_device(self, cam, params, width, height, out_ptr: int, pixel_begin=0, pixel_count=None, pixel_ids=None,
                      sample_begin=0, sample_count=None, flags=RT_OUT_MEAN, stream: int = 0):
        """Output left in device memory at `out_ptr` (e.g. a torch tensor's data_ptr())."""
        cam = np.asarray(cam, CAMERA).reshape(1); params = np.asarray(params, PARAMS).reshape(1)
        ids = None if pixel_ids is None else np.ascontiguousarray(pixel_ids, np.uint32)
        if pixel_count is None:
            pixel_count = len(ids) if ids is not None else width * height - pixel_begin
        if sample_count is None:
            sample_count = int(params[0]["min_samples"])
        cnt = np.zeros(1, COUNTERS)
        _check(self.lib.rt_render_device(self.h, _p(cam), _p(params), C.c_uint32(width), C.c_uint32(height), _p(ids),
                                         C.c_uint32(pixel_begin), C.c_uint32(pixel_count), C.c_uint32(sample_begin),
                                         C.c_uint32(sample_count), C.c_uint32(flags), C.c_void_p(out_ptr), C.c_void_p(stream),
                                         _p(cnt)), "rt_render_device")
        return cnt[0]

    # ---- TraceRay / TraceRayColor -------------------------------------------------------------
    def trace_rays(self, params, rays, mode=RT_TRACE_CLOSEST):
        rays = np.ascontiguousarray(rays, RAY); params = np.asarray(params, PARAMS).reshape(1)
        out = np.zeros(len(rays), HIT); cnt = np.zeros(1, COUNTERS)
        _check(self.lib.rt_trace_rays(self.h, _p(params), _p(rays), C.c_uint64(len(rays)), C.c_int(mode), _p(out), _p(cnt)),
               "rt_trace_rays")
        return out, cnt[0]

    def trace_primary(self, cam, params, width, height, pixel_ids=None, pixel_begin=0, pixel_count=None, sample_begin=0,
                      sample_count=1, want_rays=True, want_hits=True):
        cam = np.asarray(cam, CAMERA).reshape(1); params = np.asarray(params, PARAMS).reshape(1)
        ids = None if pixel_ids is None else np.ascontiguousarray(pixel_ids, np.uint32)
        if pixel_count is None:
            pixel_count = len(ids) if ids is not None else width * height - pixel_begin
        n = pixel_count * sample_count
        rays = np.zeros(n, RAY) if want_rays else None
        hits = np.zeros(n, HIT) if want_hits else None
        _check(self.lib.rt_trace_primary(self.h, _p(cam), _p(params), C.c_uint32(width), C.c_uint32(height), _p(ids),
                                         C.c_uint32(pixel_begin), C.c_uint32(pixel_count), C.c_uint32(sample_begin),
                                         C.c_uint32(sample_count), _p(rays), _p(hits)), "rt_trace_primary")
        return rays, hits

    def trace_color(self, params, rays, seeds):
        rays = np.ascontiguousarray(rays, RAY); seeds = np.ascontiguousarray(seeds, np.uint64)
        params = np.asarray(params, PARAMS).reshape(1)
        out = np.zeros((len(rays), 4), np.float32); cnt = np.zeros(1, COUNTERS)
        _check(self.lib.rt_trace_color(self.h, _p(params), _p(rays), _p(seeds), C.c_uint64(len(rays)), _p(out), _p(cnt)),
               "rt_trace_color")
        return out, cnt[0]


# ---- Render's partition + MPI_Gather on N GPUs (main.cpp:311-319, 345-347) -------------------------------------------
def partition_tiles(width: int, height: int, rank: int, world: int, tile: int = 32) -> np.ndarray:
    """rt_partition_tiles (host-only): linear pixel ids of the interleaved tiles `rank` owns."""
    L = load_library()
    n = C.c_uint32(0)
    _check(L.rt_partition_tiles(C.c_uint32(width), C.c_uint32(height), C.c_uint32(tile), C.c_int(rank), C.c_int(world), None, C.byref(n)),
           "rt_partition_tiles")
    ids = np.zeros(n.value, np.uint32)
    _check(L.rt_partition_tiles(C.c_uint32(width), C.c_uint32(height), C.c_uint32(tile), C.c_int(rank), C.c_int(world), _p(ids), C.byref(n)),
           "rt_partition_tiles")
    return ids


def comm_unique_id() -> bytes:
    """ncclGetUniqueId through the library (one rank calls it, the host distributes the 128 bytes)."""
    buf = (C.c_uint8 * RT_COMM_ID_BYTES)()
    _check(load_library().rt_comm_unique_id(buf), "rt_comm_unique_id")
    return bytes(buf)


class Comm:
    """One rank's rt_comm: NCCL communicator (or peer-memory group member) + the combine buffers."""

    def __init__(self, handle, lib):
        self.h, self.lib = handle, lib

    @classmethod
    def create(cls, world: int, rank: int, unique_id: Optional[bytes], device: int) -> "Comm":
        L = load_library()
        h = C.c_void_p()
        idbuf = (C.c_uint8 * RT_COMM_ID_BYTES).from_buffer_copy(unique_id) if unique_id is not None else None
        _check(L.rt_comm_create(world, rank, idbuf, device, C.byref(h)), "rt_comm_create")
        return cls(h, L)

    @classmethod
    def create_local(cls, devices) -> list:
        """One process, len(devices) GPUs: [Comm of rank 0, ...] (rt_comm_create_local)."""
        L = load_library()
        n = len(devices)
        dev = (C.c_int * n)(*devices)
        hs = (C.c_void_p * n)()
        _check(L.rt_comm_create_local(n, dev, hs), "rt_comm_create_local")
        return [cls(C.c_void_p(hs[i]), L) for i in range(n)]

    @property
    def rank(self) -> int:
        return int(self.lib.rt_comm_rank(self.h))

    @property
    def size(self) -> int:
        return int(self.lib.rt_comm_size(self.h))

    def frame_ptr(self) -> int:
        return int(self.lib.rt_comm_frame(self.h) or 0)

    def stats(self) -> dict:
        o = np.zeros(4, np.float64)
        _check(self.lib.rt_comm_get_stats(self.h, _p(o)), "rt_comm_get_stats")
        return dict(combine_ms=float(o[0]), deliver_ms=float(o[1]), reduce_bytes=float(o[2]), peer_memory=bool(o[3]))

    def close(self):
        if getattr(self, "h", None):
            self.lib.rt_comm_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def render_combined(scene: Scene, comm: Comm, cam, params, width: int, height: int, partition: str = "tiles", tile: int = 32,
                    flags: int = 0, root: int = 0, want_frame: bool = False, want_rgba8: bool = False, out=None):
    """rt_render_combined: this rank's share of Render() + the NCCL combine. Returns (frame or None, rgba8 or None, scene_luma or
    None, counters); the outputs are filled on `root` only."""
    cam = np.asarray(cam, CAMERA).reshape(1); params = np.asarray(params, PARAMS).reshape(1)
    is_root = comm.rank == root
    if not (want_frame and is_root):
        out = None
    elif out is None:
        out = np.empty((height, width, 4), np.float32)
    out8 = np.empty((height, width, 4), np.uint8) if (want_rgba8 and is_root) else None
    luma = C.c_float(0)
    cnt = np.zeros(1, COUNTERS)
    _check(scene.lib.rt_render_combined(scene.h, comm.h, _p(cam), _p(params), C.c_uint32(width), C.c_uint32(height),
                                        C.c_int(PARTITIONS[partition]), C.c_uint32(tile), C.c_uint32(flags), C.c_int(root), _p(out), _p(out8),
                                        C.byref(luma) if out8 is not None else None, _p(cnt)), "rt_render_combined")
    return out, out8, (float(luma.value) if out8 is not None else None), cnt[0]


def render_multi(scenes, comms, cam, params, width: int, height: int, partition: str = "tiles", tile: int = 32, flags: int = 0,
                 want_rgba8: bool = False):
    """rt_render_multi: Render() on len(scenes) GPUs of this process. Returns (frame (H, W, 4), rgba8 or None, scene_luma or None, counters)."""
    cam = np.asarray(cam, CAMERA).reshape(1); params = np.asarray(params, PARAMS).reshape(1)
    n = len(scenes)
    sh = (C.c_void_p * n)(*[s.h for s in scenes]); ch = (C.c_void_p * n)(*[c.h for c in comms])
    out = np.empty((height, width, 4), np.float32)
    out8 = np.empty((height, width, 4), np.uint8) if want_rgba8 else None
    luma = C.c_float(0)
    cnt = np.zeros(1, COUNTERS)
    _check(scenes[0].lib.rt_render_multi(sh, ch, C.c_int(n), _p(cam), _p(params), C.c_uint32(width), C.c_uint32(height),
                                         C.c_int(PARTITIONS[partition]), C.c_uint32(tile), C.c_uint32(flags), _p(out), _p(out8),
                                         C.byref(luma) if want_rgba8 else None, _p(cnt)), "rt_render_multi")
    return out, out8, (float(luma.value) if want_rgba8 else None), cnt[0]
