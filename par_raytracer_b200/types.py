"""POD mirrors of include/rt_b200.h as numpy structured dtypes + the SceneData container.

Every dtype here is layout-identical to the C struct of the same name (and therefore to the
reference struct it mirrors: Camera main.cpp:133-143, Ray geometry.h:9-12, BoundingSphere
bsphere.cpp:316-320, LightSource scene.h:9-15, RaycastHit raytracer.cpp:20-30).
"""
from __future__ import annotations

import dataclasses
from typing import List, Optional

import numpy as np

RAY = np.dtype([("origin", "<f4", 3), ("direction", "<f4", 3)])
CAMERA = np.dtype([("tan_a2", "<f4"), ("aspect", "<f4"), ("inv_width", "<f4"), ("inv_height", "<f4"),
                   ("position", "<f4", 3), ("forward", "<f4", 3), ("right", "<f4", 3), ("up", "<f4", 3)])
BSPHERE = np.dtype([("center", "<f4", 3), ("radius", "<f4"), ("c0", "<u4"), ("c1", "<u4")])
LIGHT = np.dtype([("type", "<i4"), ("color", "<f4", 4), ("position", "<f4", 3), ("facing", "<f4", 3),
                  ("falloff", "<f4")])
MATERIAL = np.dtype([("specular_intensity", "<f4"), ("index_of_refraction", "<f4"), ("alpha", "<f4"),
                     ("ambient_color", "<f4", 4), ("diffuse_color", "<f4", 4), ("specular_color", "<f4", 4),
                     ("emissive_color", "<f4", 4),
                     ("ambient_texture", "<i4"), ("diffuse_texture", "<i4"), ("specular_texture", "<i4"),
                     ("alpha_texture", "<i4"), ("bump_texture", "<i4")])
COUNTERS = np.dtype([("ray_count", "<u8"), ("sphere_check_count", "<u8"), ("mesh_check_count", "<u8")])
PARAMS = np.dtype([("ray_bias", "<f4"), ("reflection_samples", "<u4"), ("spec_samples", "<u4"),
                   ("bounce_depth", "<u4"), ("background_color", "<f4", 4), ("min_samples", "<u4"),
                   ("max_samples", "<u4"), ("base_seed", "<u8")])
HIT = np.dtype([("t", "<f4"), ("bw", "<f4", 3), ("vertex0", "<u4"), ("position", "<f4", 3), ("normal", "<f4", 3),
                ("object", "<i4"), ("hit", "<u4")])
STATS = np.dtype([("gpu_ms", "<f8"), ("trace_ms", "<f8"), ("shadow_ms", "<f8"), ("logic_ms", "<f8"), ("kernel_launches", "<u8"), ("waves", "<u8"),
                  ("closest_rays", "<u8"), ("shadow_rays", "<u8"), ("h2d_bytes", "<u8"), ("d2h_bytes", "<u8")])

assert RAY.itemsize == 24 and CAMERA.itemsize == 64 and BSPHERE.itemsize == 24 and LIGHT.itemsize == 48
assert MATERIAL.itemsize == 96 and PARAMS.itemsize == 48 and HIT.itemsize == 52 and COUNTERS.itemsize == 24

LIGHT_DIRECTIONAL = 0
LIGHT_POINT = 1

# gRNGInitTable[0] (main.cpp:10): default base seed of the per-(pixel, sample) contract.
DEFAULT_BASE_SEED = 0x835FDD9143716FE3
SEED_MULT = 0x9E3779B97F4A7C15


def sample_seed(base_seed: int, pixel: int, sample: int) -> int:
    """Seed of sample `sample` of linear pixel index `pixel` (rt_params.base_seed contract)."""
    return (base_seed ^ ((pixel * SEED_MULT + sample) & 0xFFFFFFFFFFFFFFFF)) & 0xFFFFFFFFFFFFFFFF


def default_params(spp: int = 10, max_spp: Optional[int] = None, base_seed: int = DEFAULT_BASE_SEED) -> np.ndarray:
    """gParams defaults of InitParams (main.cpp:419-425) + fixed spp (min == max)."""
    p = np.zeros((), dtype=PARAMS)
    p["ray_bias"] = np.float32(1e-3)
    p["reflection_samples"] = 1
    p["spec_samples"] = 1
    p["bounce_depth"] = 2
    bg = np.array([0.8275, 0.8913, 1.0, 1.0], dtype=np.float32) * np.float32(1.5)
    p["background_color"] = bg
    p["min_samples"] = spp
    p["max_samples"] = spp if max_spp is None else max_spp
    p["base_seed"] = base_seed
    return p


def default_lights() -> np.ndarray:
    """InitScene (main.cpp:519-535): light_count = 1, one directional light."""
    l = np.zeros(1, dtype=LIGHT)
    l[0]["type"] = LIGHT_DIRECTIONAL
    l[0]["color"] = np.array([0.9, 1.0, 0.95, 1.0], dtype=np.float32) * np.float32(4.0)
    f = np.array([1.0, -1.5, 0.25], dtype=np.float32)
    l[0]["facing"] = normalize3(f)
    return l


def default_material() -> np.ndarray:
    """MakeMaterial(Vector4(0.75, 0.5, 0.75, 1)) (main.cpp:506-517, 579)."""
    m = np.zeros((), dtype=MATERIAL)
    m["specular_intensity"] = 10.0
    m["index_of_refraction"] = 1.5
    m["alpha"] = 1.0
    m["ambient_color"] = (0.75, 0.5, 0.75, 1.0)
    m["diffuse_color"] = (0.75, 0.5, 0.75, 1.0)
    m["specular_color"] = (1, 1, 1, 1)
    for k in ("ambient_texture", "diffuse_texture", "specular_texture", "alpha_texture", "bump_texture"):
        m[k] = -1
    return m


def normalize3(v: np.ndarray) -> np.ndarray:
    """mathlib.h:253-262 Normalize in float32: Dot left-to-right, three divisions by sqrtf."""
    v = np.asarray(v, dtype=np.float32)
    l2 = np.float32(np.float32(v[0] * v[0]) + np.float32(v[1] * v[1])) + np.float32(v[2] * v[2])
    l2 = np.float32(l2)
    if l2 == 0:
        return v.copy()
    s = np.sqrt(l2, dtype=np.float32)
    return (v / s).astype(np.float32)


def cross3(a, b):
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    return np.array([np.float32(a[1] * b[2]) - np.float32(b[1] * a[2]),
                     np.float32(a[2] * b[0]) - np.float32(b[2] * a[0]),
                     np.float32(a[0] * b[1]) - np.float32(b[0] * a[1])], dtype=np.float32)


def make_camera(fov_deg: float, width: int, height: int, position, facing) -> np.ndarray:
    """MakeCamera (main.cpp:145-162). tanf is taken on the host exactly like the reference does;
    float32 tan here is numpy's, which can differ from glibc tanf in the last bit -- parity tests
    therefore take the camera from the fixtures / pass this same struct to both sides."""
    c = np.zeros((), dtype=CAMERA)
    half = np.float32(np.float32(fov_deg) / np.float32(2.0))
    rad = np.float32(np.float32(half / np.float32(180.0)) * np.float32(3.1415927))
    c["tan_a2"] = np.tan(rad, dtype=np.float32)
    c["aspect"] = np.float32(width) / np.float32(height)
    c["inv_width"] = np.float32(1.0) / np.float32(width)
    c["inv_height"] = np.float32(1.0) / np.float32(height)
    c["position"] = np.asarray(position, dtype=np.float32)
    fwd = normalize3(np.asarray(facing, dtype=np.float32))
    up = np.array([0, 1, 0], dtype=np.float32)
    right = normalize3(cross3(fwd, up))
    c["forward"] = fwd
    c["right"] = right
    c["up"] = normalize3(cross3(right, fwd))
    return c


@dataclasses.dataclass
class TextureData:
    size_x: int
    size_y: int
    channels: int
    texels: np.ndarray  # uint8, size_y * size_x * channels, row-major

    def __post_init__(self):
        self.texels = np.ascontiguousarray(self.texels, dtype=np.uint8).reshape(-1)
        assert self.texels.size == self.size_x * self.size_y * self.channels


@dataclasses.dataclass
class SceneData:
    """Host-side flattened scene == rt_scene_desc (include/rt_b200.h)."""
    positions: np.ndarray          # (P, 3) f32
    texcoords: np.ndarray          # (T, 2) f32
    normals: np.ndarray            # (N, 3) f32
    tangents: Optional[np.ndarray]  # (N, 3) f32 or None
    group_first: np.ndarray        # (G + 1,) u32
    idx_positions: np.ndarray      # (I,) u32
    idx_texcoords: np.ndarray
    idx_normals: np.ndarray
    group_material: np.ndarray     # (G,) i32
    spheres: np.ndarray            # (S,) BSPHERE
    sphere_group: np.ndarray       # (S,) i32
    materials: np.ndarray          # (M,) MATERIAL
    default_material: np.ndarray   # () MATERIAL
    textures: List[TextureData]
    lights: np.ndarray             # (L,) LIGHT
    name: str = "scene"

    def __post_init__(self):
        self.positions = np.ascontiguousarray(self.positions, dtype=np.float32).reshape(-1, 3)
        self.texcoords = np.ascontiguousarray(self.texcoords, dtype=np.float32).reshape(-1, 2)
        self.normals = np.ascontiguousarray(self.normals, dtype=np.float32).reshape(-1, 3)
        if self.tangents is not None:
            self.tangents = np.ascontiguousarray(self.tangents, dtype=np.float32).reshape(-1, 3)
        self.group_first = np.ascontiguousarray(self.group_first, dtype=np.uint32)
        self.idx_positions = np.ascontiguousarray(self.idx_positions, dtype=np.uint32)
        self.idx_texcoords = np.ascontiguousarray(self.idx_texcoords, dtype=np.uint32)
        self.idx_normals = np.ascontiguousarray(self.idx_normals, dtype=np.uint32)
        self.group_material = np.ascontiguousarray(self.group_material, dtype=np.int32)
        self.spheres = np.ascontiguousarray(self.spheres, dtype=BSPHERE)
        self.sphere_group = np.ascontiguousarray(self.sphere_group, dtype=np.int32)
        self.materials = np.ascontiguousarray(self.materials, dtype=MATERIAL).reshape(-1)
        self.default_material = np.asarray(self.default_material, dtype=MATERIAL).reshape(())
        self.lights = np.ascontiguousarray(self.lights, dtype=LIGHT).reshape(-1)

    @property
    def n_groups(self) -> int:
        return int(self.group_first.size - 1)

    @property
    def n_triangles(self) -> int:
        return int(self.idx_positions.size // 3)

    def validate(self) -> None:
        assert self.group_first[0] == 0 and self.group_first[-1] == self.idx_positions.size
        assert self.idx_positions.size == self.idx_texcoords.size == self.idx_normals.size
        assert np.all(np.diff(self.group_first.astype(np.int64)) % 3 == 0)
        if self.idx_positions.size:
            assert self.idx_positions.max() < len(self.positions)
            assert self.idx_texcoords.max() < len(self.texcoords)
            assert self.idx_normals.max() < len(self.normals)
        assert self.spheres.size == self.sphere_group.size
        if self.spheres.size:                        # empty: no reference hierarchy (tie-break rank = group order)
            leaves = self.sphere_group[self.sphere_group >= 0]
            assert sorted(leaves.tolist()) == list(range(self.n_groups)), "every group must be exactly one leaf"
        assert np.all(self.group_material < len(self.materials))

    # ---- (de)serialisation for tests/golden fixtures -------------------------------------------
    def to_npz_dict(self, prefix: str = "scene_") -> dict:
        d = {
            "positions": self.positions, "texcoords": self.texcoords, "normals": self.normals,
            "group_first": self.group_first, "idx_positions": self.idx_positions,
            "idx_texcoords": self.idx_texcoords, "idx_normals": self.idx_normals,
            "group_material": self.group_material, "spheres": self.spheres.view(np.uint8),
            "sphere_group": self.sphere_group, "materials": self.materials.view(np.uint8),
            "default_material": np.frombuffer(self.default_material.tobytes(), dtype=np.uint8),
            "lights": self.lights.view(np.uint8),
            "tex_info": np.array([[t.size_x, t.size_y, t.channels] for t in self.textures], dtype=np.uint32).reshape(-1, 3),
        }
        if self.tangents is not None:
            d["tangents"] = self.tangents
        for i, t in enumerate(self.textures):
            d[f"tex_{i}"] = t.texels
        return {prefix + k: v for k, v in d.items()}

    @staticmethod
    def from_npz_dict(z, prefix: str = "scene_", name: str = "scene") -> "SceneData":
        g = lambda k: z[prefix + k]
        tex_info = g("tex_info")
        textures = [TextureData(int(tex_info[i, 0]), int(tex_info[i, 1]), int(tex_info[i, 2]), g(f"tex_{i}"))
                    for i in range(tex_info.shape[0])]
        tangents = z[prefix + "tangents"] if (prefix + "tangents") in z else None
        return SceneData(
            positions=g("positions"), texcoords=g("texcoords"), normals=g("normals"), tangents=tangents,
            group_first=g("group_first"), idx_positions=g("idx_positions"), idx_texcoords=g("idx_texcoords"),
            idx_normals=g("idx_normals"), group_material=g("group_material"),
            spheres=np.frombuffer(g("spheres").tobytes(), dtype=BSPHERE),
            sphere_group=g("sphere_group"),
            materials=np.frombuffer(g("materials").tobytes(), dtype=MATERIAL),
            default_material=np.frombuffer(g("default_material").tobytes(), dtype=MATERIAL)[0],
            textures=textures, lights=np.frombuffer(g("lights").tobytes(), dtype=LIGHT), name=name)
