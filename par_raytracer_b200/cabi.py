"""ctypes declarations of include/rt_b200.h (struct layouts only; no library is loaded here)."""
from __future__ import annotations

import ctypes as C
from typing import List, Tuple

import numpy as np

from .types import SceneData

c_f = C.c_float
c_u32 = C.c_uint32
c_i32 = C.c_int32
c_u64 = C.c_uint64
P = C.POINTER


class RtTexture(C.Structure):
    _fields_ = [("size_x", c_u32), ("size_y", c_u32), ("channels", c_u32), ("texels", C.c_void_p)]


class RtMaterial(C.Structure):
    _fields_ = [("specular_intensity", c_f), ("index_of_refraction", c_f), ("alpha", c_f),
                ("ambient_color", c_f * 4), ("diffuse_color", c_f * 4), ("specular_color", c_f * 4),
                ("emissive_color", c_f * 4),
                ("ambient_texture", c_i32), ("diffuse_texture", c_i32), ("specular_texture", c_i32),
                ("alpha_texture", c_i32), ("bump_texture", c_i32)]


class RtSceneDesc(C.Structure):
    _fields_ = [
        ("n_positions", c_u32), ("positions", C.c_void_p),
        ("n_texcoords", c_u32), ("texcoords", C.c_void_p),
        ("n_normals", c_u32), ("normals", C.c_void_p),
        ("tangents", C.c_void_p),
        ("n_groups", c_u32), ("group_first", C.c_void_p),
        ("idx_positions", C.c_void_p), ("idx_texcoords", C.c_void_p), ("idx_normals", C.c_void_p),
        ("group_material", C.c_void_p),
        ("n_spheres", c_u32), ("spheres", C.c_void_p), ("sphere_group", C.c_void_p),
        ("n_materials", c_u32), ("materials", C.c_void_p),
        ("default_material", RtMaterial),
        ("n_textures", c_u32), ("textures", C.c_void_p),
        ("n_lights", c_u32), ("lights", C.c_void_p),
    ]


assert C.sizeof(RtTexture) == 24 and C.sizeof(RtMaterial) == 96


def ptr(a) -> C.c_void_p:
    if a is None:
        return C.c_void_p(None)
    return C.c_void_p(a.ctypes.data)


def make_scene_desc(scene: SceneData) -> Tuple[RtSceneDesc, List[object]]:
    """Builds rt_scene_desc pointing INTO the numpy arrays of `scene` (no copies). The returned keep-alive
    list must outlive every use of the descriptor."""
    keep: List[object] = [scene]
    d = RtSceneDesc()
    d.n_positions = len(scene.positions); d.positions = ptr(scene.positions)
    d.n_texcoords = len(scene.texcoords); d.texcoords = ptr(scene.texcoords)
    d.n_normals = len(scene.normals); d.normals = ptr(scene.normals)
    d.tangents = ptr(scene.tangents)
    d.n_groups = scene.n_groups
    d.group_first = ptr(scene.group_first)
    d.idx_positions = ptr(scene.idx_positions)
    d.idx_texcoords = ptr(scene.idx_texcoords)
    d.idx_normals = ptr(scene.idx_normals)
    d.group_material = ptr(scene.group_material)
    d.n_spheres = len(scene.spheres); d.spheres = ptr(scene.spheres); d.sphere_group = ptr(scene.sphere_group)
    d.n_materials = len(scene.materials); d.materials = ptr(scene.materials)
    dm = np.ascontiguousarray(scene.default_material).reshape(1)
    keep.append(dm)
    C.memmove(C.byref(d.default_material), dm.ctypes.data, 96)
    tex = (RtTexture * max(1, len(scene.textures)))()
    for i, t in enumerate(scene.textures):
        tex[i].size_x = t.size_x; tex[i].size_y = t.size_y; tex[i].channels = t.channels
        tex[i].texels = t.texels.ctypes.data
    keep.append(tex)
    d.n_textures = len(scene.textures); d.textures = C.cast(tex, C.c_void_p)
    d.n_lights = len(scene.lights); d.lights = ptr(scene.lights)
    return d, keep
