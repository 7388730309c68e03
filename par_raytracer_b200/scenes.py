"""Procedural scenes for the BASELINE.json configs, as flattened SceneData (+ OBJ/MTL/PNG export).

The reference renders only what obj_parser.cpp loads (no analytic spheres), so
every scene here is a triangle mesh in `g`-delimited groups with p/t/n faces, consistently wound
(the reference's triangle test is one-sided, raytracer.cpp:92-95), one material per group, written
out in exactly the dialect obj_parser.cpp accepts (obj_parser.cpp:348-426).

Host-side scene preparation only -- nothing here is on the render hot path.
"""
from __future__ import annotations

import os
import struct
import zlib
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .types import (BSPHERE, MATERIAL, SceneData, TextureData, default_lights, default_material)

F32 = np.float32


# --------------------------------------------------------------------------------------------
# group hierarchy in the reference's flattened format (bsphere.cpp:316-350)
# --------------------------------------------------------------------------------------------

def _enclose(c0, r0, c1, r1):
    """Smallest sphere around two spheres (same construction idea as bsphere.cpp:248-279), float64
    internally, padded so float32 storage still encloses both children."""
    d = np.linalg.norm(c1 - c0)
    if d + r1 <= r0:
        return c0.copy(), r0 * 1.0001
    if d + r0 <= r1:
        return c1.copy(), r1 * 1.0001
    r = 0.5 * (d + r0 + r1)
    c = c0 + (c1 - c0) * ((r - r0) / d) if d > 1e-12 else c0.copy()
    return c, r * 1.0001 + 1e-4


def build_group_hierarchy(positions: np.ndarray, group_first: np.ndarray, idx_positions: np.ndarray
                          ) -> Tuple[np.ndarray, np.ndarray]:
    """A valid BoundingHierarchy over mesh groups: leaf sphere per group, binary tree by recursive
    median split of leaf centres, flattened pre-order with child index 0 as the leaf sentinel.

    This is NOT the reference's greedy O(n^3) agglomeration (bsphere.cpp:379-428); it produces the
    same data structure (any valid hierarchy yields the same closest hits; only the leaf visit order
    -- used for the equal-t tie-break -- depends on it)."""
    G = len(group_first) - 1
    centers = np.zeros((G, 3), dtype=np.float64)
    radii = np.zeros(G, dtype=np.float64)
    for g in range(G):
        idx = idx_positions[group_first[g]:group_first[g + 1]]
        if idx.size == 0:                       # an empty group still owns a leaf (main.cpp:587 allows it)
            radii[g] = 1e-3
            continue
        p = positions[idx].astype(np.float64)
        lo, hi = p.min(axis=0), p.max(axis=0)
        c = 0.5 * (lo + hi)
        c = c.astype(np.float32).astype(np.float64)
        r = np.sqrt(((p - c) ** 2).sum(axis=1).max())
        centers[g] = c
        radii[g] = r * 1.0001 + 1e-3
    spheres: List[Tuple[np.ndarray, float, int, int]] = []
    sphere_group: List[int] = []

    def rec(ids: np.ndarray) -> Tuple[int, np.ndarray, float]:
        my = len(spheres)
        spheres.append(None)  # type: ignore
        sphere_group.append(-1)
        if len(ids) == 1:
            g = int(ids[0])
            spheres[my] = (centers[g], radii[g], 0, 0)
            sphere_group[my] = g
            return my, centers[g], radii[g]
        c = centers[ids]
        axis = int(np.argmax(c.max(axis=0) - c.min(axis=0)))
        order = ids[np.argsort(c[:, axis], kind="stable")]
        half = len(order) // 2
        i0, c0, r0 = rec(order[:half])
        i1, c1, r1 = rec(order[half:])
        cc, rr = _enclose(c0, r0, c1, r1)
        cc = cc.astype(np.float32).astype(np.float64)
        rr = max(rr, np.linalg.norm(cc - c0) + r0, np.linalg.norm(cc - c1) + r1) * 1.00001
        spheres[my] = (cc, rr, i0, i1)
        return my, cc, rr

    import sys
    sys.setrecursionlimit(max(10000, sys.getrecursionlimit()))
    rec(np.arange(G))
    out = np.zeros(len(spheres), dtype=BSPHERE)
    for i, (c, r, a, b) in enumerate(spheres):
        out[i]["center"] = c.astype(np.float32)
        out[i]["radius"] = np.nextafter(np.float32(r), np.float32(np.inf))
        out[i]["c0"] = a
        out[i]["c1"] = b
    return out, np.asarray(sphere_group, dtype=np.int32)


# --------------------------------------------------------------------------------------------
# load-time preprocessing the reference does on the host (out of scope of the hot path; restated
# in numpy only so that python-generated scenes carry the same kind of inputs)
# --------------------------------------------------------------------------------------------

def calculate_tangents(positions, texcoords, normals, group_first, idx_p, idx_t, idx_n, group_has_bump) -> np.ndarray:
    """mesh.h:59-129 CalculateTangents: per-triangle UV-delta tangents accumulated on the NORMAL index."""
    tang = np.zeros_like(normals, dtype=np.float32)
    for g in range(len(group_first) - 1):
        if not group_has_bump[g]:
            continue
        a, b = int(group_first[g]), int(group_first[g + 1])
        ip = idx_p[a:b].reshape(-1, 3)
        it = idx_t[a:b].reshape(-1, 3)
        inn = idx_n[a:b].reshape(-1, 3)
        p0, p1, p2 = positions[ip[:, 0]], positions[ip[:, 1]], positions[ip[:, 2]]
        uv0, uv1, uv2 = texcoords[it[:, 0]], texcoords[it[:, 1]], texcoords[it[:, 2]]
        dp0, dp1 = (p1 - p0).astype(F32), (p2 - p0).astype(F32)
        d0, d1 = (uv1 - uv0).astype(F32), (uv2 - uv0).astype(F32)
        f = (d0[:, 0] * d1[:, 1]).astype(F32) - (d1[:, 0] * d0[:, 1]).astype(F32)
        ok = f > 1e-7
        with np.errstate(divide="ignore", invalid="ignore"):
            inv = (F32(1.0) / f).astype(F32)
            t = (inv[:, None] * ((d1[:, 1:2] * dp0).astype(F32) - (d0[:, 1:2] * dp1).astype(F32))).astype(F32)
        for k in range(3):
            np.add.at(tang, inn[ok, k], t[ok])
    l2 = ((tang[:, 0] * tang[:, 0]).astype(F32) + (tang[:, 1] * tang[:, 1]).astype(F32)).astype(F32) + (tang[:, 2] * tang[:, 2]).astype(F32)
    nz = l2 != 0
    tang[nz] = (tang[nz] / np.sqrt(l2[nz], dtype=F32)[:, None]).astype(F32)
    return tang


def _srgb_to_linear(x: np.ndarray) -> np.ndarray:
    x = x.astype(F32)
    return np.where(x <= F32(0.04045), x / F32(12.92), np.power((x + F32(0.055)) / F32(1.055), F32(2.4), dtype=F32)).astype(F32)


def _linear_to_srgb(x: np.ndarray) -> np.ndarray:
    x = x.astype(F32)
    with np.errstate(invalid="ignore"):
        return np.where(x <= F32(0.0031308), F32(12.92) * x,
                        F32(1.055) * np.power(np.maximum(x, 0), F32(1.0 / 2.4), dtype=F32) - F32(0.055)).astype(F32)


def height_to_normal_map(height: np.ndarray) -> np.ndarray:
    """texture.cpp:102-144 ConvertHeightMapToNormalMap on a (H, W) uint8 map -> (H, W, 3) uint8
    (normals stored sRGB-encoded, texture.cpp:96-99). numpy's powf may differ from glibc's in the last
    bit at truncation boundaries; fixtures that must match the reference come from oracle/_ref instead."""
    h = _srgb_to_linear(height.astype(F32) * F32(1.0 / 255.0))
    h10 = np.roll(h, -1, axis=1)
    h01 = np.roll(h, -1, axis=0)
    a = F32(2.5)
    nx = ((h01 - h) * a).astype(F32)
    ny = ((h10 - h) * a).astype(F32)
    nz = np.ones_like(nx)
    l = np.sqrt((nx * nx + ny * ny).astype(F32) + nz * nz, dtype=F32)
    n = np.stack([nx / l, ny / l, nz / l], axis=-1).astype(F32)
    n = ((n + F32(1.0)) * F32(0.5)).astype(F32)
    return (_linear_to_srgb(n) * F32(255.0)).astype(np.uint8)


# --------------------------------------------------------------------------------------------
# mesh pieces
# --------------------------------------------------------------------------------------------

class MeshBuilder:
    def __init__(self):
        self.pos: List[np.ndarray] = []
        self.uv: List[np.ndarray] = []
        self.nrm: List[np.ndarray] = []
        self.n_pos = self.n_uv = self.n_nrm = 0
        self.groups: List[Tuple[str, int, np.ndarray, np.ndarray, np.ndarray]] = []  # name, material, ip, it, in

    def add_group(self, name: str, material: int, pos, uv, nrm, tri_p, tri_t=None, tri_n=None):
        pos = np.asarray(pos, dtype=F32).reshape(-1, 3)
        uv = np.asarray(uv, dtype=F32).reshape(-1, 2)
        nrm = np.asarray(nrm, dtype=F32).reshape(-1, 3)
        tri_p = np.asarray(tri_p, dtype=np.int64).reshape(-1, 3)
        tri_t = tri_p if tri_t is None else np.asarray(tri_t, dtype=np.int64).reshape(-1, 3)
        tri_n = tri_p if tri_n is None else np.asarray(tri_n, dtype=np.int64).reshape(-1, 3)
        self.groups.append((name, material, (tri_p + self.n_pos).astype(np.uint32).reshape(-1),
                            (tri_t + self.n_uv).astype(np.uint32).reshape(-1),
                            (tri_n + self.n_nrm).astype(np.uint32).reshape(-1)))
        self.pos.append(pos); self.uv.append(uv); self.nrm.append(nrm)
        self.n_pos += len(pos); self.n_uv += len(uv); self.n_nrm += len(nrm)

    def add_group_shared(self, name: str, material: int, tri_p, tri_t, tri_n):
        """Group whose indices refer to vertex streams already added with add_vertices()."""
        self.groups.append((name, material, np.asarray(tri_p, dtype=np.uint32).reshape(-1),
                            np.asarray(tri_t, dtype=np.uint32).reshape(-1), np.asarray(tri_n, dtype=np.uint32).reshape(-1)))

    def add_vertices(self, pos, uv, nrm) -> Tuple[int, int, int]:
        base = (self.n_pos, self.n_uv, self.n_nrm)
        pos = np.asarray(pos, dtype=F32).reshape(-1, 3); uv = np.asarray(uv, dtype=F32).reshape(-1, 2)
        nrm = np.asarray(nrm, dtype=F32).reshape(-1, 3)
        self.pos.append(pos); self.uv.append(uv); self.nrm.append(nrm)
        self.n_pos += len(pos); self.n_uv += len(uv); self.n_nrm += len(nrm)
        return base

    def finish(self, materials: np.ndarray, textures: List[TextureData], lights=None, name="scene",
               hierarchy: Optional[Tuple[np.ndarray, np.ndarray]] = None) -> SceneData:
        positions = np.concatenate(self.pos) if self.pos else np.zeros((0, 3), F32)
        texcoords = np.concatenate(self.uv) if self.uv else np.zeros((0, 2), F32)
        normals = np.concatenate(self.nrm) if self.nrm else np.zeros((0, 3), F32)
        gf = np.zeros(len(self.groups) + 1, dtype=np.uint32)
        for i, g in enumerate(self.groups):
            gf[i + 1] = gf[i] + len(g[2])
        ip = np.concatenate([g[2] for g in self.groups]).astype(np.uint32)
        it = np.concatenate([g[3] for g in self.groups]).astype(np.uint32)
        inn = np.concatenate([g[4] for g in self.groups]).astype(np.uint32)
        gm = np.array([g[1] for g in self.groups], dtype=np.int32)
        has_bump = [m >= 0 and materials[m]["bump_texture"] >= 0 for m in gm]
        tangents = calculate_tangents(positions, texcoords, normals, gf, ip, it, inn, has_bump) if any(has_bump) else None
        if isinstance(hierarchy, str) and hierarchy == "defer":     # filled in later by use_reference_hierarchy() (BuildHierarchy on the GPU)
            spheres, sg = np.zeros(0, BSPHERE), np.zeros(0, np.int32)
        else:
            spheres, sg = hierarchy if hierarchy is not None else build_group_hierarchy(positions, gf, ip)
        sd = SceneData(positions=positions, texcoords=texcoords, normals=normals, tangents=tangents, group_first=gf,
                       idx_positions=ip, idx_texcoords=it, idx_normals=inn, group_material=gm, spheres=spheres,
                       sphere_group=sg, materials=materials, default_material=default_material(), textures=textures,
                       lights=default_lights() if lights is None else lights, name=name)
        sd.group_names = [g[0] for g in self.groups]  # type: ignore[attr-defined]
        sd.validate()
        return sd


def _orient_outward(pos: np.ndarray, tris: np.ndarray, center: np.ndarray) -> np.ndarray:
    a, b, c = pos[tris[:, 0]], pos[tris[:, 1]], pos[tris[:, 2]]
    n = np.cross(b - a, c - a)
    area2 = np.linalg.norm(n, axis=1)
    out = np.einsum("ij,ij->i", n, (a + b + c) / 3.0 - center)
    tris = tris.copy()
    flip = out < 0
    tris[flip, 1], tris[flip, 2] = tris[flip, 2].copy(), tris[flip, 1].copy()
    return tris[area2 > 1e-12 * max(1.0, float(area2.max()))]


def uv_sphere(center, radius: float, nu: int = 64, nv: int = 32):
    """Tessellated sphere: (nv+1)*(nu+1) vertices (seam duplicated for UVs), 2*nu*nv - 2*nu triangles,
    outward (counter-clockwise from outside) winding, radial normals."""
    center = np.asarray(center, dtype=np.float64)
    j, i = np.meshgrid(np.arange(nv + 1), np.arange(nu + 1), indexing="ij")
    theta = np.pi * j / nv
    phi = 2.0 * np.pi * i / nu
    d = np.stack([np.sin(theta) * np.cos(phi), np.cos(theta), np.sin(theta) * np.sin(phi)], axis=-1).reshape(-1, 3)
    pos = (center + radius * d).astype(F32)
    nrm = d.astype(F32)
    uv = np.stack([i / nu, 1.0 - j / nv], axis=-1).reshape(-1, 2).astype(F32)
    jj, ii = np.meshgrid(np.arange(nv), np.arange(nu), indexing="ij")
    v00 = (jj * (nu + 1) + ii).reshape(-1)
    v01 = v00 + 1
    v10 = v00 + (nu + 1)
    v11 = v10 + 1
    tris = np.concatenate([np.stack([v00, v10, v11], axis=1), np.stack([v00, v11, v01], axis=1)], axis=0)
    order = np.argsort(np.concatenate([np.arange(len(v00)) * 2, np.arange(len(v00)) * 2 + 1]), kind="stable")
    tris = tris[order]
    tris = _orient_outward(pos.astype(np.float64), tris, center)
    return pos, uv, nrm, tris


def make_material(kd, ks=(1, 1, 1), ns=10.0, ni=1.5, d=1.0, ka=None, **tex) -> np.ndarray:
    m = np.zeros((), dtype=MATERIAL)
    m["specular_intensity"] = ns
    m["index_of_refraction"] = ni
    m["alpha"] = d
    ka = kd if ka is None else ka
    m["ambient_color"] = (*ka, 1.0)
    m["diffuse_color"] = (*kd, 1.0)
    m["specular_color"] = (*ks, 1.0)
    for k in ("ambient_texture", "diffuse_texture", "specular_texture", "alpha_texture", "bump_texture"):
        m[k] = tex.get(k, -1)
    return m


# --------------------------------------------------------------------------------------------
# procedural textures
# --------------------------------------------------------------------------------------------

def checker_texture(size: int, cells: int, c0, c1, seed: int = 0) -> TextureData:
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:size, 0:size]
    m = (((x * cells) // size + (y * cells) // size) % 2).astype(np.uint8)
    img = np.where(m[..., None] == 0, np.array(c0, dtype=np.int32), np.array(c1, dtype=np.int32))
    img = np.clip(img + rng.integers(-12, 13, size=img.shape), 0, 255).astype(np.uint8)
    return TextureData(size, size, 3, img)


def noise_height(size: int, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    y, x = np.mgrid[0:size, 0:size].astype(np.float64) / size
    h = np.zeros((size, size))
    for k in range(1, 6):
        ph = rng.uniform(0, 2 * np.pi, 2)
        h += np.sin(2 * np.pi * k * x + ph[0]) * np.cos(2 * np.pi * k * y + ph[1]) / k
    h = (h - h.min()) / (h.max() - h.min())
    return (h * 255).astype(np.uint8)


def bump_texture(size: int, seed: int = 0) -> TextureData:
    hm = noise_height(size, seed)
    t = TextureData(size, size, 3, height_to_normal_map(hm))
    t.source_height = hm  # type: ignore[attr-defined]  (what map_bump points at in the MTL)
    return t


def mask_texture(size: int, holes: int = 8) -> TextureData:
    y, x = np.mgrid[0:size, 0:size]
    cx = (x % (size // holes)) - size // holes // 2
    cy = (y % (size // holes)) - size // holes // 2
    m = ((cx * cx + cy * cy) > (size // holes // 3) ** 2).astype(np.uint8) * 255
    return TextureData(size, size, 1, m)


# --------------------------------------------------------------------------------------------
# scenes
# --------------------------------------------------------------------------------------------

def spheres_plane_scene(grid: int = 4, nu: int = 64, nv: int = 32, spacing: float = 2.6, textured: bool = False,
                        name: str = "spheres_plane") -> SceneData:
    """BASELINE config 2: grid*grid tessellated spheres (one group each) on a 2-triangle plane, diffuse /
    specular material variants (Ns in {10, 40, 200}). Defaults: 16 x 3,968 + 2 = 63,490 triangles."""
    mb = MeshBuilder()
    mats = []
    textures: List[TextureData] = []
    rng = np.random.default_rng(20170218)
    if textured:
        textures.append(checker_texture(64, 8, (230, 230, 230), (60, 90, 200), seed=1))
        textures.append(bump_texture(64, seed=2))
        textures.append(mask_texture(64, holes=4))
    ns_cycle = [10.0, 40.0, 200.0]
    k = 0
    ext = (grid - 1) * spacing * 0.5
    for gz in range(grid):
        for gx in range(grid):
            kd = tuple(float(v) for v in (0.25 + 0.7 * rng.random(3)))
            ks = tuple(float(v) for v in (0.2 + 0.8 * rng.random(3)))
            tex = {}
            d = 1.0
            if textured:
                if k % 4 == 1:
                    tex = dict(diffuse_texture=0, ambient_texture=0)
                elif k % 4 == 2:
                    tex = dict(bump_texture=1)
                elif k % 4 == 3:
                    tex = dict(alpha_texture=2)
                if k % 8 == 4:
                    d = 0.6
            mats.append(make_material(kd, ks, ns=ns_cycle[k % 3], d=d, **tex))
            r = 1.0 + 0.15 * ((k * 7) % 5 - 2) / 2.0
            c = (gx * spacing - ext, r, gz * spacing - ext)
            pos, uv, nrm, tris = uv_sphere(c, r, nu, nv)
            mb.add_group(f"sphere{k}", k, pos, uv, nrm, tris)
            k += 1
    half = ext + 3 * spacing
    mats.append(make_material((0.8, 0.8, 0.78), (0.3, 0.3, 0.3), ns=40.0,
                              **(dict(diffuse_texture=0) if textured else {})))
    ppos = np.array([[-half, 0, -half], [half, 0, -half], [half, 0, half], [-half, 0, half]], dtype=F32)
    puv = np.array([[0, 0], [8, 0], [8, 8], [0, 8]], dtype=F32)
    pn = np.array([[0, 1, 0]] * 4, dtype=F32)
    ptris = _orient_outward(ppos.astype(np.float64), np.array([[0, 1, 2], [0, 2, 3]]), np.array([0.0, -1e6, 0.0]))
    mb.add_group("plane", k, ppos, puv, pn, ptris)
    sd = mb.finish(np.array(mats, dtype=MATERIAL), textures, name=name)
    sd.camera_hint = dict(position=(0.3, ext * 0.9 + 3.0, ext + 3.2 * spacing), facing=(0.0, -0.45, -1.0), fov=60.0)  # type: ignore[attr-defined]
    return sd


def heightfield_scene(cells_x: int = 64, cells_z: int = 64, block: int = 16, size: float = 100.0, amp: float = 6.0,
                      textured: bool = True, tex_size: int = 128, name: str = "heightfield", seed: int = 20170218,
                      hierarchy=None) -> SceneData:
    """BASELINE configs 3/4/5 style: a height-field terrain of 2*cells_x*cells_z triangles, grouped in
    block x block cell tiles (2*block^2 triangles per group), ~8 cycling materials (diffuse/ambient maps,
    bump map, one alpha mask). Vertex streams are shared between groups like a real OBJ."""
    rng = np.random.default_rng(seed)
    nx, nz = cells_x + 1, cells_z + 1
    gx, gz = np.meshgrid(np.arange(nx), np.arange(nz), indexing="xy")  # (nz, nx)
    x = (gx / cells_x - 0.5) * size
    z = (gz / cells_z - 0.5) * size * (cells_z / cells_x)
    waves = [(rng.uniform(0.02, 0.25), rng.uniform(0.02, 0.25), rng.uniform(0, 6.28), rng.uniform(0.2, 1.0)) for _ in range(10)]
    y = np.zeros_like(x)
    dydx = np.zeros_like(x)
    dydz = np.zeros_like(x)
    for fx, fz, ph, a in waves:
        arg = fx * x + fz * z + ph
        y += a * np.sin(arg)
        dydx += a * fx * np.cos(arg)
        dydz += a * fz * np.cos(arg)
    scale = amp / max(1e-9, np.abs(y).max())
    y *= scale; dydx *= scale; dydz *= scale
    pos = np.stack([x, y, z], axis=-1).reshape(-1, 3).astype(F32)
    n = np.stack([-dydx, np.ones_like(x), -dydz], axis=-1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    nrm = n.reshape(-1, 3).astype(F32)
    uv = np.stack([gx / 8.0, gz / 8.0], axis=-1).reshape(-1, 2).astype(F32)

    textures: List[TextureData] = []
    mats = []
    if textured:
        textures.append(checker_texture(tex_size, 8, (220, 200, 170), (90, 120, 60), seed=3))
        textures.append(checker_texture(tex_size, 16, (200, 200, 210), (120, 60, 50), seed=4))
        textures.append(bump_texture(tex_size, seed=5))
        textures.append(mask_texture(tex_size, holes=8))
    palette = [(0.8, 0.7, 0.55), (0.45, 0.65, 0.4), (0.7, 0.7, 0.75), (0.75, 0.45, 0.4), (0.5, 0.55, 0.8),
               (0.85, 0.8, 0.4), (0.6, 0.6, 0.6), (0.7, 0.5, 0.7)]
    for k, kd in enumerate(palette):
        tex = {}
        if textured:
            if k % 4 == 0:
                tex = dict(diffuse_texture=0, ambient_texture=0)
            elif k % 4 == 1:
                tex = dict(diffuse_texture=1, bump_texture=2)
            elif k == 2:
                tex = dict(bump_texture=2)
            elif k == 7:
                tex = dict(alpha_texture=3)
        mats.append(make_material(kd, (0.5, 0.5, 0.5), ns=[10.0, 40.0, 200.0][k % 3], **tex))
    mb = MeshBuilder()
    mb.add_vertices(pos, uv, nrm)
    gi = 0
    for bz in range(0, cells_z, block):
        for bx in range(0, cells_x, block):
            cz, cx = np.meshgrid(np.arange(bz, min(bz + block, cells_z)), np.arange(bx, min(bx + block, cells_x)), indexing="ij")
            v00 = (cz * nx + cx).reshape(-1)
            v01 = v00 + 1
            v10 = v00 + nx
            v11 = v10 + 1
            # upward-facing (normal +y): cross(b-a, c-a).y > 0  <=>  a=(x,z) b=(x,z+1) c=(x+1,z+1)
            t0 = np.stack([v00, v10, v11], axis=1)
            t1 = np.stack([v00, v11, v01], axis=1)
            tris = np.empty((2 * len(v00), 3), dtype=np.int64)
            tris[0::2] = t0
            tris[1::2] = t1
            mb.add_group_shared(f"tile{gi}", gi % len(mats), tris, tris, tris)
            gi += 1
    sd = mb.finish(np.array(mats, dtype=MATERIAL), textures, name=name, hierarchy=hierarchy)
    zext = size * (cells_z / cells_x)
    sd.camera_hint = dict(position=(0.0, amp * 3.0 + 0.12 * size, 0.55 * zext), facing=(0.0, -0.45, -1.0), fov=60.0)  # type: ignore[attr-defined]
    return sd


def _grid_patch(origin, du, dv, nu: int, nv: int, uv_scale=(1.0, 1.0), bulge=None):
    """(nu x nv)-cell planar patch spanned by du, dv from `origin`; front face = cross(du, dv). Optional bulge(u, v) -> offset along
    the normal (drapery, carved relief). Returns pos, uv, nrm, tris with counter-clockwise winding seen from the front."""
    origin = np.asarray(origin, np.float64); du = np.asarray(du, np.float64); dv = np.asarray(dv, np.float64)
    n = np.cross(du, dv); n /= np.linalg.norm(n)
    j, i = np.meshgrid(np.arange(nv + 1), np.arange(nu + 1), indexing="ij")
    u, v = i / nu, j / nv
    pos = origin + u[..., None] * du + v[..., None] * dv
    nrm = np.broadcast_to(n, pos.shape).copy()
    if bulge is not None:
        h = bulge(u, v)
        pos = pos + h[..., None] * n
        hu = np.gradient(h, axis=1) * nu / np.linalg.norm(du); hv = np.gradient(h, axis=0) * nv / np.linalg.norm(dv)
        nrm = n - hu[..., None] * du / np.linalg.norm(du) - hv[..., None] * dv / np.linalg.norm(dv)
        nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    uv = np.stack([u * uv_scale[0], v * uv_scale[1]], axis=-1)
    jj, ii = np.meshgrid(np.arange(nv), np.arange(nu), indexing="ij")
    v00 = (jj * (nu + 1) + ii).reshape(-1); v01 = v00 + 1; v10 = v00 + (nu + 1); v11 = v10 + 1
    tris = np.empty((2 * len(v00), 3), np.int64)
    tris[0::2] = np.stack([v00, v01, v11], axis=1)          # cross(du, dv) side is the front
    tris[1::2] = np.stack([v00, v11, v10], axis=1)
    return pos.reshape(-1, 3).astype(F32), uv.reshape(-1, 2).astype(F32), nrm.reshape(-1, 3).astype(F32), tris


def _cylinder(center, radius: float, y0: float, y1: float, nseg: int, nring: int, flute: float = 0.0):
    """Column shaft around the y axis through (center.x, center.z), outward facing; `flute` carves shallow grooves."""
    j, i = np.meshgrid(np.arange(nring + 1), np.arange(nseg + 1), indexing="ij")
    phi = 2.0 * np.pi * i / nseg
    r = radius * (1.0 - flute * (0.5 + 0.5 * np.cos(12.0 * phi))) * (1.0 - 0.08 * j / nring)
    x = center[0] + r * np.cos(phi); z = center[1] + r * np.sin(phi); y = y0 + (y1 - y0) * j / nring
    pos = np.stack([x, y, z], axis=-1).reshape(-1, 3)
    nrm = np.stack([np.cos(phi), np.zeros_like(phi), np.sin(phi)], axis=-1).reshape(-1, 3)
    uv = np.stack([4.0 * i / nseg, 6.0 * j / nring], axis=-1).reshape(-1, 2)
    jj, ii = np.meshgrid(np.arange(nring), np.arange(nseg), indexing="ij")
    v00 = (jj * (nseg + 1) + ii).reshape(-1); v01 = v00 + 1; v10 = v00 + (nseg + 1); v11 = v10 + 1
    tris = np.concatenate([np.stack([v00, v10, v11], axis=1), np.stack([v00, v11, v01], axis=1)], axis=0)
    order = np.argsort(np.concatenate([np.arange(len(v00)) * 2, np.arange(len(v00)) * 2 + 1]), kind="stable")
    tris = tris[order]
    tris = _orient_outward(pos, tris, np.array([center[0], 0.5 * (y0 + y1), center[1]]))
    # orientation is per triangle about the axis point: good enough for a convex shaft
    return pos.astype(F32), uv.astype(F32), nrm.astype(F32), tris


def sponza_standin_scene(detail: int = 1, name: str = "sponza_standin") -> SceneData:
    """BASELINE config 1: a procedural stand-in for the Crytek Sponza atrium the reference hard-codes (main.cpp:432, 553; 262,267 triangles,
    out.txt:1), in Sponza's own coordinate range so that the reference's DEFAULT camera -- position (475, 250, 0), facing (1.25, -0.5, 1.25),
    fov 60 (main.cpp:426-431) -- looks down the nave: a tiled floor, two storeys of fluted columns, arcade walls with bump-mapped brick, a
    carved frieze, alpha-masked hanging drapery and foliage, translucent banners, open to the sky (the single directional light comes in
    from above). ~260 k triangles in ~250 OBJ groups of very different sizes (4 ... 8,192 triangles), 10 materials: diffuse + ambient maps,
    bump maps, alpha masks, a translucent material. detail = 1 is the full-size scene; smaller values give quick test scenes."""
    rng = np.random.default_rng(20170218)
    d = max(0.1, float(detail))
    q = lambda n: max(2, int(round(n * d)))                 # tessellation counts scale with detail
    textures = [checker_texture(256, 16, (214, 196, 170), (150, 128, 108), seed=11),      # 0 floor tiles
                checker_texture(256, 32, (176, 92, 70), (150, 70, 56), seed=12),          # 1 brick
                bump_texture(256, seed=13),                                                # 2 brick / stone relief
                mask_texture(256, holes=16),                                               # 3 drapery lace
                checker_texture(128, 4, (60, 130, 60), (30, 90, 40), seed=14),             # 4 foliage colour
                mask_texture(128, holes=6),                                                # 5 foliage cut-out
                bump_texture(128, seed=15)]                                                # 6 column flutes
    mats = [make_material((0.85, 0.8, 0.72), (0.25, 0.25, 0.25), ns=40.0, diffuse_texture=0, ambient_texture=0),   # 0 floor
            make_material((0.8, 0.6, 0.5), (0.1, 0.1, 0.1), ns=10.0, diffuse_texture=1, ambient_texture=1, bump_texture=2),   # 1 brick wall
            make_material((0.78, 0.74, 0.66), (0.3, 0.3, 0.3), ns=40.0, bump_texture=6),      # 2 column stone
            make_material((0.7, 0.68, 0.62), (0.2, 0.2, 0.2), ns=10.0, bump_texture=2),       # 3 carved frieze
            make_material((0.75, 0.15, 0.12), (0.2, 0.2, 0.2), ns=10.0, alpha_texture=3),     # 4 red drapery (alpha mask)
            make_material((0.2, 0.55, 0.2), (0.1, 0.1, 0.1), ns=10.0, diffuse_texture=4, alpha_texture=5),   # 5 foliage
            make_material((0.2, 0.3, 0.75), (0.6, 0.6, 0.6), ns=200.0, d=0.55),              # 6 translucent banner
            make_material((0.72, 0.7, 0.68), (0.5, 0.5, 0.5), ns=200.0),                       # 7 polished trim
            make_material((0.55, 0.38, 0.2), (0.1, 0.1, 0.1), ns=10.0),                        # 8 wood
            make_material((0.8, 0.78, 0.7), (0.15, 0.15, 0.15), ns=40.0, diffuse_texture=0)]   # 9 upper gallery floor
    mb = MeshBuilder()
    X0, X1, Z0, Z1 = -1800.0, 1800.0, -700.0, 700.0          # nave; aisles behind the columns to +-1000
    ZA = 1000.0
    H1, H2 = 420.0, 840.0                                      # storey heights

    def add(gname, mat, pos, uv, nrm, tris):
        mb.add_group(gname, mat, pos, uv, nrm, tris)

    # floor: 12 x 6 slabs, each its own group (coarse: flat), plus the aisles
    nxs, nzs = 12, 6
    for ix in range(nxs):
        for iz in range(nzs):
            x0 = X0 + (X1 - X0) * ix / nxs; z0 = -ZA + 2 * ZA * iz / nzs
            add(f"floor_{ix}_{iz}", 0, *_grid_patch((x0, 0.0, z0), (0, 0, 2 * ZA / nzs), ((X1 - X0) / nxs, 0, 0), q(6), q(6), uv_scale=(2.0, 3.0)))
    # outer walls (bump-mapped brick), two storeys, panels of very different tessellation
    for side, z, sgn in (("n", -ZA, 1.0), ("s", ZA, -1.0)):
        for storey, (y0, y1) in enumerate(((0.0, H1), (H1, H2 + 200.0))):
            for ix in range(10):
                x0 = X0 + (X1 - X0) * ix / 10; dx = (X1 - X0) / 10
                n_u = q(10 + 14 * ((ix * 7 + storey) % 3))
                if sgn > 0:
                    patch = _grid_patch((x0, y0, z), (dx, 0, 0), (0, y1 - y0, 0), n_u, q(12), uv_scale=(3.0, 3.0),
                                        bulge=lambda u, v: 6.0 * np.sin(6.28 * 3 * u) * np.sin(6.28 * 2 * v))
                else:
                    patch = _grid_patch((x0 + dx, y0, z), (-dx, 0, 0), (0, y1 - y0, 0), n_u, q(12), uv_scale=(3.0, 3.0),
                                        bulge=lambda u, v: 6.0 * np.sin(6.28 * 3 * u) * np.sin(6.28 * 2 * v))
                add(f"wall_{side}{storey}_{ix}", 1, *patch)
    for side, x, sgn in (("w", X0, 1.0), ("e", X1, -1.0)):
        for iz in range(4):
            z0 = -ZA + 2 * ZA * iz / 4; dz = 2 * ZA / 4
            if sgn > 0:
                patch = _grid_patch((x, 0.0, z0 + dz), (0, 0, -dz), (0, H2 + 200.0, 0), q(24), q(32), uv_scale=(3.0, 6.0),
                                    bulge=lambda u, v: 10.0 * np.cos(6.28 * 2 * u) * np.sin(3.14 * v))
            else:
                patch = _grid_patch((x, 0.0, z0), (0, 0, dz), (0, H2 + 200.0, 0), q(24), q(32), uv_scale=(3.0, 6.0),
                                    bulge=lambda u, v: 10.0 * np.cos(6.28 * 2 * u) * np.sin(3.14 * v))
            add(f"wall_{side}_{iz}", 1, *patch)
    # columns: 2 rows x 12 x 2 storeys, fluted, each with a polished base ring
    ncol = 12
    for row, z in enumerate((Z0, Z1)):
        for ic in range(ncol):
            x = X0 + (X1 - X0) * (ic + 0.5) / ncol
            for storey, (y0, y1, rad) in enumerate(((0.0, H1 - 40.0, 55.0), (H1, H2 - 40.0, 42.0))):
                add(f"column_{row}_{ic}_{storey}", 2, *_cylinder((x, z), rad, y0, y1, q(40), q(24), flute=0.06))
                add(f"colbase_{row}_{ic}_{storey}", 7, *_cylinder((x, z), rad * 1.35, y0, y0 + 30.0, q(24), 1))
    # gallery floors above the aisles + lintels over the columns (top faces and nave-facing faces)
    for row, (za, zb) in enumerate(((-ZA, Z0), (Z1, ZA))):
        for ix in range(6):
            x0 = X0 + (X1 - X0) * ix / 6; dx = (X1 - X0) / 6
            add(f"gallery_{row}_{ix}", 9, *_grid_patch((x0, H1, za), (0, 0, zb - za), (dx, 0, 0), q(4), q(8), uv_scale=(1.0, 2.0)))
            add(f"gallery_under_{row}_{ix}", 3, *_grid_patch((x0, H1 - 40.0, za), (dx, 0, 0), (0, 0, zb - za), q(16), q(8), uv_scale=(4.0, 2.0),
                                                               bulge=lambda u, v: 4.0 * np.sin(6.28 * 4 * u) * np.sin(6.28 * 2 * v)))
    for row, (z, sgn) in enumerate(((Z0, 1.0), (Z1, -1.0))):
        for storey, y in enumerate((H1 - 40.0, H2 - 40.0)):
            for ix in range(6):
                x0 = X0 + (X1 - X0) * ix / 6; dx = (X1 - X0) / 6
                relief = lambda u, v: 8.0 * np.sin(6.28 * 12 * u) * np.sin(3.14 * v) ** 2          # the carved frieze: finely tessellated
                if sgn > 0:
                    patch = _grid_patch((x0, y, z), (dx, 0, 0), (0, 40.0, 0), q(128), q(8), uv_scale=(8.0, 1.0), bulge=relief)
                else:
                    patch = _grid_patch((x0 + dx, y, z), (-dx, 0, 0), (0, 40.0, 0), q(128), q(8), uv_scale=(8.0, 1.0), bulge=relief)
                add(f"frieze_{row}_{storey}_{ix}", 3, *patch)
    # drapery: alpha-masked, wavy, hanging between upper columns on both sides (facing the nave) -- 8,192 triangles each at detail 1
    for row, (z, sgn) in enumerate(((Z0 + 30.0, 1.0), (Z1 - 30.0, -1.0))):
        for k in range(4):
            x0 = X0 + (X1 - X0) * (2 * k + 2.6) / ncol; dx = (X1 - X0) / ncol * 0.8
            wave = lambda u, v: 25.0 * np.sin(6.28 * 5 * u + 2.0 * v) * (0.3 + 0.7 * v)
            if sgn > 0:
                patch = _grid_patch((x0, H1 + 60.0, z), (dx, 0, 0), (0, 300.0, 0), q(64), q(64), uv_scale=(2.0, 2.0), bulge=wave)
            else:
                patch = _grid_patch((x0 + dx, H1 + 60.0, z), (-dx, 0, 0), (0, 300.0, 0), q(64), q(64), uv_scale=(2.0, 2.0), bulge=wave)
            add(f"drapery_{row}_{k}", 4, *patch)
    # translucent banners across the nave (two-sided: a front and a back sheet), low tessellation
    for k in range(3):
        x = X0 + (X1 - X0) * (k + 1.5) / 5
        add(f"banner_{k}_a", 6, *_grid_patch((x, H1 + 100.0, -300.0), (0, 0, 600.0), (0, 260.0, 0), q(8), q(6)))
        add(f"banner_{k}_b", 6, *_grid_patch((x - 2.0, H1 + 100.0, 300.0), (0, 0, -600.0), (0, 260.0, 0), q(8), q(6)))
    # planters with foliage: a wooden tub (cylinder) + alpha-cut leaf cards (each card two-sided)
    for k in range(8):
        x = X0 + (X1 - X0) * (k + 0.5) / 8; z = (-1.0 if k % 2 else 1.0) * 380.0
        add(f"tub_{k}", 8, *_cylinder((x, z), 45.0, 0.0, 70.0, q(20), q(4)))
        cards_p, cards_t, cards_n, cards_i = [], [], [], []
        base = 0
        for c in range(q(24)):
            ang = rng.uniform(0, 2 * np.pi); tilt = rng.uniform(0.2, 0.9); ln = rng.uniform(90.0, 160.0)
            du = np.array([np.cos(ang) * 30.0, 0.0, np.sin(ang) * 30.0])
            dv = np.array([-np.sin(ang) * ln * tilt, ln, np.cos(ang) * ln * tilt])
            org = np.array([x, 70.0, z]) - 0.5 * du
            for flip in (False, True):
                pp, tt, nn, ii = _grid_patch(org + (du if flip else 0), (-du if flip else du), dv, 2, 4)
                cards_p.append(pp); cards_t.append(tt); cards_n.append(nn); cards_i.append(ii + base); base += len(pp)
        add(f"foliage_{k}", 5, np.concatenate(cards_p), np.concatenate(cards_t), np.concatenate(cards_n), np.concatenate(cards_i))
    # two finely tessellated ornaments (the "lion heads"): bumpy spheres on pedestals at the far end
    for k, z in enumerate((-250.0, 250.0)):
        pos, uv, nrm, tris = uv_sphere((X1 - 260.0, 210.0, z), 90.0, q(128), q(64))
        r = 1.0 + 0.06 * np.sin(9.0 * pos[:, 0] / 90.0) * np.sin(7.0 * pos[:, 1] / 90.0) * np.cos(8.0 * pos[:, 2] / 90.0)
        ctr = np.array([X1 - 260.0, 210.0, z], F32)
        pos = (ctr + (pos - ctr) * r[:, None]).astype(F32)
        add(f"ornament_{k}", 3, pos, uv, nrm, tris)
        add(f"pedestal_{k}", 7, *_cylinder((X1 - 260.0, z), 70.0, 0.0, 125.0, q(24), q(4)))
    sd = mb.finish(np.array(mats, dtype=MATERIAL), textures, name=name)
    sd.camera_hint = dict(position=(475.0, 250.0, 0.0), facing=(1.25, -0.5, 1.25), fov=60.0)      # main.cpp:426-431  # type: ignore[attr-defined]
    return sd


def use_reference_hierarchy(sd: SceneData, device: int = 0) -> SceneData:
    """Replaces sd.spheres / sd.sphere_group by the reference's OWN hierarchy over the mesh groups -- BuildHierarchy
    (bsphere.cpp:379-444) run on the GPU by rt_build_group_hierarchy, bit-identical to the host build (tests/test_gpu_parity.py) --
    so that the equal-t tie-break order is exactly what the unmodified reference derives from the same OBJ."""
    from . import api
    sd.spheres, sd.sphere_group = api.build_group_hierarchy(sd, device)
    sd.validate()
    return sd


# --------------------------------------------------------------------------------------------
# OBJ / MTL / PNG export (input of the unmodified reference: oracle/_ref, main.cpp -d <dir>)
# --------------------------------------------------------------------------------------------

def write_png(path: str, img: np.ndarray) -> None:
    """Minimal PNG encoder (8-bit, 1/3/4 channels) -- stb_image 2.14 decodes these."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    if img.ndim == 2:
        img = img[..., None]
    h, w, c = img.shape
    ctype = {1: 0, 2: 4, 3: 2, 4: 6}[c]
    raw = np.concatenate([np.zeros((h, 1), dtype=np.uint8), img.reshape(h, w * c)], axis=1).tobytes()

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xFFFFFFFF)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(raw, 6)))
        f.write(chunk(b"IEND", b""))


def _fmt(v) -> str:
    return "%.9g" % float(v)


def write_obj(scene: SceneData, directory: str, obj_name: str = "sponza.obj") -> str:
    """Writes `scene` as <directory>/sponza.obj (+ sponza.mtl + tex_*.png). Floats are printed with 9
    significant digits so strtof (obj_parser.cpp:73-79) recovers the identical float32."""
    os.makedirs(directory, exist_ok=True)
    tex_files = {}
    for i, t in enumerate(scene.textures):
        fn = f"tex_{i}.png"
        src = getattr(t, "source_height", None)
        if src is not None:
            write_png(os.path.join(directory, fn), np.asarray(src).reshape(t.size_y, t.size_x))
        else:
            write_png(os.path.join(directory, fn), t.texels.reshape(t.size_y, t.size_x, t.channels))
        tex_files[i] = fn
    with open(os.path.join(directory, "sponza.mtl"), "w") as f:
        for i, m in enumerate(scene.materials):
            f.write(f"newmtl mat{i}\n")
            f.write(f"Ns {_fmt(m['specular_intensity'])}\nNi {_fmt(m['index_of_refraction'])}\nd {_fmt(m['alpha'])}\n")
            f.write("Ka " + " ".join(_fmt(v) for v in m["ambient_color"][:3]) + "\n")
            f.write("Kd " + " ".join(_fmt(v) for v in m["diffuse_color"][:3]) + "\n")
            f.write("Ks " + " ".join(_fmt(v) for v in m["specular_color"][:3]) + "\n")
            for key, tag in (("ambient_texture", "map_Ka"), ("diffuse_texture", "map_Kd"), ("specular_texture", "map_Ks"),
                             ("alpha_texture", "map_d"), ("bump_texture", "map_bump")):
                if m[key] >= 0:
                    f.write(f"{tag} {tex_files[int(m[key])]}\n")
            f.write("\n")
    path = os.path.join(directory, obj_name)
    names = getattr(scene, "group_names", [f"g{i}" for i in range(scene.n_groups)])
    with open(path, "w") as f:
        f.write("mtllib sponza.mtl\n")
        np.savetxt(f, scene.positions, fmt="v %.9g %.9g %.9g")
        np.savetxt(f, scene.texcoords, fmt="vt %.9g %.9g")
        np.savetxt(f, scene.normals, fmt="vn %.9g %.9g %.9g")
        for g in range(scene.n_groups):
            f.write(f"g {names[g]}\n")
            if scene.group_material[g] >= 0:
                f.write(f"usemtl mat{int(scene.group_material[g])}\n")
            a, b = int(scene.group_first[g]), int(scene.group_first[g + 1])
            tri = np.stack([scene.idx_positions[a:b], scene.idx_texcoords[a:b], scene.idx_normals[a:b]], axis=1).astype(np.int64) + 1
            tri = tri.reshape(-1, 9)
            np.savetxt(f, tri, fmt="f %d/%d/%d %d/%d/%d %d/%d/%d")
    return path
