"""Seeded synthetic scenes for the BASELINE.json configurations.

The Khronos sample assets (SciFiHelmet / Sponza / DamagedHelmet .glb) are not available offline,
so every configuration has a documented stand-in with the same triangle count class, texture
sizes and primitive structure.  Everything is numpy-only and seeded.
"""
from __future__ import annotations

import numpy as np

from ._abi import Primitive, Scene


# ---- textures --------------------------------------------------------------------------------
def _value_noise(rng: np.random.Generator, size: int, cells: int) -> np.ndarray:
    """Periodic bilinear value noise in [0,1], shape (size, size)."""
    g = rng.random((cells, cells), dtype=np.float32)
    x = (np.arange(size, dtype=np.float32) + 0.5) * (cells / size)
    i0 = np.floor(x).astype(np.int64) % cells
    i1 = (i0 + 1) % cells
    f = (x - np.floor(x)).astype(np.float32)
    f = f * f * (3.0 - 2.0 * f)
    rows = g[i0] * (1.0 - f)[:, None] + g[i1] * f[:, None]          # (size, cells)
    return rows[:, i0] * (1.0 - f)[None, :] + rows[:, i1] * f[None, :]


def _fbm(rng: np.random.Generator, size: int, base_cells: int, octaves: int) -> np.ndarray:
    out = np.zeros((size, size), np.float32)
    amp, tot = 1.0, 0.0
    for o in range(octaves):
        cells = min(size, base_cells << o)
        out += amp * _value_noise(rng, size, cells)
        tot += amp
        amp *= 0.5
    return out / tot


def _to_u8(x: np.ndarray) -> np.ndarray:
    return np.clip(np.rint(x * 255.0), 0, 255).astype(np.uint8)


def make_albedo(size: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    rgb = np.stack([_fbm(rng, size, 8, 5) for _ in range(3)], axis=-1)
    a = np.ones((size, size, 1), np.float32)
    a[:: max(1, size // 16), :, 0] = 0.75  # a few non-opaque rows so the logit path is exercised
    return _to_u8(np.concatenate([rgb, a], axis=-1))


def make_normal_map(size: int, seed: int, strength: float = 0.6) -> np.ndarray:
    rng = np.random.default_rng(seed)
    h = _fbm(rng, size, 16, 5)
    dx = (np.roll(h, -1, axis=1) - np.roll(h, 1, axis=1)) * size * strength / 32.0
    dy = (np.roll(h, -1, axis=0) - np.roll(h, 1, axis=0)) * size * strength / 32.0
    n = np.stack([-dx, -dy, np.ones_like(h)], axis=-1)
    n /= np.linalg.norm(n, axis=-1, keepdims=True)
    rgba = np.concatenate([n * 0.5 + 0.5, np.ones((size, size, 1), np.float32)], axis=-1)
    return _to_u8(rgba)


def make_mr_map(size: int, seed: int) -> np.ndarray:
    rng = np.random.default_rng(seed)
    ao = np.ones((size, size), np.float32)
    rough = _fbm(rng, size, 8, 2)
    metal = _fbm(rng, size, 4, 2)
    return _to_u8(np.stack([ao, rough, metal, np.ones_like(ao)], axis=-1))


def make_material_textures(size: int, seed: int) -> list:
    return [make_albedo(size, seed), make_normal_map(size, seed + 1), make_mr_map(size, seed + 2)]


def random_texture(w: int, h: int, seed: int) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 256, size=(h, w, 4), dtype=np.uint8)


# ---- meshes ----------------------------------------------------------------------------------
def _pack(pos, nrm, tan, uv) -> np.ndarray:
    """(T,3,3) (T,3,3) (T,3,4) (T,3,2) -> (T,36) float32."""
    return np.concatenate([pos, nrm, tan, uv], axis=-1).reshape(len(pos), 36).astype(np.float32)


def unit_quad() -> Scene:
    """BASELINE config 1: v0(0,0,0) v1(1,0,0) v2(1,1,0) v3(0,1,0); tris (0,1,2),(0,2,3); uv = xy."""
    v = np.array([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]], np.float32)
    idx = np.array([[0, 1, 2], [0, 2, 3]])
    pos = v[idx]
    nrm = np.broadcast_to(np.array([0, 0, 1], np.float32), pos.shape)
    tan = np.broadcast_to(np.array([1, 0, 0, 1], np.float32), (2, 3, 4))
    uv = pos[..., :2]
    s = Scene(_pack(pos, nrm, tan, uv))
    s.compute_bboxes()
    return s


def box(size=(1.0, 2.0, 3.0), origin=(0.0, 0.0, 0.0)) -> Scene:
    """Axis-aligned box, 12 triangles, outward normals; exercises per-face axis selection and the
    max(range_a, range_b) normalisation (SURVEY 8c KAT ii)."""
    sx, sy, sz = size
    o = np.array(origin, np.float32)
    c = np.array([[0, 0, 0], [sx, 0, 0], [sx, sy, 0], [0, sy, 0],
                  [0, 0, sz], [sx, 0, sz], [sx, sy, sz], [0, sy, sz]], np.float32) + o
    quads = [([0, 3, 2, 1], (0, 0, -1)), ([4, 5, 6, 7], (0, 0, 1)), ([0, 1, 5, 4], (0, -1, 0)),
             ([3, 7, 6, 2], (0, 1, 0)), ([0, 4, 7, 3], (-1, 0, 0)), ([1, 2, 6, 5], (1, 0, 0))]
    pos, nrm, tan, uv = [], [], [], []
    quv = np.array([[0, 0], [1, 0], [1, 1], [0, 1]], np.float32)
    for q, n in quads:
        for tri in ([0, 1, 2], [0, 2, 3]):
            p = c[[q[t] for t in tri]]
            pos.append(p)
            nrm.append(np.tile(np.array(n, np.float32), (3, 1)))
            t = p[1] - p[0]
            t = t / max(np.linalg.norm(t), 1e-20)
            tan.append(np.tile(np.append(t, 1.0).astype(np.float32), (3, 1)))
            uv.append(quv[tri])
    s = Scene(_pack(np.array(pos), np.array(nrm), np.array(tan), np.array(uv)))
    s.compute_bboxes()
    return s


def displaced_sphere(n_lon: int, n_lat: int, seed: int, amplitude: float = 0.05, radius: float = 1.0,
                     center=(0.0, 0.0, 0.0), uv_rect=(0.0, 0.0, 1.0, 1.0)) -> np.ndarray:
    """Lat-long sphere with seeded radial value-noise displacement, smooth normals, analytic
    tangents.  2*n_lon*n_lat triangles (pole rows are degenerate).  Returns (T,36)."""
    rng = np.random.default_rng(seed)
    cells = 16
    g = rng.random((cells, cells)).astype(np.float32)

    def disp(u, v):  # periodic in u
        x = u * cells
        y = v * (cells - 1)
        i0 = np.floor(x).astype(np.int64) % cells
        i1 = (i0 + 1) % cells
        j0 = np.clip(np.floor(y).astype(np.int64), 0, cells - 1)
        j1 = np.clip(j0 + 1, 0, cells - 1)
        fx = x - np.floor(x)
        fy = y - np.floor(y)
        fx = fx * fx * (3 - 2 * fx)
        fy = fy * fy * (3 - 2 * fy)
        a = g[j0, i0] * (1 - fx) + g[j0, i1] * fx
        b = g[j1, i0] * (1 - fx) + g[j1, i1] * fx
        return a * (1 - fy) + b * fy

    def surf(u, v):
        th = u * 2.0 * np.pi
        ph = v * np.pi
        # fade displacement at the poles so the surface stays closed
        r = radius * (1.0 + amplitude * (disp(u, v) - 0.5) * 2.0 * np.sin(ph) ** 2)
        d = np.stack([np.sin(ph) * np.cos(th), np.cos(ph), np.sin(ph) * np.sin(th)], axis=-1)
        return d * r[..., None]

    u = np.arange(n_lon + 1, dtype=np.float64) / n_lon
    v = np.arange(n_lat + 1, dtype=np.float64) / n_lat
    U, V = np.meshgrid(u, v)  # (n_lat+1, n_lon+1)
    P = surf(U, V)
    eps = 1e-4
    Pu = (surf(U + eps, V) - surf(U - eps, V)) / (2 * eps)
    Pv = (surf(U, np.clip(V + eps, 0, 1)) - surf(U, np.clip(V - eps, 0, 1)))
    N = np.cross(Pv, Pu)
    nn = np.linalg.norm(N, axis=-1, keepdims=True)
    D = P / np.maximum(np.linalg.norm(P, axis=-1, keepdims=True), 1e-20)
    N = np.where(nn > 1e-12, N / np.maximum(nn, 1e-20), D)
    tn = np.linalg.norm(Pu, axis=-1, keepdims=True)
    th = U * 2.0 * np.pi
    Tfallback = np.stack([-np.sin(th), np.zeros_like(th), np.cos(th)], axis=-1)
    T = np.where(tn > 1e-9, Pu / np.maximum(tn, 1e-20), Tfallback)
    T4 = np.concatenate([T, np.ones(T.shape[:-1] + (1,))], axis=-1)
    u0, v0, u1, v1 = uv_rect
    UVm = np.stack([u0 + U * (u1 - u0), v0 + V * (v1 - v0)], axis=-1)
    P = P + np.array(center, np.float64)

    j, i = np.meshgrid(np.arange(n_lat), np.arange(n_lon), indexing="ij")
    a = (j, i); b = (j, i + 1); c = (j + 1, i + 1); d = (j + 1, i)
    tris = []
    for (k0, k1, k2) in ((a, b, c), (a, c, d)):
        verts = []
        for k in (k0, k1, k2):
            verts.append(np.concatenate([P[k], N[k], T4[k], UVm[k]], axis=-1))  # (n_lat,n_lon,12)
        tris.append(np.stack(verts, axis=-2))  # (n_lat,n_lon,3,12)
    t = np.stack(tris, axis=2)  # (n_lat, n_lon, 2, 3, 12)
    return t.reshape(-1, 36).astype(np.float32)


def helmet_standin(tex_size: int = 2048, seed: int = 1234) -> Scene:
    """BASELINE config 2 stand-in for SciFiHelmet.glb: 70 074 triangles (229 x 153 x 2), one
    primitive, three tex_size^2 RGBA8 maps."""
    tri = displaced_sphere(229, 153, seed, amplitude=0.08)
    s = Scene(tri, [Primitive(0, len(tri), (1.0, 1.0, 1.0, 1.0), 0, 1, 2)],
              make_material_textures(tex_size, seed))
    s.compute_bboxes()
    return s


def damaged_helmet_standin(tex_size: int = 2048, seed: int = 4321) -> Scene:
    """BASELINE config 5 stand-in for DamagedHelmet.glb: 15 488 triangles (121 x 64 x 2)."""
    tri = displaced_sphere(121, 64, seed, amplitude=0.10)
    s = Scene(tri, [Primitive(0, len(tri), (1.0, 1.0, 1.0, 1.0), 0, 1, 2)],
              make_material_textures(tex_size, seed))
    s.compute_bboxes()
    return s


def sphere_1m(tex_size: int = 2048, seed: int = 42) -> Scene:
    """BASELINE config 4: 1000 x 500 quads = 1 000 000 triangles, seed 42, amplitude 0.05."""
    tri = displaced_sphere(1000, 500, seed, amplitude=0.05)
    s = Scene(tri, [Primitive(0, len(tri), (1.0, 1.0, 1.0, 1.0), 0, 1, 2)],
              make_material_textures(tex_size, seed))
    s.compute_bboxes()
    return s


def _grid_quad(p0, du, dv, nu, nv, normal, uv_scale=1.0) -> np.ndarray:
    """Tessellated planar rectangle p0 + s*du + t*dv, nu x nv cells -> (2*nu*nv, 36)."""
    p0 = np.asarray(p0, np.float64); du = np.asarray(du, np.float64); dv = np.asarray(dv, np.float64)
    s = np.arange(nu + 1) / nu
    t = np.arange(nv + 1) / nv
    S, T = np.meshgrid(s, t)
    P = p0 + S[..., None] * du + T[..., None] * dv
    N = np.broadcast_to(np.asarray(normal, np.float64), P.shape)
    tg = du / np.linalg.norm(du)
    T4 = np.broadcast_to(np.append(tg, 1.0), P.shape[:-1] + (4,))
    UV = np.stack([S * uv_scale, T * uv_scale], axis=-1)
    j, i = np.meshgrid(np.arange(nv), np.arange(nu), indexing="ij")
    a = (j, i); b = (j, i + 1); c = (j + 1, i + 1); d = (j + 1, i)
    tris = []
    for (k0, k1, k2) in ((a, b, c), (a, c, d)):
        tris.append(np.stack([np.concatenate([P[k], N[k], T4[k], UV[k]], axis=-1) for k in (k0, k1, k2)], axis=-2))
    return np.stack(tris, axis=2).reshape(-1, 36).astype(np.float32)


def sponza_standin(tex_size: int = 1024, seed: int = 5678, n_prims: int = 100, n_materials: int = 25,
                   target_tris: int = 262_000) -> Scene:
    """BASELINE config 3 stand-in for Sponza.glb: an axis-aligned room (floor, ceiling, 4 walls)
    plus rows of columns (spheres-as-columns), ~262 k triangles in 100 primitives sharing 25
    materials of three tex_size^2 maps each.  uv tiling > 1 exercises REPEAT wrapping."""
    rng = np.random.default_rng(seed)
    prims_tris = []
    # 6 big planar primitives: few huge triangles each (the regime that needs work splitting)
    L, W, H = 30.0, 12.0, 10.0
    planes = [((0, 0, 0), (L, 0, 0), (0, 0, W), (0, 1, 0)), ((0, H, 0), (0, 0, W), (L, 0, 0), (0, -1, 0)),
              ((0, 0, 0), (0, H, 0), (L, 0, 0), (0, 0, 1)), ((0, 0, W), (L, 0, 0), (0, H, 0), (0, 0, -1)),
              ((0, 0, 0), (0, 0, W), (0, H, 0), (1, 0, 0)), ((L, 0, 0), (0, H, 0), (0, 0, W), (-1, 0, 0))]
    for p0, du, dv, n in planes:
        prims_tris.append(_grid_quad(p0, du, dv, 8, 4, n, uv_scale=4.0))
    n_cols = n_prims - len(planes)
    per = max(2, (target_tris - sum(len(t) for t in prims_tris)) // n_cols)
    n_lat = max(2, int(np.sqrt(per / 4)))
    n_lon = max(3, per // (2 * n_lat))
    for k in range(n_cols):
        cx = 1.5 + (k % 47) * (L - 3.0) / 46.0
        cz = 2.0 + (k // 47) * (W - 4.0) / 1.0
        cy = 1.0 + 8.0 * rng.random()
        prims_tris.append(displaced_sphere(n_lon, n_lat, seed + 10 + k, amplitude=0.1,
                                           radius=0.35 + 0.25 * rng.random(), center=(cx, cy, cz)))
    mats = [make_material_textures(tex_size, seed + 1000 + 3 * m) for m in range(n_materials)]
    textures = [t for m in mats for t in m]
    prims, first = [], 0
    for k, t in enumerate(prims_tris):
        m = k % n_materials
        factor = tuple(float(x) for x in (0.6 + 0.4 * rng.random(3))) + (1.0,)
        prims.append(Primitive(first, len(t), factor, 3 * m, 3 * m + 1, 3 * m + 2, name=f"sponza_{k}"))
        first += len(t)
    s = Scene(np.concatenate(prims_tris, axis=0), prims, textures)
    s.compute_bboxes(cumulative=True)
    return s


def random_soup(n: int, seed: int, extent: float = 1.0, tri_size: float = 0.2) -> np.ndarray:
    """Random small triangles with random attributes: fuzz input for parity tests. (n,36)."""
    rng = np.random.default_rng(seed)
    c = rng.random((n, 1, 3)) * extent
    pos = c + (rng.random((n, 3, 3)) - 0.5) * tri_size
    nrm = rng.normal(size=(n, 3, 3))
    nrm /= np.linalg.norm(nrm, axis=-1, keepdims=True)
    tan = rng.normal(size=(n, 3, 3))
    tan /= np.linalg.norm(tan, axis=-1, keepdims=True)
    w = np.where(rng.random((n, 3, 1)) < 0.5, -1.0, 1.0)
    uv = rng.random((n, 3, 2)) * 2.0 - 0.5
    return _pack(pos, nrm, np.concatenate([tan, w], axis=-1), uv)
