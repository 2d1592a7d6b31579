"""Shared helpers for the parity tests: order-independent comparison keyed on fragment identity.

Tolerances (stated once, used everywhere):
  coverage (which (triangle, pixel) pairs emit a gaussian) ........ bit-exact (integer edge functions)
  position ........................................................ 1e-5 * bbox diagonal (absolute)
  raw scale / log-scale, quaternion (same sign convention) ........ 1e-5 relative (+1e-7 abs)
  colour / SH0 / opacity (compared as sigmoid(logit)) / metallic-roughness  1e-4 abs vs the oracle's fp32 sampler
                                                                    (2/255 is the bound vs a real GL driver)
  normal (TBN path) ............................................... 1e-3 abs
"""
from __future__ import annotations

import numpy as np

from mesh2splat_b200 import _abi

POS_TOL_REL_DIAG = 1e-5
REL_TOL = 1e-5
COLOR_TOL = 1e-4
NORMAL_TOL = 1e-3  # FMA-order noise is amplified by ill-conditioned TBN bases in the fuzz inputs


def scene_diag(scene: _abi.Scene) -> float:
    pos = scene.triangles.reshape(-1, 3, 12)[:, :, :3].reshape(-1, 3)
    if len(pos) == 0:
        return 1.0
    pos = pos[np.isfinite(pos).all(axis=1)]
    if len(pos) == 0:
        return 1.0
    return float(np.linalg.norm(pos.max(axis=0) - pos.min(axis=0))) or 1.0


def sort_by_key(rec: np.ndarray, keys: np.ndarray):
    order = np.argsort(keys, kind="stable")
    return rec[order], keys[order]


def _close(a, b, rtol, atol, what):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    both_inf = np.isinf(a) & np.isinf(b) & (np.sign(a) == np.sign(b))
    ok = both_nan | both_inf | (np.abs(a - b) <= atol + rtol * np.abs(b))
    if not ok.all():
        bad = np.argwhere(~ok)[0]
        raise AssertionError(f"{what}: {np.count_nonzero(~ok)} mismatches, first at {tuple(bad)}: "
                             f"got {a[tuple(bad)]!r} want {b[tuple(bad)]!r}; max abs err "
                             f"{np.nanmax(np.abs(np.where(np.isfinite(a - b), a - b, 0)))}")


def _sigmoid(x):
    x = np.asarray(x, np.float64)
    with np.errstate(over="ignore"):
        return 1.0 / (1.0 + np.exp(-x))


def assert_records_match(scene: _abi.Scene, layout: int, got, got_keys, want, want_keys):
    """Set comparison: same fragment identities (exact), same values (tolerances above)."""
    assert len(got) == len(want), f"count {len(got)} != {len(want)}"
    g, gk = sort_by_key(got, np.asarray(got_keys, np.uint64))
    w, wk = sort_by_key(want, np.asarray(want_keys, np.uint64))
    assert np.array_equal(gk, wk), "coverage differs (fragment identity sets are not equal)"
    assert len(np.unique(gk)) == len(gk), "duplicate fragment identities"
    diag = scene_diag(scene)
    if layout == _abi.LAYOUT_REF96:
        _close(g["position"][:, :3], w["position"][:, :3], 0, POS_TOL_REL_DIAG * diag, "position")
        assert np.all(g["position"][:, 3] == 1.0)
        _close(g["scale"], w["scale"], REL_TOL, 1e-12, "scale")
        _close(g["rotation"], w["rotation"], REL_TOL, 1e-6, "rotation")
        _close(g["color"], w["color"], 0, COLOR_TOL, "color")
        _close(g["normal"], w["normal"], 0, NORMAL_TOL, "normal")
        _close(g["pbr"], w["pbr"], 0, COLOR_TOL, "pbr")
    elif layout == _abi.LAYOUT_PACKED56:
        _close(g["xyz"], w["xyz"], 0, POS_TOL_REL_DIAG * diag, "xyz")
        _close(g["rot"], w["rot"], REL_TOL, 1e-6, "rot")
        _close(g["log_scale"], w["log_scale"], REL_TOL, 1e-5, "log_scale")
        _close(g["sh0"], w["sh0"], 0, COLOR_TOL / 0.28209479177387814, "sh0")
        _close(_sigmoid(g["opacity"]), _sigmoid(w["opacity"]), 0, COLOR_TOL, "sigmoid(opacity)")
    elif layout in (_abi.LAYOUT_PLY_STANDARD, _abi.LAYOUT_PLY_PBR):
        _close(g["xyz"], w["xyz"], 0, POS_TOL_REL_DIAG * diag, "xyz")
        _close(g["normal"], w["normal"], 0, NORMAL_TOL, "normal")
        _close(g["f_dc"], w["f_dc"], 0, COLOR_TOL / 0.28209479177387814, "f_dc")
        _close(_sigmoid(g["opacity"]), _sigmoid(w["opacity"]), 0, COLOR_TOL, "sigmoid(opacity)")
        _close(g["scale"], w["scale"], REL_TOL, 1e-5, "scale")
        _close(g["rot"], w["rot"], REL_TOL, 1e-6, "rot")
        if layout == _abi.LAYOUT_PLY_STANDARD:
            assert not g["f_rest"].any()
        else:
            _close(g["metallic"], w["metallic"], 0, COLOR_TOL, "metallic")
            _close(g["roughness"], w["roughness"], 0, COLOR_TOL, "roughness")
    elif layout == _abi.LAYOUT_PLY_COMPRESSED:
        _close(g["xyz"], w["xyz"], 0, POS_TOL_REL_DIAG * diag, "xyz")
        _close(g["rot"], w["rot"], REL_TOL, 1e-6, "rot")
        _close(g["scale"], w["scale"], REL_TOL, 1e-5, "scale")
        for f in ("rgba", "octa", "roughness", "metallic"):  # u8 quantisation: off by one at rounding boundaries
            d = np.abs(g[f].astype(np.int32) - w[f].astype(np.int32))
            assert d.max(initial=0) <= 1, f"{f}: max byte diff {d.max()}"
            assert np.count_nonzero(d) <= max(4, 0.002 * d.size), f"{f}: too many byte diffs"
    else:
        raise ValueError(layout)


def png_bytes(img: np.ndarray) -> bytes:
    """Minimal PNG writer (RGBA8 / RGB8 / gray8, filter 0) for loader tests."""
    import struct
    import zlib

    img = np.ascontiguousarray(img, np.uint8)
    if img.ndim == 2:
        ctype, ch = 0, 1
    else:
        ch = img.shape[2]
        ctype = {1: 0, 2: 4, 3: 2, 4: 6}[ch]
    h, w = img.shape[:2]
    raw = b"".join(b"\x00" + img[y].tobytes() for y in range(h))

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)

    return (b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, ctype, 0, 0, 0))
            + chunk(b"IDAT", zlib.compress(raw, 6)) + chunk(b"IEND", b""))


def planar_triangulation(n_points: int, seed: int) -> _abi.Scene:
    """Delaunay triangulation of the unit square (4 corners + seeded random interior points), z = 0, uv = xy.
    The triangles tile the square with shared edges, so a watertight rasteriser (top-left rule) must emit every
    pixel centre of the R x R grid exactly once — a size-independent property of the coverage rules."""
    from scipy.spatial import Delaunay
    rng = np.random.default_rng(seed)
    pts = np.vstack([[[0, 0], [1, 0], [1, 1], [0, 1]], rng.random((n_points, 2))]).astype(np.float32)
    tri = Delaunay(pts.astype(np.float64)).simplices
    v = np.zeros((len(tri), 3, 12), np.float32)
    v[:, :, 0:2] = pts[tri]
    v[:, :, 5] = 1.0            # normal (0,0,1)
    v[:, :, 6] = 1.0; v[:, :, 9] = 1.0   # tangent (1,0,0,1)
    v[:, :, 10:12] = pts[tri]
    s = _abi.Scene(v.reshape(len(tri), 36), [_abi.Primitive(0, len(tri), (1, 1, 1, 1), -1, -1, -1)], [])
    s.compute_bboxes()
    return s


def png_encode(samples: np.ndarray, ctype: int, depth: int, interlace: bool = False, plte: bytes | None = None,
               trns: bytes | None = None) -> bytes:
    """General PNG writer for loader tests.  samples: (h, w, channels) integers < 2**depth (channels: 1 gray/palette,
    2 gray+alpha, 3 rgb, 4 rgba).  Rows cycle through the five filter types; Adam7 when interlace."""
    import struct
    import zlib
    smp = np.asarray(samples)
    h, w, ch = smp.shape
    assert ch == {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]

    def pack_row(row):  # (pw, ch) -> bytes
        flat = row.reshape(-1).astype(np.uint32)
        if depth == 8:
            return flat.astype(np.uint8).tobytes()
        if depth == 16:
            return flat.astype(">u2").tobytes()
        bits = np.zeros(len(flat) * depth, np.uint8)
        for b in range(depth):
            bits[b::depth] = (flat >> (depth - 1 - b)) & 1
        return np.packbits(bits).tobytes()

    fb = max(1, ch * depth // 8)

    def filt(rows):  # list of bytes -> filtered stream
        out = bytearray()
        prev = bytes(len(rows[0])) if rows else b""
        for y, r in enumerate(rows):
            ft = y % 5
            cur = bytearray(len(r))
            for i in range(len(r)):
                a = r[i - fb] if i >= fb else 0
                b = prev[i]
                c = prev[i - fb] if i >= fb else 0
                if ft == 0: pred = 0
                elif ft == 1: pred = a
                elif ft == 2: pred = b
                elif ft == 3: pred = (a + b) >> 1
                else:
                    pp = a + b - c
                    pa, pb, pc = abs(pp - a), abs(pp - b), abs(pp - c)
                    pred = a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)
                cur[i] = (r[i] - pred) & 0xff
            out.append(ft); out += cur
            prev = r
        return bytes(out)

    passes = [(0, 0, 8, 8), (4, 0, 8, 8), (0, 4, 4, 8), (2, 0, 4, 4), (0, 2, 2, 4), (1, 0, 2, 2), (0, 1, 1, 2)] if interlace else [(0, 0, 1, 1)]
    raw = b""
    for x0, y0, dx, dy in passes:
        sub = smp[y0::dy, x0::dx]
        if sub.shape[0] == 0 or sub.shape[1] == 0:
            continue
        raw += filt([pack_row(sub[y]) for y in range(sub.shape[0])])

    def chunk(t, d):
        return struct.pack(">I", len(d)) + t + d + struct.pack(">I", zlib.crc32(t + d) & 0xffffffff)

    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, depth, ctype, 0, 0, 1 if interlace else 0))
    if plte is not None:
        out += chunk(b"PLTE", plte)
    if trns is not None:
        out += chunk(b"tRNS", trns)
    half = len(raw) // 2 or 1
    z = zlib.compress(raw, 6)
    return out + chunk(b"IDAT", z[:half]) + chunk(b"IDAT", z[half:]) + chunk(b"IEND", b"")
