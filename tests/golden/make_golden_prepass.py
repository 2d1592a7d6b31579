#!/usr/bin/env python3
"""Generates tests/golden/ref_prepass_vectors.npz from the REFERENCE's own viewer prepass shader.

Runs only in the build container (needs /root/reference): oracle/build.py compiles
src/shaders/rendering/gaussianSplattingPrepassCS.glsl + common.glsl (unmodified arithmetic, token-level GLSL -> C++
rewrites) against the reference's vendored GLM into oracle/_ref/libm2s_refprepass.so; this script feeds it seeded
gaussians and cameras and stores inputs + outputs.  The committed .npz is what travels.

    python tests/golden/make_golden_prepass.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import oracle  # noqa: E402
from oracle import build as obuild  # noqa: E402


def look_at(eye, at, up):
    f = at - eye; f /= np.linalg.norm(f)
    s = np.cross(f, up); s /= np.linalg.norm(s)
    u = np.cross(s, f)
    m = np.eye(4, dtype=np.float32)
    m[0, :3] = s; m[1, :3] = u; m[2, :3] = -f
    m[:3, 3] = -m[:3, :3] @ eye
    return m


def perspective(fovy, aspect, zn, zf):
    t = np.tan(fovy / 2)
    m = np.zeros((4, 4), np.float32)
    m[0, 0] = 1 / (aspect * t); m[1, 1] = 1 / t; m[2, 2] = -(zf + zn) / (zf - zn); m[3, 2] = -1; m[2, 3] = -(2 * zf * zn) / (zf - zn)
    return m


def column_major(m):
    return np.ascontiguousarray(np.asarray(m, np.float32).T).ravel()


def gaussians(rng, n, raw_scale):
    g = np.zeros((n, 24), np.float32)
    g[:, 0:3] = (rng.random((n, 3)) - 0.5) * 4; g[:, 3] = 1
    g[:, 4:8] = rng.random((n, 4))
    g[:, 8:11] = rng.random((n, 3)) * raw_scale + raw_scale * 1e-3
    v = rng.normal(size=(n, 3)); g[:, 12:15] = v / np.linalg.norm(v, axis=1, keepdims=True)
    q = rng.normal(size=(n, 4)); g[:, 16:20] = q / np.linalg.norm(q, axis=1, keepdims=True)
    g[:, 20:22] = rng.random((n, 2))
    return g


def main():
    assert obuild.build_ref_prepass() is not None, "needs /root/reference"
    rng = np.random.default_rng(20260923)
    rot = np.eye(4, dtype=np.float32)
    c, s = np.cos(0.7), np.sin(0.7)
    rot[:3, :3] = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]], np.float32) @ np.diag([1.2, 1.2, 1.2]).astype(np.float32)
    rot[:3, 3] = [0.3, -0.1, 0.2]
    cases = [  # (format, render mode, raw scale range, std_dev, model matrix, eye)
        (0, 0, 8.0, 0.65 / 512, np.eye(4, dtype=np.float32), [3.0, 2.0, 4.0]),
        (0, 2, 8.0, 0.65 / 256, rot, [-2.5, 1.0, 3.0]),
        (0, 1, 30.0, 0.65 / 512, np.diag([1.5, 1.5, 1.5, 1.0]).astype(np.float32), [0.5, 0.2, 2.5]),
        (1, 0, 0.02, 1.0, np.eye(4, dtype=np.float32), [3.0, 2.0, 4.0]),
        (1, 2, 0.05, 1.0, rot, [1.0, -2.0, 3.5]),
    ]
    out = {}
    for i, (fmt, mode, raw, sd, M, eye) in enumerate(cases):
        g = gaussians(rng, 500, raw)
        V = look_at(np.array(eye, np.float64), np.zeros(3), np.array([0.0, 1.0, 0.0])).astype(np.float32)
        P = perspective(np.radians(45.0), 16 / 9, 0.01, 100.0)
        quads, depths = oracle.ref_prepass(g, column_major(V), column_major(P), column_major(M), (1280.0, 720.0), (0.01, 100.0), sd, mode, fmt, 0)
        assert 50 < len(quads) < 500, len(quads)
        out[f"g{i}"] = g; out[f"view{i}"] = column_major(V); out[f"proj{i}"] = column_major(P); out[f"model{i}"] = column_major(M)
        out[f"params{i}"] = np.array([1280.0, 720.0, 0.01, 100.0, sd, mode, fmt], np.float64)
        out[f"quads{i}"] = quads; out[f"depths{i}"] = depths
    out["ncases"] = np.array(len(cases))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_prepass_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", [len(out[f"quads{i}"]) for i in range(len(cases))], "survivors")


if __name__ == "__main__":
    main()
