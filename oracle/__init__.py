"""Python bindings of the CPU checkers (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  mesh2splat_b200 never does.

  convert(...)            the plain-C restatement (oracle/m2s_oracle.c) of the whole pass
  triangle_setup(...)     its per-triangle stage, field by field
  ref_gs / ref_fs         the reference's own GLSL compiled as C++ (oracle/_ref), if built
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from mesh2splat_b200 import _abi

os.environ.setdefault("OMP_WAIT_POLICY", "passive")  # idle OpenMP threads sleep instead of burning a CPU quota
_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None
_ref = None


class orc_setup(C.Structure):
    _fields_ = [("ouv", (C.c_float * 2) * 3), ("quat", C.c_float * 4), ("scale", C.c_float * 3),
                ("X", C.c_int32 * 3), ("Y", C.c_int32 * 3), ("area2", C.c_int64),
                ("axis", C.c_int32), ("valid", C.c_int32),
                ("A", C.c_int64 * 3), ("B", C.c_int64 * 3), ("Cc", C.c_int64 * 3),
                ("incl", C.c_int32 * 3), ("x0", C.c_int32), ("y0", C.c_int32),
                ("x1", C.c_int32), ("y1", C.c_int32)]


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "libm2s_oracle.so")
        if not os.path.exists(path):
            from . import build as _b  # noqa: WPS433
            _b.build_oracle()
        _lib = C.CDLL(path)
        _lib.orc_convert.restype = C.c_uint64
        _lib.orc_convert.argtypes = [C.POINTER(_abi.m2s_scene), C.POINTER(_abi.m2s_params), C.c_void_p,
                                     C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64), C.c_int]
        _lib.orc_prepare.restype = C.c_void_p
        _lib.orc_prepare.argtypes = [C.POINTER(_abi.m2s_scene)]
        _lib.orc_release.restype = None
        _lib.orc_release.argtypes = [C.c_void_p]
        _lib.orc_convert_prepared.restype = C.c_uint64
        _lib.orc_convert_prepared.argtypes = [C.POINTER(_abi.m2s_scene), C.c_void_p, C.POINTER(_abi.m2s_params), C.c_void_p,
                                              C.c_uint64, C.c_void_p, C.POINTER(C.c_uint64), C.c_int]
        _lib.orc_triangle_setup.restype = C.c_int
        _lib.orc_triangle_setup.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(orc_setup)]
        _lib.orc_mip_level.restype = C.c_int
        _lib.orc_mip_level.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p,
                                       C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        _lib.orc_mip_count.restype = C.c_uint32
        _lib.orc_mip_count.argtypes = [C.c_uint32, C.c_uint32]
        _lib.orc_sample.restype = None
        _lib.orc_sample.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_float, C.c_float, C.c_float, C.c_void_p]
        _lib.orc_fragment.restype = None
        _lib.orc_fragment.argtypes = [C.c_void_p] * 8 + [C.c_uint32, C.c_void_p, C.c_void_p]
        _lib.orc_encode.restype = None
        _lib.orc_encode.argtypes = [C.c_uint32, C.c_void_p, C.c_float, C.c_void_p]
        _lib.orc_ply_header.restype = C.c_size_t
        _lib.orc_ply_header.argtypes = [C.c_uint32, C.c_uint64, C.c_char_p, C.c_size_t]
        _lib.orc_record_stride.restype = C.c_uint32
        _lib.orc_record_stride.argtypes = [C.c_uint32]
        _lib.orc_compute_bboxes.restype = None
        _lib.orc_compute_bboxes.argtypes = [C.c_void_p, C.POINTER(_abi.m2s_primitive), C.c_uint32, C.c_int]
        _lib.orc_max_threads.restype = C.c_int
    return _lib


def ref_lib():
    """The reference-shader library, or None if it was never built (no /root/reference)."""
    global _ref
    if _ref is None:
        path = os.path.join(_HERE, "_ref", "libm2s_refshader.so")
        if not os.path.exists(path):
            return None
        _ref = C.CDLL(path)
        _ref.ref_gs.restype = C.c_int
        _ref.ref_gs.argtypes = [C.c_void_p] * 6
        _ref.ref_fs.restype = C.c_int
        _ref.ref_fs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p,
                                C.c_uint32, C.c_int, C.c_void_p, C.POINTER(C.c_uint32)]
    return _ref


_refply = None


def ref_ply_lib():
    """The reference's own .ply writer (oracle/_ref/libm2s_refply.so), or None if it was never built."""
    global _refply
    if _refply is None:
        path = os.path.join(_HERE, "_ref", "libm2s_refply.so")
        if not os.path.exists(path):
            return None
        _refply = C.CDLL(path)
        _refply.ref_save_ply.restype = C.c_int
        _refply.ref_save_ply.argtypes = [C.c_char_p, C.c_void_p, C.c_uint64, C.c_uint, C.c_float]
    return _refply


def ref_save_ply(path: str, ref96: np.ndarray, fmt: int, mult: float) -> bool:
    """parsers::savePlyVector of the reference on REF96 records; False if the library is unavailable."""
    r = ref_ply_lib()
    if r is None:
        return False
    flat = np.ascontiguousarray(ref96).view(np.float32).reshape(-1, 24)
    r.ref_save_ply(str(path).encode(), flat.ctypes.data, len(flat), fmt, mult)
    return True


_refloader = None


def ref_loader_lib():
    """The reference's own .glb parser (oracle/_ref/libm2s_refloader.so), or None if it was never built."""
    global _refloader
    if _refloader is None:
        path = os.path.join(_HERE, "_ref", "libm2s_refloader.so")
        if not os.path.exists(path):
            return None
        L = C.CDLL(path)
        L.ref_glb_parse.restype = C.c_void_p
        L.ref_glb_parse.argtypes = [C.c_char_p]
        for f in ("ref_glb_ok", "ref_glb_mesh_count"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [C.c_void_p]
        L.ref_glb_mesh_name.restype = C.c_char_p
        L.ref_glb_mesh_name.argtypes = [C.c_void_p, C.c_int]
        L.ref_glb_face_count.restype = C.c_int
        L.ref_glb_face_count.argtypes = [C.c_void_p, C.c_int]
        L.ref_glb_faces.restype = None
        L.ref_glb_faces.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_glb_base_color.restype = None
        L.ref_glb_base_color.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.ref_glb_texture.restype = C.c_uint64
        L.ref_glb_texture.argtypes = [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                      C.POINTER(C.POINTER(C.c_ubyte))]
        L.ref_glb_free.restype = None
        L.ref_glb_free.argtypes = [C.c_void_p]
        _refloader = L
    return _refloader


def ref_load_glb(path: str):
    """SceneManager::parseGltfFile of the reference (tinygltf + stb_image).  Returns None if the library is
    unavailable, else (ok, meshes) with meshes = [{name, faces (n,36) float32 as 3 x {pos3 nrm3 tan4 uv2},
    base_color (4,), textures: {0|1|2: uint8 array (h, w, channels)}}]."""
    L = ref_loader_lib()
    if L is None:
        return None
    h = L.ref_glb_parse(str(path).encode())
    try:
        meshes = []
        for i in range(L.ref_glb_mesh_count(h)):
            n = L.ref_glb_face_count(h, i)
            faces = np.zeros((n, 36), np.float32)
            if n:
                L.ref_glb_faces(h, i, faces.ctypes.data)
            bc = np.zeros(4, np.float32)
            L.ref_glb_base_color(h, i, bc.ctypes.data)
            tex = {}
            for which in range(3):
                w, hh, ch = C.c_int(0), C.c_int(0), C.c_int(0)
                data = C.POINTER(C.c_ubyte)()
                nbytes = L.ref_glb_texture(h, i, which, C.byref(w), C.byref(hh), C.byref(ch), C.byref(data))
                if nbytes:
                    arr = np.ctypeslib.as_array(data, shape=(int(nbytes),)).copy()
                    per = int(nbytes) // max(1, w.value * hh.value)
                    tex[which] = arr.reshape(hh.value, w.value, per) if per * w.value * hh.value == nbytes else arr
            meshes.append({"name": L.ref_glb_mesh_name(h, i).decode("utf-8", "replace"), "faces": faces, "base_color": bc, "textures": tex})
        return bool(L.ref_glb_ok(h)), meshes
    finally:
        L.ref_glb_free(h)


def max_threads() -> int:
    return int(lib().orc_max_threads())


_usable = None


def usable_threads() -> int:
    """Threads worth starting: the CPUs this process may run on, capped by the number of physical cores
    among them and by the cgroup CPU quota.  (Measured on a GPU box: 128 hardware threads allowed, 64
    of them usable -> 128 OpenMP threads ran the pass 30x slower than 64 did.)"""
    global _usable
    if _usable is not None:
        return _usable
    try:
        cpus = sorted(os.sched_getaffinity(0))
    except AttributeError:
        cpus = list(range(os.cpu_count() or 1))
    n = len(cpus)
    cores = set()
    for c in cpus:
        try:
            with open(f"/sys/devices/system/cpu/cpu{c}/topology/thread_siblings_list") as f:
                cores.add(f.read().strip())
        except OSError:
            cores.add(str(c))
    n = min(n, max(1, len(cores)))
    for path in ("/sys/fs/cgroup/cpu.max",):
        try:
            with open(path) as f:
                quota, period = f.read().split()[:2]
            if quota != "max":
                n = min(n, max(1, int(int(quota) / int(period))))
        except (OSError, ValueError):
            pass
    try:
        with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f, open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as g:
            q, per = int(f.read()), int(g.read())
        if q > 0 and per > 0:
            n = min(n, max(1, q // per))
    except (OSError, ValueError):
        pass
    _usable = max(1, n)
    return _usable


def calibrate_threads(prep: "Prepared", resolution: int = 128, layout: int = _abi.LAYOUT_PACKED56) -> int:
    """Thread count that runs the pass fastest on this host: usable_threads() and its neighbours are timed
    on a small conversion (hyper-threads and container quotas make the best count a measurement, not a lookup)."""
    import time
    try:
        allowed = len(os.sched_getaffinity(0))
    except AttributeError:
        allowed = os.cpu_count() or 1
    base = usable_threads()
    cand = sorted({max(1, base // 2), base, min(allowed, base * 2)})
    best, best_t = base, float("inf")
    out = None
    for c in cand:
        _, _, out = prep.convert(resolution, layout, out=out, threads=c)
        t0 = time.perf_counter()
        for _ in range(2):
            _, _, out = prep.convert(resolution, layout, out=out, threads=c)
        t = time.perf_counter() - t0
        if t < best_t:
            best, best_t = c, t
    return best


def convert(scene: _abi.Scene, resolution: int, layout: int = _abi.LAYOUT_REF96, gaussian_std: float = 0.65,
            max_gaussians: int = 0, flags: int = 0, first_triangle: int = 0, triangle_count: int = 0,
            capacity: int | None = None, want_keys: bool = True, threads: int = 0, row_begin: int = 0, row_end: int = 0):
    """Whole pass on the CPU. Returns (records[structured], keys|None, total)."""
    cs, keep = scene.c_struct()
    p = _abi.make_params(resolution, layout, gaussian_std, max_gaussians, flags, first_triangle, triangle_count,
                         row_begin, row_end)
    if capacity is None:
        capacity = max_gaussians if max_gaussians else (
            _abi.reference_capacity(resolution, len(scene.primitives)) if not (flags & _abi.FLAG_UNCAPPED)
            else 6 * resolution * resolution * max(1, len(scene.primitives)))
    stride = _abi.STRIDES[layout]
    out = np.zeros(capacity * stride, np.uint8)
    keys = np.zeros(capacity, np.uint64) if want_keys else None
    total = C.c_uint64(0)
    n = lib().orc_convert(C.byref(cs), C.byref(p), out.ctypes.data, capacity,
                          keys.ctypes.data if want_keys else None, C.byref(total), threads or usable_threads())
    del keep
    rec = out[: n * stride].view(_abi.record_dtype(layout))
    return rec, (keys[:n] if want_keys else None), int(total.value)


class Prepared:
    """Scene with its mip chains built once; convert() can then be timed like the reference's
    steady-state ConversionPass::execute (mesh + textures already resident)."""

    def __init__(self, scene: _abi.Scene):
        self.scene = scene
        self.cs, self._keep = scene.c_struct()
        self.handle = lib().orc_prepare(C.byref(self.cs))

    def convert(self, resolution: int, layout: int = _abi.LAYOUT_REF96, gaussian_std: float = 0.65, max_gaussians: int = 0,
                flags: int = 0, capacity: int | None = None, out: np.ndarray | None = None, threads: int = 0):
        p = _abi.make_params(resolution, layout, gaussian_std, max_gaussians, flags)
        if capacity is None:
            capacity = max_gaussians or (_abi.reference_capacity(resolution, len(self.scene.primitives))
                                         if not (flags & _abi.FLAG_UNCAPPED)
                                         else 6 * resolution * resolution * max(1, len(self.scene.primitives)))
        stride = _abi.STRIDES[layout]
        if out is None:
            out = np.empty(capacity * stride, np.uint8)
        total = C.c_uint64(0)
        n = lib().orc_convert_prepared(C.byref(self.cs), self.handle, C.byref(p), out.ctypes.data, capacity, None,
                                       C.byref(total), threads or usable_threads())
        return int(n), int(total.value), out

    def close(self):
        if self.handle:
            lib().orc_release(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # noqa: BLE001
            pass


def triangle_setup(tri36: np.ndarray, bmin, bmax, resolution: int) -> orc_setup:
    t = np.ascontiguousarray(tri36, np.float32)
    a = np.asarray(bmin, np.float32); b = np.asarray(bmax, np.float32)
    s = orc_setup()
    lib().orc_triangle_setup(t.ctypes.data, a.ctypes.data, b.ctypes.data, resolution, C.byref(s))
    return s


def mip_level(img: np.ndarray, level: int) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape[:2]
    dst = np.zeros((h, w, 4), np.uint8)
    ow, oh = C.c_uint32(0), C.c_uint32(0)
    r = lib().orc_mip_level(img.ctypes.data, w, h, level, dst.ctypes.data, C.byref(ow), C.byref(oh))
    if r != 0:
        raise ValueError("level out of range")
    return dst.reshape(-1)[: ow.value * oh.value * 4].reshape(oh.value, ow.value, 4).copy()


def mip_count(w: int, h: int) -> int:
    return int(lib().orc_mip_count(w, h))


def sample(img: np.ndarray, u: float, v: float, lam: float) -> np.ndarray:
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros(4, np.float32)
    lib().orc_sample(img.ctypes.data, img.shape[1], img.shape[0], u, v, lam, out.ctypes.data)
    return out


def fragment(P, N, T, scale, quat, albedo, nrm, mr, flags: int, factor) -> np.ndarray:
    arrs = [np.ascontiguousarray(x, np.float32) for x in (P, N, T, scale, quat, albedo, nrm, mr)]
    f = np.ascontiguousarray(factor, np.float32)
    rec = np.zeros(24, np.float32)
    lib().orc_fragment(*[a.ctypes.data for a in arrs], flags, f.ctypes.data, rec.ctypes.data)
    return rec


def encode(layout: int, rec24: np.ndarray, mult: float) -> np.ndarray:
    r = np.ascontiguousarray(rec24, np.float32)
    dst = np.zeros(_abi.STRIDES[layout], np.uint8)
    lib().orc_encode(layout, r.ctypes.data, mult, dst.ctypes.data)
    return dst


def ply_header(fmt: int, count: int) -> bytes:
    buf = C.create_string_buffer(8192)
    n = lib().orc_ply_header(fmt, count, buf, 8192)
    return buf.raw[:n]


def ply_bytes(ref96: np.ndarray, fmt: int, mult: float) -> bytes:
    """Whole .ply file (header + body) from REF96 records, as parsers::savePlyVector writes it."""
    layout = _abi.PLY_FORMAT_LAYOUT.get(fmt, _abi.LAYOUT_PLY_STANDARD)
    flat = np.ascontiguousarray(ref96).view(np.float32).reshape(-1, 24)
    stride = _abi.STRIDES[layout]
    body = np.zeros(len(flat) * stride, np.uint8)
    L = lib()
    for i in range(len(flat)):
        L.orc_encode(layout, flat[i].ctypes.data, mult, body.ctypes.data + i * stride)
    return ply_header(fmt, len(flat)) + body.tobytes()


def ref_gs(tri36, bmin, bmax):
    """Reference geometry shader. Returns (glpos[3,4], scale[3], quat_wxyz[4]) or None if unavailable."""
    r = ref_lib()
    if r is None:
        return None
    t = np.ascontiguousarray(tri36, np.float32)
    a = np.asarray(bmin, np.float32); b = np.asarray(bmax, np.float32)
    glpos = np.zeros((3, 4), np.float32); sc = np.zeros(3, np.float32); q = np.zeros(4, np.float32)
    n = r.ref_gs(t.ctypes.data, a.ctypes.data, b.ctypes.data, glpos.ctypes.data, sc.ctypes.data, q.ctypes.data)
    assert n == 3
    return glpos, sc, q


def ref_fs(varyings19, albedo, nrm, mr, flags: int, factor, counter_start: int = 0, max_gaussians: int = 1 << 30):
    """Reference fragment shader. Returns (written, rec24, counter_after) or None if unavailable."""
    r = ref_lib()
    if r is None:
        return None
    arrs = [np.ascontiguousarray(x, np.float32) for x in (varyings19, albedo, nrm, mr)]
    f = np.ascontiguousarray(factor, np.float32)
    rec = np.zeros(24, np.float32)
    cnt = C.c_uint32(0)
    w = r.ref_fs(arrs[0].ctypes.data, arrs[1].ctypes.data, arrs[2].ctypes.data, arrs[3].ctypes.data, flags,
                 f.ctypes.data, counter_start, max_gaussians, rec.ctypes.data, C.byref(cnt))
    return bool(w), rec, int(cnt.value)


# ---- viewer prepass ----------------------------------------------------------------------------
_refpre = None


def ref_prepass_lib():
    """The reference's own prepass compute shader compiled as C++ (oracle/_ref/libm2s_refprepass.so), or None."""
    global _refpre
    if _refpre is None:
        path = os.path.join(_HERE, "_ref", "libm2s_refprepass.so")
        if not os.path.exists(path):
            return None
        _refpre = C.CDLL(path)
        _refpre.ref_prepass.restype = C.c_uint32
        _refpre.ref_prepass.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int,
                                        C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    return _refpre


def _prepass_call(fn, gaussians24, world_to_view, view_to_clip, model_to_world, resolution, near_far, std_dev, render_mode, fmt, ply_has_pbr):
    g = np.ascontiguousarray(gaussians24, np.float32).reshape(-1, 24)
    mats = [np.ascontiguousarray(np.asarray(m, np.float32).reshape(4, 4).T.ravel() if transpose else m, np.float32)
            for m, transpose in ((world_to_view, False), (view_to_clip, False), (model_to_world, False))]
    res, nf = np.asarray(resolution, np.float32), np.asarray(near_far, np.float32)
    quads = np.zeros((len(g), 24), np.float32)
    depths = np.zeros(len(g), np.float32)
    n = fn(g.ctypes.data, len(g), mats[0].ctypes.data, mats[1].ctypes.data, mats[2].ctypes.data, res.ctypes.data, nf.ctypes.data,
           C.c_float(std_dev), int(render_mode), int(fmt), int(ply_has_pbr), quads.ctypes.data, depths.ctypes.data)
    return quads[:n].copy(), depths[:n].copy()


def prepass(gaussians24, world_to_view, view_to_clip, model_to_world, resolution, near_far, std_dev, render_mode=0, fmt=0, ply_has_pbr=0):
    """orc_prepass: the C restatement of gaussianSplattingPrepassCS.glsl.  Matrices: 16 floats, column-major (glm::mat4).
    Returns (quads [m, 24], depths [m]) in input order of the survivors."""
    L = lib()
    L.orc_prepass.restype = C.c_uint32
    L.orc_prepass.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_uint32,
                              C.c_uint32, C.c_void_p, C.c_void_p]
    return _prepass_call(L.orc_prepass, gaussians24, world_to_view, view_to_clip, model_to_world, resolution, near_far, std_dev, render_mode, fmt, ply_has_pbr)


def ref_prepass(gaussians24, world_to_view, view_to_clip, model_to_world, resolution, near_far, std_dev, render_mode=0, fmt=0, ply_has_pbr=0):
    """The reference's own shader on the same inputs (needs oracle/_ref, i.e. the build container)."""
    L = ref_prepass_lib()
    if L is None:
        raise RuntimeError("oracle/_ref/libm2s_refprepass.so is not built")
    return _prepass_call(L.ref_prepass, gaussians24, world_to_view, view_to_clip, model_to_world, resolution, near_far, std_dev, render_mode, fmt, ply_has_pbr)


def packed56_as_gaussian_vertex(packed: np.ndarray) -> np.ndarray:
    """PACKED56 records as the GaussianVertex array the reference builds when it loads a standard .ply without PBR values
    (parsers.cpp:560-622: scale = exp(log scale), colour = SH0 * C0 + 0.5, alpha = sigmoid(opacity), no normal, pbr = 0)."""
    f = np.ascontiguousarray(packed).view(np.float32).reshape(-1, 14)
    g = np.zeros((len(f), 24), np.float32)
    g[:, 0:3] = f[:, 0:3]; g[:, 3] = 1.0
    g[:, 4:7] = f[:, 10:13] * np.float32(0.28209479177387814) + np.float32(0.5)
    g[:, 7] = (1.0 / (1.0 + np.exp(-f[:, 13].astype(np.float64)))).astype(np.float32)
    g[:, 8:11] = np.exp(f[:, 7:10].astype(np.float32))
    g[:, 16:20] = f[:, 3:7]
    return g
