""".glb -> Scene through the C-ABI loader (m2s_glb_load: SceneManager::parseGltfFile +
setupMeshBuffers bbox rule + loadTextures, src/utils/SceneManager.cpp:195-649)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _abi
from ._lib import M2SError, check, lib


def load_glb(path: str, cumulative_bbox: bool = True) -> _abi.Scene:
    """Raises OSError if the file cannot be read, ValueError if it is not a loadable .glb."""
    h = C.c_void_p(0)
    try:
        check(lib().m2s_glb_load(str(path).encode(), 1 if cumulative_bbox else 0, C.byref(h)))
    except M2SError as e:
        if e.status == _abi.M2S_E_IO:
            raise OSError(e.message) from None
        raise ValueError(e.message) from None
    try:
        v = lib().m2s_hscene_view(h).contents
        nt = int(v.triangle_count)
        tris = np.ctypeslib.as_array(C.cast(v.triangles, C.POINTER(C.c_float)), shape=(nt * 36,)).copy().reshape(nt, 36) \
            if nt else np.zeros((0, 36), np.float32)
        texs = []
        for i in range(v.texture_count):
            t = v.textures[i]
            a = np.ctypeslib.as_array(C.cast(t.rgba, C.POINTER(C.c_uint8)), shape=(t.height, t.width, 4)).copy()
            texs.append(a)
        prims = []
        for i in range(v.primitive_count):
            p = v.primitives[i]
            prims.append(_abi.Primitive(int(p.first_triangle), int(p.triangle_count), tuple(p.base_color_factor),
                                        int(p.albedo_texture), int(p.normal_texture), int(p.metallic_roughness_texture),
                                        tuple(p.bbox_min), tuple(p.bbox_max),
                                        lib().m2s_hscene_primitive_name(h, i).decode("utf-8", "replace")))
        scene = _abi.Scene(tris, prims, texs)
        if not v.primitive_count:
            scene.primitives = []
        return scene
    finally:
        lib().m2s_hscene_free(h)
