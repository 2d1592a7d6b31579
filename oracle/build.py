#!/usr/bin/env python3
"""Build the checkers under oracle/ (TEST INFRASTRUCTURE — never linked into the product).

  libm2s_oracle.so            the plain-C restatement (m2s_oracle.c), always built (gcc, OpenMP)
  _ref/libm2s_refshader.so    the REFERENCE's own conversion shaders — converterGS.glsl and
                              converterFS.glsl, read where they lie under /root/reference and
                              turned into a C++ translation unit by the mechanical token rewrites
                              below — compiled against the reference's vendored GLM.  Only built
                              when /root/reference exists (this container); the GPU box uses the
                              prebuilt .so that travels with the snapshot.  Nothing from the
                              reference is written into the repository: generated files and the
                              .so live in oracle/_ref/ (git-ignored).

The reference application itself (Win32 + OpenGL 4.6 + GLFW/GLEW .lib files, CMake non-Windows
branch is a todo) cannot be built here; see DESIGN.md.  What _ref pins is the shader ARITHMETIC
(per-triangle stage and per-fragment stage); the rasteriser and the texture unit are GL-driver
behaviour and have no reference source to compile.
"""
from __future__ import annotations

import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("M2S_REFERENCE", "/root/reference")
REF_OUT = os.path.join(HERE, "_ref")
CFLAGS = ["-O2", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-fvisibility=hidden"]


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise RuntimeError("build failed: " + cmd[0])


def _newer(target: str, *deps: str) -> bool:
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(d) <= t for d in deps if os.path.exists(d))


def build_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "m2s_oracle.c")
    out = os.path.join(HERE, "libm2s_oracle.so")
    hdr = os.path.join(HERE, "..", "include", "m2s.h")
    if force or not _newer(out, src, hdr, __file__):
        _run(["gcc", "-std=c11", "-fopenmp", *CFLAGS, "-o", out, src, "-lm"])
    return out


# ---- GLSL -> C++ token rewrites (mechanical; the arithmetic is untouched) -------------------
_FLOAT_LIT = re.compile(r"(?<![\w.])((?:\d+\.\d*|\.\d+)(?:[eE][-+]?\d+)?|\d+[eE][-+]?\d+)(?![\w.])")


def _common(src: str) -> str:
    src = re.sub(r"^\s*#version.*$", "", src, flags=re.M)
    src = _FLOAT_LIT.sub(lambda m: m.group(1) + "f", src)  # GLSL literals are float
    return src


def glsl_gs_to_cpp(src: str) -> str:
    src = _common(src)
    src = re.sub(r"^\s*layout\s*\([^)]*\)\s*(in|out)\s*;\s*$", "", src, flags=re.M)
    src = re.sub(r"in\s+VS_OUT\s*\{(.*?)\}\s*gs_in\[\]\s*;", r"struct VS_OUT {\1}; static VS_OUT gs_in[3];", src, flags=re.S)
    src = re.sub(r"^\s*uniform\s+", "static ", src, flags=re.M)
    src = re.sub(r"^\s*(?:flat\s+)?out\s+(\w+\s+\w+\s*;)", r"static \1", src, flags=re.M)
    # parameter qualifiers
    src = re.sub(r"([(,]\s*)in\s+(\w+\s+\w+)", r"\1\2", src)
    src = re.sub(r"([(,]\s*)out\s+(\w+)\s+(\w+)", r"\1\2& \3", src)
    src = src.replace("void main()", "void gs_main()")
    return src


def glsl_fs_to_cpp(src: str) -> str:
    src = _common(src)
    src = re.sub(r"layout\s*\(std430[^)]*\)\s*buffer\s+GaussianBuffer\s*\{.*?\}\s*gaussianBuffer\s*;",
                 "static struct { GaussianVertex* vertices; } gaussianBuffer;", src, flags=re.S)
    src = re.sub(r"layout\s*\([^)]*\)\s*uniform\s+atomic_uint", "static atomic_uint", src)
    src = re.sub(r"^\s*uniform\s+", "static ", src, flags=re.M)
    src = re.sub(r"^\s*in\s+(\w+\s+\w+\s*;)", r"static \1", src, flags=re.M)
    src = re.sub(r"\bdiscard\s*;", "return;", src)
    # swizzles used by the shader: .xyz of vec3/vec4 -> vec3(...), .bg -> helper
    src = re.sub(r"(texture\([^()]*\))\.xyz", r"vec3(\1)", src)
    src = re.sub(r"(texture\([^()]*\))\.bg", r"swz_bg(\1)", src)
    src = re.sub(r"\b(\w+)\.xyz\b", r"vec3(\1)", src)
    src = src.replace("void main()", "void fs_main()")
    return src


def build_ref(force: bool = False) -> str | None:
    gs = os.path.join(REF, "src", "shaders", "conversion", "converterGS.glsl")
    fs = os.path.join(REF, "src", "shaders", "conversion", "converterFS.glsl")
    glm = os.path.join(REF, "thirdParty", "glm")
    out = os.path.join(REF_OUT, "libm2s_refshader.so")
    if not (os.path.exists(gs) and os.path.exists(fs) and os.path.isdir(glm)):
        return out if os.path.exists(out) else None
    harness = os.path.join(HERE, "ref_harness.cpp")
    if not force and _newer(out, gs, fs, harness, __file__):
        return out
    os.makedirs(REF_OUT, exist_ok=True)
    with open(gs) as f:
        gs_cpp = glsl_gs_to_cpp(f.read())
    with open(fs) as f:
        fs_cpp = glsl_fs_to_cpp(f.read())
    with open(os.path.join(REF_OUT, "converterGS.inc"), "w") as f:
        f.write(gs_cpp)
    with open(os.path.join(REF_OUT, "converterFS.inc"), "w") as f:
        f.write(fs_cpp)
    _run(["g++", "-std=gnu++17", *CFLAGS, "-w", "-I", glm, "-I", REF_OUT, "-o", out, harness])
    return out


_STUBS = {
    "crtdbg.h": "/* empty stand-in: the reference includes <crtdbg.h> (MSVC debug CRT) */\n",
    "windows.h": ("/* stand-in for <windows.h> on Linux: the .ply writer uses nothing from it */\n#pragma once\n"
                  "typedef unsigned long DWORD; typedef char* LPSTR; typedef void* HMODULE;\n#ifndef MAX_PATH\n#define MAX_PATH 260\n#endif\n"
                  "static inline DWORD GetModuleFileNameA(HMODULE, LPSTR buf, DWORD n) { if (n) buf[0] = 0; return 0; }\n"),
    "compat.h": "/* forced include: MSVC lets the reference call isnan() unqualified */\n#include <cmath>\nusing std::isnan;\n",
}


def build_ref_ply(force: bool = False) -> str | None:
    """The reference's own .ply writer (src/parsers/parsers.cpp + src/utils/utils.cpp, compiled where they lie)."""
    srcs = [os.path.join(REF, "src", "parsers", "parsers.cpp"), os.path.join(REF, "src", "utils", "utils.cpp")]
    out = os.path.join(REF_OUT, "libm2s_refply.so")
    if not all(os.path.exists(x) for x in srcs):
        return out if os.path.exists(out) else None
    harness = os.path.join(HERE, "ref_ply_harness.cpp")
    if not force and _newer(out, *srcs, harness, __file__):
        return out
    stubs = os.path.join(REF_OUT, "stubs")
    os.makedirs(stubs, exist_ok=True)
    for name, text in _STUBS.items():
        with open(os.path.join(stubs, name), "w") as f:
            f.write(text)
    tp = os.path.join(REF, "thirdParty")
    inc = ["-I", stubs, "-I", os.path.join(REF, "src"), "-I", os.path.join(REF, "src", "utils"), "-I", tp, "-I", os.path.join(tp, "glm"),
           "-I", os.path.join(tp, "glew", "include"), "-I", os.path.join(tp, "GLFW", "include"), "-I", os.path.join(tp, "imgui"),
           "-I", os.path.join(tp, "imgui", "backends")]
    _run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-w", "-ffp-contract=off", "-DGLEW_NO_GLU", "-include", os.path.join(stubs, "compat.h"),
          *inc, "-o", out, *srcs, harness, "-lstdc++fs"])
    return out


def _ref_includes(stubs: str) -> list:
    tp = os.path.join(REF, "thirdParty")
    return ["-I", stubs, "-I", os.path.join(REF, "src"), "-I", os.path.join(REF, "src", "utils"), "-I", tp, "-I", os.path.join(tp, "glm"),
            "-I", os.path.join(tp, "glew", "include"), "-I", os.path.join(tp, "GLFW", "include"), "-I", os.path.join(tp, "imgui"),
            "-I", os.path.join(tp, "imgui", "backends"), "-I", os.path.join(tp, "xatlas"), "-I", os.path.join(tp, "ImGuiFileDialog"),
            "-I", os.path.join(tp, "imguizmo")]


def build_ref_loader(force: bool = False) -> str | None:
    """The reference's own .glb parser: SceneManager::parseGltfFile with tinygltf + stb_image (compiled where they lie)."""
    srcs = [os.path.join(REF, "src", "utils", "SceneManager.cpp"), os.path.join(REF, "src", "utils", "utils.cpp"),
            os.path.join(REF, "src", "parsers", "parsers.cpp")]
    out = os.path.join(REF_OUT, "libm2s_refloader.so")
    if not all(os.path.exists(x) for x in srcs):
        return out if os.path.exists(out) else None
    harness = os.path.join(HERE, "ref_loader_harness.cpp")
    if not force and _newer(out, *srcs, harness, __file__):
        return out
    stubs = os.path.join(REF_OUT, "stubs")
    os.makedirs(stubs, exist_ok=True)
    for name, text in _STUBS.items():
        with open(os.path.join(stubs, name), "w") as f:
            f.write(text)
    flags = ["-std=c++17", "-O1", "-fPIC", "-w", "-ffp-contract=off", "-DGLEW_NO_GLU", "-include", os.path.join(stubs, "compat.h"), *_ref_includes(stubs)]
    objs = []
    for src in srcs + [harness]:
        obj = os.path.join(REF_OUT, "ld_" + os.path.splitext(os.path.basename(src))[0] + ".o")
        _run(["g++", *flags, "-c", src, "-o", obj])
        objs.append(obj)
    # the GLEW entry points SceneManager.cpp references are data symbols of glew.c (only a Windows .lib is vendored):
    # null pointers, generated from the object file's undefined symbols — the GL half is never called
    und = subprocess.run(["nm", "-u", *objs], capture_output=True, text=True).stdout
    glew = sorted({w for line in und.splitlines() for w in line.split() if w.startswith("__glew")})
    gl_null = os.path.join(REF_OUT, "glew_null.c")
    with open(gl_null, "w") as f:
        f.write("/* generated by oracle/build.py: null GLEW entry points (never called) */\n")
        for g in glew:
            f.write(f"void* {g} = 0;\n")
    gl_obj = os.path.join(REF_OUT, "ld_glew_null.o")
    _run(["gcc", "-fPIC", "-c", gl_null, "-o", gl_obj])
    _run(["g++", "-shared", "-o", out, *objs, gl_obj, "-lstdc++fs"])
    for o in objs + [gl_obj]:
        os.remove(o)
    return out


def glsl_prepass_to_cpp(src: str, common: str) -> str:
    """gaussianSplattingPrepassCS.glsl + common.glsl -> one C++ include (token rewrites only)."""
    def rw(s: str) -> str:
        s = re.sub(r"^\s*#version.*$", "", s, flags=re.M)
        s = s.replace("highp ", "")
        s = _FLOAT_LIT.sub(lambda m: m.group(1) + "f", s)
        s = re.sub(r"([(,]\s*)in\s+(\w+\s+\w+)", r"\1\2", s)
        s = re.sub(r"([(,]\s*)out\s+(\w+)\s+(\w+)", r"\1\2& \3", s)
        s = re.sub(r"clamp\(([^,()]+(?:\([^()]*\))?[^,()]*),\s*0,\s*1\)", r"clamp(\1, 0.0f, 1.0f)", s)   # GLSL converts the int bounds
        # the one swizzle used as an lvalue
        s = s.replace("pos2d.xyz = pos2d.xyz / pos2d.w;", "{ vec3 t_ = vec3(pos2d) / pos2d.w; pos2d.x = t_.x; pos2d.y = t_.y; pos2d.z = t_.z; }")
        s = re.sub(r"\.(xyz|xy|yx)\b(?!\s*\()", r".\1()", s)                         # rvalue swizzles -> GLM swizzle functions
        s = re.sub(r"gl_GlobalInvocationID\.(xy|yx)\(\)", r"vec2(gl_GlobalInvocationID.\1())", s)   # GLSL converts uvec2 -> vec2
        return s
    src = src.replace('#include "common.glsl"', "")
    src = re.sub(r"layout\s*\(std430,\s*binding\s*=\s*\d+\)\s*(?:readonly|writeonly)?\s*buffer\s+(\w+)\s*\{\s*(\w+)\s+(\w+)\[\];\s*\}\s*(\w+)\s*;",
                 r"static struct { \2* \3; } \4;", src)
    src = re.sub(r"layout\s*\(binding\s*=\s*\d+\)\s*uniform\s+atomic_uint", "static atomic_uint", src)
    src = re.sub(r"layout\s*\(local_size_x[^)]*\)\s*in\s*;", "", src)
    src = re.sub(r"^\s*uniform\s+", "static ", src, flags=re.M)
    src = src.replace("void main()", "void prepass_main()")
    return rw(common) + "\n" + rw(src)


def build_ref_prepass(force: bool = False) -> str | None:
    """The reference's viewer prepass compute shader (row f-4), compiled where it lies against the reference's GLM."""
    cs = os.path.join(REF, "src", "shaders", "rendering", "gaussianSplattingPrepassCS.glsl")
    cm = os.path.join(REF, "src", "shaders", "rendering", "common.glsl")
    glm = os.path.join(REF, "thirdParty", "glm")
    out = os.path.join(REF_OUT, "libm2s_refprepass.so")
    if not (os.path.exists(cs) and os.path.exists(cm) and os.path.isdir(glm)):
        return out if os.path.exists(out) else None
    harness = os.path.join(HERE, "ref_prepass_harness.cpp")
    if not force and _newer(out, cs, cm, harness, __file__):
        return out
    os.makedirs(REF_OUT, exist_ok=True)
    with open(cs) as f, open(cm) as g:
        inc = glsl_prepass_to_cpp(f.read(), g.read())
    with open(os.path.join(REF_OUT, "prepassCS.inc"), "w") as f:
        f.write(inc)
    _run(["g++", "-std=gnu++17", *CFLAGS, "-w", "-I", glm, "-I", REF_OUT, "-o", out, harness])
    return out


def build_all(force: bool = False) -> dict:
    return {"oracle": build_oracle(force), "ref": build_ref(force), "ref_ply": build_ref_ply(force), "ref_loader": build_ref_loader(force),
            "ref_prepass": build_ref_prepass(force)}


if __name__ == "__main__":
    print(build_all(force="--force" in sys.argv))
