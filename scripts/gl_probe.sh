#!/bin/bash
# f-3: is there ANY OpenGL implementation (Mesa llvmpipe / OSMesa / EGL) on the GPU box that could replay the
# reference's conversion shaders?  Writes a log that is committed under profiles/.
out=${1:-gpurun_out/gl_probe.log}
{
echo "== date: $(date -u)"; echo "== uname: $(uname -a)"
echo "== ldconfig GL/EGL/OSMesa/GLX/gbm/glapi libraries:"
ldconfig -p | grep -Ei 'libEGL|libGL\.|libGLX|libOpenGL|libOSMesa|libgbm|libglapi|libGLESv2|libGLdispatch|swrast|llvmpipe|libvulkan' || echo "(none)"
echo "== find mesa/dri/egl vendor files:"
find / -xdev \( -name 'libOSMesa*' -o -name '*swrast*' -o -name 'libgallium*' -o -name '*_dri.so' -o -name 'libEGL_*' -o -name 'libGLX_*' -o -name '*nvidia*egl*' -o -name '10_nvidia.json' -o -name '50_mesa.json' \) 2>/dev/null | head -40 || true
echo "== /usr/share/glvnd/egl_vendor.d:"; ls -la /usr/share/glvnd/egl_vendor.d /etc/glvnd/egl_vendor.d 2>&1 | head
echo "== nvidia driver GL libs:"; ls /usr/lib/x86_64-linux-gnu | grep -Ei 'nvidia.*(gl|egl)|libnvidia-(egl|gl)' | head -20 || echo "(none)"
echo "== NVIDIA_DRIVER_CAPABILITIES=$NVIDIA_DRIVER_CAPABILITIES"
echo "== python GL bindings:"
python - <<'PY'
import importlib
for m in ("OpenGL", "glfw", "moderngl", "pyglet", "vispy", "glcontext", "pyrender", "open3d", "vtk", "trimesh", "pygame", "wgpu", "vulkan"):
    try:
        importlib.import_module(m); print("  import", m, ": OK")
    except Exception as e:
        print("  import", m, ":", type(e).__name__)
import ctypes, ctypes.util
for lib in ("EGL", "GL", "OSMesa", "OpenGL", "GLX", "vulkan"):
    print("  find_library(%s) = %s" % (lib, ctypes.util.find_library(lib)))
for name in ("libEGL.so.1", "libEGL_nvidia.so.0", "libOSMesa.so.8", "libGL.so.1"):
    try:
        ctypes.CDLL(name); print("  dlopen", name, ": OK")
    except OSError as e:
        print("  dlopen", name, ":", e)
# try an EGL device-platform context on the NVIDIA driver (headless), if libEGL exists
try:
    egl = ctypes.CDLL("libEGL.so.1")
    egl.eglGetDisplay.restype = ctypes.c_void_p
    egl.eglGetDisplay.argtypes = [ctypes.c_void_p]
    d = egl.eglGetDisplay(None)
    major, minor = ctypes.c_int(), ctypes.c_int()
    ok = egl.eglInitialize(ctypes.c_void_p(d), ctypes.byref(major), ctypes.byref(minor)) if d else 0
    print("  eglGetDisplay(DEFAULT) =", d, "eglInitialize ->", ok, major.value, minor.value)
except OSError as e:
    print("  EGL probe skipped:", e)
PY
echo "== pip wheels with GL in /opt/wheelhouse:"; ls /opt/wheelhouse 2>/dev/null | grep -Ei 'gl|mesa|egl|vulkan|wgpu' || echo "(none)"
echo "== apt packages:"; dpkg -l 2>/dev/null | grep -Ei 'mesa|libgl|libegl|osmesa|glvnd' | head -20 || echo "(none)"
} > "$out" 2>&1
echo "gl probe written to $out"
