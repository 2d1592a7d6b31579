"""Timeline of the raster kernel's warps (needs a -DM2S_TRACE build selected with M2S_LIB)."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mesh2splat_b200 import synth, _abi, _lib
from mesh2splat_b200.api import Context
layout = {"ref96": 0, "packed56": 1}[sys.argv[1] if len(sys.argv) > 1 else "packed56"]
R = int(sys.argv[2]) if len(sys.argv) > 2 else 512
which = sys.argv[3] if len(sys.argv) > 3 else "helmet"
ctx = Context(0)
scene = {"helmet": lambda: synth.helmet_standin(2048), "quad": synth.unit_quad, "dh": lambda: synth.damaged_helmet_standin(2048)}[which]()
ds = ctx.upload(scene)
nw = 148 * 16
tr = torch.zeros(nw * 16, dtype=torch.int64, device="cuda")
_lib.lib().m2s_debug_set_trace(C.c_void_p(tr.data_ptr()))
out = None
for i in range(5):
    tr.zero_()
    out = ctx.convert(ds, R, layout, flags=_abi.FLAG_UNCAPPED, capacity=6 * R * R, out=out.data if out else None)
t = tr.cpu().numpy().reshape(nw, 16).astype(np.float64)
t0 = t[:, 11][t[:, 11] > 0].min() if (t[:, 11] > 0).any() else t[:, 0][t[:, 0] > 0].min()
names = ["start", "tma_done", "setup_done", "walk_done", "scan_done", "flush_done", "unit_end", "units_done", "list_done", "direct_done", "-", "kernel_entry", "help_done", "stores_done", "cta_synced"]
if os.environ.get("TRACE_SETUP"):
    names += ["s:prim_loaded", "s:quat_done", "s:scale_done", "s:raster_done"]
print(f"{which} R={R} layout={layout}: device_ms={out.device_ms:.4f} total={out.total}")
for k, n in enumerate(names):
    v = t[:, k]; v = v[v > 0]
    if len(v):
        r = (v - t0) / 1e3
        print(f"{n:12s} n={len(v):5d}  min {r.min():7.2f}  p50 {np.median(r):7.2f}  p90 {np.percentile(r, 90):7.2f}  max {r.max():7.2f} us")
d = (t[:, 9] - t[:, 8]) / 1e3; n = t[:, 10]; m = (t[:, 9] > 0)
if m.any():
    print("direct shading per warp: us p50 %.2f p90 %.2f max %.2f; fragments p50 %d max %d; us per 32-fragment group p50 %.2f" % (np.median(d[m]), np.percentile(d[m], 90), d[m].max(), np.median(n[m]), n[m].max(), np.median(d[m] / np.maximum(1, np.ceil(n[m] / 32)))))
sys.exit(0)
it = t[:, 12]
print("drain items/warp: mean %.1f max %d; per-item us: load %.2f setup %.2f raster+flush %.2f" % (
    it.mean(), it.max(), t[:, 13].sum() / max(it.sum(), 1) / 1e3, t[:, 14].sum() / max(it.sum(), 1) / 1e3, t[:, 15].sum() / max(it.sum(), 1) / 1e3))
