"""Target for an ncu launch list (gpu__time_duration.sum): two conversions of each BASELINE configuration, in a fixed
order, so the per-kernel times can be attributed by position.  usage: launch_configs.py [names...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesh2splat_b200 import synth, _abi
from mesh2splat_b200.api import Context

P, Rf, U = _abi.LAYOUT_PACKED56, _abi.LAYOUT_REF96, _abi.FLAG_UNCAPPED
CONFIGS = {
    "quad64": (synth.unit_quad, 64, Rf, 0),
    "helmet512": (lambda: synth.helmet_standin(2048), 512, P, U),
    "helmet512_ref96": (lambda: synth.helmet_standin(2048), 512, Rf, U),
    "sponza1024": (lambda: synth.sponza_standin(1024), 1024, P, U),
    "sphere1m_256": (lambda: synth.sphere_1m(2048), 256, P, U),
    "dh1024": (lambda: synth.damaged_helmet_standin(2048), 1024, P, U),
    "dh2048": (lambda: synth.damaged_helmet_standin(2048), 2048, P, U),
}
names = sys.argv[1:] or list(CONFIGS)
ctx = Context(0)
for name in names:
    make, R, layout, flags = CONFIGS[name]
    scene = make()
    ds = ctx.upload(scene)
    cap = 6 * R * R * max(1, len(scene.primitives)) if flags & U else _abi.reference_capacity(R, len(scene.primitives))
    cap = min(cap, 60_000_000)
    out = None
    for i in range(2):
        out = ctx.convert(ds, R, layout, flags=flags, capacity=cap, out=out.data if out else None)
    print(f"CONFIG {name} total {out.total} ms {out.device_ms:.4f}", flush=True)
    ds.free()
