"""Quick device-time probe (not the contract bench): helmet stand-in at a few densities."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mesh2splat_b200 import synth, _abi
from mesh2splat_b200.api import Context

ctx = Context(0)
print("SMs", ctx.sm_count)
scenes = {"helmet": synth.helmet_standin(2048), "quad": synth.unit_quad()}
if len(sys.argv) > 1 and sys.argv[1] == "all":
    scenes["sphere1m"] = synth.sphere_1m(2048)
for name, s in scenes.items():
    ds = ctx.upload(s)
    for layout in (_abi.LAYOUT_REF96, _abi.LAYOUT_PACKED56):
        for R in (64, 256, 512, 1024, 2048):
            cap = 6 * R * R
            out = None
            ts = []
            for i in range(8):
                out = ctx.convert(ds, R, layout, flags=_abi.FLAG_UNCAPPED, capacity=cap, out=out.data if out else None)
                ts.append(out.device_ms)
            t = float(np.median(ts[3:]))
            stride = _abi.STRIDES[layout]
            nmaps = 3 if layout == 0 else 1
            alg = out.total * stride + s.triangle_count * 144 + sum(min(out.total * 4, t_.nbytes) for t_ in s.textures[:nmaps])
            print(f"{name:9s} layout={layout} R={R:5d} N={out.total:9d} {t*1e3:9.1f} us  {out.total/t/1e3:9.1f} Mg/s  {alg/t/1e6:8.1f} GB/s")
    ds.free()
