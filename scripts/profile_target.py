"""Small target for ncu: a few conversions of the bench workload (helmet stand-in, density 512)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesh2splat_b200 import synth, _abi
from mesh2splat_b200.api import Context

layout = {"ref96": 0, "packed56": 1}[sys.argv[1] if len(sys.argv) > 1 else "packed56"]
R = int(sys.argv[2]) if len(sys.argv) > 2 else 512
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
ctx = Context(0)
ds = ctx.upload(synth.helmet_standin(2048))
out = None
for i in range(n):
    out = ctx.convert(ds, R, layout, flags=_abi.FLAG_UNCAPPED, capacity=6 * R * R, out=out.data if out else None)
print(out.total, out.device_ms)
