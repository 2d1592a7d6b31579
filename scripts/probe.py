import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mesh2splat_b200 import synth, _abi
from mesh2splat_b200.api import Context
ctx = Context(0)
which, R = sys.argv[1], int(sys.argv[2])
scene = {"dh": synth.damaged_helmet_standin, "helmet": synth.helmet_standin}[which](2048)
ds = ctx.upload(scene)
out = None
for i in range(4):
    out = ctx.convert(ds, R, 1, flags=1, capacity=6 * R * R, out=out.data if out else None)
print(which, R, out.total, out.device_ms)
