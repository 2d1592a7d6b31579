"""Short A/B probe: median device time of a few BASELINE conversions for the library selected by M2S_LIB (L2 flushed
between launches).  usage: quick_ab.py [config ...]   configs: helmet512 helmet512_ref96 dh2048 dh1024 sphere1m sponza1024 quad64"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mesh2splat_b200 import synth, _abi
from mesh2splat_b200.api import Context
P, Rf, U = _abi.LAYOUT_PACKED56, _abi.LAYOUT_REF96, _abi.FLAG_UNCAPPED
CONFIGS = {"helmet512": ("helmet", 512, P), "helmet512_ref96": ("helmet", 512, Rf), "dh2048": ("dh", 2048, P), "dh1024": ("dh", 1024, P), "dh512": ("dh", 512, P),
           "sphere1m": ("sphere1m", 256, P), "sponza1024": ("sponza", 1024, P), "quad64": ("quad", 64, Rf)}
SCENES = {"helmet": lambda: synth.helmet_standin(2048), "dh": lambda: synth.damaged_helmet_standin(2048),
          "sphere1m": lambda: synth.sphere_1m(2048), "sponza": lambda: synth.sponza_standin(1024), "quad": synth.unit_quad}
names = sys.argv[1:] or ["helmet512", "helmet512_ref96", "dh2048", "sphere1m", "sponza1024"]
ctx = Context(0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tag = os.path.basename(os.environ.get("M2S_LIB", "") or "default")
cache = {}
for name in names:
    sc, R, layout = CONFIGS[name]
    if sc not in cache:
        scene = SCENES[sc]()
        cache = {sc: (scene, ctx.upload(scene))}  # one scene resident at a time
    scene, ds = cache[sc]
    cap = min(6 * R * R * max(1, len(scene.primitives)), 60_000_000)
    out, ts = None, []
    for i in range(20):
        flush.zero_(); torch.cuda.synchronize()
        out = ctx.convert(ds, R, layout, flags=U, capacity=cap, out=out.data if out else None)
        ts.append(out.device_ms)
    print(f"{tag:14s} {name:16s} median {np.median(ts[4:]) * 1e3:8.2f} us  min {min(ts[4:]) * 1e3:8.2f} us  N={out.total}", flush=True)
