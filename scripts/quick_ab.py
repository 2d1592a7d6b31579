"""Very short A/B probe: median device time of the bench conversion (helmet stand-in, R = 512) for the library
selected by M2S_LIB, L2 flushed between launches.  ~6 s per run."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mesh2splat_b200 import synth, _abi
from mesh2splat_b200.api import Context
ctx = Context(0)
ds = ctx.upload(synth.helmet_standin(2048))
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tag = os.path.basename(os.environ.get("M2S_LIB", "") or "default")
for layout, name in ((_abi.LAYOUT_PACKED56, "packed56"), (_abi.LAYOUT_REF96, "ref96")):
    out, ts = None, []
    for i in range(24):
        flush.zero_(); torch.cuda.synchronize()
        out = ctx.convert(ds, 512, layout, flags=_abi.FLAG_UNCAPPED, capacity=6 * 512 * 512, out=out.data if out else None)
        ts.append(out.device_ms)
    print(f"{tag:18s} {name:9s} median {np.median(ts[4:]) * 1e3:7.2f} us  min {min(ts[4:]) * 1e3:7.2f} us  N={out.total}", flush=True)
