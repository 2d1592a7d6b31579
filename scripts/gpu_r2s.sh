#!/bin/bash
mkdir -p gpurun_out
M2S_LIB=$PWD/mesh2splat_b200/variants/trace.so timeout 300 python scripts/trace_raster.py packed56 512 helmet > gpurun_out/r2s_trace.txt 2>&1; cat gpurun_out/r2s_trace.txt
cat > /tmp/shares.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from mesh2splat_b200 import synth, _abi
from mesh2splat_b200.api import Context
ctx = Context(0); scene = synth.helmet_standin(2048); ds = ctx.upload(scene)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
p = _abi.make_params(512, _abi.LAYOUT_PACKED56, 0.65, 0, _abi.FLAG_UNCAPPED, 0, 0, 0, 0)
cap = 6 * 512 * 512
out = torch.empty(cap * 56, dtype=torch.uint8, device="cuda")
sh = []
for i in range(12):
    flush.zero_(); torch.cuda.synchronize()
    sh.append(ctx.convert_timed(ds, p, out, cap))
tot = []
o = None
for i in range(16):
    flush.zero_(); torch.cuda.synchronize()
    o = ctx.convert(ds, 512, _abi.LAYOUT_PACKED56, flags=_abi.FLAG_UNCAPPED, capacity=cap, out=o.data if o else None)
    tot.append(o.device_ms)
print(os.path.basename(os.environ.get("M2S_LIB", "default")), "raster %.2f us fragment %.2f us | step median %.2f min %.2f" % (np.median([s[0] for s in sh[3:]]) * 1e3, np.median([s[1] for s in sh[3:]]) * 1e3, np.median(tot[4:]) * 1e3, min(tot[4:]) * 1e3))
PY
for lib in "" mesh2splat_b200/variants/noplanes.so mesh2splat_b200/variants/nodirect.so; do
  M2S_LIB=${lib:+$PWD/$lib} timeout 200 python /tmp/shares.py 2>&1 | tail -1 | tee -a gpurun_out/r2s_shares.txt
done
