#!/bin/bash
mkdir -p gpurun_out
for la in 1 2 3; do for c in 4 6 8; do echo "== lookahead $la chunks $c: $(M2S_HOST_LOOKAHEAD=$la M2S_HOST_CHUNKS=$c timeout 300 python scripts/e2e_probe.py 2>&1 | tail -1)"; done; done | tee gpurun_out/r2v_e2e.txt
echo "== trace, lookahead 2, 8 chunks"; M2S_HOST_LOOKAHEAD=2 M2S_HOST_CHUNKS=8 M2S_HOST_TRACE=1 timeout 300 python scripts/e2e_probe.py 2>&1 | tail -12 | tee -a gpurun_out/r2v_e2e.txt
