#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "host or upload_range or file or pipel" > gpurun_out/r2u_pytest.log 2>&1; tail -3 gpurun_out/r2u_pytest.log
for c in 0 2 4 8; do if [ $c = 0 ]; then export -n M2S_HOST_CHUNKS; unset M2S_HOST_CHUNKS; else export M2S_HOST_CHUNKS=$c; fi; echo "== chunks ${c} (0 = default): $(timeout 300 python scripts/e2e_probe.py 2>&1 | tail -1)"; done | tee gpurun_out/r2u_e2e.txt
unset M2S_HOST_CHUNKS
echo "== trace, default"; M2S_HOST_TRACE=1 timeout 300 python scripts/e2e_probe.py 2>&1 | tail -12 | tee -a gpurun_out/r2u_e2e.txt
