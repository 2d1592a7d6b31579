#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "host or upload_range or file or pipel" > gpurun_out/r2e_pytest.log 2>&1; tail -3 gpurun_out/r2e_pytest.log
for c in 2 4 8; do echo "== eager, chunks $c: $(M2S_HOST_CHUNKS=$c timeout 300 python scripts/e2e_probe.py 2>&1 | tail -1)"; done | tee gpurun_out/r2e_e2e.txt
for c in 4; do echo "== on demand, chunks $c: $(M2S_HOST_NO_EAGER=1 M2S_HOST_CHUNKS=$c timeout 300 python scripts/e2e_probe.py 2>&1 | tail -1)"; done | tee -a gpurun_out/r2e_e2e.txt
echo "== trace, defaults"; M2S_HOST_TRACE=1 timeout 300 python scripts/e2e_probe.py 2>&1 | tail -12 | tee -a gpurun_out/r2e_e2e.txt
