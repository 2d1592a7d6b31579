#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "host or upload_range or file or pipel" > gpurun_out/r2w_pytest.log 2>&1; tail -3 gpurun_out/r2w_pytest.log
for la in 2 3 4 99; do for c in 4 8; do echo "== lookahead $la chunks $c: $(M2S_HOST_LOOKAHEAD=$la M2S_HOST_CHUNKS=$c timeout 300 python scripts/e2e_probe.py 2>&1 | tail -1)"; done; done | tee gpurun_out/r2w_e2e.txt
echo "== trace, defaults"; M2S_HOST_TRACE=1 timeout 300 python scripts/e2e_probe.py 2>&1 | tail -12 | tee -a gpurun_out/r2w_e2e.txt
