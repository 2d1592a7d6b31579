#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2i_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2i_smoke.log
if ! grep -q "smoke ok" gpurun_out/r2i_smoke.log; then tail -30 gpurun_out/r2i_smoke.log; exit 1; fi
timeout 1200 python -m pytest tests -m gpu -x -q -k "host or upload_range or file" > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2i_pytest.log
tail -5 gpurun_out/r2i_pytest.log
for c in 2 4 6 8; do echo "== chunks $c"; M2S_HOST_CHUNKS=$c M2S_HOST_TRACE=1 timeout 300 python scripts/e2e_probe.py 2>&1 | tail -14; done > gpurun_out/r2i_e2e_chunks.log 2>&1
grep -E "chunks|e2e ms" gpurun_out/r2i_e2e_chunks.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2i_bench_p56.json 2> gpurun_out/r2i_bench_p56.err; python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_p56.json')); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],d['e2e']['resident_scene']['ms_per_step'],'frac',d['roofline']['frac'])"; tail -3 gpurun_out/r2i_bench_p56.err
timeout 600 python bench.py --steps 30 --warmup 5 --layout ref96 > gpurun_out/r2i_bench_ref96.json 2> gpurun_out/r2i_bench_ref96.err; python -c "
import json; d=json.load(open('gpurun_out/r2i_bench_ref96.json')); print('ref96 value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],'frac',d['roofline']['frac'])"
