#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2h_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2h_smoke.log
if ! grep -q "smoke ok" gpurun_out/r2h_smoke.log; then tail -30 gpurun_out/r2h_smoke.log; exit 1; fi
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2h_pytest.log
tail -8 gpurun_out/r2h_pytest.log
timeout 300 python scripts/quick_ab.py helmet512 helmet512_ref96 dh1024 dh2048 sphere1m sponza1024 quad64 2>&1 | grep median | tee gpurun_out/r2h_ab.txt
for c in 2 4 6; do echo "== chunks $c"; M2S_HOST_CHUNKS=$c M2S_HOST_TRACE=1 timeout 300 python scripts/e2e_probe.py 2>&1 | tail -12; done > gpurun_out/r2h_e2e_chunks.log 2>&1
grep -E "chunks|e2e ms" gpurun_out/r2h_e2e_chunks.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2h_bench_p56.json 2> gpurun_out/r2h_bench_p56.err; python -c "
import json; d=json.load(open('gpurun_out/r2h_bench_p56.json')); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],d['e2e']['resident_scene']['ms_per_step'],'frac',d['roofline']['frac'],d['roofline']['launch_shares'], 'cpu', d['cpu_baseline']['value'])"; tail -3 gpurun_out/r2h_bench_p56.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2h_bench_reference.json 2>&1; tail -c 600 gpurun_out/r2h_bench_reference.json
