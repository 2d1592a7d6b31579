#!/bin/bash
# round 2, GPU call A: GL probe, smoke, parity suite, bench, config sweep (each under its own timeout)
mkdir -p gpurun_out
bash scripts/gl_probe.sh gpurun_out/r02_gl_probe.log
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/r2a_smi.txt 2>&1
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2a_smoke.log
if ! grep -q "smoke ok" gpurun_out/r2a_smoke.log; then tail -30 gpurun_out/r2a_smoke.log; exit 1; fi
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2a_pytest.log
tail -15 gpurun_out/r2a_pytest.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2a_bench_p56.json 2> gpurun_out/r2a_bench_p56.err; tail -c 1500 gpurun_out/r2a_bench_p56.json
timeout 600 python bench.py --steps 30 --warmup 5 --layout ref96 > gpurun_out/r2a_bench_ref96.json 2> gpurun_out/r2a_bench_ref96.err
timeout 900 python scripts/config_sweep.py > gpurun_out/r2a_sweep.md 2>&1; cat gpurun_out/r2a_sweep.md | tail -14
