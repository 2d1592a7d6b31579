#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2d_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2d_smoke.log
if ! grep -q "smoke ok" gpurun_out/r2d_smoke.log; then tail -30 gpurun_out/r2d_smoke.log; exit 1; fi
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2d_pytest.log
tail -12 gpurun_out/r2d_pytest.log
timeout 900 python scripts/config_sweep.py > gpurun_out/r2d_sweep.md 2>&1; tail -13 gpurun_out/r2d_sweep.md
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"raster_kernel|fragment_kernel" --csv --log-file gpurun_out/r2d_launches.csv python scripts/launch_configs.py > gpurun_out/r2d_launch_configs.log 2>&1
for c in 1 2 4 8; do echo "== chunks $c"; M2S_HOST_CHUNKS=$c M2S_HOST_TRACE=1 timeout 300 python scripts/e2e_probe.py 2>&1 | tail -14; done > gpurun_out/r2d_e2e_chunks.log 2>&1
grep -E "chunks|e2e ms" gpurun_out/r2d_e2e_chunks.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2d_bench_p56.json 2> gpurun_out/r2d_bench_p56.err; python -c "
import json; d=json.load(open('gpurun_out/r2d_bench_p56.json')); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['ms_per_step'],d['e2e']['resident_scene']['ms_per_step'],'frac',d['roofline']['frac'],d['roofline']['launch_shares'])"; tail -3 gpurun_out/r2d_bench_p56.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fragment_kernel" -s 2 -c 1 -o gpurun_out/r2d_fragment_p56 -f python scripts/profile_target.py packed56 512 4 > gpurun_out/r2d_ncu_frag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fragment_kernel" -s 2 -c 1 -o gpurun_out/r2d_fragment_p56_2048 -f python scripts/profile_target.py packed56 2048 4 > gpurun_out/r2d_ncu_frag2048.log 2>&1
ls gpurun_out/r2d*.ncu-rep
