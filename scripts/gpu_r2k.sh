#!/bin/bash
# round-2 checkpoint run: smoke, full GPU parity suite, A/B table, bench lines, ncu launch list + full captures
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2k_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2k_smoke.log
if ! grep -q "smoke ok" gpurun_out/r2k_smoke.log; then tail -30 gpurun_out/r2k_smoke.log; exit 1; fi
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2k_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2k_pytest.log
tail -16 gpurun_out/r2k_pytest.log
timeout 300 python scripts/quick_ab.py helmet512 helmet512_ref96 dh512 dh1024 dh2048 sphere1m sponza1024 quad64 2>&1 | grep -E "median|Error|error" | tee gpurun_out/r2k_ab.txt
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2k_bench_p56.json 2> gpurun_out/r2k_bench_p56.err; tail -c 1500 gpurun_out/r2k_bench_p56.json; tail -3 gpurun_out/r2k_bench_p56.err
timeout 600 python bench.py --steps 30 --warmup 5 --layout ref96 > gpurun_out/r2k_bench_ref96.json 2> gpurun_out/r2k_bench_ref96.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/r2k_bench_reference.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2k_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2k_ncu_bench.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"raster_kernel|fragment_kernel" --csv --log-file gpurun_out/r2k_launches_configs.csv python scripts/launch_configs.py > gpurun_out/r2k_launch_configs.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"raster_kernel|fragment_kernel" -s 4 -c 2 -o gpurun_out/r2k_helmet512_p56 -f python scripts/profile_target.py packed56 512 4 > gpurun_out/r2k_ncu_full.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"raster_kernel|fragment_kernel" -s 4 -c 2 -o gpurun_out/r2k_helmet2048_p56 -f python scripts/profile_target.py packed56 2048 4 >> gpurun_out/r2k_ncu_full.log 2>&1
ls -la gpurun_out | tail -20
