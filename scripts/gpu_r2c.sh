#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2c_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2c_smoke.log
if ! grep -q "smoke ok" gpurun_out/r2c_smoke.log; then tail -30 gpurun_out/r2c_smoke.log; exit 1; fi
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2c_pytest.log
tail -25 gpurun_out/r2c_pytest.log
timeout 900 python scripts/config_sweep.py > gpurun_out/r2c_sweep.md 2>&1; tail -13 gpurun_out/r2c_sweep.md
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"raster_kernel|fragment_kernel" --csv --log-file gpurun_out/r2c_launches.csv python scripts/launch_configs.py > gpurun_out/r2c_launch_configs.log 2>&1
tail -8 gpurun_out/r2c_launch_configs.log
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2c_bench_p56.json 2> gpurun_out/r2c_bench_p56.err; tail -c 2500 gpurun_out/r2c_bench_p56.json; tail -5 gpurun_out/r2c_bench_p56.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fragment_kernel" -s 2 -c 1 -o gpurun_out/r2c_fragment_p56 -f python scripts/profile_target.py packed56 512 4 > gpurun_out/r2c_ncu_frag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"raster_kernel" -s 2 -c 1 -o gpurun_out/r2c_raster_p56 -f python scripts/profile_target.py packed56 512 4 > gpurun_out/r2c_ncu_rast.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fragment_kernel" -s 2 -c 1 -o gpurun_out/r2c_fragment_p56_2048 -f python scripts/profile_target.py packed56 2048 4 > gpurun_out/r2c_ncu_frag2048.log 2>&1
ls -la gpurun_out/r2c*.ncu-rep
