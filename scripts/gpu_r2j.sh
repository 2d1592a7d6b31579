#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2j_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2j_smoke.log
if ! grep -q "smoke ok" gpurun_out/r2j_smoke.log; then tail -30 gpurun_out/r2j_smoke.log; exit 1; fi
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2j_pytest.log
tail -8 gpurun_out/r2j_pytest.log
timeout 300 python scripts/quick_ab.py helmet512 helmet512_ref96 dh1024 dh2048 sphere1m sponza1024 quad64 2>&1 | grep median | tee gpurun_out/r2j_ab.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"raster_kernel|fragment_kernel" --csv --log-file gpurun_out/r2j_launches.csv python scripts/launch_configs.py > gpurun_out/r2j_launch_configs.log 2>&1
