#!/bin/bash
mkdir -p gpurun_out
M2S_LIB=$PWD/mesh2splat_b200/variants/trace.so timeout 300 python scripts/trace_raster.py packed56 512 helmet > gpurun_out/r2n_trace.txt 2>&1; cat gpurun_out/r2n_trace.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; tail -3 gpurun_out/r2n_pytest.log
timeout 300 python scripts/quick_ab.py helmet512 helmet512_ref96 dh512 dh1024 dh2048 sphere1m sponza1024 quad64 2>&1 | grep -E "median|rror" | tee gpurun_out/r2n_ab.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"raster_kernel" -s 2 -c 1 -o gpurun_out/r2n_helmet512_p56_raster -f python scripts/profile_target.py packed56 512 4 > gpurun_out/r2n_ncu.log 2>&1
tail -2 gpurun_out/r2n_ncu.log
