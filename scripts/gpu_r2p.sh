#!/bin/bash
mkdir -p gpurun_out
M2S_LIB=$PWD/mesh2splat_b200/variants/trace.so timeout 300 python scripts/trace_raster.py packed56 512 helmet > gpurun_out/r2p_trace.txt 2>&1; cat gpurun_out/r2p_trace.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2p_pytest.log 2>&1; tail -3 gpurun_out/r2p_pytest.log
for lib in "" mesh2splat_b200/variants/nopipe.so mesh2splat_b200/variants/d0.so; do
  M2S_LIB=${lib:+$PWD/$lib} timeout 300 python scripts/quick_ab.py helmet512 helmet512_ref96 dh512 dh2048 sphere1m sponza1024 quad64 2>&1 | grep -E "median|rror" | tee -a gpurun_out/r2p_ab.txt
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"raster_kernel" -s 2 -c 1 -o gpurun_out/r2p_helmet512_p56_raster -f python scripts/profile_target.py packed56 512 4 > gpurun_out/r2p_ncu.log 2>&1
M2S_LIB=$PWD/mesh2splat_b200/variants/d0.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:"raster_kernel|fragment_kernel" -s 4 -c 2 -o gpurun_out/r2p_helmet512_p56_d0 -f python scripts/profile_target.py packed56 512 4 >> gpurun_out/r2p_ncu.log 2>&1
tail -2 gpurun_out/r2p_ncu.log
