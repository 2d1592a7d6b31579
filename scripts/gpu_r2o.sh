#!/bin/bash
mkdir -p gpurun_out
M2S_LIB=$PWD/mesh2splat_b200/variants/trace.so timeout 300 python scripts/trace_raster.py packed56 512 helmet > gpurun_out/r2o_trace.txt 2>&1; cat gpurun_out/r2o_trace.txt
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2o_pytest.log 2>&1; tail -3 gpurun_out/r2o_pytest.log
for lib in "" mesh2splat_b200/variants/t1024.so mesh2splat_b200/variants/d0.so; do
  M2S_LIB=${lib:+$PWD/$lib} timeout 300 python scripts/quick_ab.py helmet512 helmet512_ref96 dh512 dh1024 dh2048 sphere1m sponza1024 quad64 2>&1 | grep -E "median|rror" | tee -a gpurun_out/r2o_ab.txt
done
