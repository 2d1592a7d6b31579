#!/bin/bash
# parity + A/B of the gather/plane fragment stage
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2l_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2l_smoke.log
tail -5 gpurun_out/r2l_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2l_pytest.log
tail -30 gpurun_out/r2l_pytest.log
for lib in "" $(ls mesh2splat_b200/variants/*.so 2>/dev/null); do
  M2S_LIB=${lib:+$PWD/$lib} timeout 300 python scripts/quick_ab.py helmet512 helmet512_ref96 dh512 dh1024 dh2048 sphere1m sponza1024 quad64 2>&1 | grep -E "median|rror" | tee -a gpurun_out/r2l_ab.txt
done
