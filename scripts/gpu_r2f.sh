#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2f_smoke.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/r2f_smoke.log
if ! grep -q "smoke ok" gpurun_out/r2f_smoke.log; then tail -30 gpurun_out/r2f_smoke.log; exit 1; fi
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/r2f_pytest.log
tail -8 gpurun_out/r2f_pytest.log
timeout 200 python scripts/pcie_probe.py > gpurun_out/r2f_pcie.log 2>&1; cat gpurun_out/r2f_pcie.log
: > gpurun_out/r2f_ab.txt
for lib in "" build_variants/bufs1.so build_variants/early.so build_variants/early_bufs1.so build_variants/fw8.so build_variants/fw8_early.so; do
  M2S_LIB=$lib timeout 300 python scripts/quick_ab.py helmet512 helmet512_ref96 dh2048 sphere1m sponza1024 2>&1 | grep median >> gpurun_out/r2f_ab.txt
done
cat gpurun_out/r2f_ab.txt
