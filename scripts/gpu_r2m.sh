#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fragment_kernel" -s 2 -c 1 -o gpurun_out/r2m_helmet512_p56 -f python scripts/profile_target.py packed56 512 4 > gpurun_out/r2m_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fragment_kernel" -s 2 -c 1 -o gpurun_out/r2m_helmet512_ref96 -f python scripts/profile_target.py ref96 512 4 >> gpurun_out/r2m_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fragment_kernel" -s 2 -c 1 -o gpurun_out/r2m_helmet2048_p56 -f python scripts/profile_target.py packed56 2048 4 >> gpurun_out/r2m_ncu.log 2>&1
tail -3 gpurun_out/r2m_ncu.log
