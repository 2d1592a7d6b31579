#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/rn_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/rn_ncu_bench.log 2>&1
for k in raster_kernel fragment_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/rn_helmet512_p56_$k -f python scripts/profile_target.py packed56 512 4 > gpurun_out/rn_ncu_$k.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fragment_kernel -s 2 -c 1 -o gpurun_out/rn_helmet2048_p56_fragment_kernel -f python scripts/profile_target.py packed56 2048 4 > gpurun_out/rn_ncu_2048.log 2>&1
ls -la gpurun_out | grep rn_
