#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"raster_kernel|fragment_kernel" --csv --log-file gpurun_out/r2b_launches.csv python scripts/launch_configs.py > gpurun_out/r2b_launch_configs.log 2>&1
tail -8 gpurun_out/r2b_launch_configs.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"fragment_kernel" -s 2 -c 1 -o gpurun_out/r2b_fragment_p56 -f python scripts/profile_target.py packed56 512 4 > gpurun_out/r2b_ncu_frag.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"raster_kernel" -s 2 -c 1 -o gpurun_out/r2b_raster_p56 -f python scripts/profile_target.py packed56 512 4 > gpurun_out/r2b_ncu_rast.log 2>&1
ls -la gpurun_out/*.ncu-rep
