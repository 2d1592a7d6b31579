#!/bin/bash
# checkpoint: parity suite, A/B table, bench lines, ncu launch list + full captures (one report per kernel)
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2q_smoke.log 2>&1; tail -1 gpurun_out/r2q_smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; tail -3 gpurun_out/r2q_pytest.log
timeout 300 python scripts/quick_ab.py helmet512 helmet512_ref96 dh512 dh1024 dh2048 sphere1m sponza1024 quad64 2>&1 | grep -E "median|rror" | tee gpurun_out/r2q_ab.txt
timeout 600 python bench.py --steps 30 --warmup 5 > gpurun_out/r2q_bench_p56.json 2> gpurun_out/r2q_bench_p56.err; tail -c 900 gpurun_out/r2q_bench_p56.json; tail -3 gpurun_out/r2q_bench_p56.err
timeout 600 python bench.py --steps 30 --warmup 5 --layout ref96 > gpurun_out/r2q_bench_ref96.json 2> gpurun_out/r2q_bench_ref96.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2q_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/r2q_ncu_bench.log 2>&1
for k in raster_kernel fragment_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$k -s 2 -c 1 -o gpurun_out/r2q_helmet512_p56_$k -f python scripts/profile_target.py packed56 512 4 > gpurun_out/r2q_ncu_$k.log 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fragment_kernel -s 2 -c 1 -o gpurun_out/r2q_helmet2048_p56_fragment_kernel -f python scripts/profile_target.py packed56 2048 4 > gpurun_out/r2q_ncu_2048.log 2>&1
ls gpurun_out | grep r2q
