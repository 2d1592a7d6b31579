#!/bin/bash
# 2-GPU check: fused-gather parity tests + the bench at N = 2 (both workloads)
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2r_pytest_multi.log 2>&1; tail -4 gpurun_out/r2r_pytest_multi.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r2r_bench_n2.json 2> gpurun_out/r2r_bench_n2.err; tail -c 1800 gpurun_out/r2r_bench_n2.json; tail -3 gpurun_out/r2r_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --workload sphere_1m > gpurun_out/r2r_bench_config4_n2.json 2> gpurun_out/r2r_bench_config4_n2.err; tail -c 1200 gpurun_out/r2r_bench_config4_n2.json; tail -3 gpurun_out/r2r_bench_config4_n2.err
