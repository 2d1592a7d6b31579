#!/bin/bash
# 2-GPU call: fused-gather parity, bench N=2 (both strategies measured, gather parity, multi-GPU e2e), N=1 e2e probe
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q > gpurun_out/r2g_pytest_multi.log 2>&1; tail -5 gpurun_out/r2g_pytest_multi.log
N=$(nvidia-smi -L | wc -l)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r2g_bench_n$N.json 2> gpurun_out/r2g_bench_n$N.err; tail -c 3000 gpurun_out/r2g_bench_n$N.json; tail -5 gpurun_out/r2g_bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 --workload sphere_1m > gpurun_out/r2g_bench_sphere_n$N.json 2> gpurun_out/r2g_bench_sphere_n$N.err; tail -c 1500 gpurun_out/r2g_bench_sphere_n$N.json; tail -3 gpurun_out/r2g_bench_sphere_n$N.err
for c in 4 6; do echo "== chunks $c"; M2S_HOST_CHUNKS=$c M2S_HOST_TRACE=1 timeout 300 python scripts/e2e_probe.py 2>&1 | tail -12; done > gpurun_out/r2g_e2e_chunks.log 2>&1
grep -E "chunks|e2e ms" gpurun_out/r2g_e2e_chunks.log
