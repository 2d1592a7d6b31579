#!/bin/bash
# final single-GPU validation of the round: smoke, full parity suite, A/B table, bench lines (ours, ref96, reference arm)
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/rf_smoke.log 2>&1; tail -1 gpurun_out/rf_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/rf_pytest.log 2>&1; tail -3 gpurun_out/rf_pytest.log
timeout 300 python scripts/quick_ab.py helmet512 helmet512_ref96 dh512 dh1024 dh2048 sphere1m sponza1024 quad64 2>&1 | grep -E "median|rror" | tee gpurun_out/rf_ab.txt
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/rf_bench_reference.json 2>&1
timeout 600 python bench.py > gpurun_out/rf_bench_p56.json 2> gpurun_out/rf_bench_p56.err; tail -c 600 gpurun_out/rf_bench_p56.json; tail -3 gpurun_out/rf_bench_p56.err
timeout 600 python bench.py --steps 30 --warmup 5 --layout ref96 > gpurun_out/rf_bench_ref96.json 2> gpurun_out/rf_bench_ref96.err
for w in sphere_1m sponza_standin damaged_helmet_standin; do timeout 600 python bench.py --steps 20 --warmup 3 --workload $w > gpurun_out/rf_bench_$w.json 2> gpurun_out/rf_bench_$w.err; done
echo "e2e probe: $(timeout 300 python scripts/e2e_probe.py 2>&1 | tail -1)"
