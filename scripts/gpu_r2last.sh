#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/rl_smoke.log 2>&1; tail -1 gpurun_out/rl_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/rl_pytest.log 2>&1; tail -3 gpurun_out/rl_pytest.log
timeout 600 python bench.py > gpurun_out/rl_bench_p56.json 2> gpurun_out/rl_bench_p56.err; python -c "
import json; d=json.load(open('gpurun_out/rl_bench_p56.json')); print('value',round(d['value']),'us',round(d['ms_per_step']*1e3,2),'frac',round(d['roofline']['frac'],4),'e2e ms',round(d['e2e']['ms_per_step'],3),round(d['e2e']['value']),'cpu',round(d['cpu_baseline']['value']))"; tail -2 gpurun_out/rl_bench_p56.err
