#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "prepass" > gpurun_out/rpp_pytest.log 2>&1; tail -6 gpurun_out/rpp_pytest.log
timeout 300 python scripts/prepass_bench.py 2>&1 | tail -6 | tee gpurun_out/rpp_bench.txt
