#!/bin/bash
set -x
mkdir -p gpurun_out
rm -f gpurun_out/precision_sweep.json
timeout 900 python -m pytest tests/test_gpu_sweep.py -m gpu -q > gpurun_out/j1_sweep.txt 2>&1; tail -5 gpurun_out/j1_sweep.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "golden or stream or reference_eval or shard or permut" > gpurun_out/j1_pytest.txt 2>&1; tail -5 gpurun_out/j1_pytest.txt
timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/j1_bench.json 2> gpurun_out/j1_bench.err; tail -2 gpurun_out/j1_bench.err
