#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sinkhorn or golden" > gpurun_out/l1_pytest.txt 2>&1; tail -3 gpurun_out/l1_pytest.txt
timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/l1_bench.json 2> gpurun_out/l1_bench.err; tail -2 gpurun_out/l1_bench.err
timeout 300 python bench.py --n 2048 --steps 3 --warmup 3 --no-cpu-baseline --no-eager --no-latency > gpurun_out/l1_bench_cfg4.json 2> gpurun_out/l1_bench_cfg4.err
