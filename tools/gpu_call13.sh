#!/bin/bash
set -x
mkdir -p gpurun_out
MDGAT_OZ_EVEN_ITEMS=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden and n512" > gpurun_out/n1_pytest.txt 2>&1; tail -3 gpurun_out/n1_pytest.txt
MDGAT_OZ_EVEN_ITEMS=1 timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/n1_bench_even.json 2> gpurun_out/n1_bench_even.err
timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/n1_bench.json 2> gpurun_out/n1_bench.err
