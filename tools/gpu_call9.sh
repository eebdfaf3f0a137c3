#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sinkhorn" > gpurun_out/i1_pytest.txt 2>&1; tail -5 gpurun_out/i1_pytest.txt
MDGAT_SK_CTAS=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sinkhorn" > gpurun_out/i1_pytest2.txt 2>&1; tail -5 gpurun_out/i1_pytest2.txt
timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/i1_bench.json 2> gpurun_out/i1_bench.err; tail -2 gpurun_out/i1_bench.err
MDGAT_SK_CTAS=1 timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/i1_bench_c1.json 2> gpurun_out/i1_bench_c1.err
MDGAT_SK_CTAS=2 timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/i1_bench_c2.json 2> gpurun_out/i1_bench_c2.err
timeout 600 python -m pytest tests/test_gpu_sweep.py -m gpu -q -x > gpurun_out/i1_sweep.txt 2>&1; tail -5 gpurun_out/i1_sweep.txt
