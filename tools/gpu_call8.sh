#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sinkhorn or nonfinite or golden or match" > gpurun_out/h1_pytest.txt 2>&1; tail -5 gpurun_out/h1_pytest.txt
timeout 600 python -m pytest tests/test_gpu_sweep.py -m gpu -q -x > gpurun_out/h1_sweep.txt 2>&1; tail -5 gpurun_out/h1_sweep.txt
timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/h1_bench.json 2> gpurun_out/h1_bench.err; tail -2 gpurun_out/h1_bench.err
timeout 300 python bench.py --precision exact --no-cpu-baseline --no-eager --no-latency > gpurun_out/h1_bench_exact.json 2> gpurun_out/h1_bench_exact.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"sinkhorn_fused" -c 1 -o gpurun_out/h1_prof_sk -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-latency > gpurun_out/h1_ncu.log 2>&1
