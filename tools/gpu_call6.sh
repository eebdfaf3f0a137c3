#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c6_pytest.txt 2>&1; tail -5 gpurun_out/c6_pytest.txt
timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err
timeout 300 python bench.py --attention tcgen05_i8_all --no-cpu-baseline --no-eager --no-latency > gpurun_out/c6_bench_i8all.json 2> gpurun_out/c6_bench_i8all.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c6_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-eager --no-latency > /dev/null 2>&1
ls -la gpurun_out | tail -5
