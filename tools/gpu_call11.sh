#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python bench.py --n 2048 --steps 3 --warmup 3 --no-cpu-baseline --no-eager --no-latency > gpurun_out/k1_bench_cfg4.json 2> gpurun_out/k1_bench_cfg4.err; tail -2 gpurun_out/k1_bench_cfg4.err
timeout 600 python bench.py --n 2048 --steps 3 --warmup 3 --precision exact --no-cpu-baseline --no-eager --no-latency > gpurun_out/k1_bench_cfg4_exact.json 2> gpurun_out/k1_bench_cfg4_exact.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/k1_launches_cfg4.csv python bench.py --n 2048 --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-latency > /dev/null 2>&1
timeout 600 python bench.py --no-cpu-baseline --no-eager > gpurun_out/k1_bench_lat.json 2> gpurun_out/k1_bench_lat.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/k1_launches_b1.csv python bench.py --batch 1 --n 256 --sinkhorn 20 --steps 3 --warmup 3 --no-cpu-baseline --no-eager --no-latency > gpurun_out/k1_b1.json 2>gpurun_out/k1_b1.err
