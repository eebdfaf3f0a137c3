#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c5_pytest.txt 2>&1; tail -5 gpurun_out/c5_pytest.txt
timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/c5_bench.json 2> gpurun_out/c5_bench.err
MDGAT_SLICE_TILED=0 timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/c5_bench_oldslice.json 2> gpurun_out/c5_bench_oldslice.err
timeout 300 python bench.py --attention tcgen05_i8_all --no-cpu-baseline --no-eager --no-latency > gpurun_out/c5_bench_i8all.json 2> gpurun_out/c5_bench_i8all.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c5_launches.csv python bench.py --attention tcgen05_i8_all --steps 2 --warmup 1 --no-cpu-baseline --no-eager --no-latency > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"topk_threshold_kernel|slice_rows_tiled_kernel|slice_qk_sides_kernel|slice_v_sides_kernel|attn_i8_kernel" -s 40 -c 12 -o gpurun_out/c5_prof_misc -f python bench.py --attention tcgen05_i8_all --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-latency > gpurun_out/c5_ncu_misc.log 2>&1
ls -la gpurun_out | tail -8
