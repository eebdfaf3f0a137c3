#!/bin/bash
# round-2 checkpoint: fused top-k test + slot sweep, full default bench line, launch list, ncu captures for profiles/r2_*
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "topk" > gpurun_out/f1_pytest.txt 2>&1; tail -5 gpurun_out/f1_pytest.txt
MDGAT_TOPK_FUSED=1 MDGAT_TOPK_SLOTS=4 timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/f1_bench_fused4.json 2> gpurun_out/f1_bench_fused4.err
MDGAT_TOPK_FUSED=1 MDGAT_TOPK_SLOTS=8 timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/f1_bench_fused8.json 2> gpurun_out/f1_bench_fused8.err
timeout 900 python bench.py > gpurun_out/f1_bench.json 2> gpurun_out/f1_bench.err; tail -2 gpurun_out/f1_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/f1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-eager --no-latency > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_i8_kernel|ozaki_gemm_kernel|slice_rows|slice_qk|slice_v|topk_softmax_pv|attn_full_kernel|sinkhorn_fused" -s 60 -c 24 -o gpurun_out/f1_prof_all -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-latency > gpurun_out/f1_ncu.log 2>&1
ls -la gpurun_out | tail -8
