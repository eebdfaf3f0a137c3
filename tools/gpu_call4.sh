#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c4_pytest.txt 2>&1; tail -5 gpurun_out/c4_pytest.txt
for cvt in 0 3 4; do
  MDGAT_ATTN_CVT=$cvt timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/c4_bench_cvt${cvt}.json 2> gpurun_out/c4_bench_cvt${cvt}.err
done
MDGAT_ATTN_CVT=4 MDGAT_ATTN_CW=16 timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/c4_bench_cvt4_cw16.json 2> gpurun_out/c4_bench_cvt4_cw16.err
MDGAT_ATTN_CVT=4 timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_i8_kernel -s 6 -c 1 -o gpurun_out/c4_prof_attn_i8 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-latency > gpurun_out/c4_ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ozaki_gemm_kernel|slice_rows_kernel|slice_sides_kernel|topk_softmax_pv" -s 12 -c 8 -o gpurun_out/c4_prof_gemm -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-latency > gpurun_out/c4_ncu_gemm.log 2>&1
ls -la gpurun_out | tail -8
