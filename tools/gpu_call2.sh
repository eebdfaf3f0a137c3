#!/bin/bash
# GPU call 2: full parity run (no -x), attention epilogue variants, default bench with both reference arms, ncu of the attention kernel
set -x
mkdir -p gpurun_out
rm -f gpurun_out/precision_sweep.json
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c2_pytest.txt 2>&1; tail -15 gpurun_out/c2_pytest.txt
for cvt in 0 3; do for cw in 8 16; do
  MDGAT_ATTN_CVT=$cvt MDGAT_ATTN_CW=$cw timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/c2_bench_cvt${cvt}_cw${cw}.json 2> gpurun_out/c2_bench_cvt${cvt}_cw${cw}.err
done; done
timeout 900 python bench.py > gpurun_out/c2_bench.json 2> gpurun_out/c2_bench.err; tail -3 gpurun_out/c2_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c2_bench_ref.json 2> gpurun_out/c2_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-eager --no-latency > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_i8_kernel -s 6 -c 1 -o gpurun_out/c2_prof_attn_i8 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-latency > gpurun_out/c2_ncu_attn.log 2>&1
ls -la gpurun_out | tail -12
