#!/bin/bash
# GPU call 3: parity, attention variants after the MMA-ahead fix, tcgen05 top-k path, both reference arms, ncu
set -x
mkdir -p gpurun_out
rm -f gpurun_out/precision_sweep.json
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/c3_pytest.txt 2>&1; tail -8 gpurun_out/c3_pytest.txt
for cvt in 0 3; do
  MDGAT_ATTN_CVT=$cvt timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/c3_bench_cvt${cvt}.json 2> gpurun_out/c3_bench_cvt${cvt}.err
done
timeout 300 python bench.py --attention tcgen05_i8_all --no-cpu-baseline --no-eager --no-latency > gpurun_out/c3_bench_i8all.json 2> gpurun_out/c3_bench_i8all.err
timeout 900 python bench.py > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err; tail -3 gpurun_out/c3_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c3_bench_ref.json 2> gpurun_out/c3_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c3_launches.csv python bench.py --attention tcgen05_i8_all --steps 2 --warmup 1 --no-cpu-baseline --no-eager --no-latency > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_i8_kernel -s 6 -c 1 -o gpurun_out/c3_prof_attn_i8 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-latency > gpurun_out/c3_ncu_attn.log 2>&1
ls -la gpurun_out | tail -12
