#!/bin/bash
# GPU call 1 of round 2: parity (incl. the 131k-row sweep), bench with the unmodified reference timed beside it, engine variants
set -x
mkdir -p gpurun_out
rm -f gpurun_out/precision_sweep.json
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/c1_gpu.txt 2>&1
nproc >> gpurun_out/c1_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.txt 2>&1; tail -5 gpurun_out/c1_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c1_smoke.txt 2>&1; tail -2 gpurun_out/c1_smoke.txt
timeout 900 python bench.py > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; tail -3 gpurun_out/c1_bench.err
for cvt in 0 2; do
  MDGAT_ATTN_CVT=$cvt timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/c1_bench_cvt$cvt.json 2> gpurun_out/c1_bench_cvt$cvt.err
done
MDGAT_ATTN_CW=16 timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/c1_bench_cw16.json 2> gpurun_out/c1_bench_cw16.err
MDGAT_ATTN_CW=16 MDGAT_ATTN_CVT=0 timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/c1_bench_cw16_cvt0.json 2> gpurun_out/c1_bench_cw16_cvt0.err
timeout 300 python bench.py --precision exact --no-cpu-baseline --no-eager --no-latency > gpurun_out/c1_bench_exact.json 2> gpurun_out/c1_bench_exact.err
timeout 300 python bench.py --attention tcgen05_i8_all --no-cpu-baseline --no-eager --no-latency > gpurun_out/c1_bench_i8all.json 2> gpurun_out/c1_bench_i8all.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c1_bench_ref.json 2> gpurun_out/c1_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c1_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-eager --no-latency > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_i8_kernel -s 6 -c 1 -o gpurun_out/c1_prof_attn_i8 -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager --no-latency > gpurun_out/c1_ncu_attn.log 2>&1
ls -la gpurun_out | tail -20
