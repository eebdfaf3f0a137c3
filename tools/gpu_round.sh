#!/bin/bash
# one GPU call at the end of a round: parity tests, the two bench arms, ncu launch list + full captures of the top kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.txt 2>&1; tail -3 gpurun_out/pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.txt 2>&1; tail -2 gpurun_out/smoke.txt
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_final_ref.json 2> gpurun_out/bench_final_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-eager > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_i8_kernel -s 6 -c 1 -o gpurun_out/prof_attn_i8_final -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager > gpurun_out/ncu_attn_i8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ozaki_gemm_kernel -s 30 -c 3 -o gpurun_out/prof_oz_final -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager > gpurun_out/ncu_oz.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"topk_softmax_pv_kernel|sinkhorn_fused_kernel|slice_rows_kernel|attn_full_kernel" -s 20 -c 6 -o gpurun_out/prof_misc_final -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-eager > gpurun_out/ncu_misc.log 2>&1
timeout 120 tools/ubench/umma_i8_pattern > gpurun_out/umma_i8_pattern.txt 2>&1
ls -la gpurun_out | tail -15
