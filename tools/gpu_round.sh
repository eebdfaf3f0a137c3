#!/bin/bash
# GPU calls at the end of a round. gpurun copies back at most 64 MiB, and one ncu --set full record with source is 5-12 MB:
#   bash tools/gpu_round.sh <prefix> tests     parity tests, smoke (+ its first 1000 launches), both bench arms, cfg4 line, launch list
#   bash tools/gpu_round.sh <prefix> ncu       ncu --set full captures of one GNN layer's kernels and of the top-k / Sinkhorn kernels;
#                                              every report is also exported as raw CSV (what tools/summarize_profiles.py reads)
# then `python tools/summarize_profiles.py <tag> <prefix>`
P=${1:-final}
MODE=${2:-tests}
set -x
mkdir -p gpurun_out
if [ "$MODE" = tests ]; then
rm -f gpurun_out/precision_sweep.json
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/${P}_pytest.txt 2>&1; tail -3 gpurun_out/${P}_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${P}_smoke.txt 2>&1; tail -2 gpurun_out/${P}_smoke.txt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1000 --csv --log-file gpurun_out/${P}_smoke_launches.csv python -c "import __graft_entry__ as g; g.smoke()" > /dev/null 2>&1
timeout 900 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; tail -2 gpurun_out/${P}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${P}_bench_ref.json 2> gpurun_out/${P}_bench_ref.err
timeout 600 python bench.py --n 2048 --steps 3 --warmup 3 --no-cpu-baseline --no-eager --no-latency > gpurun_out/${P}_bench_cfg4.json 2> gpurun_out/${P}_bench_cfg4.err
timeout 300 python bench.py --no-cuda-graph --no-cpu-baseline --no-eager --no-latency > gpurun_out/${P}_bench_nograph.json 2> gpurun_out/${P}_bench_nograph.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${P}_launches.csv python bench.py --steps 2 --warmup 1 --no-cuda-graph --no-cpu-baseline --no-eager --no-latency > /dev/null 2>&1
else
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_i8_kernel" -s 6 -c 2 -o gpurun_out/${P}_prof_attn -f python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline --no-eager --no-latency > gpurun_out/${P}_ncu_attn.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"ozaki_gemm_kernel" -s 30 -c 3 -o gpurun_out/${P}_prof_gemm -f python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline --no-eager --no-latency > gpurun_out/${P}_ncu_gemm.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"topk_softmax_pv|attn_full_kernel|sinkhorn_fused" -c 5 -o gpurun_out/${P}_prof_topk -f python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline --no-eager --no-latency > gpurun_out/${P}_ncu_topk.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:"slice_rows|slice_qk|slice_v|gemm_f64" -s 30 -c 10 -o gpurun_out/${P}_prof_misc -f python bench.py --steps 1 --warmup 1 --no-cuda-graph --no-cpu-baseline --no-eager --no-latency > gpurun_out/${P}_ncu_misc.log 2>&1
for r in gpurun_out/${P}_prof_*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}_raw.csv 2>/dev/null; done
ls -la gpurun_out
# keep the copy-back under 64 MiB: drop the largest report(s) if needed (their raw CSV stays)
while [ $(du -sm gpurun_out | cut -f1) -gt 55 ]; do big=$(ls -S gpurun_out/*.ncu-rep | head -1); [ -z "$big" ] && break; rm -f "$big"; done
fi
ls -la gpurun_out | tail -14
