#!/bin/bash
# one GPU call: triage, parity tests, bench A/B, ncu launch list + full captures
set -x
mkdir -p gpurun_out
timeout 300 python tools/debug_attn_i8.py > gpurun_out/debug_attn.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q -k "attention" > gpurun_out/pytest_attn.txt 2>&1; rc=$?
tail -5 gpurun_out/pytest_attn.txt
if [ $rc -ne 0 ]; then echo ATTN_TESTS_FAILED; tail -40 gpurun_out/debug_attn.txt; exit 0; fi
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest.txt 2>&1
tail -5 gpurun_out/pytest.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/bench_i8.json 2> gpurun_out/bench_i8.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager --attention dmma > gpurun_out/bench_dmma_attn.json 2> gpurun_out/bench_dmma_attn.err
python - <<'PY'
import json
for f in ('gpurun_out/bench_i8.json','gpurun_out/bench_dmma_attn.json'):
    try:
        d=json.load(open(f)); print(f, round(d['value']), {k:round(v,2) for k,v in d['roofline']['stage_ms_per_step'].items()})
    except Exception as e: print(f, 'ERR', e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1p.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_i8_kernel -s 6 -c 2 -o gpurun_out/prof_attn_i8_r1p -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/ncu_attn_i8.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ozaki_gemm_kernel -s 30 -c 3 -o gpurun_out/prof_oz_r1p -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-eager > gpurun_out/ncu_oz.log 2>&1
ls -la gpurun_out | tail -12
