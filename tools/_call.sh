set -x
mkdir -p gpurun_out
P=s2
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${P}_pytest.txt 2>&1; tail -3 gpurun_out/${P}_pytest.txt
cp gpurun_out/precision_sweep.json gpurun_out/${P}_precision_sweep.json 2>/dev/null
MDGAT_SK_THREADS=1024 timeout 900 python -m pytest tests -m gpu -q -x -k "sinkhorn or golden or sweep or error_table" > gpurun_out/${P}_pytest_sk1024.txt 2>&1; tail -3 gpurun_out/${P}_pytest_sk1024.txt
Q="--no-cpu-baseline --no-eager --no-latency"
timeout 300 python bench.py $Q > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err
MDGAT_SLICE_ONEFMA=0 timeout 300 python bench.py $Q > gpurun_out/${P}_bench_tele.json 2> gpurun_out/${P}_bench_tele.err
MDGAT_SK_THREADS=1024 timeout 300 python bench.py $Q > gpurun_out/${P}_bench_sk1024.json 2> gpurun_out/${P}_bench_sk1024.err
timeout 300 python bench.py $Q --no-cuda-graph > gpurun_out/${P}_bench_nograph.json 2> gpurun_out/${P}_bench_nograph.err
MDGAT_SK_THREADS=1024 timeout 300 python bench.py --n 2048 --steps 3 --warmup 3 $Q > gpurun_out/${P}_bench_cfg4_sk1024.json 2> gpurun_out/${P}_bench_cfg4_sk1024.err
timeout 300 python bench.py --n 2048 --steps 3 --warmup 3 $Q > gpurun_out/${P}_bench_cfg4.json 2> gpurun_out/${P}_bench_cfg4.err
python - <<'PY'
import json
for t in ['','_tele','_sk1024','_nograph','_cfg4','_cfg4_sk1024']:
    try:
        d=json.loads(open('gpurun_out/s2_bench%s.json'%t).read().strip().splitlines()[-1])
        st=d['roofline']['stage_ms_per_step']
        print(t, round(d['value'],1), round(d['e2e']['value'],1), d['gpu_launches'], {k:round(v,3) for k,v in st.items()}, d['config'].get('sinkhorn',{}).get('iterations_run_per_pair'), (d.get('parity') or {}).get('max_score_err'))
    except Exception as e: print(t, 'ERR', e)
PY
tail -5 gpurun_out/s2_bench.err
