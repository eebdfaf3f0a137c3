set -x
mkdir -p gpurun_out
P=s1
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/${P}_pytest.txt 2>&1; tail -3 gpurun_out/${P}_pytest.txt
timeout 600 python bench.py > gpurun_out/${P}_bench.json 2> gpurun_out/${P}_bench.err; tail -2 gpurun_out/${P}_bench.err
MDGAT_SK_TOL=0 timeout 300 python bench.py --no-cpu-baseline --no-eager --no-latency > gpurun_out/${P}_bench_tol0.json 2> gpurun_out/${P}_bench_tol0.err
timeout 300 python bench.py --cuda-graph --no-cpu-baseline --no-eager --no-latency > gpurun_out/${P}_bench_graph.json 2> gpurun_out/${P}_bench_graph.err
timeout 300 python bench.py --n 2048 --steps 3 --warmup 3 --no-cpu-baseline --no-eager --no-latency > gpurun_out/${P}_bench_cfg4.json 2> gpurun_out/${P}_bench_cfg4.err
cp gpurun_out/precision_sweep.json gpurun_out/${P}_precision_sweep.json 2>/dev/null
python - <<'PY'
import json
for t in ['','_tol0','_graph','_cfg4']:
    try:
        d=json.loads(open('gpurun_out/s1_bench%s.json'%t).read().strip().splitlines()[-1])
        print(t, round(d['value'],1), d['roofline']['stage_ms_per_step'], d['config'].get('sinkhorn'), d.get('parity'))
    except Exception as e: print(t, 'ERR', e)
PY
