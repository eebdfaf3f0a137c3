"""Prints a one-line summary of bench.py JSON result files (development helper)."""
import json
import sys

for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().split('\n')[-1])
    except Exception as e:
        print('%-40s ERR %s' % (f, e))
        continue
    r = d.get('roofline') or {}
    st = {k: round(v, 3) for k, v in (r.get('stage_ms_per_step') or {}).items()}
    p = d.get('parity') or {}
    print('%-40s %8.1f pairs/s %7.3f ms  e2e %8.1f  launches %s  flips %s err %s' % (
        f.split('/')[-1], d['value'], d['ms_per_step'], (d.get('e2e') or {}).get('value', 0), d.get('gpu_launches'),
        p.get('index_flips'), p.get('max_score_err')))
    print('    stages', st)
    for k in ('cpu_baseline', 'gpu_eager_reference', 'latency_batch1'):
        if d.get(k):
            v = d[k]
            print('   ', k, {kk: vv for kk, vv in v.items() if kk not in ('what', 'sample', 'shape')})
