"""Turns the ncu outputs of tools/gpu_round.sh (gpurun_out/) into the tracked summaries under profiles/.

  python tools/summarize_profiles.py <tag>      # e.g. r1_final
Reads gpurun_out/launches_final.csv (gpu__time_duration.sum per launch), the --set full reports
prof_{attn_i8,oz,misc}_final.ncu-rep (through `ncu -i ... --page raw --csv`) and bench_final*.json.
"""
import collections
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'gpurun_out')
METRICS = [
    ('gpu__time_duration.sum', 'duration'),
    ('launch__grid_size', 'grid'),
    ('launch__registers_per_thread', 'regs'),
    ('sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active', 'IMMA pipe %'),
    ('sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'FP64 pipe %'),
    ('smsp__issue_active.avg.pct_of_peak_sustained_active', 'issue slots %'),
    ('sm__warps_active.avg.pct_of_peak_sustained_active', 'warps active %'),
    ('dram__bytes_read.sum', 'DRAM read'),
    ('dram__bytes_write.sum', 'DRAM write'),
    ('gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'DRAM throughput %'),
    ('lts__throughput.avg.pct_of_peak_sustained_elapsed', 'L2 throughput %'),
    ('lts__t_sector_hit_rate.pct', 'L2 hit %'),
    ('smsp__inst_executed.sum', 'warp instructions'),
]


def launch_table(path, forwards):
    lines = [l for l in open(path) if not l.startswith('==')]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get('Metric Name') != 'gpu__time_duration.sum':
            continue
        k = row['Kernel Name']
        if 'mdgat' not in k:
            continue
        v = float(row['Metric Value'].replace(',', ''))
        v = {'ns': v / 1000, 'us': v, 'ms': v * 1000}[row['Metric Unit']]
        a = agg.setdefault(k.split('(')[0][-60:], [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for k, a in agg.items() if 'peak_kernel' not in k)
    out = ['| kernel | launches / forward | us / forward | avg us | share |', '|---|---|---|---|---|']
    for k, a in agg.items():
        if 'peak_kernel' in k:
            continue
        out.append('| `%s` | %.1f | %.1f | %.1f | %.3f |' % (k, a[0] / forwards, a[1] / forwards, a[1] / a[0], a[1] / tot))
    out.append('')
    out.append('Sum over the path kernels: %.1f us per forward (serialised, cold-cache launches; compare shares).' % (tot / forwards))
    return '\n'.join(out), agg


def ncu_raw(rep):
    """Raw-page CSV of a report: the export made on the GPU box (<name>_raw.csv) when present, else `ncu -i`."""
    pre = rep[:-len('.ncu-rep')] + '_raw.csv'
    if os.path.isfile(pre):
        return open(pre).read()
    return subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout


def ncu_table(rep):
    raw = ncu_raw(rep)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    out = ['| kernel | ' + ' | '.join(n for _, n in METRICS) + ' |', '|---|' + '---|' * len(METRICS)]
    for r in rows[2:]:
        cells = []
        for m, _ in METRICS:
            if m in ix:
                v, u = r[ix[m]], units[ix[m]]
                try:
                    v = '%.1f' % float(v) if '.' in v else v
                except ValueError:
                    pass
                cells.append((v + ' ' + u).strip())
            else:
                cells.append('-')
        out.append('| `%s` | ' % r[ix['Kernel Name']][:48] + ' | '.join(cells) + ' |')
    return '\n'.join(out)


def ncu_traffic(rep):
    """{kernel name: {'launches', 'dram_bytes_per_launch', 'us_per_launch'}} averaged over the launches captured in `rep`."""
    raw = ncu_raw(rep)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    mult = {'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}
    tmul = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 's': 1e6}
    out = {}
    for r in rows[2:]:
        name = r[ix['Kernel Name']].split('(')[0].replace('void ', '')
        b = 0.0
        for m in ('dram__bytes_read.sum', 'dram__bytes_write.sum'):
            b += float(r[ix[m]].replace(',', '')) * mult[units[ix[m]]]
        t = float(r[ix['gpu__time_duration.sum']].replace(',', '')) * tmul[units[ix['gpu__time_duration.sum']]]
        a = out.setdefault(name, {'launches': 0, 'dram_bytes_per_launch': 0.0, 'us_per_launch': 0.0})
        a['launches'] += 1
        a['dram_bytes_per_launch'] += b
        a['us_per_launch'] += t
    for a in out.values():
        a['dram_bytes_per_launch'] /= a['launches']
        a['us_per_launch'] /= a['launches']
    return out


def main():
    """python tools/summarize_profiles.py <tag> [<prefix>]: gpurun_out/<prefix>_bench.json, <prefix>_bench_ref.json (optional),
    <prefix>_launches.csv and every <prefix>_prof_*.ncu-rep -> profiles/<tag>_bench.json, _bench_reference_arm.json, _ncu_summary.md.
    Without a prefix: the round-1 file names (bench_final.json, launches_final.csv, prof_*_final.ncu-rep)."""
    import glob
    tag = sys.argv[1] if len(sys.argv) > 1 else 'r1_final'
    prefix = sys.argv[2] if len(sys.argv) > 2 else None
    if prefix:
        fb, fr, fl = prefix + '_bench.json', prefix + '_bench_ref.json', prefix + '_launches.csv'
        # a report may have been dropped on the box to fit the copy-back limit: its raw CSV export stands in for it
        reps = sorted({f[:-len('_raw.csv')] + '.ncu-rep' for f in glob.glob(os.path.join(OUT, prefix + '_prof_*_raw.csv'))} |
                      set(glob.glob(os.path.join(OUT, prefix + '_prof_*.ncu-rep'))))
    else:
        fb, fr, fl = 'bench_final.json', 'bench_final_ref.json', 'launches_final.csv'
        reps = [os.path.join(OUT, n + '.ncu-rep') for n in ('prof_attn_i8_final', 'prof_oz_final', 'prof_misc_final')]
    bench = json.loads(open(os.path.join(OUT, fb)).read().strip().splitlines()[-1])
    json.dump(bench, open(os.path.join(ROOT, 'profiles', tag + '_bench.json'), 'w'), indent=1)
    ref = None
    if os.path.isfile(os.path.join(OUT, fr)):
        ref = json.loads(open(os.path.join(OUT, fr)).read().strip().splitlines()[-1])
        json.dump(ref, open(os.path.join(ROOT, 'profiles', tag + '_bench_reference_arm.json'), 'w'), indent=1)
    # the launch list covers warm-up + timed + e2e forwards of `bench.py --steps 2 --warmup 1`: count them by a once-per-forward kernel
    table, agg = launch_table(os.path.join(OUT, fl), 1)
    forwards = [a[0] for k, a in agg.items() if 'pack_inputs_kernel' in k][0]
    table, _ = launch_table(os.path.join(OUT, fl), forwards)
    cb, eg = bench.get('cpu_baseline'), bench.get('gpu_eager_reference') or bench.get('gpu_eager_port')
    md = ['# %s -- ncu launch list and top-kernel captures (cfg2: B=32, N=M=512, L=9, T=100, one B200)' % tag, '',
          'Commands: `tools/gpu_round.sh` (launch list: `ncu --metrics gpu__time_duration.sum --clock-control none` on',
          '`bench.py --steps 2 --warmup 1`; captures: `ncu --set full --clock-control none --import-source on -k <kernel>`).', '',
          'Bench of the same build (`%s_bench.json`): **%.1f pairs/s** resident, %.1f pairs/s end to end%s%s.'
          % (tag, bench['value'], bench['e2e']['value'],
             ', CPU %s %.2f pairs/s on %d threads' % (cb['kind'], cb['value'], cb['cores']) if cb else '',
             ', unmodified reference as eager PyTorch fp64 on the same GPU %.1f pairs/s' % eg['pairs_per_s'] if eg else ''), '',
          'Live stage times (CUDA events on the launch stream, ms per forward): `%s`' % json.dumps({k: round(v, 3) for k, v in bench['roofline']['stage_ms_per_step'].items()}), '',
          '## Launch list (%d forwards in the capture)' % forwards, '', table, '']
    traffic = {}
    for rep in reps:
        if os.path.isfile(rep) or os.path.isfile(rep[:-len('.ncu-rep')] + '_raw.csv'):
            md += ['## ncu --set full: %s' % os.path.basename(rep), '', ncu_table(rep), '']
            traffic.update(ncu_traffic(rep))
    if traffic:
        # per-launch DRAM bytes of each captured kernel (dram__bytes_read.sum + dram__bytes_write.sum): bench.py reports the
        # dominant kernel's figure as roofline.traffic
        json.dump({'source': 'ncu --set full --clock-control none captures of `bench.py --steps 1 --warmup 1` (%s)' % ', '.join(os.path.basename(r) for r in reps),
                   'kernels': traffic}, open(os.path.join(ROOT, 'profiles', tag + '_traffic.json'), 'w'), indent=1)
    open(os.path.join(ROOT, 'profiles', tag + '_ncu_summary.md'), 'w').write('\n'.join(md))
    print('\n'.join(md)[:3000])


if __name__ == '__main__':
    main()
