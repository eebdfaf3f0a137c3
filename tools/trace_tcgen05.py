"""Timeline of one CTA of the tcgen05 kernels (mdgat_debug_trace): per role the (tag, clock64) records of CTA (0,0,0).

Ozaki GEMM tags: loader 1000+u issue start / 2000+u copies issued; MMA thread 3000+u unit start / 4000+u operands
landed / 5000+u accumulator set free / 6000+u MMAs issued; epilogue warps 7000+u before the wait on the tensor core /
8000+u accumulators ready / 9000+u TMEM drained (then the float64 Horner pass + stores until the next 7000).
Usage on a GPU box: python tools/trace_tcgen05.py [R]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mdgat_matcher_b200 import ops, _capi

ROLES = ('loader', 'mma', 'epi warp0', 'epi warp4')
CAP = 1024


def run(name, fn, flags=0, full=True):
    dev = torch.device('cuda:0')
    fn()
    torch.cuda.synchronize()
    buf = torch.zeros(8 * (2 + 2 * CAP), dtype=torch.int64, device=dev)
    _capi.check(_capi.lib.mdgat_debug_trace(buf.data_ptr()))
    _capi.check(_capi.lib.mdgat_debug_flags(flags))
    fn()
    torch.cuda.synchronize()
    _capi.check(_capi.lib.mdgat_debug_flags(0))
    _capi.check(_capi.lib.mdgat_debug_trace(None))
    b = buf.cpu().view(8, 2 + 2 * CAP)
    recs = {}
    t0 = None
    for r, role in enumerate(ROLES):
        n = int(b[r, 0])
        rr = [(int(b[r, 2 + 2 * i]), int(b[r, 3 + 2 * i])) for i in range(n)]
        recs[role] = rr
        if rr:
            t0 = rr[0][1] if t0 is None else min(t0, rr[0][1])
    print('==== %s (debug flags %d)' % (name, flags))
    for role in ROLES:
        rr = recs[role]
        if not rr or not full:
            continue
        print('-- %s (%d records)' % (role, len(rr)))
        prev = None
        line = []
        for tag, clk in rr:
            line.append('%d@%d(+%d)' % (tag, clk - t0, 0 if prev is None else clk - prev))
            prev = clk
            if len(line) == 6:
                print('   ' + '  '.join(line)); line = []
        if line:
            print('   ' + '  '.join(line))
    # per-unit summary of the MMA thread and the first epilogue warp
    mma = dict(recs['mma']); epi = dict(recs['epi warp0'])
    units = sorted(t - 3000 for t in mma if 3000 <= t < 4000)
    if units:
        print('-- per unit: wait operands | wait tmem | issue | epi: wait mma | drain tmem | math+stores')
        for u in units:
            if 8000 + u not in epi:
                print('   u=%2d  %6d %6d %6d' % (u, mma[4000 + u] - mma[3000 + u], mma[5000 + u] - mma[4000 + u], mma[6000 + u] - mma[5000 + u]))
                continue
            try:
                nxt = epi.get(7000 + u + 1, None)
                print('   u=%2d  %6d %6d %6d | %6d %6d %s' % (
                    u, mma[4000 + u] - mma[3000 + u], mma[5000 + u] - mma[4000 + u], mma[6000 + u] - mma[5000 + u],
                    epi[8000 + u] - epi[7000 + u], epi[9000 + u] - epi[8000 + u],
                    '%6d' % (nxt - epi[9000 + u]) if nxt else '     -'))
            except KeyError:
                pass
        last = max(c for r in recs.values() for _, c in r)
        print('   CTA span (first record -> last epilogue record): %d cycles for %d units = %.0f cycles/unit' %
              (last - t0, len(units), (last - t0) / len(units)))


def main():
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    dev = torch.device('cuda:0')
    g = torch.Generator(device='cpu').manual_seed(0)
    x = torch.randn(R, 128, generator=g, dtype=torch.float64).to(dev)
    m = torch.randn(R, 128, generator=g, dtype=torch.float64).to(dev)
    h = torch.randn(R, 256, generator=g, dtype=torch.float64).to(dev)
    wqkv = torch.randn(384, 128, generator=g, dtype=torch.float64).to(dev)
    w1 = torch.randn(256, 256, generator=g, dtype=torch.float64).to(dev)
    w2 = torch.randn(128, 256, generator=g, dtype=torch.float64).to(dev)
    b1 = torch.randn(256, generator=g, dtype=torch.float64).to(dev)
    run('plain K=128 N=384 (q/k/v shape)', lambda: ops.linear_i8(x, wqkv))
    run('MLP 256->256 relu (cat[x, msg])', lambda: ops.linear_i8(x, w1, bias=b1, relu=True, x2=m))
    run('MLP 256->128 + residual', lambda: ops.linear_i8(h, w2, residual=x))
    # bits 8..13 switch off the loader / MMA-thread marks 1000..6000: fewer probes, less perturbation
    only6000 = (1 + 2 + 4 + 8 + 16) << 8
    for flags in (only6000, only6000 + 1, only6000 + 2, only6000 + 3):
        run('plain K=128 N=384', lambda: ops.linear_i8(x, wqkv), flags, True)
    for flags in ():
        run('plain K=128 N=384', lambda: ops.linear_i8(x, wqkv), flags, False)
        run('MLP 256->256 relu', lambda: ops.linear_i8(x, w1, bias=b1, relu=True, x2=m), flags, False)


if __name__ == '__main__':
    main()
