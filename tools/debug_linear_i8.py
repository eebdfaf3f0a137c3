"""Triage of the Ozaki GEMM on a GPU box: where (rows, columns) the error against numpy sits."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mdgat_matcher_b200 import ops
dev = torch.device('cuda:0')
def t(a): return torch.from_numpy(np.ascontiguousarray(a)).to(dev)
for (R, K, Nout, relu, useres) in [(333, 256, 256, True, True), (333, 256, 256, False, False), (333, 256, 256, True, False), (333, 256, 256, False, True), (4096, 256, 128, True, True), (1000, 128, 384, True, True)]:
    rng = np.random.default_rng(R + K + Nout)
    x = rng.normal(size=(R, K)); w = rng.normal(size=(Nout, K)) / np.sqrt(K); b = rng.normal(size=Nout); res = rng.normal(size=(R, Nout))
    want = x @ w.T + b
    if relu: want = np.maximum(want, 0)
    if useres: want = want + res
    got = ops.linear_i8(t(x), t(w), t(b), relu=relu, residual=t(res) if useres else None).cpu().numpy()
    err = np.abs(got - want)
    bad = err > 1e-9
    print(R, K, Nout, 'relu', relu, 'res', useres, 'max err %.3e' % err.max(), 'bad frac %.4f' % bad.mean(),
          'bad rows%%8:', np.bincount(np.nonzero(bad)[0] % 8, minlength=8).tolist(), 'bad cols%%32:', np.bincount(np.nonzero(bad)[1] % 32, minlength=32).tolist(), flush=True)
    if bad.any():
        r, c = np.nonzero(bad); print('   first bad', r[0], c[0], got[r[0], c[0]], want[r[0], c[0]], (x[r[0], :128] @ w[c[0], :128]), (x[r[0], 128:] @ w[c[0], 128:]), b[c[0]], res[r[0], c[0]])
