"""Triage of the tcgen05 attention engine on a GPU box: error of each stage against the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from mdgat_matcher_b200 import ops
from oracle import mdgat_oracle as O

dev = torch.device('cuda:0')
def t(a): return torch.from_numpy(np.ascontiguousarray(a)).to(dev)
for (N, M, qs, vs) in [(128, 32, 1, 1), (128, 64, 1, 1), (128, 128, 3, 1), (200, 77, 3, 1), (512, 512, 5, 3)]:
    rng = np.random.default_rng(N + M)
    q = rng.normal(size=(2, 128, N)) * qs; k = rng.normal(size=(2, 128, M)) * qs; v = rng.normal(size=(2, 128, M)) * vs
    want, _ = O.attention(q.reshape(2, 32, 4, N), k.reshape(2, 32, 4, M), v.reshape(2, 32, 4, M))
    want = want.reshape(2, 128, N)
    res = {}
    for name, kw in [('dmma', dict(engine='dmma')), ('i8_logits+topk(M)', dict(engine='tcgen05_i8', topk=M)), ('i8_full', dict(engine='tcgen05_i8'))]:
        try:
            got = ops.attention(t(q), t(k), t(v), **kw).cpu().numpy()
            torch.cuda.synchronize()
            err = np.abs(got - want)
            res[name] = '%.2e (nan %d, argmax %s)' % (np.nanmax(err), int(np.isnan(got).sum()), np.unravel_index(np.nanargmax(err), err.shape))
        except Exception as e:
            res[name] = 'EXC ' + repr(e)[:200]
    print(N, M, res, flush=True)
