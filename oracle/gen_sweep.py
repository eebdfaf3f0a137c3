"""Generates tests/golden/sweep/*.npz: outputs of the UNMODIFIED reference (oracle/ref_loader.py) on a
large seeded sample of the cfg2 workload. Run in the build container only (needs /root/reference):

    python oracle/gen_sweep.py            # all cases (about 1 min of CPU per case)

The sweep is the gate for every precision decision of the CUDA path (number of digit planes per operand,
DESIGN.md section 2): >= 100 k keypoint rows at the headline shape (N = M = 512, L = 9, T = 100, default k list,
pre-trained weights), i.e. SURVEY.md 7.3's ">= 50 k rows before any precision claim".

Only OUTPUTS are stored (matches as int16, scores and the row / column maxima of Z as float64, about 0.3 MB per
16-pair case); the inputs are regenerated from the seed by mdgat_matcher_b200.synth.make_batch (CPU torch generator,
deterministic) and pinned by a float64 checksum per input tensor stored next to the outputs.
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader as RL            # noqa: E402
from mdgat_matcher_b200 import synth           # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'sweep')

# (name, seed, B). seed 1000 / B = 32 is exactly the batch bench.py times on rank 0.
CASES = [('sweep_s%d_b16' % s, s, 16) for s in (100, 101, 102, 103, 104, 105)] + [('cfg2_s1000_b32', 1000, 32)]
N = 512


def input_checksums(data):
    return {k: float(v.double().sum()) for k, v in sorted(data.items())}


def run_case(net, zcap, name, seed, B):
    data = synth.make_batch(seed, B, N)
    t = time.time()
    out = RL.run_reference(net, data)
    dt = time.time() - t
    Z = zcap['Z'].numpy()
    rec = {
        'case': json.dumps({'name': name, 'seed': seed, 'B': B, 'N': N, 'M': N, 'L': 9, 'T': 100, 'weights': 'checkpoint'}),
        'torch_version': torch.__version__, 'threads': torch.get_num_threads(),
        'input_checksums': json.dumps(input_checksums(data)),
        'matches0': out['matches0'].numpy().astype(np.int16), 'matches1': out['matches1'].numpy().astype(np.int16),
        'matching_scores0': out['matching_scores0'].numpy(), 'matching_scores1': out['matching_scores1'].numpy(),
        'loss': np.asarray(out['loss'].numpy()),
        'Z_rowmax': Z[:, :-1, :].max(2), 'Z_colmax': Z[:, :, :-1].max(1),
    }
    # arg-max margin of every row / column of Z (best minus second best): how close the reference itself is to a flip
    zr = np.sort(Z[:, :-1, :], axis=2)
    zc = np.sort(Z[:, :, :-1], axis=1)
    rec['Z_rowmargin'] = zr[:, :, -1] - zr[:, :, -2]
    rec['Z_colmargin'] = zc[:, -1, :] - zc[:, -2, :]
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **rec)
    print('%-18s %.1fs matched %.3f min margin %.2e loss %s' %
          (name, dt, float((out['matches0'] >= 0).double().mean()),
           float(min(rec['Z_rowmargin'].min(), rec['Z_colmargin'].min())), rec['loss']), flush=True)


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    cfg = RL.net_config(L=9, sinkhorn_iterations=100)
    net, mod, zcap = RL.build_reference_net(cfg, 'checkpoint')
    for name, seed, B in CASES:
        if only and name not in only:
            continue
        run_case(net, zcap, name, seed, B)
