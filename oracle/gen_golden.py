"""Generates tests/golden/*.npz by running the UNMODIFIED reference (oracle/ref_loader.py)
on seeded synthetic inputs. Run in the build container only (needs /root/reference):

    python oracle/gen_golden.py

Each file stores the case description, the inputs, and the reference outputs
(matches0/1, matching_scores0/1, loss, and Z / per-layer descriptors where small).
torch version and thread count are recorded because the arithmetic lives in torch.
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

OUT = os.path.join(ROOT, 'tests', 'golden')

CASES = [
    # name, weights, cfg overrides, input spec, what to store
    dict(name='cfg1_seeded_L4_n128', weights='seeded', L=4, T=20, B=1, N=128, M=128, seed=1,
         store_Z=True, store_layers=True),
    dict(name='ckpt_L9_n512_T100', weights='checkpoint', L=9, T=100, B=2, N=512, M=512, seed=0,
         store_Z=True),
    dict(name='ckpt_L9_ragged_gap', weights='checkpoint', L=9, T=20, B=2, N=256, M=200, seed=2,
         loss_method='gap_loss', store_Z=True),
    dict(name='ckpt_L9_superglue_mode', weights='checkpoint', L=9, T=50, B=1, N=384, M=320, seed=3,
         k=[], loss_method='gap_loss', store_Z=True),
    dict(name='seeded_L9_n512', weights='seeded', L=9, T=100, B=2, N=512, M=512, seed=4,
         store_Z=False),
    dict(name='ckpt_L9_duplicates', weights='checkpoint', L=9, T=20, B=1, N=256, M=256, seed=5,
         duplicates=40, store_Z=True),
    dict(name='ckpt_L9_sgloss_mutual', weights='checkpoint', L=9, T=20, B=1, N=192, M=192, seed=6,
         loss_method='superglue', mutual_check=True, store_Z=True, no_loss=True),
    dict(name='ckpt_L9_n2048', weights='checkpoint', L=9, T=100, B=1, N=2048, M=2048, seed=7,
         store_Z=False),
    dict(name='ckpt_L9_n512_b8', weights='checkpoint', L=9, T=100, B=8, N=512, M=512, seed=8,
         store_Z=False),
    # non-finite inputs (a zero-norm FPFH row becomes NaN in load_data.py:290): 'poke' = (key, pair, row, column or None, value).
    # The whole pair turns NaN upstream; these pin what match extraction and the losses then report, and that the other
    # pairs of the batch are untouched
    dict(name='seeded_L2_nan_triplet', weights='seeded', L=2, T=20, B=3, N=64, M=64, seed=11, k=[16, None, 16, None],
         store_Z=False, poke=[['descriptors0', 1, 7, 3, 'nan']]),
    dict(name='seeded_L2_nan_gap_ragged', weights='seeded', L=2, T=20, B=3, N=80, M=64, seed=12, k=[16, None, 16, None],
         loss_method='gap_loss', store_Z=False, poke=[['keypoints1', 0, 5, 1, 'nan'], ['scores0', 2, 9, None, 'inf']]),
    dict(name='seeded_L2_nan_sg_mutual', weights='seeded', L=2, T=20, B=3, N=64, M=64, seed=13, k=[16, None, 16, None],
         loss_method='superglue', mutual_check=True, store_Z=False, poke=[['descriptors1', 2, 0, 0, '-inf']]),
    dict(name='seeded_L2_nan_sg', weights='seeded', L=2, T=20, B=2, N=64, M=64, seed=14, k=[16, None, 16, None],
         loss_method='superglue', mutual_check=False, store_Z=False, poke=[['descriptors0', 0, 63, 32, 'nan']]),
]


def run_case(c):
    k = c.get('k', None)
    cfg = RL.net_config(L=c['L'], k=k, sinkhorn_iterations=c['T'],
                        loss_method=c.get('loss_method', 'triplet_loss'),
                        mutual_check=c.get('mutual_check', False))
    if c['weights'] == 'seeded':
        sd = synth.seeded_state_dict(c['L'], c['seed'])
        weights = {'module.' + kk: v for kk, v in sd.items()}
    else:
        weights = 'checkpoint'
    net, mod, zcap = RL.build_reference_net(cfg, weights)
    data = synth.make_batch(c['seed'], c['B'], c['N'], c['M'], duplicates=c.get('duplicates', 0))
    for key, b, r, col, val in c.get('poke', []):
        if col is None:
            data[key][b, r] = float(val)
        else:
            data[key][b, r, col] = float(val)
    layers = []
    if c.get('store_layers'):
        def hook(m, inp, out):
            layers.append(out.detach().numpy().copy())
        for l in net.module.gnn.layers:
            l.register_forward_hook(hook)
    t = time.time()
    if c.get('no_loss'):
        # the reference's 'superglue' loss block reads gt through different indexing; keep gt
        pass
    out = RL.run_reference(net, data)
    dt = time.time() - t
    rec = {'case': json.dumps({kk: v for kk, v in c.items()}),
           'torch_version': torch.__version__, 'threads': torch.get_num_threads()}
    for kk, v in data.items():
        rec['in_' + kk] = v.numpy()
    for kk in ('matches0', 'matches1', 'matching_scores0', 'matching_scores1'):
        rec[kk] = out[kk].numpy()
    rec['loss'] = np.asarray(out['loss'].numpy())
    Z = zcap['Z'].numpy()
    if c.get('store_Z'):
        rec['Z'] = Z
    rec['Z_rowmax'] = Z[:, :-1, :].max(2)
    rec['Z_colmax'] = Z[:, :, :-1].max(1)
    rec['scores_in_sum'] = np.asarray(zcap['scores_in'].numpy().sum())
    if layers:
        # hook fires for side 0 then side 1 of every layer: store the deltas
        rec['layer_deltas'] = np.stack(layers)
    np.savez_compressed(os.path.join(OUT, c['name'] + '.npz'), **rec)
    frac = float((out['matches0'] >= 0).double().mean())
    print('%-28s %.2fs matched %.3f loss %s' % (c['name'], dt, frac, rec['loss']))


if __name__ == '__main__':
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for c in CASES:
        if only and c['name'] not in only:
            continue
        run_case(c)
