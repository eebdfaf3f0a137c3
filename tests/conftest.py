import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    rec = {k: z[k] for k in z.files}
    rec['case'] = json.loads(str(rec['case']))
    return rec


def golden_inputs(rec):
    return {k[3:]: v for k, v in rec.items() if k.startswith('in_')}


def case_cfg(case):
    from oracle.ref_loader import DEFAULT_K
    return {
        'L': case['L'], 'k': case.get('k', list(DEFAULT_K)), 'sinkhorn_iterations': case['T'],
        'loss_method': case.get('loss_method', 'triplet_loss'),
        'mutual_check': case.get('mutual_check', False), 'match_threshold': 0.2,
        'triplet_loss_gamma': 0.5, 'descriptor': 'FPFH', 'lr': 1e-4, 'train_step': 3,
    }


def case_weights(case):
    """numpy fp64 state dict for a golden case; skips when the checkpoint fixture is absent."""
    from mdgat_matcher_b200 import synth
    if case['weights'] == 'seeded':
        sd = synth.seeded_state_dict(case['L'], case['seed'])
        return {k: v.numpy() for k, v in sd.items()}
    from oracle.build_ref import load_checkpoint_state_dict
    sd = load_checkpoint_state_dict()
    if sd is None:
        pytest.skip('oracle/_ref/best_model_fp32.npz missing (run __graft_entry__.build() where '
                    '/root/reference exists)')
    return sd
