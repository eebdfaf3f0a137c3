"""Builds oracle/_ref/ (git-ignored, travels to the GPU box with the snapshot).

The reference is pure Python, so there is nothing to compile; what cannot travel is the
reference tree itself. This script extracts the one fixture the parity tests need from it:
the pre-trained weights (/root/reference/pre-trained/best_model.pth, 348 fp64 tensors),
rounded to fp32 exactly as test.py:156-159 does when it loads them into an fp32 module.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')
WEIGHTS = os.path.join(REF_DIR, 'best_model_fp32.npz')
CHECKPOINT = os.environ.get('MDGAT_REFERENCE_ROOT', '/root/reference') + '/pre-trained/best_model.pth'


def build(force=False):
    if os.path.isfile(WEIGHTS) and not force:
        return WEIGHTS
    if not os.path.isfile(CHECKPOINT):
        return None
    import torch
    ck = torch.load(CHECKPOINT, map_location='cpu', weights_only=True)
    out = {}
    for k, v in ck['net'].items():
        k = k[7:] if k.startswith('module.') else k
        a = v.numpy()
        out[k] = a.astype(np.float32) if a.dtype.kind == 'f' else a
    os.makedirs(REF_DIR, exist_ok=True)
    np.savez(WEIGHTS, **out)
    return WEIGHTS


def load_checkpoint_state_dict():
    """Returns {name: float64 ndarray} = fp64(fp32(checkpoint)), or None if unavailable."""
    path = build()
    if path is None:
        return None
    with np.load(path) as z:
        return {k: (z[k].astype(np.float64) if z[k].dtype.kind == 'f' else z[k]) for k in z.files}


if __name__ == '__main__':
    print(build(force='--force' in sys.argv))
