"""Builds oracle/_ref/ (git-ignored, travels to the GPU box with the gpurun snapshot).

TEST / BASELINE INFRASTRUCTURE ONLY: nothing under mdgat-matcher_b200/ reads oracle/_ref/.

The reference is pure Python, so there is nothing to compile; what cannot travel is the reference tree itself
(/root/reference does not exist on the GPU box). This script places into oracle/_ref/:

* best_model_fp32.npz -- the pre-trained weights (/root/reference/pre-trained/best_model.pth, 348 fp64 tensors)
  rounded to fp32 exactly as test.py:156-159 does when it loads them into an fp32 module;
* reference/ -- byte-identical copies of the reference files the baseline arms and the cfg5 test execute UNMODIFIED
  on the GPU box: models/mdgat.py (+ models/superglue.py and the models/pointnet/pointnet_util.py it imports) for
  `bench.py --impl reference` (CPU) and the eager-GPU denominator; test.py, test_registration_metric.py, load_data.py
  and utils/utils_test.py for the unchanged-script run through the launcher (tests/test_reference_scripts_gpu.py).
  The copies never enter git history (`oracle/_ref/` is in .gitignore) and a sha256 manifest records what was copied.
"""
import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, '_ref')
WEIGHTS = os.path.join(REF_DIR, 'best_model_fp32.npz')
REF_COPY = os.path.join(REF_DIR, 'reference')
REFERENCE_ROOT = os.environ.get('MDGAT_REFERENCE_ROOT', '/root/reference')
CHECKPOINT = REFERENCE_ROOT + '/pre-trained/best_model.pth'
REF_FILES = ('models/mdgat.py', 'models/superglue.py', 'models/pointnet/pointnet_util.py',
             'test.py', 'test_registration_metric.py', 'train.py', 'load_data.py', 'utils/utils_test.py')


def build_weights(force=False):
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


def build_reference_copy(force=False):
    """Copies REF_FILES byte for byte; returns the directory, or None when the reference tree is absent and no
    earlier copy exists."""
    manifest = os.path.join(REF_COPY, 'MANIFEST.json')
    have = os.path.isfile(manifest)
    if not os.path.isfile(os.path.join(REFERENCE_ROOT, REF_FILES[0])):
        return REF_COPY if have else None
    if have and not force:
        return REF_COPY
    sums = {}
    for rel in REF_FILES:
        src, dst = os.path.join(REFERENCE_ROOT, rel), os.path.join(REF_COPY, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, 'rb') as f:
            sums[rel] = hashlib.sha256(f.read()).hexdigest()
    with open(manifest, 'w') as f:
        json.dump({'source': REFERENCE_ROOT, 'sha256': sums}, f, indent=1)
    return REF_COPY


def reference_root():
    """Where an unmodified reference tree can be imported from: /root/reference in the build container, the copy under
    oracle/_ref/reference on the GPU box; None if neither exists."""
    if os.path.isfile(os.path.join(REFERENCE_ROOT, REF_FILES[0])):
        return REFERENCE_ROOT
    if os.path.isfile(os.path.join(REF_COPY, REF_FILES[0])):
        return REF_COPY
    return None


def build(force=False):
    build_reference_copy(force)
    return build_weights(force)


def load_checkpoint_state_dict():
    """Returns {name: float64 ndarray} = fp64(fp32(checkpoint)), or None if unavailable."""
    path = build_weights()
    if path is None:
        return None
    with np.load(path) as z:
        return {k: (z[k].astype(np.float64) if z[k].dtype.kind == 'f' else z[k]) for k in z.files}


def write_checkpoint_pth(path):
    """A torch checkpoint in the layout of pre-trained/best_model.pth (train.py:288-304: net with the DataParallel
    'module.' prefix, optimizer, epoch, lr_schedule, loss) rebuilt from best_model_fp32.npz, for running the reference's
    unchanged scripts where the 71 MB original did not travel. Loading it into the fp32 module test.py builds gives
    bit-identical parameters (the npz already holds the fp32 rounding)."""
    import torch
    sd = load_checkpoint_state_dict()
    if sd is None:
        return None
    net = {'module.' + k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}
    torch.save({'net': net, 'optimizer': {}, 'epoch': 17, 'lr_schedule': 1e-4, 'loss': 0.29647}, path)
    return path


if __name__ == '__main__':
    print(build(force='--force' in sys.argv), reference_root())
