"""BASELINE config 5 for real: the reference's UNCHANGED evaluation script test_registration_metric.py
(/root/reference/test_registration_metric.py:130-286) run end to end against the drop-in through the launcher, on a
KITTI-format sequence written by kitti_io.write_synthetic_sequence (the real keypoint files are a separate download,
SURVEY.md fact 10), batch 1, 256 keypoints, T = 20, the shipped checkpoint.

The script's printed per-pair lines and summary line must equal what this package's own GPU-resident pipeline computes for
the same pairs: PairBatcher (prepare_pairs_kernel) -> MDGAT.forward -> register_pairs_kernel, i.e. the batched
replacement of the script's loader, numpy Kabsch and metric code.

The script file comes from oracle/_ref/reference (byte-identical copy made by oracle/build_ref.py; /root/reference itself
in the build container). Nothing here reads /root/reference at run time on the GPU box.
"""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _script(name):
    from oracle.build_ref import reference_root
    root = reference_root()
    if root is None or not os.path.isfile(os.path.join(root, name)):
        pytest.skip('reference scripts not available (oracle/_ref/reference missing)')
    return os.path.join(root, name)


def test_registration_metric_script_runs_unchanged_and_agrees(tmp_path):
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from oracle.build_ref import write_checkpoint_pth
    from mdgat_matcher_b200 import kitti_io, ops
    from mdgat_matcher_b200.models.mdgat import MDGAT
    script = _script('test_registration_metric.py')
    ckpt = write_checkpoint_pth(str(tmp_path / 'best_model.pth'))
    if ckpt is None:
        pytest.skip('oracle/_ref/best_model_fp32.npz missing')
    frames, n = 14, 256
    dirs = kitti_io.write_synthetic_sequence(str(tmp_path / 'KITTI'), seq=10, frames=frames, n_kpts=n, n_landmarks=400, seed=7,
                                             step=1.0, saliency_scale=0.01)

    # ---- the unchanged script, through the launcher (one visible GPU: DataParallel then calls the module directly)
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES=os.environ.get('CUDA_VISIBLE_DEVICES', '0').split(',')[0])
    cmd = [sys.executable, '-m', 'mdgat_matcher_b200.launcher', script,
           '--train_path', dirs['train_path'], '--txt_path', dirs['txt_path'], '--keypoints_path', dirs['keypoints_path'],
           '--resume_model', ckpt, '--max_keypoints', str(n)]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=str(tmp_path), timeout=900)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    per_pair, failed = {}, set()
    for line in r.stdout.splitlines():
        m = re.match(r'idx(\d+), rep ([\d.]+), inlier (\d+), precision\(inlier ratio\) ([\d.]+), recall ([\d.]+), '
                     r'fp_rate ([\d.]+), tp_rate ([\d.]+), RTE ([\d.]+), RRE ([\d.]+)', line)
        if m:
            per_pair[int(m.group(1))] = [float(x) for x in m.groups()[1:]]
        m = re.match(r'idx(\d+), rep ([\d.]+), registration fail', line)
        if m:
            failed.add(int(m.group(1)))
    summary = [l for l in r.stdout.splitlines() if '||' in l and not l.startswith('repeat')]
    assert len(per_pair) + len(failed) == frames - 3 and len(summary) == 1, r.stdout[-3000:]
    s = [float(x) for x in re.findall(r'[-\d.]+(?:e-?\d+)?|nan', summary[0])]

    # ---- this package's own pipeline on the same pairs
    dev = torch.device('cuda:0')
    from oracle.build_ref import load_checkpoint_state_dict
    from oracle.ref_loader import net_config
    net = MDGAT(net_config(L=9, sinkhorn_iterations=20))
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in load_checkpoint_state_dict().items()})
    net = net.double().eval().to(dev)
    pb = kitti_io.PairBatcher(dirs['train_path'], dirs['txt_path'], dirs['keypoints_path'], 10, max_keypoints=n, device=dev)
    acc = {k: [] for k in ('rep', 'inlier', 'prec', 'recall', 'fp', 'tp', 'rte', 'rre', 'rr')}
    with torch.no_grad():
        for i in range(len(pb)):
            batch = pb.batch(i, 1)
            out = net(batch)
            T, st = ops.register_pairs(batch['keypoints0'], batch['keypoints1'], out['matches0'], batch['gt_matches0'], batch['T_gt'])
            g = {k: float(v[0]) for k, v in st.items()}
            rep = g['n_valid_gt'] / n
            prec = g['tp'] / g['n_valid'] if g['n_valid'] > 0 else 0.0
            recall = g['tp'] / g['n_valid_gt'] if g['n_valid'] > 0 else 0.0
            fp_rate, tp_rate = g['fp'] / (g['fp'] + g['tn']), g['tp'] / (g['tp'] + g['fn'])
            got = [rep, g['tp'], prec, recall, fp_rate, tp_rate, g['rte'], g['rre']]
            ok = g['rte'] < 2 and g['rre'] < np.pi / 180 * 5
            assert ok == (pb.pairs[i][0] in per_pair), (i, got)          # the script prints metrics only for registered pairs
            if ok:
                for a, b in zip(got, per_pair[pb.pairs[i][0]]):
                    assert abs(a - b) <= 0.00051, (i, got, per_pair[pb.pairs[i][0]])      # three printed decimals
            acc['rr'].append(1.0 if ok else 0.0)
            for k, v in zip(('rep', 'inlier', 'prec', 'recall', 'fp', 'tp'), got[:6]):
                acc[k].append(v)
            if g['rte'] < 2:
                acc['rte'].append(g['rte'])
            if g['rre'] < np.pi / 180 * 5:
                acc['rre'].append(g['rre'])
    mean = {k: float(np.mean(v)) for k, v in acc.items()}
    f1 = 2 * mean['prec'] * mean['recall'] / (mean['prec'] + mean['recall'])
    # summary: repeatibility, inlier, RR || precision, recall, F1 || fp_rate, tp_rate || RTE, RRE
    mine = [mean['rep'], mean['inlier'], mean['rr'], mean['prec'], mean['recall'], f1, mean['fp'], mean['tp'], mean['rte'], mean['rre']]
    assert len(s) == len(mine), (summary, s)
    for a, b, tol in zip(mine, s, [6e-4, 6e-2, 6e-4, 6e-4, 6e-4, 6e-4, 6e-4, 6e-4, 6e-4, 6e-4]):
        assert abs(a - b) <= tol, (mine, s)
    assert mean['rr'] > 0.5 and mean['prec'] > 0.5, mean                  # the matcher actually registers the synthetic pairs


def test_train_script_runs_unchanged_for_one_epoch(tmp_path):
    """The reference's UNCHANGED train.py (/root/reference/train.py:125-300) for one epoch through the launcher on synthetic
    KITTI-format sequences (train split 00, 02-07, validation split 09): the training steps run the drop-in's differentiable
    path (attention and Sinkhorn forward + hand-written backward on the CUDA kernels, gap_loss), the validation pass its CUDA
    inference path (gap_loss on the device), and the checkpoint the script saves loads back into the drop-in."""
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    from mdgat_matcher_b200 import kitti_io
    from mdgat_matcher_b200.models.mdgat import MDGAT
    from oracle.ref_loader import net_config
    script = _script('train.py')
    n = 128
    for seq in (0, 2, 3, 4, 5, 6, 7, 9):
        # saliency 32 +- 10 as in the real keypoint files: train.py's loader keeps saliency > 10 only (ensure_kpts_num,
        # load_data.py:180-211) and loops forever on a pair it leaves without keypoints
        dirs = kitti_io.write_synthetic_sequence(str(tmp_path / 'KITTI'), seq=seq, frames=6, n_kpts=n, n_landmarks=200, seed=20 + seq,
                                                 step=1.0, saliency_scale=1.0)
    env = dict(os.environ, PYTHONPATH=ROOT, CUDA_VISIBLE_DEVICES=os.environ.get('CUDA_VISIBLE_DEVICES', '0').split(',')[0])
    cmd = [sys.executable, '-m', 'mdgat_matcher_b200.launcher', script,
           '--train_path', dirs['train_path'], '--txt_path', dirs['txt_path'], '--keypoints_path', dirs['keypoints_path'],
           '--max_keypoints', str(n), '--batch_size', '2', '--epoch', '1', '--l', '9', '--sinkhorn_iterations', '20']
    r = subprocess.run(cmd, capture_output=True, text=True, env=env, cwd=str(tmp_path), timeout=600)
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    m = re.search(r'Validation loss: ([-\d.naninf]+), epoch_loss: ([-\d.naninf]+)', r.stdout)
    assert m, r.stdout[-2000:]
    val_loss, epoch_loss = float(m.group(1)), float(m.group(2))
    assert np.isfinite(val_loss) and np.isfinite(epoch_loss) and epoch_loss > 0
    saved = [os.path.join(dp, f) for dp, _, fs in os.walk(str(tmp_path / 'checkpoint')) for f in fs if f.endswith('.pth')]
    assert len(saved) == 1, saved
    ck = torch.load(saved[0], map_location='cpu', weights_only=False)
    net = torch.nn.DataParallel(MDGAT(net_config(L=9, sinkhorn_iterations=20, loss_method='gap_loss')))
    net.load_state_dict(ck['net'])                                           # strict: the drop-in's parameter names are the reference's
    assert ck['epoch'] == 1 and abs(ck['loss'] - val_loss) < 1e-3
