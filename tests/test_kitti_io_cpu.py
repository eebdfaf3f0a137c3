"""Host side of the input format: keypoint .bin round trip, the loader's keypoint-count fix-up and
descriptor normalisation, and the synthetic sequence in the reference's directory layout."""
import os

import numpy as np

from mdgat_matcher_b200 import kitti_io


def test_bin_round_trip_and_synthetic_sequence(tmp_path):
    dirs = kitti_io.write_synthetic_sequence(str(tmp_path), seq=10, frames=6, n_kpts=64, n_landmarks=400, seed=1)
    kp, score, desc = kitti_io.read_keypoint_bin(os.path.join(dirs['keypoints_path'], '10', '000002.bin'))
    assert kp.shape == (64, 3) and score.shape == (64,) and desc.shape == (64, 33) and kp.dtype == np.float32
    assert (score > 10).all()
    pairs = kitti_io.read_pairs(dirs['txt_path'], 10)
    assert pairs == [(0, 3), (1, 4), (2, 5)]
    poses = kitti_io.read_poses(dirs['train_path'], 10)
    calib = kitti_io.read_calib(dirs['train_path'], 10)
    assert len(poses) == 6 and poses[0].shape == (4, 4) and calib.shape == (4, 4) and abs(calib[0, 1] + 0.99996) < 1e-4
    # the reference's own parsers read the same files (load_data.py:9-29 logic restated in the oracle is not needed:
    # the format is plain text); consecutive frames overlap: most keypoints of frame 0 reappear in frame 3
    from oracle import mdgat_oracle as O
    k0 = kitti_io.read_keypoint_bin(os.path.join(dirs['keypoints_path'], '10', '000000.bin'))[0].astype(np.float64)
    k3 = kitti_io.read_keypoint_bin(os.path.join(dirs['keypoints_path'], '10', '000003.bin'))[0].astype(np.float64)
    m1, m2, T_gt, rep = O.prepare_pair(k0, k3, poses[0], poses[3], calib, 0.5)
    assert rep > 20 and (m1 >= 0).sum() == rep


def test_keypoint_count_fix_and_descriptor_norm():
    rng = np.random.default_rng(0)
    kp = rng.normal(size=(10, 3)).astype(np.float32)
    score = np.array([50, 5, 40, 30, 9, 60, 70, 80, 20, 15], dtype=np.float32)      # two are <= 10 and dropped
    desc = np.abs(rng.normal(size=(10, 33))).astype(np.float32)
    k, s, d = kitti_io.fix_keypoint_count(kp, score, desc, 6)
    assert len(k) == 6 and np.array_equal(s, score[score > 10][:6])
    k, s, d = kitti_io.fix_keypoint_count(kp, score, desc, 13)                    # pad by duplication in front
    assert len(k) == 13 and np.array_equal(k[5:], kp[score > 10]) and np.array_equal(k[:5], kp[score > 10][:5])
    n = kitti_io.normalize_descriptors(desc)
    assert n.dtype == np.float64 and np.abs(np.linalg.norm(n, axis=1) - 1).max() < 1e-6


def test_reference_loader_reads_synthetic_sequence_and_agrees(tmp_path):
    """With the real reference present (build container): its unmodified SparseDataset reads the
    synthetic sequence, and the oracle's pair preparation + this package's descriptor normalisation
    reproduce its ground truth, T_gt, repeatability and descriptors item by item."""
    import sys
    import types
    import pytest
    if not os.path.isfile('/root/reference/load_data.py'):
        pytest.skip('reference tree not present')
    from mdgat_matcher_b200 import launcher
    from oracle import mdgat_oracle as O
    launcher.install_optional_stubs()
    sys.dont_write_bytecode = True
    sys.path.insert(0, '/root/reference')
    try:
        import importlib
        load_data = importlib.import_module('load_data')
    finally:
        sys.path.remove('/root/reference')
    dirs = kitti_io.write_synthetic_sequence(str(tmp_path), seq=10, frames=6, n_kpts=96, n_landmarks=500, seed=3)
    for mutual in (False, True):
        opt = types.SimpleNamespace(train_path=dirs['train_path'], keypoints='USIP', keypoints_path=dirs['keypoints_path'],
                                    descriptor='FPFH', max_keypoints=96, threshold=0.5, ensure_kpts_num=False,
                                    mutual_check=mutual, memory_is_enough=False, txt_path=dirs['txt_path'])
        ds = load_data.SparseDataset(opt, 'test')
        assert len(ds) == 3
        poses, calib = kitti_io.read_poses(dirs['train_path'], 10), kitti_io.read_calib(dirs['train_path'], 10)
        for i, (a, b) in enumerate(kitti_io.read_pairs(dirs['txt_path'], 10)):
            item = ds[i]
            k0, s0, d0 = kitti_io.read_keypoint_bin(os.path.join(dirs['keypoints_path'], '10', '%06d.bin' % a))
            k1, s1, d1 = kitti_io.read_keypoint_bin(os.path.join(dirs['keypoints_path'], '10', '%06d.bin' % b))
            m1, m2, T_gt, rep = O.prepare_pair(k0.astype(np.float64), k1.astype(np.float64), poses[a], poses[b], calib, 0.5, mutual)
            assert np.array_equal(item['gt_matches0'], m1) and np.array_equal(item['gt_matches1'], m2)
            assert np.abs(item['T_gt'].numpy() - T_gt).max() < 1e-9 and item['rep'] == rep
            assert np.array_equal(item['descriptors0'].numpy(), kitti_io.normalize_descriptors(d0))
            assert np.array_equal(item['keypoints0'].numpy(), k0.astype(np.float64))
            assert np.array_equal(item['scores1'].numpy(), s1.astype(np.float64))
