"""KITTI keypoint files and batches for the matcher -- the data format on the input side of the
hot path.

File format (/root/reference/load_data.py:146-165): `<keypoints_path>/<seq>/<%06d>.bin`, float32,
n x 37 = [x y z | saliency | 33-bin FPFH]. Pair lists: `<txt_path>/<seq>/groundtruths.txt` (a header
line, then `anc_idx pos_idx ...`, load_data.py:9-29); poses `<train_path>/poses/<seq>.txt` (12 floats
per line) and `<train_path>/calib/sequences/<seq>/calib.txt` (last line = Tr, load_data.py:72-91).

`write_synthetic_sequence` fabricates a sequence in exactly that layout from a random landmark
field (the real keypoint files are a separate download), so that the reference's
unchanged scripts -- and `PairBatcher` below -- have something to read. `PairBatcher` is the
batched, device-side replacement of SparseDataset.__getitem__: file reads on the host, world
transform / ground-truth matching / T_gt on the GPU (`ops.prepare_pairs`).
"""
import os

import numpy as np
import torch

from . import synth

RECORD = 37


def read_keypoint_bin(path):
    a = np.fromfile(path, dtype=np.float32).reshape(-1, RECORD)
    return a[:, :3], a[:, 3], a[:, 4:]


def write_keypoint_bin(path, kp, score, desc):
    rec = np.concatenate([kp, score[:, None], desc], axis=1).astype(np.float32)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    rec.tofile(path)


def normalize_descriptors(desc):
    """load_data.py:290-292: float32 row norms, multiply by the float32 reciprocal, then to double."""
    desc = np.asarray(desc, dtype=np.float32)
    norm = np.linalg.norm(desc, axis=1).reshape(-1, 1)
    return np.multiply(desc, 1 / norm).astype(np.float64)


def fix_keypoint_count(kp, score, desc, n):
    """ensure_kpts_num (load_data.py:180-211): drop keypoints with saliency <= 10 (sic), truncate to
    n, or pad by stacking leading keypoints in front until there are n (this creates exact duplicates)."""
    valid = score > 10
    kp, score, desc = kp[valid], score[valid], desc[valid]
    if n < len(kp):
        return kp[:n], score[:n], desc[:n]
    while n > len(kp):
        k = n - len(kp)
        kp, score, desc = np.vstack((kp[:k], kp)), np.hstack((score[:k], score)), np.vstack((desc[:k], desc))
    return kp, score, desc


def read_pairs(txt_path, seq):
    pairs = []
    with open(os.path.join(txt_path, '%02d' % seq, 'groundtruths.txt')) as f:
        for i, line in enumerate(f):
            if i == 0 or not line.strip():
                continue
            t = line.split()
            pairs.append((int(t[0]), int(t[1])))
    return pairs


def read_poses(train_path, seq):
    poses = []
    with open(os.path.join(train_path, 'poses', '%02d.txt' % seq)) as f:
        for line in f:
            if line.strip():
                poses.append(np.vstack((np.array(line.split(), dtype=np.float64).reshape(3, 4), [0, 0, 0, 1])))
    return poses


def read_calib(train_path, seq):
    calib = None
    with open(os.path.join(train_path, 'calib', 'sequences', '%02d' % seq, 'calib.txt')) as f:
        for line in f:
            if ':' in line:
                vals = line.split(':', 1)[1].split()
                if len(vals) == 12:
                    calib = np.vstack((np.array(vals, dtype=np.float64).reshape(3, 4), [0, 0, 0, 1]))
    return calib                       # the last 12-value line (Tr), as the reference's loop leaves it


def write_synthetic_sequence(root, seq=10, frames=12, n_kpts=256, n_landmarks=3000, seed=0, step=2.0, saliency_scale=1.0):
    """Creates <root>/{poses,calib,preprocess,keypoints}: a straight-ish drive past a static landmark
    field; every frame sees the n_kpts nearest landmarks in its own LiDAR frame with small noise, so
    consecutive frames share most keypoints. Returns the directory names the scripts' flags expect.

    Saliency is written as 32 +- 10 (times saliency_scale): the loader's optional keypoint-count fix-up keeps only
    saliency > 10 (load_data.py:180-211). The pre-trained matcher saw saliencies around 0.32 and keypoint clouds about
    18 m wide: with saliency_scale = 0.01, a sparse field (n_landmarks ~ 1.5 n_kpts) and
    step <= 1 m it registers these synthetic pairs (>= 95 % of its matches correct); with the dense default field the
    clouds are a few metres wide, far from its training distribution, and it finds no correct match."""
    rng = np.random.default_rng(seed)
    g = torch.Generator().manual_seed(seed)
    calib = np.array([[4.276802385584e-04, -9.999672484946e-01, -8.084491683471e-03, -1.198459927713e-02],
                      [-7.210626507497e-03, 8.081198471645e-03, -9.999413164504e-01, -5.403984729748e-02],
                      [9.999738645903e-01, 4.859485810390e-04, -7.206933692422e-03, -2.921968648686e-01],
                      [0, 0, 0, 1.0]])
    poses = []
    for f in range(frames):
        yaw = 0.02 * f
        T = np.eye(4)
        T[:3, :3] = [[np.cos(yaw), 0, np.sin(yaw)], [0, 1, 0], [-np.sin(yaw), 0, np.cos(yaw)]]
        T[:3, 3] = [0.1 * f, 0.0, step * f]
        poses.append(T)
    # landmarks in cam0-world coordinates around the path; descriptors / saliency persistent per landmark
    lm = np.stack([rng.normal(0, 18, n_landmarks), rng.normal(0, 0.7, n_landmarks),
                   rng.uniform(-15, step * frames + 15, n_landmarks)], axis=1)
    desc = synth._desc(g, n_landmarks).numpy() * 100.0              # raw FPFH bins are unnormalised counts
    score = np.clip(32 + 10 * rng.normal(size=n_landmarks), 11, 100)  # the loader keeps saliency > 10 only
    kp_dir = os.path.join(root, 'keypoints')
    for f, T in enumerate(poses):
        to_velo = np.linalg.inv(T @ calib)
        local = (to_velo @ np.concatenate([lm, np.ones((n_landmarks, 1))], 1).T).T[:, :3]
        near = np.argsort(np.linalg.norm(local, axis=1))[:n_kpts]
        near = near[rng.permutation(len(near))]
        write_keypoint_bin(os.path.join(kp_dir, '%02d' % seq, '%06d.bin' % f),
                           local[near] + rng.normal(0, 0.03, (len(near), 3)), score[near] * saliency_scale,
                           desc[near] * (1 + 0.05 * rng.normal(size=(len(near), 33))))
    os.makedirs(os.path.join(root, 'poses'), exist_ok=True)
    with open(os.path.join(root, 'poses', '%02d.txt' % seq), 'w') as fh:
        for T in poses:
            fh.write(' '.join('%.12e' % v for v in T[:3].reshape(-1)) + '\n')
    cdir = os.path.join(root, 'calib', 'sequences', '%02d' % seq)
    os.makedirs(cdir, exist_ok=True)
    with open(os.path.join(cdir, 'calib.txt'), 'w') as fh:
        for name in ('P0', 'P1', 'P2', 'P3'):
            fh.write(name + ': ' + ' '.join(['7.0e+02', '0', '6.0e+02', '0', '0', '7.0e+02', '1.8e+02', '0', '0', '0', '1', '0']) + '\n')
        fh.write('Tr: ' + ' '.join('%.12e' % v for v in calib[:3].reshape(-1)) + '\n')
    tdir = os.path.join(root, 'preprocess', '%02d' % seq)
    os.makedirs(tdir, exist_ok=True)
    with open(os.path.join(tdir, 'groundtruths.txt'), 'w') as fh:
        fh.write('idx1 idx2 t1 t2 t3 q1 q2 q3 q4\n')
        for f in range(frames - 3):
            fh.write('%d %d 0 0 0 0 0 0 1\n' % (f, f + 3))
    return {'train_path': root, 'txt_path': os.path.join(root, 'preprocess'), 'keypoints_path': kp_dir}


class PairBatcher:
    """Batched replacement of SparseDataset.__getitem__ + DataLoader collation for one sequence:
    yields dicts in the layout MDGAT.forward expects, ground truth computed on the device."""

    def __init__(self, train_path, txt_path, keypoints_path, seq, max_keypoints=256, threshold=0.5,
                 mutual_check=False, ensure_kpts_num=False, device='cuda'):
        self.kp_path, self.seq, self.n = keypoints_path, seq, max_keypoints
        self.threshold, self.mutual, self.ensure = threshold, mutual_check, ensure_kpts_num
        self.pairs = read_pairs(txt_path, seq)
        self.poses = read_poses(train_path, seq)
        self.calib = read_calib(train_path, seq)
        self.device = torch.device(device)

    def __len__(self):
        return len(self.pairs)

    def _frame(self, idx):
        kp, score, desc = read_keypoint_bin(os.path.join(self.kp_path, '%02d' % self.seq, '%06d.bin' % idx))
        if self.ensure:
            kp, score, desc = fix_keypoint_count(kp, score, desc, self.n)
        return kp.astype(np.float64), score.astype(np.float64), normalize_descriptors(desc)

    def batch(self, start, size):
        """Pairs [start, start+size) as one batch (all frames must hold the same number of keypoints,
        which `ensure_kpts_num` guarantees)."""
        from . import ops
        sel = self.pairs[start:start + size]
        f0 = [self._frame(a) for a, _ in sel]
        f1 = [self._frame(b) for _, b in sel]
        dev = self.device
        def stack(items, i):
            return torch.from_numpy(np.stack([it[i] for it in items])).to(dev)
        out = {'keypoints0': stack(f0, 0), 'keypoints1': stack(f1, 0), 'scores0': stack(f0, 1), 'scores1': stack(f1, 1),
               'descriptors0': stack(f0, 2), 'descriptors1': stack(f1, 2)}
        p0 = torch.from_numpy(np.stack([self.poses[a] for a, _ in sel])).to(dev)
        p1 = torch.from_numpy(np.stack([self.poses[b] for _, b in sel])).to(dev)
        m0, m1, T_gt, rep = ops.prepare_pairs(out['keypoints0'], out['keypoints1'], p0, p1,
                                              torch.from_numpy(self.calib).to(dev), self.threshold, self.mutual)
        out.update(gt_matches0=m0, gt_matches1=m1, T_gt=T_gt, rep=rep,
                   sequence=['%02d' % self.seq] * len(sel), idx0=[a for a, _ in sel])
        return out
