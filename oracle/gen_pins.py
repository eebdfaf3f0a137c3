"""Generates tests/golden/pins_knn_registration.npz by running the UNMODIFIED reference functions that sit beside the
forward on the hot path's edges (build container only, needs /root/reference):

* models/mdgat.py:8-32   knn(), get_graph_feature()   (dead code upstream, named by the north star)
* utils/utils_test.py:27-39, 73-110   calculate_error2(), solve_icp()   (registration metrics)

The fixture pins oracle/mdgat_oracle.py's knn / get_graph_feature / solve_icp / registration_error (tests/test_oracle_golden.py)
and, through them and directly, the CUDA kernels mdgat_knn and mdgat_register_pairs (tests/test_gpu_parity.py).

    python oracle/gen_pins.py
"""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_loader as RL                   # noqa: E402
from mdgat_matcher_b200 import launcher               # noqa: E402

OUT = os.path.join(ROOT, 'tests', 'golden', 'pins_knn_registration.npz')


def main():
    mod = RL.load_reference_module('mdgat', 'cpu')              # unmodified file; the shim only redirects torch.device('cuda')
    launcher.install_optional_stubs()                            # utils_test.py imports open3d for its plotting helpers only
    sys.dont_write_bytecode = True
    sys.path.insert(0, RL.REFERENCE_ROOT)
    try:
        utils_test = importlib.import_module('utils.utils_test')
    finally:
        sys.path.remove(RL.REFERENCE_ROOT)
        for k in [k for k in sys.modules if k == 'utils' or k.startswith('utils.')]:
            sys.modules.pop(k)
    rec = {'torch_version': torch.__version__}
    g = torch.Generator().manual_seed(20)
    # ---- knn / get_graph_feature: (B,3,n) queries against (B,3,m) sources, several k
    for i, (B, n, m, k) in enumerate([(2, 96, 128, 8), (1, 200, 77, 16), (3, 64, 64, 64), (1, 33, 2048, 20)]):
        x = torch.randn(B, 3, n, generator=g, dtype=torch.float64) * torch.tensor([18.6, 12.1, 0.72], dtype=torch.float64)[None, :, None]
        src = torch.randn(B, 3, m, generator=g, dtype=torch.float64) * torch.tensor([18.6, 12.1, 0.72], dtype=torch.float64)[None, :, None]
        idx = mod.knn(x, src, k)
        adj = mod.get_graph_feature(x, src, k)
        rec['knn%d_x' % i], rec['knn%d_src' % i], rec['knn%d_k' % i] = x.numpy(), src.numpy(), np.asarray(k)
        rec['knn%d_idx' % i], rec['knn%d_adj' % i] = idx.numpy(), adj.numpy().astype(np.int8)
    # ---- solve_icp / calculate_error2: matched keypoint sets under a known rigid transform + noise, incl. a reflective
    # degenerate (planar) case where the reference's missing det(R) fix shows
    rng = np.random.default_rng(21)
    for i, (npts, noise, planar) in enumerate([(60, 0.05, False), (12, 0.3, False), (200, 0.0, False), (5, 0.02, True), (3, 0.5, False)]):
        P = rng.normal(size=(npts, 3)) * [18.6, 12.1, 0.0 if planar else 0.72]
        yaw, pitch = rng.uniform(-0.3, 0.3), rng.uniform(-0.05, 0.05)
        Rz = np.array([[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]])
        Ry = np.array([[np.cos(pitch), 0, np.sin(pitch)], [0, 1, 0], [-np.sin(pitch), 0, np.cos(pitch)]])
        T_gt = np.eye(4)
        T_gt[:3, :3] = Rz @ Ry
        T_gt[:3, 3] = rng.normal(size=3) * [3.0, 0.5, 0.05]
        # mkpts0 = T_gt applied to mkpts1 (+ noise): calculate_error2 solves mkpts1 -> mkpts0
        mk1 = P
        mk0 = (T_gt[:3, :3] @ P.T).T + T_gt[:3, 3] + noise * rng.normal(size=(npts, 3))
        T, rte, rre = utils_test.calculate_error2(mk0, mk1, 0, torch.from_numpy(T_gt))
        Ticp = utils_test.solve_icp(mk1, mk0)
        rec['reg%d_mkpts0' % i], rec['reg%d_mkpts1' % i], rec['reg%d_T_gt' % i] = mk0, mk1, T_gt
        rec['reg%d_T' % i], rec['reg%d_rte' % i], rec['reg%d_rre' % i] = T.numpy(), np.asarray(rte), np.asarray(rre)
        rec['reg%d_Ticp' % i] = Ticp
        print('reg%d: n=%d rte=%.3e rre=%.3e det(R)=%.3f' % (i, npts, rte, rre, np.linalg.det(Ticp[:3, :3])))
    np.savez_compressed(OUT, **rec)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
