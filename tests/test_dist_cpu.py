"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the batch sharding and the
one all-gather of results (mdgat_matcher_b200/dist.py)."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from mdgat_matcher_b200 import dist as D
    D.init_process_group('gloo')
    n = 7                                              # ragged: 4 + 3
    # 'calib' has as many rows as there are pairs but is shared, not per pair: it must not be sliced
    full = {'keypoints0': torch.arange(n * 5 * 3, dtype=torch.float64).reshape(n, 5, 3), 'tag': 'x', 'calib': torch.eye(n)}
    mine = D.shard_batch(full)
    lo, hi = D.shard_bounds(n, rank, world)
    assert mine['keypoints0'].shape[0] == hi - lo and mine['tag'] == 'x' and mine['calib'].shape == (n, n)
    out = {'matches0': (mine['keypoints0'][:, :, 0] * 2).long(), 'matching_scores0': mine['keypoints0'][:, :, 1]}
    g = D.all_gather_outputs(out, keys=('matches0', 'matching_scores0'), n_total=n)
    loss = D.all_reduce_mean_loss(torch.tensor(float(rank + 1), dtype=torch.float64), hi - lo)
    q.put((rank, g['matches0'], g['matching_scores0'], float(loss)))
    torch.distributed.destroy_process_group()


def test_shard_and_all_gather_world2():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = torch.arange(7 * 5 * 3, dtype=torch.float64).reshape(7, 5, 3)
    for rank, m0, s0, loss in res:
        assert torch.equal(m0, (full[:, :, 0] * 2).long())
        assert torch.equal(s0, full[:, :, 1])
        assert abs(loss - (1 * 4 + 2 * 3) / 7) < 1e-12


def test_shard_bounds_cover_everything():
    from mdgat_matcher_b200.dist import shard_bounds
    for n in (0, 1, 7, 32, 1199):
        for w in (1, 2, 4, 8):
            spans = [shard_bounds(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(w - 1))
