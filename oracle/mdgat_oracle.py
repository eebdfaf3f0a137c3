"""TEST INFRASTRUCTURE ONLY -- CPU (numpy, float64) restatement of the reference hot path.

Parity pin: every function here is checked against outputs of the *unmodified* reference
(/root/reference/models/mdgat.py run through oracle/ref_loader.py in the build container);
the generated vectors live in tests/golden/ together with oracle/gen_golden.py, and
tests/test_oracle_golden.py re-checks the oracle against them on every run.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; it is the checker, never the product path. The arithmetic the reference
delegates to PyTorch (einsum, softmax, topk, logsumexp, Conv1d(k=1), BatchNorm1d; torch is
unpinned upstream, 2.11.0 was used to generate the golden vectors) is restated with numpy.

Layout follows the reference: activations are channel-major (B, C, N).
All line numbers cite /root/reference/models/mdgat.py unless stated otherwise.
"""
import math
from concurrent.futures import ThreadPoolExecutor

import numpy as np

BN_EPS = 1e-5          # torch.nn.BatchNorm1d default, used by MLP() at :43
NUM_HEADS = 4          # AttentionalPropagation(feature_dim, 4) at :255


# --------------------------------------------------------------------------- blocks

def conv1x1(x, w, b):
    """nn.Conv1d(kernel_size=1) (:40): x (B,Cin,N), w (Cout,Cin[,1]), b (Cout,) -> (B,Cout,N)."""
    w = w.reshape(w.shape[0], w.shape[1])
    return np.matmul(w[None], x) + b[None, :, None]


def batchnorm_eval(x, gamma, beta, mean, var):
    """nn.BatchNorm1d in eval mode (:43): per-channel affine from the running statistics."""
    inv = 1.0 / np.sqrt(var + BN_EPS)
    return (x - mean[None, :, None]) * inv[None, :, None] * gamma[None, :, None] + beta[None, :, None]


def mlp(sd, prefix, x, n_conv):
    """MLP() (:34-46): conv, then BN + ReLU after every conv but the last.
    nn.Sequential indices: conv i sits at 3*i, its BN at 3*i+1."""
    for i in range(n_conv):
        p = '%s.%d' % (prefix, 3 * i)
        x = conv1x1(x, sd[p + '.weight'], sd[p + '.bias'])
        if i < n_conv - 1:
            q = '%s.%d' % (prefix, 3 * i + 1)
            x = batchnorm_eval(x, sd[q + '.weight'], sd[q + '.bias'],
                               sd[q + '.running_mean'], sd[q + '.running_var'])
            x = np.maximum(x, 0.0)
    return x


def keypoint_encoder(sd, kpts, scores):
    """KeypointEncoder.forward (:184-188): cat[kpts^T (3), score (1)] -> MLP 4-32-64-128-128."""
    inp = np.concatenate([kpts.transpose(0, 2, 1), scores[:, None, :]], axis=1)
    return mlp(sd, 'kenc.encoder', inp, 4)


def descriptor_encoder(sd, desc):
    """DescriptorEncoder.forward (:152-155): desc^T -> MLP 33-64-128-128."""
    return mlp(sd, 'denc.encoder', desc.transpose(0, 2, 1), 3)


def _softmax_last(x):
    m = x.max(axis=-1, keepdims=True)
    e = np.exp(x - m)
    return e / e.sum(axis=-1, keepdims=True)


def attention(q, k, v):
    """attention() (:190-194). q (B,d,H,N), k/v (B,d,H,M) -> (B,d,H,N), prob (B,H,N,M)."""
    dim = q.shape[1]
    scores = np.einsum('bdhn,bdhm->bhnm', q, k, optimize=True) / dim ** .5
    prob = _softmax_last(scores)
    return np.einsum('bhnm,bdhm->bdhn', prob, v, optimize=True), prob


def topk_indices(scores, k):
    """scores.topk(k, dim=3) (:202). Exactly k entries per row; ties are broken towards the
    lowest index (torch leaves the choice implementation-defined; the result of the layer
    does not depend on it when tied columns carry identical key/value vectors)."""
    if k > scores.shape[-1]:
        raise RuntimeError('selected index k out of range')      # what torch.topk raises
    order = np.argsort(-scores, axis=-1, kind='stable')
    return order[..., :k]


def dynamic_attention(q, k, v, topk):
    """dynamic_attention() (:196-210): dense logits, keep the top-k per row, softmax over the
    kept k, scatter into a zero (B,H,N,M) matrix, contract with v."""
    dim = q.shape[1]
    scores = np.einsum('bdhn,bdhm->bhnm', q, k, optimize=True) / dim ** .5
    idx = topk_indices(scores, topk)
    kept = np.take_along_axis(scores, idx, axis=-1)
    s = _softmax_last(kept)
    prob = np.zeros_like(scores)
    np.put_along_axis(prob, idx, s, axis=-1)
    return np.einsum('bhnm,bdhm->bdhn', prob, v, optimize=True), prob


def multi_headed_attention(sd, prefix, x, source, topk):
    """MultiHeadedAttention.forward (:223-237). proj.0/1/2 -> q/k/v, .view(B, 32, 4, N)
    (channel c = d*4 + h), attention or dynamic_attention, merge conv."""
    b, c, _ = x.shape
    d = c // NUM_HEADS
    q = conv1x1(x, sd[prefix + '.proj.0.weight'], sd[prefix + '.proj.0.bias']).reshape(b, d, NUM_HEADS, -1)
    k = conv1x1(source, sd[prefix + '.proj.1.weight'], sd[prefix + '.proj.1.bias']).reshape(b, d, NUM_HEADS, -1)
    v = conv1x1(source, sd[prefix + '.proj.2.weight'], sd[prefix + '.proj.2.bias']).reshape(b, d, NUM_HEADS, -1)
    if topk is None:
        out, prob = attention(q, k, v)
    else:
        out, prob = dynamic_attention(q, k, v, topk)
    out = np.ascontiguousarray(out).reshape(b, c, -1)
    return conv1x1(out, sd[prefix + '.merge.weight'], sd[prefix + '.merge.bias']), prob


def attentional_propagation(sd, prefix, x, source, topk):
    """AttentionalPropagation.forward (:246-248): mlp(cat[x, attn(x, source)]), 256-256-128."""
    msg, prob = multi_headed_attention(sd, prefix + '.attn', x, source, topk)
    return mlp(sd, prefix + '.mlp', np.concatenate([x, msg], axis=1), 2), prob


def layer_topk(i, k_list, L):
    """The k schedule of AttentionalGNN.forward (:268-272)."""
    if i > 2 * L - 1 - len(k_list):
        return k_list[i - 2 * L + len(k_list)]
    return None


def attentional_gnn(sd, desc0, desc1, k_list, L, trace=None):
    """AttentionalGNN.forward (:259-276); layer names are ['self','cross']*L (:353)."""
    for i in range(2 * L):
        cross = (i % 2 == 1)
        src0, src1 = (desc1, desc0) if cross else (desc0, desc1)
        topk = layer_topk(i, k_list, L)
        p = 'gnn.layers.%d' % i
        delta0, _ = attentional_propagation(sd, p, desc0, src0, topk)
        delta1, _ = attentional_propagation(sd, p, desc1, src1, topk)
        desc0, desc1 = desc0 + delta0, desc1 + delta1
        if trace is not None:
            trace.append((desc0.copy(), desc1.copy()))
    return desc0, desc1


def logsumexp(x, axis):
    """torch.logsumexp as used at :283-284."""
    m = x.max(axis=axis, keepdims=True)
    m = np.where(np.isfinite(m), m, 0.0)
    return np.squeeze(m, axis=axis) + np.log(np.exp(x - m).sum(axis=axis))


def log_sinkhorn_iterations(Z, log_mu, log_nu, iters):
    """log_sinkhorn_iterations() (:279-285)."""
    u, v = np.zeros_like(log_mu), np.zeros_like(log_nu)
    for _ in range(iters):
        u = log_mu - logsumexp(Z + v[:, None, :], axis=2)
        v = log_nu - logsumexp(Z + u[:, :, None], axis=1)
    return Z + u[:, :, None] + v[:, None, :]


def log_optimal_transport(scores, alpha, iters):
    """log_optimal_transport() (:288-308): dustbin row/column/corner = alpha, marginals
    log_mu = [norm]*m + [log n + norm], norm = -log(m+n); returns Z - norm."""
    b, m, n = scores.shape
    couplings = np.full((b, m + 1, n + 1), float(alpha), dtype=np.float64)
    couplings[:, :m, :n] = scores
    norm = -math.log(m + n)
    log_mu = np.concatenate([np.full(m, norm), [math.log(n) + norm]])
    log_nu = np.concatenate([np.full(n, norm), [math.log(m) + norm]])
    log_mu = np.broadcast_to(log_mu[None], (b, m + 1)).copy()
    log_nu = np.broadcast_to(log_nu[None], (b, n + 1)).copy()
    Z = log_sinkhorn_iterations(couplings, log_mu, log_nu, iters)
    return Z - norm


def extract_matches(Z, loss_method='triplet_loss', mutual_check=False, match_threshold=0.2):
    """Match extraction (:442-483). Returns matches0/1 (int64, -1 invalid) and
    matching_scores0/1 (float64; int64 zeros when nothing is valid, as :465-467 does)."""
    if loss_method == 'superglue':
        inner = Z[:, :-1, :-1]
        idx0, idx1 = inner.argmax(2), inner.argmax(1)
        max0, max1 = inner.max(2), inner.max(1)
        if mutual_check:
            ar0 = np.arange(idx0.shape[1])[None]
            ar1 = np.arange(idx1.shape[1])[None]
            mutual0 = ar0 == np.take_along_axis(idx1, idx0, 1)
            mutual1 = ar1 == np.take_along_axis(idx0, idx1, 1)
            ms0 = np.where(mutual0, np.exp(max0), 0.0)
            ms1 = np.where(mutual1, np.take_along_axis(ms0, idx1, 1), 0.0)
            valid0 = mutual0 & (ms0 > match_threshold)
            valid1 = mutual1 & np.take_along_axis(valid0, idx1, 1)
        else:
            valid0 = np.exp(max0) > match_threshold
            valid1 = np.exp(max1) > match_threshold
            ms0 = np.where(valid0, np.exp(max0), 0.0)
            ms1 = np.where(valid1, np.exp(max1), 0.0)
    else:
        r, c = Z[:, :-1, :], Z[:, :, :-1]
        idx0, idx1 = r.argmax(2), c.argmax(1)
        max0, max1 = r.max(2), c.max(1)
        valid0, valid1 = idx0 < (Z.shape[2] - 1), idx1 < (Z.shape[1] - 1)
        if valid0.sum() == 0:
            ms0 = np.zeros_like(idx0)
            ms1 = np.zeros_like(idx1)
        elif mutual_check:
            batch = idx0.shape[0]
            ar0 = np.broadcast_to(np.arange(idx0.shape[1], dtype=np.float64)[None], idx0.shape)
            ar1 = np.broadcast_to(np.arange(idx1.shape[1], dtype=np.float64)[None], idx1.shape)
            # :471-472 -- .view(batch,-1) needs the same number of valid entries per element
            a0 = ar0[valid0].reshape(batch, -1) == np.take_along_axis(idx1, idx0[valid0].reshape(batch, -1), 1)
            a1 = ar1[valid1].reshape(batch, -1) == np.take_along_axis(idx0, idx1[valid1].reshape(batch, -1), 1)
            mutual0 = np.zeros(idx0.shape, dtype=bool)
            mutual1 = np.zeros(idx1.shape, dtype=bool)
            mutual0[valid0] = a0.reshape(-1)
            mutual1[valid1] = a1.reshape(-1)
            ms0 = np.where(mutual0, np.exp(max0), 0.0)
            ms1 = np.where(mutual1, np.exp(max1), 0.0)
        else:
            ms0 = np.where(valid0, np.exp(max0), 0.0)
            ms1 = np.where(valid1, np.exp(max1), 0.0)
    m0 = np.where(valid0, idx0, -1).astype(np.int64)
    m1 = np.where(valid1, idx1, -1).astype(np.int64)
    return m0, m1, ms0, ms1


def _top2_indices(x, axis):
    order = np.argsort(-x, axis=axis, kind='stable')
    return np.take(order, [0, 1], axis=axis)


def triplet_loss(Z, gt0, gt1, gamma):
    """triplet_loss branch (:512-546). gt arrays use m / n for "no match" (after :519-520)."""
    b, n = gt0.shape
    m = gt1.shape[1]
    max0 = _top2_indices(Z[:, :-1, :], 2)            # (b, n, 2)
    max1 = _top2_indices(Z[:, :, :-1], 1)            # (b, 2, m)
    bi = np.arange(b)[:, None]
    ii = np.arange(n)[None]
    neg = (max0[:, :, 0] == gt0).astype(np.int64)
    neg_idx = max0[bi, ii, neg]
    anc_neg = Z[bi, ii, neg_idx]
    anc_pos = Z[bi, ii, gt0]
    jj = np.arange(m)[None]
    neg = (max1[:, 0, :] == gt1).astype(np.int64)
    neg_idx = max1[bi, neg, jj]
    anc_neg = np.concatenate([anc_neg, Z[bi, neg_idx, jj]], axis=1)
    anc_pos = np.concatenate([anc_pos, Z[bi, gt1, jj]], axis=1)
    anc_neg = -np.log(np.exp(anc_neg))
    anc_pos = -np.log(np.exp(anc_pos))
    return np.mean(np.maximum(anc_pos - anc_neg + gamma, 0.0))


def gap_loss(Z, gt0, gt1, gamma):
    """gap_loss branch (:547-594); returns shape (B,)."""
    b, n = gt0.shape
    m = gt1.shape[1]
    bi = np.arange(b)[:, None]
    rows = Z[:, :-1, :]                                # (b, n, m+1)
    pos = rows[bi, np.arange(n)[None], gt0]            # (b, n)
    lp = -np.log(np.exp(pos))[:, :, None]
    ln = -np.log(np.exp(rows))
    g = np.maximum(lp - ln + gamma, 0.0)
    g[bi, np.arange(n)[None], gt0] = 0.0               # the positive itself is excluded (:565-567)
    l0 = np.mean(2 * np.log(g.sum(axis=2) + 1), axis=1)
    # pc1 -> pc0 (:576-592). The reference selects with boolean masks over (n+1, m), which
    # flattens row-major: positives come out ordered by ROW index (not by column) and the
    # n*m negatives are re-viewed as (n, m) without regard to which column they came from.
    # That quirk is part of the observable loss value, so it is restated literally.
    cols = Z[:, :, :-1]                                # (b, n+1, m)
    pos_match = np.arange(n + 1)[None, :, None] == gt1[:, None, :]
    l1 = np.empty(b)
    for i in range(b):
        pos = cols[i][pos_match[i]].reshape(m)
        neg = cols[i][~pos_match[i]].reshape(n, m)
        lp = -np.log(np.exp(pos))[None, :]
        ln = -np.log(np.exp(neg))
        g = np.maximum(lp - ln + gamma, 0.0)
        l1[i] = np.mean(2 * np.log(g.sum(axis=0) + 1))
    return (l0 + l1) / 2


def superglue_loss(Z, gt0, gt1):
    """'superglue' loss branch (:487-511); gt arrays keep -1 for "no match", which indexes the dustbin column / row as a
    negative index does upstream (:493, :503). Needs n == m (:501 masks an (b, n) index tensor with a (b, m) mask)."""
    b, n = gt0.shape
    m = gt1.shape[1]
    if n != m:
        raise IndexError('superglue loss needs N == M')
    bi = np.arange(b)[:, None]
    tp = Z[bi, np.arange(n)[None], gt0].sum(axis=1)                    # :494-495
    inv = gt1 == -1                                                    # :498-499
    xx = inv.sum(axis=1)
    tn = np.array([Z[i, -1, :m][inv[i]].sum() for i in range(b)])      # :501-509: rows gt1 = -1 -> last row
    return np.mean((-tp - tn) / (xx + m))                              # :510


def knn(x, src, k):
    """knn() (:8-15): x (B,3,n), src (B,3,m) -> indices (B,n,k) of the k nearest sources."""
    inner = -2 * np.matmul(x.transpose(0, 2, 1), src)
    xx = (x ** 2).sum(axis=1, keepdims=True)
    ss = (src ** 2).sum(axis=1, keepdims=True)
    pd = -xx.transpose(0, 2, 1) - inner - ss
    return topk_indices(pd, k)


def get_graph_feature(x, src, k):
    """get_graph_feature() (:17-32): one-hot adjacency (B,n,m) int64 of the kNN graph."""
    idx = knn(x, src, k)
    adj = np.zeros((x.shape[0], x.shape[2], src.shape[2]), dtype=np.int64)
    np.put_along_axis(adj, idx, 1, axis=2)
    return adj


# --------------------------------------------------------------------------- forward

def _np(t):
    if hasattr(t, 'detach'):
        t = t.detach().cpu().numpy()
    return np.asarray(t)


def state_dict_to_numpy(sd):
    """Accepts a torch or numpy state dict, strips a DataParallel 'module.' prefix and applies
    the test.py weight convention fp64(fp32(w))."""
    out = {}
    for k, v in sd.items():
        k = k[7:] if k.startswith('module.') else k
        a = _np(v)
        if a.dtype.kind == 'f':
            a = a.astype(np.float32).astype(np.float64)
        out[k] = a
    return out


def forward(sd, data, cfg, trace=None):
    """MDGAT.forward (:369-603) for descriptor='FPFH', eval mode, float64.

    sd: numpy state dict (state_dict_to_numpy); data: dict of arrays in the loader's layout;
    cfg keys: L, k, sinkhorn_iterations, loss_method, mutual_check, match_threshold,
    triplet_loss_gamma. Returns the reference's output dict plus 'Z', 'scores_in'."""
    kpts0 = _np(data['keypoints0']).astype(np.float64)
    kpts1 = _np(data['keypoints1']).astype(np.float64)
    if kpts0.shape[1] == 0 or kpts1.shape[1] == 0:                       # :374-382
        return {'matches0': np.full(kpts0.shape[:-1], -1, dtype=np.int32)[0],
                'matches1': np.full(kpts1.shape[:-1], -1, dtype=np.int32)[0],
                'matching_scores0': np.zeros(kpts0.shape[:-1])[0],
                'matching_scores1': np.zeros(kpts1.shape[:-1])[0],
                'skip_train': True}
    d0 = _np(data['descriptors0']).astype(np.float64)
    d1 = _np(data['descriptors1']).astype(np.float64)
    s0 = _np(data['scores0']).astype(np.float64)
    s1 = _np(data['scores1']).astype(np.float64)
    desc0 = descriptor_encoder(sd, d0) + keypoint_encoder(sd, kpts0, s0)     # :392
    desc1 = descriptor_encoder(sd, d1) + keypoint_encoder(sd, kpts1, s1)     # :393
    if trace is not None:
        trace.append((desc0.copy(), desc1.copy()))
    desc0, desc1 = attentional_gnn(sd, desc0, desc1, cfg['k'], cfg['L'], trace)   # :395
    md0 = conv1x1(desc0, sd['final_proj.weight'], sd['final_proj.bias'])     # :397
    md1 = conv1x1(desc1, sd['final_proj.weight'], sd['final_proj.bias'])
    scores = np.matmul(md0.transpose(0, 2, 1), md1) / 128 ** .5              # :430-431
    Z = log_optimal_transport(scores, float(sd['bin_score']), cfg.get('sinkhorn_iterations', 100))
    m0, m1, ms0, ms1 = extract_matches(Z, cfg['loss_method'], cfg['mutual_check'],
                                       cfg.get('match_threshold', 0.2))
    out = {'matches0': m0, 'matches1': m1, 'matching_scores0': ms0, 'matching_scores1': ms1,
           'Z': Z, 'scores_in': scores}
    if 'gt_matches0' in data and cfg['loss_method'] in ('triplet_loss', 'gap_loss'):
        n, m = kpts0.shape[1], kpts1.shape[1]
        gt0 = _np(data['gt_matches0']).astype(np.int64)
        gt1 = _np(data['gt_matches1']).astype(np.int64)
        gt0 = np.where(gt0 == -1, m, gt0)                                    # :519-520
        gt1 = np.where(gt1 == -1, n, gt1)
        gamma = cfg.get('triplet_loss_gamma', 0.5)
        if cfg['loss_method'] == 'triplet_loss':
            out['loss'] = triplet_loss(Z, gt0, gt1, gamma)
        else:
            out['loss'] = gap_loss(Z, gt0, gt1, gamma)
    elif 'gt_matches0' in data and cfg['loss_method'] == 'superglue':
        out['loss'] = superglue_loss(Z, _np(data['gt_matches0']).astype(np.int64), _np(data['gt_matches1']).astype(np.int64))
    return out


def forward_threaded(sd, data, cfg, threads):
    """Batch elements are independent in eval mode: one element per task on a
    thread pool (numpy releases the GIL in matmul / ufunc loops). Used only as the CPU
    baseline leg of bench.py. Match-level outputs are concatenated; 'loss' is dropped."""
    b = _np(data['keypoints0']).shape[0]
    keys = ('keypoints0', 'keypoints1', 'descriptors0', 'descriptors1', 'scores0', 'scores1')
    arrs = {k: _np(data[k]) for k in keys}
    cfg = dict(cfg)

    def one(i):
        d = {k: arrs[k][i:i + 1] for k in keys}
        o = forward(sd, d, cfg)
        return {k: o[k] for k in ('matches0', 'matches1', 'matching_scores0', 'matching_scores1')}

    with ThreadPoolExecutor(max_workers=threads) as ex:
        parts = list(ex.map(one, range(b)))
    return {k: np.concatenate([p[k] for p in parts], axis=0) for k in parts[0]}


# --------------------------------------------------------------------------- output side (f-4)

def solve_icp(P, Q):
    """solve_icp() (/root/reference/utils/utils_test.py:73-110): one-shot Kabsch, P -> Q.
    R = U V^T straight from the SVD of Q_c^T P_c (no reflection fix upstream)."""
    up, uq = P.mean(axis=0), Q.mean(axis=0)
    U, _, Vt = np.linalg.svd((Q - uq).T @ (P - up), full_matrices=True)
    R = U @ Vt
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = uq - R @ up
    return T


def registration_error(mkpts0, mkpts1, T_gt):
    """calculate_error2() (utils_test.py:27-39): T from solve_icp(mkpts1, mkpts0), RTE and RRE
    of inv(T) T_gt."""
    T = solve_icp(mkpts1, mkpts0)
    E = np.linalg.inv(T) @ T_gt
    rte = np.linalg.norm(E[:3, 3])
    with np.errstate(invalid='ignore'):
        rre = np.arccos((E[0, 0] + E[1, 1] + E[2, 2] - 1) / 2)
    return T, rte, rre


def match_statistics(matches, matches_gt, m):
    """TP / FP / TN / FN counts of /root/reference/test_registration_metric.py:216-246 for one pair;
    gt value m (or -1) means 'no match' (the forward rewrites -1 to m, the script maps it back)."""
    g = np.where(matches_gt == m, -1, matches_gt)
    valid, valid_gt = matches > -1, g > -1
    return {'n_valid': int(valid.sum()), 'n_valid_gt': int(valid_gt.sum()),
            'tp': int((valid & (matches == g)).sum()), 'fp': int((valid & (matches != g)).sum()),
            'tn': int((~valid & (g == -1)).sum()), 'fn': int((~valid & (g > -1)).sum())}


# --------------------------------------------------------------------------- input side (f-2)

def prepare_pair(kp1, kp2, pose1, pose2, T_cam0_velo, threshold, mutual_check=False):
    """Ground truth of one pair as SparseDataset.__getitem__ builds it
    (/root/reference/load_data.py:213-285): world coordinates, nearest neighbours under a
    threshold (optionally mutual), T_gt and the repeatability count."""
    h1 = np.concatenate([kp1, np.ones((len(kp1), 1))], axis=1)
    h2 = np.concatenate([kp2, np.ones((len(kp2), 1))], axis=1)
    T_gt = np.linalg.inv(T_cam0_velo) @ np.linalg.inv(pose1) @ pose2 @ T_cam0_velo           # :238
    w1 = (pose1 @ T_cam0_velo @ h1.T).T[:, :3]                                               # :241-245
    w2 = (pose2 @ T_cam0_velo @ h2.T).T[:, :3]
    dists = np.sqrt(((w1[:, None, :] - w2[None, :, :]) ** 2).sum(-1))                        # cdist, :257
    min1, min2 = np.argmin(dists, axis=0), np.argmin(dists, axis=1)
    min1v = dists.min(axis=1)
    min1f = min2[min1v < threshold]
    rep = len(min1f)
    match1 = -np.ones(len(kp1), dtype=np.int16)
    match2 = -np.ones(len(kp2), dtype=np.int16)
    if mutual_check:
        xx = np.where(min2[min1] == np.arange(min1.shape[0]))[0]
        matches = np.intersect1d(min1f, xx)
        match1[min1[matches]] = matches
        match2[matches] = min1[matches]
    else:
        match1[min1v < threshold] = min1f
        min2v = dists.min(axis=0)
        match2[min2v < threshold] = min1[min2v < threshold]
    return match1, match2, T_gt, rep
