"""CPU model (torch float64) of the digit-plane arithmetic of the CUDA path, used to CHOOSE the number of int8 digit
planes per operand before spending GPU time. Development tool: not imported by the product.

The model restates the forward exactly as the kernels compute it -- BatchNorm and merge conv folded, operands cut into
S digit planes with the kernels' scaling rules (per row and 128-column chunk, base 128, for the GEMMs; per head row /
per value channel, base 256, for attention), only the plane pairs (s, t) with s + t <= S - 1 multiplied, P cut into SP
unsigned bytes below a row bound c_i >= max -- and compares matches / scores with the outputs of the UNMODIFIED
reference stored by oracle/gen_sweep.py.

    python tools/precision_model.py --cases sweep_s100_b16 --configs 7/7/6 5/5/4 4/4/3

A config is gemm_S/attn_S/attn_SP, optionally followed by :lo=<gemm_S>/<attn_S>/<attn_SP>@<first layer> to use
a second setting from that layer on, and / or by :t to run the top-k layers on the digit-plane kernel too (TOPK mode)
instead of float64 logits + exact selection.
"""
import argparse
import json
import math
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from mdgat_matcher_b200 import synth, packing     # noqa: E402
from oracle.build_ref import load_checkpoint_state_dict      # noqa: E402

BN_EPS = 1e-5
DEFAULT_K = [128, None, 128, None, 64, None, 64, None]


def frexp_exp(mx):
    """e with mx = m 2^e, m in [0.5, 1); 0 where mx == 0."""
    _, e = torch.frexp(mx)
    return torch.where(mx > 0, e, torch.zeros_like(e)).to(torch.float64)


GEMM_CHUNK = [128]          # columns that share one digit scale per row (128 = the kernels' k chunk; 0 = the whole row)
GEMM_BITS = [7]             # bits per GEMM digit plane (7 = the kernels' base-128 digits; 8: what-if base 256)
MSG_VBOUND = [False]        # message planes of the full-attention layers cut with the per-pair bound max |v| instead of the row maximum


def gemm_planes(x, S, e_min=None):
    """x [R, K] (K multiple of 128) -> (planes [S][R, K], trunc [S+1][R, K]) in value units; trunc[j] = first j planes.
    e_min [R, chunks, 1]: lower limit of the chunk exponents (-inf where the row maximum decides)."""
    R, K = x.shape
    ch = GEMM_CHUNK[0] or K
    xc = x.reshape(R, K // ch, ch)
    e = frexp_exp(xc.abs().amax(dim=2, keepdim=True))
    if e_min is not None:
        e = torch.maximum(e, e_min)
    b = float(GEMM_BITS[0])
    t = xc * torch.exp2(b - 1.0 - e)
    unit = torch.exp2(e - b + 1.0)
    trunc = [torch.zeros_like(x)]
    for j in range(1, S + 1):
        f = (2.0 ** b) ** (j - 1)
        trunc.append((torch.round(t * f) / f * unit).reshape(R, K))
    planes = [trunc[j + 1] - trunc[j] for j in range(S)]
    return planes, trunc


def gemm_emul(x, w, S, e_min=None):
    """x [R, K] @ w[Nout, K]^T with S digit planes per operand, pairs s + t <= S - 1 (0-based); S == 0: exact float64."""
    if S == 0:
        return x @ w.t()
    px, _ = gemm_planes(x, S, e_min)
    _, tw = gemm_planes(w, S)
    y = px[0] @ tw[S].t()
    for s in range(1, S):
        y = y + px[s] @ tw[S - s].t()
    return y


def bal_trunc(I, S):
    """I integer-valued float64 (|I| < 2^(8S-1)); returns trunc[j] = value of the first j balanced base-256 digits."""
    out = [torch.zeros_like(I)]
    for j in range(1, S + 1):
        f = 256.0 ** (S - j)
        out.append(torch.round(I / f) * f)
    return out


def attn_emul(q, k, v, S, SP, slack=1.5, topk=0):
    """q [B,H,N,32], k/v [B,H,M,32] float64 -> messages [B,H,N,32] as attn_i8_kernel computes them (S planes of q, k, v,
    SP byte planes of P); S == 0: exact float64 softmax attention. topk > 0: TOPK mode -- the kept set is chosen on the
    digit-plane logits themselves, the shift is the exact row maximum, probabilities outside the kept set are 0."""
    if S == 0:
        z = q @ k.transpose(2, 3) / math.sqrt(32.0)
        return torch.softmax(z, dim=-1) @ v
    eq = frexp_exp(q.abs().amax(dim=3, keepdim=True))
    ek = frexp_exp(k.abs().amax(dim=3, keepdim=True))
    Iq = torch.round(q * torch.exp2(8.0 * S - 2 - eq))
    Ik = torch.round(k * torch.exp2(8.0 * S - 2 - ek))
    tq, tk = bal_trunc(Iq, S), bal_trunc(Ik, S)
    z = None
    for s in range(S):
        term = (tq[s + 1] - tq[s]) @ tk[S - s].transpose(2, 3)
        z = term if z is None else z + term
    z = z * torch.exp2(eq - (8.0 * S - 2)) * torch.exp2(ek - (8.0 * S - 2)).transpose(2, 3) / math.sqrt(32.0)
    c = z.amax(dim=3, keepdim=True) + (0.0 if topk else slack)
    ph = torch.round(torch.exp(z - c) * 2.0 ** (8 * SP - 1))
    if topk:
        kth = z.topk(topk, dim=3).values[..., -1:]
        ph = torch.where(z >= kth, ph, torch.zeros_like(ph))
    rowsum = ph.sum(dim=3, keepdim=True)
    ev = frexp_exp(v.abs().amax(dim=2, keepdim=True))                 # per (b, h, channel)
    Iv = torch.round(v * torch.exp2(8.0 * S - 2 - ev))
    tv = bal_trunc(Iv, S)
    out = None
    for a in range(SP):
        if S - a < 1:
            break
        lo = torch.floor(ph / 256.0 ** (SP - 1 - a)) * 256.0 ** (SP - 1 - a)
        hi = torch.floor(ph / 256.0 ** (SP - a)) * 256.0 ** (SP - a)
        term = (lo - hi) @ tv[S - a]
        out = term if out is None else out + term
    return out * torch.exp2(ev - (8.0 * S - 2)) / rowsum


def fold_bn(w, b, sd, name):
    s = sd[name + '.weight'] / torch.sqrt(sd[name + '.running_var'] + BN_EPS)
    return w * s[:, None], (b - sd[name + '.running_mean']) * s + sd[name + '.bias']


def conv(sd, name):
    w = sd[name + '.weight']
    return w.reshape(w.shape[0], w.shape[1]), sd[name + '.bias']


def mlp_exact(sd, prefix, x, n_conv):
    for i in range(n_conv):
        w, b = conv(sd, '%s.%d' % (prefix, 3 * i))
        if i < n_conv - 1:
            w, b = fold_bn(w, b, sd, '%s.%d' % (prefix, 3 * i + 1))
        x = x @ w.t() + b
        if i < n_conv - 1:
            x = torch.relu(x)
    return x


def sinkhorn(scores, alpha, iters):
    b, m, n = scores.shape
    Z = scores.new_empty(b, m + 1, n + 1)
    Z[:, :m, :n] = scores
    Z[:, m, :] = alpha
    Z[:, :, n] = alpha
    norm = -math.log(m + n)
    log_mu = torch.cat([scores.new_full((m,), norm), scores.new_tensor([math.log(n) + norm])])[None].expand(b, -1)
    log_nu = torch.cat([scores.new_full((n,), norm), scores.new_tensor([math.log(m) + norm])])[None].expand(b, -1)
    u, v = torch.zeros_like(log_mu), torch.zeros_like(log_nu)
    for _ in range(iters):
        u = log_mu - torch.logsumexp(Z + v.unsqueeze(1), dim=2)
        v = log_nu - torch.logsumexp(Z + u.unsqueeze(2), dim=1)
    return Z + u.unsqueeze(2) + v.unsqueeze(1) - norm


def forward_model(sd, data, L, T, k_list, prec):
    """prec(layer) -> (gemm_S, attn_S, attn_SP, topk_S); topk_S = planes of q/k for the dense logits of a top-k layer
    (0 = float64 logits from the float64 q/k the GEMM produced)."""
    B, N, _ = data['keypoints0'].shape
    M = data['keypoints1'].shape[1]
    xs = []
    for side in ('0', '1'):
        kin = torch.cat([data['keypoints' + side], data['scores' + side][..., None]], dim=2)
        x = mlp_exact(sd, 'denc.encoder', data['descriptors' + side], 3) + mlp_exact(sd, 'kenc.encoder', kin, 4)
        xs.append(x.reshape(-1, 128))
    X = torch.cat(xs, 0)                       # [B*N + B*M, 128], point-major like the kernels
    R0 = B * N
    sched = packing.layer_k_schedule(k_list, L)
    for l in range(2 * L):
        gS, aS, aSP, tS = prec(l)
        p = 'gnn.layers.%d.' % l
        wq, bq = conv(sd, p + 'attn.proj.0')
        wk, bk = conv(sd, p + 'attn.proj.1')
        wv, bv = conv(sd, p + 'attn.proj.2')
        qkv = gemm_emul(X, torch.cat([wq, wk, wv], 0), gS) + torch.cat([bq, bk, bv])
        def heads(t, n):      # [B*n, 128] (c = d*4 + h) -> [B, 4, n, 32]
            return t.reshape(B, n, 32, 4).permute(0, 3, 1, 2).contiguous()
        q0, k0, v0 = (heads(qkv[:R0, i * 128:(i + 1) * 128], N) for i in range(3))
        q1, k1, v1 = (heads(qkv[R0:, i * 128:(i + 1) * 128], M) for i in range(3))
        cross = l % 2 == 1
        msgs = []
        ebs = []
        for (q, k, v) in ((q0, k1 if cross else k0, v1 if cross else v0), (q1, k0 if cross else k1, v0 if cross else v1)):
            kk = sched[l]
            if kk > 0:
                if tS == 0:
                    z = q @ k.transpose(2, 3) / math.sqrt(32.0)
                    idx = z.topk(kk, dim=3).indices
                    pr = torch.zeros_like(z).scatter(3, idx, torch.softmax(z.gather(3, idx), dim=-1))
                    o = pr @ v
                else:
                    o = attn_emul(q, k, v, aS, aSP, topk=kk)
            else:
                o = attn_emul(q, k, v, aS, aSP)
            n = o.shape[2]
            msgs.append(o.permute(0, 2, 3, 1).reshape(B * n, 128))        # back to c = d*4 + h
            eb = frexp_exp(v.abs().amax(dim=(1, 2, 3)))                   # per pair: |message| <= max |v| < 2^eb
            if not (MSG_VBOUND[0] and kk == 0):
                eb = torch.full_like(eb, -1e9)
            ebs.append(eb[:, None].expand(B, n).reshape(B * n))
        msg = torch.cat(msgs, 0)
        e_min = torch.stack([torch.full_like(torch.cat(ebs, 0), -1e9), torch.cat(ebs, 0)], 1)[:, :, None]
        wm, bm = conv(sd, p + 'attn.merge')
        w1, b1 = conv(sd, p + 'mlp.0')
        w1, b1 = fold_bn(w1, b1, sd, p + 'mlp.1')
        w1f = torch.cat([w1[:, :128], w1[:, 128:] @ wm], 1)
        b1f = b1 + w1[:, 128:] @ bm
        h = torch.relu(gemm_emul(torch.cat([X, msg], 1), w1f, gS, e_min if GEMM_CHUNK[0] == 128 else None) + b1f)
        w2, b2 = conv(sd, p + 'mlp.3')
        X = X + gemm_emul(h, w2, gS) + b2
    wf, bf = conv(sd, 'final_proj')
    MD = X @ wf.t() + bf
    md0, md1 = MD[:R0].reshape(B, N, 128), MD[R0:].reshape(B, M, 128)
    scores = md0 @ md1.transpose(1, 2) / math.sqrt(128.0)
    Z = sinkhorn(scores, sd['bin_score'], T)
    max0, max1 = Z[:, :-1, :].max(2), Z[:, :, :-1].max(1)
    v0, v1 = max0.indices < M, max1.indices < N
    return {
        'matches0': torch.where(v0, max0.indices, -1), 'matches1': torch.where(v1, max1.indices, -1),
        'matching_scores0': torch.where(v0, max0.values.exp(), 0.0), 'matching_scores1': torch.where(v1, max1.values.exp(), 0.0),
        'Z_rowmax': max0.values, 'Z_colmax': max1.values,
    }


def parse_config(s):
    parts = s.split(':')
    base = tuple(int(x) for x in parts[0].split('/'))
    lo, first = None, 10 ** 9
    for p in parts[1:]:
        if p.startswith('lo='):
            v, at = p[3:].split('@')
            lo, first = tuple(int(x) for x in v.split('/')), int(at)
    topk_i8 = any(p == 't' for p in parts[1:])          # ':t' = top-k layers on the digit-plane kernel too (TOPK mode)

    def prec(l):
        g, a, sp = (lo if l >= first else base)
        return g, a, sp, (a if topk_i8 else 0)
    return prec


def compare(out, ref):
    res = {}
    flips = 0
    errs = []
    for side in ('0', '1'):
        m = out['matches' + side].numpy()
        flips += int((m != ref['matches' + side].astype(np.int64)).sum())
        errs.append(np.abs(out['matching_scores' + side].numpy() - ref['matching_scores' + side]).ravel())
    e = np.concatenate(errs)
    zr = np.abs(out['Z_rowmax'].numpy() - ref['Z_rowmax']).max()
    res.update(rows=int(e.size), flips=flips, max_score_err=float(e.max()), p99=float(np.quantile(e, 0.99)),
               p999=float(np.quantile(e, 0.999)), frac_gt_1e5=float((e > 1e-5).mean()), max_Zrow_err=float(zr))
    return res


def run_golden(names, configs):
    """The same comparison on the committed golden cases of tests/golden (other shapes, weights, k lists)."""
    sys.path.insert(0, os.path.join(ROOT, 'tests'))
    from conftest import load_golden, golden_inputs, case_weights
    for name in names:
        rec = load_golden(name)
        case = rec['case']
        sd = {k: torch.from_numpy(np.asarray(v)) for k, v in case_weights(case).items()}
        data = {k: torch.from_numpy(v) for k, v in golden_inputs(rec).items()}
        k_list = case.get('k', DEFAULT_K)
        if case.get('loss_method') == 'superglue':
            print(json.dumps({'case': name, 'skipped': 'threshold match variant not modelled'}))
            continue
        ref = {k: rec[k] for k in ('matches0', 'matches1', 'matching_scores0', 'matching_scores1', 'Z_rowmax')}
        for cs in configs:
            t = time.time()
            with torch.no_grad():
                out = forward_model(sd, data, case['L'], case['T'], k_list, parse_config(cs))
            r = compare(out, ref)
            r.update(case=name, config=cs, seconds=round(time.time() - t, 1))
            print(json.dumps(r), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--golden', nargs='+', default=None, help='golden case names (tests/golden) instead of sweep cases')
    ap.add_argument('--cases', nargs='+', default=['sweep_s100_b16'])
    ap.add_argument('--configs', nargs='+', default=['7/7/6', '5/5/4', '4/4/3'])
    ap.add_argument('--pairs', type=int, default=0, help='only the first n pairs of each case (0 = all)')
    ap.add_argument('--threads', type=int, default=0)
    ap.add_argument('--msg-vbound', action='store_true', help='message planes of full-attention layers scaled by the per-pair bound max |v|')
    ap.add_argument('--gemm-bits', type=int, default=7, help='bits per GEMM digit plane (what-if; the kernels use 7)')
    ap.add_argument('--gemm-chunk', type=int, default=128, help='columns sharing one digit scale per row (0 = whole row)')
    args = ap.parse_args()
    if args.threads:
        torch.set_num_threads(args.threads)
    GEMM_CHUNK[0] = args.gemm_chunk
    GEMM_BITS[0] = args.gemm_bits
    MSG_VBOUND[0] = args.msg_vbound
    if args.golden:
        run_golden(args.golden, args.configs)
        return
    sd_np = load_checkpoint_state_dict()
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in sd_np.items()}
    for cname in args.cases:
        ref = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'sweep', cname + '.npz')))
        case = json.loads(str(ref['case']))
        data = synth.make_batch(case['seed'], case['B'], case['N'])
        chk = json.loads(str(ref['input_checksums']))
        for k, v in data.items():
            assert abs(float(v.double().sum()) - chk[k]) <= 1e-12 * abs(chk[k]), 'input %s differs from the generator run' % k
        nb = args.pairs or case['B']
        data = {k: v[:nb] for k, v in data.items()}
        refc = {k: (v[:nb] if isinstance(v, np.ndarray) and v.ndim >= 1 and v.shape[0] == case['B'] else v) for k, v in ref.items()}
        for cs in args.configs:
            t = time.time()
            with torch.no_grad():
                out = forward_model(sd, data, case['L'], case['T'], DEFAULT_K, parse_config(cs))
            r = compare(out, refc)
            r.update(case=cname, config=cs, seconds=round(time.time() - t, 1))
            print(json.dumps(r), flush=True)


if __name__ == '__main__':
    main()
