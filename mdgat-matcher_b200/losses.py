"""Loss terms of MDGAT.forward (mdgat.py:486-594) restated on the assignment matrix Z.

The triplet loss (test.py default) is computed inside the CUDA match kernel; these torch
versions serve the training path (autograd) and the eval-mode 'gap_loss' / 'superglue'
variants, which the device path evaluates from a materialised Z. They reproduce the
reference's observable values, including two quirks: -log(exp(z)) is kept literal, and
gap_loss's second direction gathers through boolean masks that flatten row-major
(mdgat.py:580-584), which misaligns positives and negatives whenever gt is not the identity.
"""
import torch


def _nle(z):
    return -torch.log(z.exp())


def triplet_loss(Z, gt0, gt1, gamma):
    """gt0 (B,N) / gt1 (B,M) long, 'no match' already mapped to M / N. mdgat.py:512-546."""
    b, n = gt0.shape
    top0 = Z[:, :-1, :].topk(2, dim=2).indices            # (B,N,2)
    top1 = Z[:, :, :-1].topk(2, dim=1).indices            # (B,2,M)
    pick0 = (top0[:, :, 0] == gt0).long()
    neg0 = top0.gather(2, pick0[:, :, None])[:, :, 0]
    an0 = Z[:, :-1, :].gather(2, neg0[:, :, None])[:, :, 0]
    ap0 = Z[:, :-1, :].gather(2, gt0[:, :, None])[:, :, 0]
    pick1 = (top1[:, 0, :] == gt1).long()
    neg1 = top1.gather(1, pick1[:, None, :])[:, 0, :]
    an1 = Z[:, :, :-1].gather(1, neg1[:, None, :])[:, 0, :]
    ap1 = Z[:, :, :-1].gather(1, gt1[:, None, :])[:, 0, :]
    an = _nle(torch.cat([an0, an1], dim=1))
    ap = _nle(torch.cat([ap0, ap1], dim=1))
    return torch.clamp(ap - an + gamma, min=0).mean()


def gap_loss(Z, gt0, gt1, gamma):
    """Returns shape (B,). mdgat.py:547-594."""
    b, n = gt0.shape
    m = gt1.shape[1]
    rows = Z[:, :-1, :]
    is_pos = torch.arange(m + 1, device=Z.device)[None, None, :] == gt0[:, :, None]
    pos = rows[is_pos].view(b, n)
    neg = rows[~is_pos].view(b, n, m)
    g = torch.clamp(_nle(pos)[:, :, None] - _nle(neg) + gamma, min=0)
    l0 = torch.mean(2 * torch.log(g.sum(dim=2) + 1), dim=1)
    cols = Z[:, :, :-1]
    is_pos = torch.arange(n + 1, device=Z.device)[None, :, None] == gt1[:, None, :]
    pos = cols[is_pos].view(b, m)                 # row-major flattening: ordered by row index
    neg = cols[~is_pos].view(b, n, m)
    g = torch.clamp(_nle(pos)[:, None, :] - _nle(neg) + gamma, min=0)
    l1 = torch.mean(2 * torch.log(g.sum(dim=1) + 1), dim=1)
    return (l0 + l1) / 2


def superglue_loss(Z, gt0, gt1):
    """gt with -1 for 'no match' (not remapped): index -1 addresses the dustbin. mdgat.py:487-511."""
    b, n = gt0.shape
    m = gt1.shape[1]
    if n != m:
        raise IndexError('superglue loss needs N == M (mdgat.py:501 indexes an N-shaped mask with M)')
    tp = Z.gather(2, (gt0.long() % (m + 1))[:, :, None])[:, :n, 0].sum(dim=1)
    unmatched = gt1 == -1
    last_row = Z[:, -1, :m]
    tn = (last_row * unmatched).sum(dim=1)
    xx = unmatched.sum(dim=1)
    return torch.mean((-tp - tn) / (xx + m))
