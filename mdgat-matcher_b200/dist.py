"""Multi-GPU plumbing: one process per GPU, the batch of pairs sharded by rank, weights
replicated once at start-up, and a single all-gather of the match results per batch
(SURVEY.md section 8e). Batch elements are independent in eval mode, so no kernel of the
path is followed by a collective; this replaces the per-forward parameter broadcast +
scatter + gather of nn.DataParallel (/root/reference/train.py:192-196, test.py:158).
"""
import os

import torch
import torch.distributed as dist


def init_process_group(backend=None, device_id=None):
    """Reads RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT from the environment (torchrun)."""
    if dist.is_initialized():
        return
    if backend is None:
        backend = 'nccl' if torch.cuda.is_available() else 'gloo'
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    os.environ.setdefault('MASTER_PORT', '29511')
    kw = {'device_id': device_id} if (device_id is not None and backend == 'nccl') else {}
    dist.init_process_group(backend=backend, rank=int(os.environ.get('RANK', '0')),
                            world_size=int(os.environ.get('WORLD_SIZE', '1')), **kw)


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)


def shard_bounds(n_items, rank, world_size):
    """Contiguous shard [lo, hi) of n_items for this rank; the first n_items % world ranks get one extra."""
    base, extra = divmod(n_items, world_size)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(data, rank=None, world_size=None):
    """Slice every tensor of a loader dict along dim 0 (the pair index)."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    n = next(v.shape[0] for v in data.values() if isinstance(v, torch.Tensor))
    lo, hi = shard_bounds(n, rank, world_size)
    return {k: (v[lo:hi] if isinstance(v, torch.Tensor) and v.dim() > 0 and v.shape[0] == n else v)
            for k, v in data.items()}


def all_gather_outputs(out, keys=('matches0', 'matches1', 'matching_scores0', 'matching_scores1')):
    """Concatenates the per-rank results along the pair dimension on every rank (rank order).
    Shards may be ragged (n_items not divisible by the world size)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return {k: out[k] for k in keys if k in out}
    w = dist.get_world_size()
    res = {}
    for k in keys:
        if k not in out:
            continue
        t = out[k].contiguous()
        cnt = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
        counts = [torch.zeros_like(cnt) for _ in range(w)]
        dist.all_gather(counts, cnt)
        counts = [int(c.item()) for c in counts]
        mx = max(counts)
        pad = t.new_zeros((mx,) + tuple(t.shape[1:]))
        pad[:t.shape[0]] = t
        parts = [torch.empty_like(pad) for _ in range(w)]
        dist.all_gather(parts, pad)
        res[k] = torch.cat([p[:c] for p, c in zip(parts, counts)], dim=0)
    return res


def all_reduce_mean_loss(loss, n_local):
    """Mean loss over the global batch from per-rank means (triplet loss is a mean over pairs)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return loss
    t = torch.stack([loss.double() * n_local, loss.new_tensor(float(n_local), dtype=torch.float64)])
    dist.all_reduce(t)
    return t[0] / t[1]
