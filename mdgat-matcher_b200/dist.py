"""Multi-GPU plumbing: one process per GPU, the batch of pairs sharded by rank, weights
replicated once at start-up, and a single all-gather of the match results per batch
(ONE all_gather_into_tensor of a packed int64 buffer, no host synchronisation). Batch elements are independent in eval mode, so no kernel of the
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


# entries of a loader batch that carry one item per pair (load_data.py:299-321 + what test.py adds); everything else in
# the dict (shared calibration, flags, ...) is passed through unsliced even if its first dimension happens to equal the
# batch size
PAIR_KEYS = ('keypoints0', 'keypoints1', 'descriptors0', 'descriptors1', 'scores0', 'scores1', 'gt_matches0',
             'gt_matches1', 'T_gt', 'rep', 'idx0', 'idx1', 'sequence', 'cloud0', 'cloud1')


def shard_batch(data, rank=None, world_size=None, keys=None):
    """Slice the per-pair entries of a loader dict (tensors or lists, `keys` or PAIR_KEYS) along the pair index."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    keys = [k for k in (PAIR_KEYS if keys is None else keys) if k in data]
    if not keys:
        return dict(data)
    n = len(data[keys[0]])
    lo, hi = shard_bounds(n, rank, world_size)
    return {k: (v[lo:hi] if k in keys else v) for k, v in data.items()}


OUTPUT_KEYS = ('matches0', 'matches1', 'matching_scores0', 'matching_scores1')


def pack_outputs(out, keys=OUTPUT_KEYS):
    """One (pairs, columns) int64 buffer holding every requested result (float64 columns are carried as their bit
    patterns), plus the layout needed to take it apart again."""
    cols, meta = [], []
    for k in keys:
        if k not in out:
            continue
        t = out[k]
        if t.dtype not in (torch.int64, torch.float64):
            raise TypeError('%s: only int64 / float64 results are packed (got %s)' % (k, t.dtype))
        flat = t.contiguous().reshape(t.shape[0], -1)
        cols.append(flat if t.dtype == torch.int64 else flat.view(torch.int64))
        meta.append((k, t.dtype, tuple(t.shape[1:]), flat.shape[1]))
    return torch.cat(cols, dim=1), meta


def unpack_outputs(buf, meta):
    res, c = {}, 0
    for k, dtype, shape, width in meta:
        col = buf[:, c:c + width].contiguous()
        res[k] = (col if dtype == torch.int64 else col.view(torch.float64)).reshape((buf.shape[0],) + shape)
        c += width
    return res


def all_gather_outputs(out, keys=OUTPUT_KEYS, n_total=None):
    """Concatenates the per-rank results along the pair dimension on every rank (rank order) with ONE collective: the
    results are packed into a single int64 buffer and exchanged by all_gather_into_tensor; no host synchronisation.
    Equal shards by default; ragged shards (n_total pairs split by shard_bounds) are padded to the largest shard, whose
    size every rank derives from n_total without communicating."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return {k: out[k] for k in keys if k in out}
    w = dist.get_world_size()
    buf, meta = pack_outputs(out, keys)
    b = buf.shape[0]
    if n_total is None:
        sizes = [b] * w
    else:
        sizes = [hi - lo for lo, hi in (shard_bounds(n_total, r, w) for r in range(w))]
        if sizes[dist.get_rank()] != b:
            raise ValueError('rank %d holds %d pairs, shard_bounds(%d) says %d' % (dist.get_rank(), b, n_total, sizes[dist.get_rank()]))
    bmax = max(sizes)
    if b < bmax:
        buf = torch.cat([buf, buf.new_zeros((bmax - b, buf.shape[1]))], dim=0)
    gathered = buf.new_empty((w * bmax, buf.shape[1]))
    dist.all_gather_into_tensor(gathered, buf)
    if min(sizes) < bmax:
        gathered = torch.cat([gathered[r * bmax:r * bmax + sizes[r]] for r in range(w)], dim=0)
    return unpack_outputs(gathered, meta)


def all_reduce_mean_loss(loss, n_local):
    """Mean loss over the global batch from per-rank means (triplet loss is a mean over pairs)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return loss
    t = torch.stack([loss.double() * n_local, loss.new_tensor(float(n_local), dtype=torch.float64)])
    dist.all_reduce(t)
    return t[0] / t[1]
