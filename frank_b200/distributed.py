"""Multi-GPU plumbing for the mapping path (one process per GPU, torch.distributed).

`M`, `j` and `H0` are sums over independent visibilities (frank/statistical_models.py:200-218 already accumulates
them chunk by chunk), so the path shards by visibility with no data-path collective: every rank maps its slice and
one all-reduce of the packed (M, j, H0) buffer -- 0.72 MB at N = 300 -- combines them; the q-range check needs a
min/max all-reduce.  NCCL on GPUs (NVLink/NVSwitch), gloo in the CPU tests.
"""
import numpy as np

__all__ = ['shard_bounds', 'allreduce_mapping', 'map_visibilities_sharded']


def shard_bounds(n, rank, world):
    """Contiguous slice [lo, hi) of n visibilities owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_mapping(mapping, group=None, device=None):
    """Sum 'M', 'j', 'null_likelihood' of a map_visibilities() result over the ranks of `group`, in place.

    The three are packed into one buffer so a single collective is issued."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return mapping
    M, j = np.asarray(mapping['M']), np.asarray(mapping['j'])
    packed = np.concatenate([M.reshape(-1), j.reshape(-1), [mapping['null_likelihood']]])
    backend = dist.get_backend(group)
    t = torch.from_numpy(packed)
    if backend == 'nccl':
        t = t.cuda(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    out = t.cpu().numpy()
    mapping['M'] = out[:M.size].reshape(M.shape).copy()
    mapping['j'] = out[M.size:M.size + j.size].reshape(j.shape).copy()
    mapping['null_likelihood'] = float(out[-1])
    return mapping


def map_visibilities_sharded(vis_map, u, v, V, weights, group=None):
    """Every rank passes the FULL arrays (or views of them); each maps its own contiguous slice and the partial
    normal equations are all-reduced.  The q-range check (statistical_models.py:512-535) is made globally: a rank
    whose slice is out of range raises on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(len(u), rank, world)
    w = weights if np.ndim(weights) == 0 else weights[lo:hi]
    err = None
    try:
        mapping = vis_map.map_visibilities(u[lo:hi], v[lo:hi], V[lo:hi], w)
    except ValueError as e:                      # out-of-range baselines on this rank
        err, mapping = e, None
    if world > 1:
        flag = torch.tensor([1.0 if err is not None else 0.0], dtype=torch.float64)
        if dist.get_backend(group) == 'nccl':
            flag = flag.cuda()
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        if flag.item() > 0:
            raise err if err is not None else ValueError("a peer rank found baselines beyond the last collocation point")
    elif err is not None:
        raise err
    return allreduce_mapping(mapping, group)
