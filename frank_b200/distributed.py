"""Multi-GPU plumbing (one process per GPU, torch.distributed for the rendezvous).

Two things shard on this path:

* the **visibilities**: `M`, `j` and `H0` are sums over independent visibilities (frank/statistical_models.py:200-218
  already accumulates them chunk by chunk), so every rank maps its slice and ONE collective combines the partial
  normal equations.  With `init_library_comm()` the library owns an NCCL communicator and issues that all-reduce itself,
  on its own stream, right behind the Gram kernel (no host synchronisation in between); without it
  `allreduce_mapping()` does the same through torch.distributed (NCCL on GPUs, gloo in the CPU tests);
* the **hyper-parameter sweep** (BASELINE config 4; the reference's `run_multiple_fits`, frank/fit.py:493-563, loops over
  (alpha, w_smooth) pairs and even re-maps the visibilities for each): the grid points are independent given `M`, `j`, so
  `sweep_sharded()` deals them to the ranks, every rank runs its points as ONE batched device loop, and an all-gather
  returns every point's (p, mu, niter, converged) to every rank.
"""
import numpy as np

__all__ = ['shard_bounds', 'init_library_comm', 'allreduce_mapping', 'channel_major_order', 'map_visibilities_sharded', 'sweep_shard',
           'sweep_sharded']


def shard_bounds(n, rank, world):
    """Contiguous slice [lo, hi) of n items owned by `rank` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _dist():
    import torch.distributed as dist
    return dist if (dist.is_available() and dist.is_initialized()) else None


def init_library_comm(ctx=None, group=None):
    """Attach an NCCL communicator of libfrankb200 to `ctx` (default: this rank's context).

    Rank 0 creates the NCCL unique id (fb_comm_unique_id), torch.distributed broadcasts its 128 bytes, every rank joins
    (fb_comm_init).  From then on `VisibilityMapping.map_visibilities` returns the all-reduced normal equations on every
    rank.  No-op for a single process.  Returns the context."""
    from frank_b200 import _lib
    if ctx is None:
        ctx = _lib.get_context()
    dist = _dist()
    if dist is None or dist.get_world_size(group) == 1:
        return ctx
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    box = [ctx.comm_unique_id().tobytes() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    ctx.comm_init(world, rank, np.frombuffer(box[0], dtype=np.uint8))
    return ctx


def allreduce_mapping(mapping, group=None, device=None):
    """Sum 'M', 'j', 'null_likelihood' of a map_visibilities() result over the ranks of `group`, in place, through
    torch.distributed (for callers that have not attached a library communicator).

    The three are packed into one buffer so a single collective is issued; with the NCCL backend the buffer lives on
    `device` (default: this rank's context device, NOT torch's current device)."""
    import torch
    dist = _dist()
    if dist is None or dist.get_world_size(group) == 1:
        return mapping
    M, j = np.asarray(mapping['M']), np.asarray(mapping['j'])
    packed = np.concatenate([M.reshape(-1), j.reshape(-1), [mapping['null_likelihood']]])
    t = torch.from_numpy(packed)
    if dist.get_backend(group) == 'nccl':
        if device is None:
            from frank_b200 import _lib
            device = _lib.get_context().device
        t = t.to(torch.device('cuda', device))
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    out = t.cpu().numpy()
    mapping['M'] = out[:M.size].reshape(M.shape).copy()
    mapping['j'] = out[M.size:M.size + j.size].reshape(j.shape).copy()
    mapping['null_likelihood'] = float(out[-1])
    return mapping


def channel_major_order(frequencies):
    """Stable permutation that groups the visibilities by frequency channel.  Sharding the permuted arrays hands every rank
    as few channels as possible -- with C channels over R >= C ranks, 1 / (R / C) of ONE channel instead of 1 / R of every
    channel -- which keeps the rank's tiles of 64 baseline-sorted visibilities C times narrower in baseline (the Gram
    kernel's one-row-per-(mode, tile) regime, DESIGN.md K3 'sparse regime')."""
    return np.argsort(np.asarray(frequencies).reshape(-1), kind='stable')


def map_visibilities_sharded(vis_map, u, v, V, weights, group=None, frequencies=None, channel_major=False):
    """Every rank passes the FULL arrays (or views of them); each maps its own contiguous slice and the partial
    normal equations are summed over the ranks.

    With a library communicator attached to the mapping's context (`init_library_comm`) the sum -- and the global
    q-range check of statistical_models.py:512-535 -- happen inside the device call.  Otherwise torch.distributed does
    both: a rank whose slice is out of range raises on every rank.

    Multi-frequency data: the channel list is the GLOBAL np.unique(frequencies), whatever a rank's slice contains;
    `channel_major=True` slices the channel-sorted order (`channel_major_order`) instead of the given one."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group) if dist else 1
    rank = dist.get_rank(group) if dist else 0
    lo, hi = shard_bounds(len(u), rank, world)
    sel = slice(lo, hi)
    channels = None
    if frequencies is not None:
        frequencies = np.asarray(frequencies).reshape(-1)
        channels = np.unique(frequencies)
        if channel_major:
            sel = channel_major_order(frequencies)[lo:hi]
    w = weights if np.ndim(weights) == 0 else weights[sel]
    kw = {} if frequencies is None else {'frequencies': frequencies[sel], 'channels': channels}
    in_library = False
    if world > 1 and hasattr(vis_map, '_context'):
        in_library = vis_map._context().comm_info()[0]
    if in_library:
        return vis_map.map_visibilities(u[sel], v[sel], V[sel], w, **kw)
    err = None
    try:
        mapping = vis_map.map_visibilities(u[sel], v[sel], V[sel], w, **kw)
    except ValueError as e:                      # out-of-range baselines on this rank
        err, mapping = e, None
    if world > 1:
        flag = torch.tensor([1.0 if err is not None else 0.0], dtype=torch.float64)
        if dist.get_backend(group) == 'nccl':
            from frank_b200 import _lib
            flag = flag.to(torch.device('cuda', _lib.get_context().device))
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        if flag.item() > 0:
            raise err if err is not None else ValueError("a peer rank found baselines beyond the last collocation point")
    elif err is not None:
        raise err
    return allreduce_mapping(mapping, group)


def sweep_shard(n_points, rank, world):
    """Grid points of rank `rank`: every world-th point of the flattened (alpha-major) grid starting at `rank`.  Strided, not
    contiguous: the iteration count of a point grows steeply as alpha -> 1 (frank/radial_fitters.py:804-808 says as much), so
    a contiguous block would hand one rank all the slow points."""
    return np.arange(rank, n_points, world)


def sweep_sharded(solve_points, n_points, N, group=None, ctx=None):
    """Run a hyper-parameter sweep sharded by grid point and gather every point's result on every rank.

    solve_points(idx) -> dict(p [k, N], mu [k, N], niter [k], converged [k]) solves the grid points `idx` (this rank's
    share) -- on the GPU path one batched fb_frank_normal_loop call.  Results travel as one packed float64 row per
    point (p | mu | niter | converged) through the library's NCCL all-gather when a communicator is attached to `ctx`,
    else through torch.distributed.all_gather (gloo in the CPU tests).  Returns dict(p [n_points, N], mu, niter,
    converged) identical on every rank and, by construction, identical to the single-rank batch: the per-point
    arithmetic does not depend on which other points share the batch."""
    import torch
    dist = _dist()
    world = dist.get_world_size(group) if dist else 1
    rank = dist.get_rank(group) if dist else 0
    idx = sweep_shard(n_points, rank, world)
    res = solve_points(idx) if len(idx) else {'p': np.zeros((0, N)), 'mu': np.zeros((0, N)), 'niter': np.zeros(0), 'converged': np.zeros(0)}
    width = 2 * N + 2
    per = -(-n_points // world)                                   # rows per rank, padded to the largest share
    mine = np.zeros((per, width))
    k = len(idx)
    mine[:k, :N], mine[:k, N:2 * N] = res['p'], res['mu']
    mine[:k, 2 * N], mine[:k, 2 * N + 1] = res['niter'], res['converged']
    if world == 1:
        rows = mine[None]
    elif ctx is not None and ctx.comm_info()[0]:
        rows = ctx.comm_allgather(mine).reshape(world, per, width)
    else:
        t = torch.from_numpy(mine)
        if dist.get_backend(group) == 'nccl':
            from frank_b200 import _lib
            t = t.to(torch.device('cuda', _lib.get_context().device))
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t, group=group)
        rows = np.stack([x.cpu().numpy() for x in parts])
    out = np.zeros((n_points, width))
    for r in range(world):
        sel = sweep_shard(n_points, r, world)
        out[sel] = rows[r, :len(sel)]
    return {'p': out[:, :N].copy(), 'mu': out[:, N:2 * N].copy(), 'niter': out[:, 2 * N].astype(np.int64),
            'converged': out[:, 2 * N + 1].astype(bool)}
