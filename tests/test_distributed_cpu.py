"""World-size-2 gloo tests (CPU) of the multi-GPU plumbing: visibility sharding + all-reduce of the partial normal
equations.  The per-rank mapper here is the CPU oracle standing in for the CUDA kernel (no GPU in this test)."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from oracle import frank_oracle as fo


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _OracleMapper(object):
    """Duck-types VisibilityMapping.map_visibilities with the oracle (test double)."""

    def __init__(self, N):
        self.dht = fo.DHTTables(1.6 / fo.RAD_TO_ARCSEC, N)

    def map_visibilities(self, u, v, V, w):
        m = fo.map_visibilities(self.dht, u, v, V, w, 30., 40., 1e-3, -2e-3)
        return {'M': m['M'], 'j': m['j'], 'null_likelihood': m['null_likelihood']}


def _worker(rank, world, port, out_dir, scale):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from frank_b200.distributed import map_visibilities_sharded, shard_bounds
    u, v, V, w, _ = fo.synthetic_disc(4001, 40, seed=77)
    u = u * scale
    vm = _OracleMapper(40)
    try:
        m = map_visibilities_sharded(vm, u, v, V, w)
        np.savez(os.path.join(out_dir, f'r{rank}.npz'), M=m['M'], j=m['j'], H0=m['null_likelihood'], lohi=shard_bounds(4001, rank, world))
    except ValueError:
        np.savez(os.path.join(out_dir, f'r{rank}.npz'), raised=1)
    dist.destroy_process_group()


def test_sharded_mapping_allreduce_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), 1.0), nprocs=world, join=True)
    u, v, V, w, dht = fo.synthetic_disc(4001, 40, seed=77)
    ref = fo.map_visibilities(dht, u, v, V, w, 30., 40., 1e-3, -2e-3)
    r0, r1 = np.load(tmp_path / 'r0.npz'), np.load(tmp_path / 'r1.npz')
    assert list(r0['lohi']) == [0, 2001] and list(r1['lohi']) == [2001, 4001]
    assert np.array_equal(r0['M'], r1['M']) and np.array_equal(r0['j'], r1['j'])          # every rank holds the sum
    assert np.max(np.abs(r0['M'] - ref['M'])) <= 1e-14 * np.max(np.abs(ref['M']))
    assert np.max(np.abs(r0['j'] - ref['j'])) <= 1e-13 * np.max(np.abs(ref['j']))
    assert abs(float(r0['H0']) - ref['null_likelihood']) <= 1e-12 * abs(ref['null_likelihood'])


def test_sharded_mapping_range_error_is_global(tmp_path):
    """Baselines beyond the last collocation point on ONE rank raise ValueError on every rank."""
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), 3.0), nprocs=world, join=True)
    assert 'raised' in np.load(tmp_path / 'r0.npz').files and 'raised' in np.load(tmp_path / 'r1.npz').files


class _OracleChannelMapper(_OracleMapper):
    """Multi-frequency stand-in with the product's `channels` contract: one (M, j) per entry of the GLOBAL channel list."""

    def map_visibilities(self, u, v, V, w, frequencies=None, channels=None):
        assert channels is not None, "a sharded multi-frequency call must pass the global channel list"
        N = self.dht.N if hasattr(self.dht, 'N') else len(self.dht.j_nk)
        Ms, js, H0 = np.zeros((len(channels), N, N)), np.zeros((len(channels), N)), 0.0
        for c, f in enumerate(channels):
            s = frequencies == f
            if s.any():
                m = fo.map_visibilities(self.dht, u[s], v[s], V[s], w[s], 30., 40., 1e-3, -2e-3)
                Ms[c], js[c] = m['M'], m['j']
                H0 += m['null_likelihood']
        self.seen = np.unique(frequencies)
        return {'M': Ms, 'j': js, 'null_likelihood': H0}


def _chan_worker(rank, world, port, out_dir, channel_major):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from frank_b200.distributed import map_visibilities_sharded
    u, v, V, w, _ = fo.synthetic_disc(3000, 24, seed=5)
    # frequency-sorted data: the second rank's slice has no visibility of the first channel, and the other way round
    freqs = np.repeat([1.0e11, 2.3e11], 1500) if not channel_major else np.tile([1.0e11, 2.3e11], 1500)
    vm = _OracleChannelMapper(24)
    m = map_visibilities_sharded(vm, u, v, V, w, frequencies=freqs, channel_major=channel_major)
    np.savez(os.path.join(out_dir, f'r{rank}.npz'), M=m['M'], j=m['j'], H0=m['null_likelihood'], seen=vm.seen)
    dist.destroy_process_group()


@pytest.mark.parametrize('channel_major', [False, True])
def test_sharded_multifrequency_uses_the_global_channel_list(tmp_path, channel_major):
    """A rank whose slice misses a channel still returns that channel (zeros), so the all-reduce lines up; with
    channel_major=True interleaved frequencies are dealt so that every rank sees ONE channel."""
    world = 2
    mp.spawn(_chan_worker, args=(world, _free_port(), str(tmp_path), channel_major), nprocs=world, join=True)
    u, v, V, w, dht = fo.synthetic_disc(3000, 24, seed=5)
    freqs = np.repeat([1.0e11, 2.3e11], 1500) if not channel_major else np.tile([1.0e11, 2.3e11], 1500)
    r0, r1 = np.load(tmp_path / 'r0.npz'), np.load(tmp_path / 'r1.npz')
    assert r0['M'].shape == (2, 24, 24) and np.array_equal(r0['M'], r1['M'])
    assert len(r0['seen']) == 1 and len(r1['seen']) == 1 and r0['seen'][0] != r1['seen'][0]
    for c, f in enumerate([1.0e11, 2.3e11]):
        s = freqs == f
        ref = fo.map_visibilities(dht, u[s], v[s], V[s], w[s], 30., 40., 1e-3, -2e-3)
        assert np.max(np.abs(r0['M'][c] - ref['M'])) <= 1e-14 * np.max(np.abs(ref['M']))
        assert np.max(np.abs(r0['j'][c] - ref['j'])) <= 1e-13 * np.max(np.abs(ref['j']))


def test_shard_bounds_cover():
    from frank_b200.distributed import shard_bounds
    for n in [0, 1, 7, 64, 1000003]:
        for world in [1, 2, 3, 8]:
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def _sweep_worker(rank, world, port, out_dir):
    """Sweep sharded by grid point: every rank solves its share of a 3 x 3 (alpha, wsmooth) grid (oracle solver standing in
    for the batched device loop) and the all-gather returns the whole grid to every rank."""
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port))
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from frank_b200.distributed import sweep_sharded
    N = 24
    u, v, V, w, dht = fo.synthetic_disc(3000, N, seed=5)
    m = fo.map_visibilities(dht, u, v, V, w, 30., 40., 1e-3, -2e-3)
    grid = [(a, ws) for a in (1.05, 1.2, 1.4) for ws in (1e-3, 1e-2, 1e-1)]
    solved = []

    def solve_points(idx):
        solved.extend(int(i) for i in idx)
        fits = [fo.frank_fit(dht, m['M'], m['j'], alpha=grid[i][0], weights_smooth=grid[i][1], max_iter=60) for i in idx]
        return {'p': np.array([f['power_spectrum'] for f in fits]), 'mu': np.array([f['MAP'] for f in fits]),
                'niter': np.array([f['num_iterations'] for f in fits]), 'converged': np.array([f['converged'] for f in fits])}

    res = sweep_sharded(solve_points, len(grid), N)
    np.savez(os.path.join(out_dir, f's{rank}.npz'), solved=np.array(solved), **res)
    dist.destroy_process_group()


def test_sweep_sharded_by_grid_point_gloo(tmp_path):
    world = 2
    mp.spawn(_sweep_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r0, r1 = np.load(tmp_path / 's0.npz'), np.load(tmp_path / 's1.npz')
    assert list(r0['solved']) == [0, 2, 4, 6, 8] and list(r1['solved']) == [1, 3, 5, 7]      # each point solved exactly once
    for k in ('p', 'mu', 'niter', 'converged'):
        assert np.array_equal(r0[k], r1[k])                                                  # every rank holds the whole grid
    # equal to the unsharded sweep, bit for bit
    N = 24
    u, v, V, w, dht = fo.synthetic_disc(3000, N, seed=5)
    m = fo.map_visibilities(dht, u, v, V, w, 30., 40., 1e-3, -2e-3)
    grid = [(a, ws) for a in (1.05, 1.2, 1.4) for ws in (1e-3, 1e-2, 1e-1)]
    for i, (a, ws) in enumerate(grid):
        f = fo.frank_fit(dht, m['M'], m['j'], alpha=a, weights_smooth=ws, max_iter=60)
        assert np.array_equal(r0['mu'][i], f['MAP']) and np.array_equal(r0['p'][i], f['power_spectrum'])
        assert int(r0['niter'][i]) == f['num_iterations'] and bool(r0['converged'][i]) == f['converged']


def test_sweep_sharded_single_process():
    from frank_b200.distributed import sweep_sharded, sweep_shard
    assert list(sweep_shard(64, 3, 8)) == list(range(3, 64, 8))
    res = sweep_sharded(lambda idx: {'p': np.outer(idx, np.ones(3)), 'mu': np.outer(idx, 2 * np.ones(3)),
                                     'niter': idx * 10, 'converged': idx % 2}, 5, 3)
    assert np.array_equal(res['p'][:, 0], np.arange(5)) and np.array_equal(res['niter'], np.arange(5) * 10)
    assert list(res['converged']) == [False, True, False, True, False]
