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


def test_shard_bounds_cover():
    from frank_b200.distributed import shard_bounds
    for n in [0, 1, 7, 64, 1000003]:
        for world in [1, 2, 3, 8]:
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1
