"""On-hardware multi-rank tests (need >= 2 GPUs on the box: `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`).

Each rank is one process on one GPU; the library's own NCCL communicator (fb_comm_init) all-reduces the partial normal
equations inside the mapping call.  What is asserted is the KERNELS' result, not plumbing: the all-reduced M, j, H0 of R
ranks equal the single-GPU mapping of the concatenated visibilities within the parity floor, on every rank, and the sweep
sharded by grid point equals the single-rank batch bit for bit."""
import os
import socket
import sys

import numpy as np
import pytest

from oracle import frank_oracle as fo

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _ngpu():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


N_VIS, N_MODES, GEOM = 200_001, 120, (30., 40., 1e-3, -2e-3)


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), LOCAL_RANK=str(rank), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from frank_b200 import _lib, distributed
    from frank_b200.constants import rad_to_arcsec
    from frank_b200.geometry import FixedGeometry
    from frank_b200.hankel import DiscreteHankelTransform
    from frank_b200.radial_fitters import FrankFitter
    from frank_b200.statistical_models import VisibilityMapping
    u, v, V, w, _ = fo.synthetic_disc(N_VIS, N_MODES, seed=21)
    freqs = np.random.default_rng(3).choice(np.array([1., 2.]), N_VIS)
    geom = FixedGeometry(*GEOM)
    dht = DiscreteHankelTransform(1.6 / rad_to_arcsec, N_MODES)
    vm = VisibilityMapping(dht, geom, verbose=False, device=rank)
    out = {}
    # (1) torch.distributed all-reduce of per-rank mappings (no library communicator yet)
    m = distributed.map_visibilities_sharded(vm, u, v, V, w)
    out.update(M_torch=m['M'], j_torch=m['j'], H0_torch=m['null_likelihood'])
    # (2) the library's communicator: all-reduce inside the device call, host and device entry points, multi-channel
    ctx = distributed.init_library_comm(_lib.get_context(rank))
    assert ctx.comm_info() == (True, rank, world)
    m = distributed.map_visibilities_sharded(vm, u, v, V, w)
    out.update(M_lib=m['M'], j_lib=m['j'], H0_lib=m['null_likelihood'])
    lo, hi = distributed.shard_bounds(N_VIS, rank, world)
    td = [torch.from_numpy(x[lo:hi]).cuda() for x in (u, v, V, w)]
    md = vm.map_visibilities(*td)
    out.update(M_dev=md['M'], j_dev=md['j'], H0_dev=md['null_likelihood'])
    mm = distributed.map_visibilities_sharded(vm, u, v, V, w, frequencies=freqs)
    out.update(M_multi=mm['M'], j_multi=mm['j'], H0_multi=mm['null_likelihood'])
    # (3) the global q-range check: only the last rank's slice reaches beyond the collocation range
    u2 = u.copy()
    u2[-10:] *= 3.0
    try:
        distributed.map_visibilities_sharded(vm, u2, v, V, w)
        out['raised'] = 0
    except ValueError:
        out['raised'] = 1
    # (4) hyper-parameter sweep sharded by grid point over the library's all-gather
    FF = FrankFitter(1.6, N_MODES, geom, verbose=False, device=rank, max_iter=300, convergence_failure='ignore')
    pre = {'M': out['M_lib'], 'j': out['j_lib'], 'null_likelihood': out['H0_lib'], 'hash': [False, dht, geom, 'opt_thick', None]}
    sols = FF.fit_sweep_preprocessed(pre, alphas=[1.05, 1.2, 1.4], weights_smooths=[1e-3, 1e-2, 1e-1])
    out['sweep_MAP'] = np.array([s.MAP for s in sols])
    out['sweep_p'] = np.array([s.power_spectrum for s in sols])
    out['sweep_niter'] = np.array(FF.sweep_diagnostics['num_iterations'])
    ctx.comm_destroy()
    np.savez(os.path.join(out_dir, f'r{rank}.npz'), **out)
    dist.destroy_process_group()


@pytest.mark.skipif(_ngpu() < 2, reason="needs two GPUs")
def test_two_rank_allreduce_equals_single_gpu(tmp_path):
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [np.load(tmp_path / f'r{k}.npz') for k in range(world)]
    # single-GPU result on the concatenated data, in this process
    from frank_b200.constants import rad_to_arcsec
    from frank_b200.geometry import FixedGeometry
    from frank_b200.hankel import DiscreteHankelTransform
    from frank_b200.radial_fitters import FrankFitter
    from frank_b200.statistical_models import VisibilityMapping
    u, v, V, w, odht = fo.synthetic_disc(N_VIS, N_MODES, seed=21)
    freqs = np.random.default_rng(3).choice(np.array([1., 2.]), N_VIS)
    geom = FixedGeometry(*GEOM)
    dht = DiscreteHankelTransform(1.6 / rad_to_arcsec, N_MODES)
    vm = VisibilityMapping(dht, geom, verbose=False, device=0)
    one = vm.map_visibilities(u, v, V, w)
    onem = vm.map_visibilities(u, v, V, w, frequencies=freqs)
    ref = fo.map_visibilities(odht, u, v, V, w, *GEOM)
    d = np.sqrt(np.diag(one['M']))
    eps = np.finfo(float).eps
    for tag in ('torch', 'lib', 'dev'):
        assert np.array_equal(r[0][f'M_{tag}'], r[1][f'M_{tag}']) and np.array_equal(r[0][f'j_{tag}'], r[1][f'j_{tag}'])
        M = r[0][f'M_{tag}']
        assert np.max(np.abs(M - one['M']) / np.outer(d, d)) < 64 * eps, tag          # a different summation order, nothing more
        assert np.max(np.abs(M - ref['M']) / (1e-10 * np.abs(ref['M']) + 16 * eps * np.outer(d, d))) <= 1.0, tag
        assert np.max(np.abs(r[0][f'j_{tag}'] - one['j'])) <= 1e-13 * np.max(np.abs(one['j']))
        assert abs(float(r[0][f'H0_{tag}']) - one['null_likelihood']) <= 1e-13 * abs(one['null_likelihood'])
    assert np.array_equal(r[0]['M_lib'], r[0]['M_dev'])                                # host and device entry points agree
    for c in range(2):
        dc = np.sqrt(np.diag(onem['M'][c]))
        assert np.max(np.abs(r[0]['M_multi'][c] - onem['M'][c]) / np.outer(dc, dc)) < 64 * eps
    assert np.array_equal(r[0]['M_multi'], r[1]['M_multi'])
    assert int(r[0]['raised']) == 1 and int(r[1]['raised']) == 1                        # the range error is global
    # sweep: sharded == single-rank batch, bit for bit
    FF = FrankFitter(1.6, N_MODES, geom, verbose=False, device=0, max_iter=300, convergence_failure='ignore')
    pre = {'M': r[0]['M_lib'], 'j': r[0]['j_lib'], 'null_likelihood': float(r[0]['H0_lib']), 'hash': [False, dht, geom, 'opt_thick', None]}
    sols = FF.fit_sweep_preprocessed(pre, alphas=[1.05, 1.2, 1.4], weights_smooths=[1e-3, 1e-2, 1e-1])
    for k in range(world):
        assert np.array_equal(r[k]['sweep_MAP'], np.array([s.MAP for s in sols]))
        assert np.array_equal(r[k]['sweep_p'], np.array([s.power_spectrum for s in sols]))
        assert np.array_equal(r[k]['sweep_niter'], np.array(FF.sweep_diagnostics['num_iterations']))
