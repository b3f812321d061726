"""BASELINE.json configs[3]: 64 (alpha, wsmooth) pairs x power-spectrum iteration at N = 300, sharded by grid point.

    python scripts/run_sweep.py                                        # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node R --master-addr 127.0.0.1 scripts/run_sweep.py

Every rank maps the same 1e6 visibilities (M, j do not depend on the hyper-parameters), solves its 64 / R grid points as one
batched device loop and the results are all-gathered through the library's NCCL communicator.  Prints one JSON line and, on
R > 1 ranks, checks that the gathered grid equals the single-rank batch bit for bit."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from frank_b200 import _lib, distributed
from frank_b200.constants import rad_to_arcsec
from frank_b200.geometry import FixedGeometry
from frank_b200.hankel import DiscreteHankelTransform
from frank_b200.radial_fitters import FrankFitter

rank, world, local = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1)), int(os.environ.get('LOCAL_RANK', 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
N = 300
geom = FixedGeometry(*bench.GEOM)
dht = DiscreteHankelTransform(1.6 / rad_to_arcsec, N)
u, v, V, w = bench.synthetic_visibilities_device(1_000_000, dht, seed=1)
FF = FrankFitter(1.6, N, geom, verbose=False, convergence_failure='ignore', device=local)
pre = FF.preprocess_visibilities(u, v, V, w)
alphas, wss = np.linspace(1.01, 1.5, 8), np.logspace(-4, -1, 8)
ctx = _lib.get_context(local)
# single-rank batch first (reference for the bit-for-bit check and for the scaling ratio)
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    one = FF.fit_sweep_preprocessed(pre, alphas, wss, group=None, on_cholesky_failure='flag') if world == 1 else None
    t_one = time.perf_counter() - t0
if world > 1:
    # every rank alone (no process group visible to the sweep): the 64-point batch on one GPU
    import frank_b200.distributed as fd
    saved = fd._dist
    fd._dist = lambda: None
    for _ in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        one = FF.fit_sweep_preprocessed(pre, alphas, wss, on_cholesky_failure='flag')
        t_one = time.perf_counter() - t0
    fd._dist = saved
    distributed.init_library_comm(ctx)
    for _ in range(2):
        dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
        sols = FF.fit_sweep_preprocessed(pre, alphas, wss, on_cholesky_failure='flag')
        torch.cuda.synchronize(); t_sh = time.perf_counter() - t0
    tt = torch.tensor([t_sh], dtype=torch.float64, device='cuda')
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    same = all(np.array_equal(a.MAP, b.MAP) and np.array_equal(a.power_spectrum, b.power_spectrum) for a, b in zip(one, sols))
    it = np.array(FF.sweep_diagnostics['num_iterations'])
    if rank == 0:
        print(json.dumps({'config': '4 sharded by grid point', 'ranks': world, 'grid_points': 64, 'N': N, 'sweep_s_one_rank': t_one,
                          'sweep_s_sharded_max_over_ranks': float(tt.item()), 'speedup': t_one / float(tt.item()),
                          'equals_single_rank_batch_bit_for_bit': bool(same), 'iterations_sum': int(it.sum()),
                          'gather': 'library NCCL all-gather (fb_comm_allgather)'}))
    ctx.comm_destroy()
    dist.barrier()
    dist.destroy_process_group()
else:
    it = np.array(FF.sweep_diagnostics['num_iterations'])
    print(json.dumps({'config': '4 one rank', 'ranks': 1, 'grid_points': 64, 'N': N, 'sweep_s_one_rank': t_one, 'iterations_sum': int(it.sum()),
                      'us_per_point_iteration': t_one / it.sum() * 1e6}))
