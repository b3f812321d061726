"""Does a grid point's result depend on the batch it is solved in?  (dev tool)"""
import os, sys, time, json
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from frank_b200 import _lib
from frank_b200.constants import rad_to_arcsec
from frank_b200.filter import CriticalFilter
from frank_b200.geometry import FixedGeometry
from frank_b200.hankel import DiscreteHankelTransform
from frank_b200.radial_fitters import FrankFitter

N = 300
geom = FixedGeometry(*bench.GEOM)
dht = DiscreteHankelTransform(1.6 / rad_to_arcsec, N)
u, v, V, w = bench.synthetic_visibilities_device(1_000_000, dht, seed=1)
FF = FrankFitter(1.6, N, geom, verbose=False, convergence_failure='ignore')
pre = FF.preprocess_visibilities(u, v, V, w)
FF._build_matrices(pre)
pI = FF._starting_spectrum()
grid = [(float(a), float(ws)) for a in np.linspace(1.01, 1.5, 8) for ws in np.logspace(-4, -1, 8)]
ctx = _lib.get_context()
filters = [CriticalFilter(dht, a, 1e-15, ws, 1e-3) for a, ws in grid]

def run(idx):
    out = ctx.frank_normal_loop(FF._M, FF._j, np.tile(pI, (len(idx), 1)), [filters[i]._alpha for i in idx], [filters[i]._p_0 for i in idx],
                                np.stack([filters[i]._Tinv for i in idx]), 1e-3, 2000, want_chol=False)
    return out

t0 = time.perf_counter(); full = run(list(range(64))); t_full = time.perf_counter() - t0
full2 = run(list(range(64)))
print('B=64: info nonzero at', np.nonzero(full['info'])[0].tolist(), 'info', full['info'][np.nonzero(full['info'])[0]].tolist(), 'niter', full['niter'][np.nonzero(full['info'])[0]].tolist(), f'{t_full:.2f} s')
print('B=64 repeat identical:', np.array_equal(full['p'], full2['p']), np.array_equal(full['info'], full2['info']))
for B in (8, 1):
    bad, diff = [], []
    t0 = time.perf_counter()
    for r in range(0, 64 // B if B > 1 else 8):
        idx = list(range(r, 64, 8)) if B == 8 else [r * 8]
        o = run(idx)
        for k, i in enumerate(idx):
            if o['info'][k]: bad.append(i)
            if not (np.array_equal(o['p'][k], full['p'][i]) and np.array_equal(o['mu'][k], full['mu'][i]) and o['niter'][k] == full['niter'][i]):
                diff.append((i, int(o['niter'][k]), int(full['niter'][i]), float(np.max(np.abs(o['p'][k] - full['p'][i]) / full['p'][i]))))
    print(f'B={B}: info nonzero at', bad, 'points differing from the B=64 run:', diff, f'{time.perf_counter() - t0:.2f} s')
