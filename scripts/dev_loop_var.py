"""Run-to-run variation of the solver loop (same M, j): forked vs serial graph, device time vs wall time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from frank_b200.constants import rad_to_arcsec
from frank_b200.geometry import FixedGeometry
from frank_b200.hankel import DiscreteHankelTransform
from frank_b200.radial_fitters import FrankFitter

os.environ['FB_SOLVER_TRACE'] = '1'
dht = DiscreteHankelTransform(1.6 / rad_to_arcsec, 300)
u, v, V, w = bench.synthetic_visibilities_device(10_000_000, dht, 12345)
geom = FixedGeometry(*bench.GEOM)
FF = FrankFitter(1.6, 300, geom, alpha=1.05, weights_smooth=1e-4, verbose=False, store_iteration_diagnostics=True)
pre = FF.preprocess_visibilities(u, v, V, w)
for diag in (False,):
    FF = FrankFitter(1.6, 300, geom, alpha=1.05, weights_smooth=1e-4, verbose=False, store_iteration_diagnostics=diag)
    for fork in ("1",):
        os.environ['FB_SOLVER_FORK'] = fork
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); FF.fit_preprocessed(pre); ts.append(time.perf_counter() - t0)
        print(f'diagnostics={diag} fork={fork} wall s:', ' '.join(f'{t:.3f}' for t in ts), flush=True)
