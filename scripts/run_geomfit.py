"""Time FitGeometryFourierBessel on synthetic visibilities: python scripts/run_geomfit.py <n_vis>."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from frank_b200.constants import rad_to_arcsec, deg_to_rad
from frank_b200.geometry import FixedGeometry, FitGeometryFourierBessel

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1000000
rng = np.random.default_rng(777)
gt = FixedGeometry(32.0, 47.0, dRA=0.021, dDec=-0.034)
q = 1.2e6 * np.sqrt(rng.uniform(1e-4, 1, n))
th = rng.uniform(0, 2 * np.pi, n)
ud, vd = q * np.cos(th), q * np.sin(th)
sig = 0.25 / rad_to_arcsec
Vd = np.cos(32.0 * deg_to_rad) * 2 * np.pi * sig * sig * 4e10 * np.exp(-2 * np.pi ** 2 * sig * sig * q * q)
u, v, V = gt.undo_correction(ud, vd, Vd.astype(complex))
w = np.full(n, 2.5e3)
V = V + (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(w)
# one-off process start-up (CUDA context, library load, pinned allocations) is paid by a small warm-up fit, as bench.py does
FitGeometryFourierBessel(1.6, 20, guess=[28., 44., 0.015, -0.03], solver='device').fit(u[:5000], v[:5000], V[:5000], w[:5000])
gf = FitGeometryFourierBessel(1.6, 20, guess=[28., 44., 0.015, -0.03])
t0 = time.perf_counter()
gf.fit(u, v, V, w)
dt = time.perf_counter() - t0
print(json.dumps({'what': 'FitGeometryFourierBessel(1.6, 20)', 'n_vis': n, 'fit_s': dt, 'residual_evaluations': int(gf._nfev),
                  's_per_evaluation': dt / gf._nfev, 'inc': gf.inc, 'PA': gf.PA, 'dRA': gf.dRA, 'dDec': gf.dDec,
                  'truth': [32.0, 47.0, 0.021, -0.034]}))
