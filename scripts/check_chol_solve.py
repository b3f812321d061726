"""One-off device check of fb_chol_solve against scipy.linalg.cho_solve (run under gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.linalg
from frank_b200 import _lib
from frank_b200.constants import rad_to_arcsec
from frank_b200.hankel import DiscreteHankelTransform

ctx = _lib.get_context(0)
rng = np.random.default_rng(3)
worst = 0.0
for N in (40, 100, 300, 417, 500):
    ctx.dht_setup(DiscreteHankelTransform(1.6 / rad_to_arcsec, N))
    G = rng.standard_normal((N, N))
    A = G @ G.T + N * np.eye(N)
    U = np.triu(scipy.linalg.cho_factor(A)[0])
    for b in (rng.standard_normal(N), rng.standard_normal((N, 7)), np.eye(N)):
        ref = scipy.linalg.cho_solve((U, False), b)
        got = ctx.chol_solve(U, b)
        err = np.max(np.abs(got - ref)) / np.max(np.abs(ref))
        worst = max(worst, err)
        print(N, b.shape, err, flush=True)
        assert got.shape == ref.shape
print('CHOL_SOLVE_OK' if worst < 1e-12 else 'CHOL_SOLVE_BAD', worst)
