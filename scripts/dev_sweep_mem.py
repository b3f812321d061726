import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import frank_oracle as fo
from frank_b200 import _lib
from frank_b200.constants import rad_to_arcsec
from frank_b200.filter import CriticalFilter
from frank_b200.hankel import DiscreteHankelTransform
N, B, iters = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
u, v, V, w, odht = fo.synthetic_disc(20000, N, seed=3)
m = fo.map_visibilities(odht, u, v, V, w, 30., 40., 1e-3, -2e-3)
dht = DiscreteHankelTransform(1.6 / rad_to_arcsec, N)
ctx = _lib.get_context(); ctx.dht_setup(dht)
f = CriticalFilter(dht, 1.3, 1e-15, 1e-2, 1e-3)
p0 = 1e10 * (dht.q / dht.q[0]) ** -2
out = ctx.frank_normal_loop(m['M'], m['j'], np.tile(p0, (B, 1)), np.full(B, 1.3), np.full(B, 1e-15), np.tile(f._Tinv, (B, 1, 1)), 1e-3, iters, want_chol=False)
same = all(np.array_equal(out['p'][0], out['p'][b]) for b in range(B))
print('B', B, 'info', out['info'].tolist()[-8:], 'niter', out['niter'].tolist()[-8:], 'all problems identical:', same)
if not same:
    print([b for b in range(B) if not np.array_equal(out['p'][0], out['p'][b])])
