"""A/B timing of the Gram kernel for the library named by FRANK_B200_LIB (dev tool): python scripts/ab_gram.py [N[:n_vis] ...]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import frank_oracle as fo
from frank_b200 import _lib
from frank_b200.hankel import DiscreteHankelTransform
from frank_b200.geometry import FixedGeometry
from frank_b200.statistical_models import VisibilityMapping
from frank_b200.constants import rad_to_arcsec
import bench

g = FixedGeometry(*bench.GEOM)
res = {'lib': os.path.basename(_lib.LIB_PATH)}
# parity on a small case first
u, v, V, w, odht = fo.synthetic_disc(20000, 300, seed=3)
vm = VisibilityMapping(DiscreteHankelTransform(1.6 / rad_to_arcsec, 300), g, verbose=False)
m = vm.map_visibilities(u, v, V, w)
o = fo.map_visibilities(odht, u, v, V, w, *bench.GEOM)
d = np.sqrt(np.diag(o['M']))
res['parity_cs'] = float(np.max(np.abs(m['M'] - o['M']) / np.outer(d, d)))
# arguments: N or N:n_vis (default 1e7 visibilities)
for arg in sys.argv[1:] or ['300']:
    N, n = (int(float(x)) for x in (arg.split(':') + ['1e7'])[:2])
    dht = DiscreteHankelTransform(1.6 / rad_to_arcsec, N)
    vm = VisibilityMapping(dht, g, verbose=False)
    ud, vd, Vd, wd = bench.synthetic_visibilities_device(n, dht, seed=1)
    t = []
    for _ in range(5):
        vm.map_visibilities(ud, vd, Vd, wd)
        t.append(vm.last_timing['gram_ms'])
    res[f'gram_ms_N{N}_n{n}'] = [round(float(x), 3) for x in t]
print(json.dumps(res))
