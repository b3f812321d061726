"""Small mappings for compute-sanitizer (dev tool): opt_thick N=64 and N=300, debris multi-channel N=96, sparse and dense tiles."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import frank_oracle as fo
from frank_b200.hankel import DiscreteHankelTransform
from frank_b200.geometry import FixedGeometry
from frank_b200.statistical_models import VisibilityMapping
from frank_b200.constants import rad_to_arcsec
import bench

g = FixedGeometry(*bench.GEOM)
for N, n in ((64, 3000), (300, 20000), (300, 700)):
    u, v, V, w, odht = fo.synthetic_disc(n, N, seed=3)
    vm = VisibilityMapping(DiscreteHankelTransform(1.6 / rad_to_arcsec, N), g, verbose=False)
    m = vm.map_visibilities(u, v, V, w)
    o = fo.map_visibilities(odht, u, v, V, w, *bench.GEOM)
    d = np.sqrt(np.diag(o['M']))
    print(N, n, 'opt_thick', float(np.max(np.abs(m['M'] - o['M']) / np.outer(d, d))))
N, n = 96, 6000
u, v, V, w, odht = fo.synthetic_disc(n, N, seed=5)
freq = np.repeat([1.0e11, 1.1e11, 1.2e11], n // 3)
vm = VisibilityMapping(DiscreteHankelTransform(1.6 / rad_to_arcsec, N), g, vis_model='debris',
                       scale_height=lambda R: 0.05 * R, verbose=False)
m = vm.map_visibilities(u, v, V, w, frequencies=freq)
print(N, n, 'debris, 3 channels', m['M'].shape, float(np.abs(m['M']).max()))
