"""Development diagnostics for the solver path on a B200 (run under gpurun)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import frank_oracle as fo
from frank_b200 import _lib
from frank_b200.hankel import DiscreteHankelTransform
from frank_b200.geometry import FixedGeometry
from frank_b200.filter import CriticalFilter
from frank_b200.radial_fitters import FrankFitter
from frank_b200.constants import rad_to_arcsec

g = np.load('tests/golden/mapping.npz'); f = np.load('tests/golden/fit_normal.npz')
N = int(g['N'])
dht = DiscreteHankelTransform(1.6 / rad_to_arcsec, N)
geom = FixedGeometry(*[float(x) for x in g['geom']])
ctx = _lib.get_context(); ctx.dht_setup(dht)
FF = FrankFitter(1.6, N, geom, verbose=False)
FF._build_matrices({'hash': [False, dht, geom, 'opt_thick', None], 'M': g['M_opt_thick'], 'j': g['j_opt_thick'], 'null_likelihood': 0.0})
p_init = FF._starting_spectrum()
out = ctx.frank_normal_loop(g['M_opt_thick'], g['j_opt_thick'], p_init, 1.05, 1e-15, FF._filter._Tinv, 1e-3, 2000)
peak = np.abs(f['MAP']).max()
print('solver on reference M,j: iters', out['niter'][0], int(f['num_iterations']), 'MAP diff/peak', np.abs(out['mu'][0] - f['MAP']).max() / peak,
      'p rel', np.abs(out['p'][0] / f['power_spectrum'] - 1).max())
sol = FrankFitter(1.6, N, geom, verbose=False).fit(g['u'], g['v'], g['V'], g['w'])
print('full path: MAP diff/peak', np.abs(sol.MAP - f['MAP']).max() / peak)

# timing of the solver loop at N = 300 (config 1/2 shape)
for n, NN in [(200000, 300), (200000, 500)]:
    u, v, V, w, odht = fo.synthetic_disc(n, NN, analytic=True)
    FF = FrankFitter(1.6, NN, FixedGeometry(30., 40., 1e-3, -2e-3), verbose=False, store_iteration_diagnostics=False)
    t = time.time(); pre = FF.preprocess_visibilities(u, v, V, w); t1 = time.time() - t
    t = time.time(); sol = FF.fit_preprocessed(pre); t2 = time.time() - t
    t = time.time(); sol = FF.fit_preprocessed(pre); t2b = time.time() - t
    FF2 = FrankFitter(1.6, NN, FixedGeometry(30., 40., 1e-3, -2e-3), verbose=False, store_iteration_diagnostics=True)
    FF2.fit_preprocessed(pre)
    it = FF2.iteration_diagnostics['num_iterations']
    print(f'n={n} N={NN}: map {t1*1e3:.1f} ms, solver loop {t2*1e3:.1f} / {t2b*1e3:.1f} ms, {it} iterations -> {t2b/it*1e6:.1f} us/iter')
    if NN == 300:
        t = time.time(); o = fo.map_visibilities(odht, u, v, V, w, 30., 40., 1e-3, -2e-3); r = fo.frank_fit(odht, o['M'], o['j']); to = time.time() - t
        print('   oracle: iterations', r['num_iterations'], 'MAP diff/peak', np.abs(sol.MAP - r['MAP']).max() / np.abs(r['MAP']).max(), f'({to:.1f} s)')
