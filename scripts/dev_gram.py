"""Development diagnostics for the mapping path on a B200 (run under gpurun). Prints parity and timing."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from oracle import frank_oracle as fo
from frank_b200 import _lib
from frank_b200.hankel import DiscreteHankelTransform
from frank_b200.geometry import FixedGeometry
from frank_b200.statistical_models import VisibilityMapping
from frank_b200.constants import rad_to_arcsec

def relerr(a, b):
    with np.errstate(divide='ignore', invalid='ignore'):
        r = np.abs(a - b) / np.abs(b)
    return np.nanmax(r), np.max(np.abs(a - b)) / np.max(np.abs(b))

def run(n, N, check=True, reps=1):
    u, v, V, w, odht = fo.synthetic_disc(n, N, analytic=(n > 200000))
    g = FixedGeometry(30., 40., 1e-3, -2e-3)
    dht = DiscreteHankelTransform(1.6 / rad_to_arcsec, N)
    vm = VisibilityMapping(dht, g, verbose=False)
    t = time.time(); m = vm.map_visibilities(u, v, V, w); th = time.time() - t
    print(f"n={n} N={N} host-path wall {th*1e3:.1f} ms timing {vm.last_timing}")
    ud, vd, Vd, wd = [torch.from_numpy(x).cuda() for x in (u, v, V, w)]
    for _ in range(reps):
        torch.cuda.synchronize(); t = time.time(); md = vm.map_visibilities(ud, vd, Vd, wd); td = time.time() - t
        tm = vm.last_timing
        print(f"   dev-path wall {td*1e3:.1f} ms  prep {tm['prep_ms']:.3f} gram {tm['gram_ms']:.3f} fin {tm['finalize_ms']:.3f} ms"
              f" -> {n*N/tm['gram_ms']/1e6:.2f} Gvis.mode/s (gram only)")
    print("   host vs dev identical:", np.array_equal(m['M'], md['M']), np.array_equal(m['j'], md['j']))
    if check:
        t = time.time(); o = fo.map_visibilities(odht, u, v, V, w, 30., 40., 1e-3, -2e-3); to = time.time() - t
        pe, mx = relerr(m['M'], o['M']); pj, mj = relerr(m['j'], o['j'])
        print(f"   oracle {to:.2f} s | M per-entry rel {pe:.3e} max-norm {mx:.3e} | j per-entry {pj:.3e} max-norm {mj:.3e} | "
              f"H0 rel {abs(m['null_likelihood']-o['null_likelihood'])/abs(o['null_likelihood']):.3e} | sym {np.abs(m['M']-m['M'].T).max()}")
        a, kz, Vre, perm = _lib.get_context().debug_prepped(n)
        aq = o['q'] * (1. / odht.Qmax)
        isperm = np.array_equal(np.sort(perm), np.arange(n))
        print(f"   prep: perm valid {isperm} sorted-inversions {np.sum(np.diff(a) < -a.max()/65000)} a bit-equal {np.mean(a == aq[perm]):.6f} kz bit-equal {np.mean(kz == o['k'][perm]):.6f} Vre max rel {np.max(np.abs(Vre-o['Vre'][perm])/np.abs(o['Vre']).max()):.2e}")
    return m

if __name__ == '__main__':
    ctx = _lib.get_context()
    dht = DiscreteHankelTransform(1.6 / rad_to_arcsec, 300)
    ctx.dht_setup(dht)
    import ctypes, subprocess
    subprocess.check_call(['make', '-s', '-C', 'oracle'])
    lib = ctypes.CDLL('oracle/_build/liboracle_j0.so')
    x = np.random.default_rng(0).uniform(0, 940, 2000000)
    ex = np.empty_like(x); lib.oracle_j0_exact_array(x.ctypes.data_as(ctypes.c_void_p), ex.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(x.size))
    got = ctx.debug_j0(x)
    import scipy.special
    print(f"J0 device vs exact: max {np.abs(got-ex).max():.2e} rms {np.sqrt(np.mean((got-ex)**2)):.2e}; vs scipy max {np.abs(got-scipy.special.j0(x)).max():.2e}")
    run(6000, 60)
    run(5000, 300)
    run(100000, 300)
    run(200000, 500)
    run(1000000, 300, check=False, reps=3)
    run(10000000, 300, check=False, reps=3)
