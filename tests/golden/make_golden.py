#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (discsim/frank 1.2.3).

Run in the build container only (the GPU box has no /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Inputs are seeded synthetic visibilities (generated with NumPy only, stored in the
fixture so tests never need the reference) plus the 196-row AS 209 subsample shipped in
the reference's docs (docs/tutorials/test_datafile.txt, data not source).  Outputs are
whatever the reference's public API returns for them:

  j0_golden.npz      scipy.special.j0 on sample arguments (the reference's J0, hankel.py:23)
  dht.npz            DiscreteHankelTransform tables / coefficients / transform (hankel.py)
  mapping.npz        VisibilityMapping.map_visibilities M, j, H0 for opt_thick / opt_thin /
                     debris / multi-channel (statistical_models.py:109-237)
  fit_normal.npz     FrankFitter(method='Normal').fit  MAP, power spectrum, iterations
  fit_lognormal.npz  FrankFitter(method='LogNormal').fit
  fit_as209sub.npz   FrankFitter on the 196-visibility AS 209 subsample, N=20
  uvbin.npz          UVDataBinner bin indices, counts, means, errors (utilities.py:180-400)
  estweights.npz     estimate_weights in its five call forms (utilities.py:515-631)
  geomfit.npz        FitGeometryGaussian / FitGeometryFourierBessel results (geometry.py:404-763)
  svd_fallback.npz   GaussianModel solutions through the reference's SVD pseudo-inverse branch (indefinite and
                     rank-deficient systems, statistical_models.py:747-755)
  gauss_kat.npz      the reference's analytic Gaussian Hankel-pair test inputs (tests.py:37-130)
  config1_normal_1e6_N300.npz   BASELINE.json configs[0]: FrankFitter Normal fit, 1e6 visibilities, N=300, alpha=1.05,
                     wsmooth=1e-4 -- the reference's M, j, H0, MAP, power spectrum, iteration count and its own
                     self-noise (M under permutation / block_size, profile under permutation).  Inputs are NOT stored:
                     they are regenerated bit for bit by oracle.frank_oracle.synthetic_disc (NumPy/SciPy only; input
                     checksums stored)
  config3_lognormal_N500.npz    configs[2] shape: LogNormal MAP fit, N=500, Rmax=1.0", alpha=1.3, wsmooth=1e-2 on 5e4 visibilities
                     (at Rmax=1.6" the reference itself aborts, see gen_config3):
                     M (upper triangle), j, H0, s_MAP, MAP, power spectrum, iteration count, Newton statistics
  config5_debris_N2000.npz      configs[4] shape: N=2000, 4 channels, debris scale height, 2e4 visibilities:
                     j, H0, channels, the diagonal and 6000 seeded sample entries of every channel's M
"""
import os
import sys
import warnings

import numpy as np

REF = '/root/reference'
sys.path.insert(0, REF)
sys.dont_write_bytecode = True
warnings.filterwarnings('ignore')

import scipy.special  # noqa: E402
import frank  # noqa: E402
from frank.constants import rad_to_arcsec, deg_to_rad  # noqa: E402
from frank.geometry import FixedGeometry, FitGeometryGaussian, FitGeometryFourierBessel  # noqa: E402
from frank.hankel import DiscreteHankelTransform  # noqa: E402
from frank.radial_fitters import FrankFitter, FourierBesselFitter  # noqa: E402
from frank.debris_fitters import FrankDebrisFitter  # noqa: E402
from frank.statistical_models import VisibilityMapping  # noqa: E402
from frank.utilities import UVDataBinner, estimate_weights  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
assert frank.__version__ == '1.2.3', frank.__version__


def synthetic(n_vis, N, Rmax=1.6, seed=12345, geom=(30., 40., 1e-3, -2e-3)):
    """Seeded Gaussian-ring disc of SURVEY.md section 8(d), built through the reference."""
    rng = np.random.default_rng(seed)
    g = FixedGeometry(*geom)
    dht = DiscreteHankelTransform(Rmax / rad_to_arcsec, N)
    vm = VisibilityMapping(dht, g, verbose=False)
    q = 0.98 * dht.q[-1] * np.sqrt(rng.uniform(1e-5, 1, n_vis))
    th = rng.uniform(0, 2 * np.pi, n_vis)
    ud, vd = q * np.cos(th), q * np.sin(th)
    u, v = g.reproject(ud, vd)
    r = dht.r * rad_to_arcsec
    I = 1e10 * np.exp(-0.5 * ((r - 0.6) / 0.08) ** 2) + 3e9 * np.exp(-0.5 * (r / 0.3) ** 2)
    Vd = vm.predict_visibilities(I, q, k=None)
    _, _, V = g.undo_correction(ud, vd, Vd.astype(complex))
    w = 1e4 * rng.uniform(0.5, 2, n_vis)
    V = V + (rng.standard_normal(n_vis) + 1j * rng.standard_normal(n_vis)) / np.sqrt(w)
    return u, v, V, w, g


def self_noise(make_fitter, u, v, V, w, MAP, nperm=3):
    """How far the REFERENCE's fitted profile moves (relative to its peak) when the same visibilities are
    presented in another order or accumulated with another block_size -- the round-off floor any
    re-implementation inherits (the posterior precision matrices have condition numbers 1e7..1e10)."""
    worst = 0.0
    for k in range(nperm):
        perm = np.random.default_rng(1000 + k).permutation(len(u))
        sol = make_fitter({}).fit(u[perm], v[perm], V[perm], w[perm] if np.ndim(w) else w)
        worst = max(worst, np.max(np.abs(sol.MAP - MAP)) / np.max(np.abs(MAP)))
    sol = make_fitter({'block_size': 20000}).fit(u, v, V, w)
    worst = max(worst, np.max(np.abs(sol.MAP - MAP)) / np.max(np.abs(MAP)))
    return worst


def gen_geomfit():
    """Geometry fitters (geometry.py:404-763) on a noisy Gaussian disc seen at inc=32, PA=47 with an offset source."""
    rng = np.random.default_rng(777)
    n = 3000
    gt = FixedGeometry(32.0, 47.0, dRA=0.021, dDec=-0.034)
    q = 1.2e6 * np.sqrt(rng.uniform(1e-4, 1, n))
    th = rng.uniform(0, 2 * np.pi, n)
    ud, vd = q * np.cos(th), q * np.sin(th)
    sig = 0.25 / rad_to_arcsec
    Vd = np.cos(32.0 * deg_to_rad) * 2 * np.pi * sig * sig * 4e10 * np.exp(-2 * np.pi ** 2 * sig * sig * q * q)
    u, v, V = gt.undo_correction(ud, vd, Vd.astype(complex))
    w = np.full(n, 2.5e3)
    V = V + (rng.standard_normal(n) + 1j * rng.standard_normal(n)) / np.sqrt(w)
    gg = FitGeometryGaussian()
    gg.fit(u, v, V, w)
    gg2 = FitGeometryGaussian(phase_centre=(0.021, -0.034), guess=[20., 60., 0., 0.])
    gg2.fit(u, v, V, w)
    gf = FitGeometryFourierBessel(1.6, 20, guess=[28., 44., 0.015, -0.03])
    gf.fit(u, v, V, w)
    gf2 = FitGeometryFourierBessel(1.6, 20, inc_pa=(32.0, 47.0), guess=[0., 0., 0.015, -0.03])
    gf2.fit(u, v, V, w)
    np.savez_compressed(os.path.join(OUT, 'geomfit.npz'), u=u, v=v, V=V, w=w,
                        gauss=np.array([gg.inc, gg.PA, gg.dRA, gg.dDec]),
                        gauss_fixed_centre=np.array([gg2.inc, gg2.PA, gg2.dRA, gg2.dDec]),
                        fb=np.array([gf.inc, gf.PA, gf.dRA, gf.dDec]),
                        fb_fixed_incpa=np.array([gf2.inc, gf2.PA, gf2.dRA, gf2.dDec]))
    print('geomfit:', gg.inc, gg.PA, gg.dRA, gg.dDec, '|', gf.inc, gf.PA, gf.dRA, gf.dDec, '|', gf2.dRA, gf2.dDec)


def gen_estweights():
    # ---- estimate_weights (utilities.py:515-631) on deprojected baselines ---------------------------
    rng_w = np.random.default_rng(2024)
    ne = 20000
    ue = rng_w.uniform(-2e6, 2e6, ne)
    ve = rng_w.uniform(-2e6, 2e6, ne)
    sig = 0.02 * (1 + np.hypot(ue, ve) / 1e6)                          # baseline-dependent noise
    Ve = np.exp(-(np.hypot(ue, ve) / 8e5) ** 2) + sig * (rng_w.standard_normal(ne) + 1j * 2 * rng_w.standard_normal(ne))
    ew = {}
    ew['log'] = np.ma.filled(estimate_weights(ue, ve, Ve, nbins=300, verbose=False), np.nan)
    ew['lin'] = np.ma.filled(estimate_weights(ue, ve, Ve, nbins=300, log=False, verbose=False), np.nan)
    ew['fine'] = np.ma.filled(estimate_weights(ue, ve, Ve, nbins=8000, verbose=False), np.nan)      # many single-count bins
    ew['median'] = np.ma.filled(estimate_weights(ue, ve, Ve, nbins=300, use_median=True, verbose=False), np.nan)
    ew['uV'] = np.ma.filled(estimate_weights(np.hypot(ue, ve), Ve.real, nbins=100, verbose=False), np.nan)   # (u, V) call form, real V
    np.savez_compressed(os.path.join(OUT, 'estweights.npz'), u=ue, v=ve, V=Ve, **ew)


def gen_svd_fallback():
    """GaussianModel on systems whose Cholesky factorisation fails, so that the reference takes its SVD
    pseudo-inverse branch (statistical_models.py:747-755): (a) a well-conditioned symmetric INDEFINITE M (no prior),
    (b) the same M plus a prior p, still indefinite, (c) a rank-deficient M = H^T W H from fewer visibilities than
    collocation points (FourierBesselFitter, no prior) -- its solution is round-off dominated along the null space,
    so only the fitted visibilities H mu are stored for comparison."""
    from frank.statistical_models import GaussianModel
    N = 40
    dht = DiscreteHankelTransform(1.6 / rad_to_arcsec, N)
    rng = np.random.default_rng(99)
    Q, _ = np.linalg.qr(rng.standard_normal((N, N)))
    lam = rng.uniform(1, 10, N) * np.where(rng.uniform(size=N) < 0.3, -1, 1)
    M = (Q * lam) @ Q.T
    M = 0.5 * (M + M.T)
    j = rng.standard_normal(N)
    ga = GaussianModel(dht, M, j)
    assert ga._Dsvd is not None
    p = np.full(N, 1e3) * rng.uniform(0.5, 2, N)
    gb = GaussianModel(dht, M, j, p=p)
    assert gb._Dsvd is not None
    # (c)
    u, v, V, w, g = synthetic(25, N, seed=5)
    FB = FourierBesselFitter(1.6, N, g, verbose=False)
    sol = FB.fit(u, v, V, w)
    assert sol._fit._Dsvd is not None
    np.savez_compressed(os.path.join(OUT, 'svd_fallback.npz'), N=N, M=M, j=j, p=p, mu_a=ga.mean, s1_a=ga._Dsvd[1],
                        Dj_a=ga.Dsolve(j), mu_b=gb.mean, s1_b=gb._Dsvd[1],
                        u=u, v=v, V=V, w=w, geom=[g.inc, g.PA, g.dRA, g.dDec], Mc=FB._M, jc=FB._j, mu_c=sol.mean,
                        Vfit_c=sol.predict(u, v), s1_c=sol._fit._Dsvd[1])


def gen_lognormal(g):
    """FrankFitter(method='LogNormal') at N = 40 with the whole iteration history (the stopping iteration of this path is
    round-off sensitive: the history lets the tests compare trajectories at equal iteration count)."""
    Nl = 40
    ul, vl, Vl, wl, _ = synthetic(4000, Nl, seed=777)
    FL = FrankFitter(1.6, Nl, g, alpha=1.3, weights_smooth=1e-2, method='LogNormal', verbose=False,
                     store_iteration_diagnostics=True)
    sl = FL.fit(ul, vl, Vl, wl)
    dl = FL.iteration_diagnostics
    np.savez_compressed(os.path.join(OUT, 'fit_lognormal.npz'), N=Nl, u=ul, v=vl, V=Vl, w=wl, M=FL._M, j=FL._j, MAP=sl.MAP, power_spectrum=sl.power_spectrum,
                        num_iterations=dl['num_iterations'], s_MAP=sl._fit.MAP,
                        p_first=np.array(dl['power_spectrum'][:3]), MAP_first=np.array(dl['MAP'][:3]),
                        p_hist=np.array(dl['power_spectrum']), s_hist=np.array(dl['MAP']),
                        self_noise=self_noise(lambda kw: FrankFitter(1.6, Nl, g, alpha=1.3, weights_smooth=1e-2, method='LogNormal',
                                                                     verbose=False, **kw), ul, vl, Vl, wl, sl.MAP, nperm=1))
    return ul, vl, Vl, wl


def _upper(M):
    return M[np.triu_indices(M.shape[0])]


def gen_config1():
    """BASELINE.json configs[0]: the reference's own CPU-runnable case (1e6 visibilities, N=300, Normal)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from oracle import frank_oracle as fo
    import time
    n, N = 1_000_000, 300
    u, v, V, w, _ = fo.synthetic_disc(n, N)               # seed 12345, SURVEY 8(d) shape; regenerated by the tests
    g = FixedGeometry(30., 40., 1e-3, -2e-3)
    FF = FrankFitter(1.6, N, g, alpha=1.05, weights_smooth=1e-4, verbose=False, store_iteration_diagnostics=True)
    t0 = time.time()
    sol = FF.fit(u, v, V, w)
    print('config1: reference fit', time.time() - t0, 's', FF.iteration_diagnostics['num_iterations'], 'iterations', flush=True)
    M, j, H0 = FF._M.copy(), FF._j.copy(), FF._H0
    # the reference's own round-off floor on this data set: same visibilities permuted / another block_size
    perm = np.random.default_rng(1000).permutation(n)
    FP = FrankFitter(1.6, N, g, alpha=1.05, weights_smooth=1e-4, verbose=False)
    solp = FP.fit(u[perm], v[perm], V[perm], w[perm])
    Mp = FP._M
    vm = VisibilityMapping(FF._DHT, g, verbose=False, block_size=20000)
    Mb = vm.map_visibilities(u, v, V, w)['M']
    d = np.sqrt(np.diag(M))
    self_M_entry = max(np.max(np.abs(Mp - M) / np.abs(M)), np.max(np.abs(Mb - M) / np.abs(M)))
    self_M_cs = max(np.max(np.abs(Mp - M) / np.outer(d, d)), np.max(np.abs(Mb - M) / np.outer(d, d)))
    self_M_max = max(np.max(np.abs(Mp - M)), np.max(np.abs(Mb - M))) / np.max(np.abs(M))
    self_prof = np.max(np.abs(solp.MAP - sol.MAP)) / np.max(np.abs(sol.MAP))
    print('config1: self-noise M per-entry %.3e  /sqrt(MkkMll) %.3e  max-norm %.3e  profile %.3e' %
          (self_M_entry, self_M_cs, self_M_max, self_prof), flush=True)
    np.savez_compressed(os.path.join(OUT, 'config1_normal_1e6_N300.npz'), n_vis=n, N=N, Rmax=1.6,
                        geom=np.array([30., 40., 1e-3, -2e-3]), alpha=1.05, wsmooth=1e-4,
                        in_check=np.array([u.sum(), v.sum(), V.real.sum(), V.imag.sum(), w.sum(), u[123456], V[654321].real]),
                        M=M, j=j, H0=H0, MAP=sol.MAP, power_spectrum=sol.power_spectrum,
                        num_iterations=FF.iteration_diagnostics['num_iterations'],
                        self_noise_M_entry=self_M_entry, self_noise_M_cs=self_M_cs, self_noise_M_max=self_M_max,
                        self_noise=self_prof, num_iterations_permuted=-1)


def gen_config3():
    """BASELINE.json configs[2] shape: LogNormal MAP fit at N=500 on 5e4 visibilities, Rmax = 1.0".

    With the SURVEY disc at Rmax = 1.6" the REFERENCE cannot run this configuration: for N >= 300 the Hessian of its first
    log-normal fit is numerically indefinite, cho_factor fails, the SVD pseudo-inverse branch (statistical_models.py:1152-1158)
    takes over and the first power-spectrum update returns zeros and infinities -> ValueError('Bad value in power
    spectrum'), for every (alpha, wsmooth) and every n_vis tried (2e4 .. 1e6; N = 250 is the largest N that completes).
    At Rmax = 1.0" (the disc ends at ~0.85") and 5e4 visibilities (seed 31) the reference completes at N = 500, which is what
    is pinned here -- when generated with OMP_NUM_THREADS=2.  With 1e5 visibilities of the same seed, or with the same 5e4
    visibilities and 1 or 4 BLAS threads (a different summation order inside dgemm), it aborts again: at this N the
    reference's log-normal path sits on the edge of its own numerical failure, which is why the GPU test compares the solver
    on the reference's M and j at the authors' tolerance and not bit for bit.."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from oracle import frank_oracle as fo
    import time
    import frank.minimizer as fm
    n, N, Rmax = int(os.environ.get('C3_NVIS', 50_000)), 500, 1.0
    u, v, V, w, _ = fo.synthetic_disc(n, N, Rmax, seed=31)
    g = FixedGeometry(30., 40., 1e-3, -2e-3)
    stats = []
    orig = fm.MinimizeNewton

    def spy(*a, **k):
        x, st = orig(*a, **k)
        stats.append(st)
        return x, st
    import frank.statistical_models as fsm
    fsm.MinimizeNewton = spy
    FL = FrankFitter(Rmax, N, g, alpha=1.3, weights_smooth=1e-2, method='LogNormal', verbose=False,
                     store_iteration_diagnostics=True)
    t0 = time.time()
    sl = FL.fit(u, v, V, w)
    fsm.MinimizeNewton = orig
    dl = FL.iteration_diagnostics
    print('config3: reference LogNormal fit', time.time() - t0, 's', dl['num_iterations'], 'iterations', flush=True)
    st = np.array(stats)
    np.savez_compressed(os.path.join(OUT, 'config3_lognormal_N500.npz'), n_vis=n, N=N, Rmax=Rmax, seed=31, alpha=1.3, wsmooth=1e-2,
                        in_check=np.array([u.sum(), v.sum(), V.real.sum(), V.imag.sum(), w.sum()]),
                        M_upper=_upper(FL._M), j=FL._j, H0=FL._H0, MAP=sl.MAP, s_MAP=sl._fit.MAP,
                        power_spectrum=sl.power_spectrum, num_iterations=dl['num_iterations'],
                        p_first=np.array(dl['power_spectrum'][:3]), MAP_first=np.array(dl['MAP'][:3]),
                        p_hist=np.array(dl['power_spectrum'])[::4], s_hist=np.array(dl['MAP'])[::4],
                        newton_stats=st)


def gen_config5():
    """BASELINE.json configs[4] shape: N=2000, 4 channels, debris model (2e4 visibilities)."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
    from oracle import frank_oracle as fo
    n, N = 20_000, 2000
    u, v, V, w, _ = fo.synthetic_disc(n, N, seed=55)
    freqs = np.random.default_rng(56).choice(np.array([2.1e11, 2.3e11, 3.3e11, 3.4e11]), n)
    g = FixedGeometry(30., 40., 1e-3, -2e-3)
    dht = DiscreteHankelTransform(1.6 / rad_to_arcsec, N)
    vm = VisibilityMapping(dht, g, vis_model='debris', scale_height=lambda r: 0.05 * r, verbose=False)
    m = vm.map_visibilities(u, v, V, w, frequencies=freqs)
    M = m['M']
    rng = np.random.default_rng(57)
    rows, cols = rng.integers(0, N, 6000), rng.integers(0, N, 6000)
    np.savez_compressed(os.path.join(OUT, 'config5_debris_N2000.npz'), n_vis=n, N=N, seed=55,
                        in_check=np.array([u.sum(), v.sum(), V.real.sum(), V.imag.sum(), w.sum(), freqs.sum()]),
                        channels=m['channels'], j=m['j'], H0=m['null_likelihood'], H2=vm._H2,
                        M_diag=np.array([np.diag(Mc) for Mc in M]), rows=rows, cols=cols,
                        M_sample=np.array([Mc[rows, cols] for Mc in M]),
                        M_absmax=np.array([np.max(np.abs(Mc)) for Mc in M]),
                        M_sum=np.array([Mc.sum() for Mc in M]))
    print('config5 written', flush=True)


def main():
    for name, fn in (('config1', gen_config1), ('config3', gen_config3), ('config5', gen_config5),
                     ('lognormal', lambda: gen_lognormal(FixedGeometry(30., 40., 1e-3, -2e-3)))):
        if sys.argv[1:] == [name]:
            fn()
            return
    if sys.argv[1:] == ['svd_fallback']:
        gen_svd_fallback()
        return
    if sys.argv[1:] == ['estweights']:           # regenerate this fixture alone
        gen_estweights()
        return
    if sys.argv[1:] == ['geomfit']:
        gen_geomfit()
        return
    # ---- J0 -------------------------------------------------------------------------------
    rng = np.random.default_rng(7)
    x = np.concatenate([rng.uniform(0, 1e-5, 200), rng.uniform(0, 5, 3000), rng.uniform(5, 30, 3000),
                        rng.uniform(30, 950, 6000), rng.uniform(950, 6300, 3000),
                        scipy.special.jn_zeros(0, 300), [0.0, 5.0, 1e-5]])
    np.savez_compressed(os.path.join(OUT, 'j0_golden.npz'), x=x, j0=scipy.special.j0(x))

    # ---- DHT ------------------------------------------------------------------------------
    N = 64
    dht = DiscreteHankelTransform(1.6 / rad_to_arcsec, N)
    qs = np.random.default_rng(3).uniform(0, dht.q[-1], 40)
    f = np.exp(-0.5 * (dht.r * rad_to_arcsec / 0.4) ** 2)
    np.savez_compressed(os.path.join(OUT, 'dht.npz'), Rmax=dht.Rmax, N=N, r=dht.r, q=dht.q, Qmax=dht.Qmax,
                        Ykm=dht._Ykm, scale_factor=dht._scale_factor, j_nk=dht._j_nk, j_nN=dht._j_nN,
                        coeff=dht.coefficients(), qs=qs, coeff_qs=dht.coefficients(qs),
                        f=f, Hf=dht.transform(f), Hf_qs=dht.transform(f, qs))

    # ---- mapping --------------------------------------------------------------------------
    n_vis, N = 6000, 60
    u, v, V, w, g = synthetic(n_vis, N)
    dht = DiscreteHankelTransform(1.6 / rad_to_arcsec, N)
    out = dict(u=u, v=v, V=V, w=w, N=N, Rmax=1.6, geom=np.array([g.inc, g.PA, g.dRA, g.dDec]))
    for model in ['opt_thick', 'opt_thin']:
        vm = VisibilityMapping(dht, g, vis_model=model, verbose=False)
        m = vm.map_visibilities(u, v, V, w)
        out.update({f'M_{model}': m['M'], f'j_{model}': m['j'], f'H0_{model}': m['null_likelihood']})
    up, vp, wp, Vp = g.apply_correction(u, v, V, use3D=True)
    out.update(up=up, vp=vp, wp=wp, Vp=Vp, q=np.hypot(up, vp))
    vm = VisibilityMapping(dht, g, vis_model='debris', scale_height=lambda r: 0.05 * r, verbose=False)
    m = vm.map_visibilities(u, v, V, w)
    out.update(M_debris=m['M'], j_debris=m['j'], H0_debris=m['null_likelihood'], H2_debris=vm._H2)
    freqs = np.random.default_rng(5).choice(np.array([2.1e11, 2.3e11, 3.4e11]), n_vis)
    vm = VisibilityMapping(dht, g, verbose=False)
    m = vm.map_visibilities(u, v, V, w, frequencies=freqs)
    out.update(freqs=freqs, M_multi=m['M'], j_multi=m['j'], H0_multi=m['null_likelihood'], channels=m['channels'])
    m = vm.map_visibilities(u, v, V, 2.5)       # scalar weights (radial_fitters.py:544)
    out.update(M_scalar_w=m['M'], j_scalar_w=m['j'], H0_scalar_w=m['null_likelihood'])
    Itest = 1e10 * np.exp(-0.5 * ((dht.r * rad_to_arcsec - 0.6) / 0.08) ** 2)
    out.update(I_pred=Itest, V_pred=vm.predict_visibilities(Itest, out['q'], wp))
    np.savez_compressed(os.path.join(OUT, 'mapping.npz'), **out)

    # ---- Normal fit -----------------------------------------------------------------------
    FF = FrankFitter(1.6, N, g, alpha=1.05, weights_smooth=1e-4, verbose=False,
                     store_iteration_diagnostics=True)
    sol = FF.fit(u, v, V, w)
    diag = FF.iteration_diagnostics
    uu = np.random.default_rng(11).uniform(-2e6, 2e6, 50)
    vv = np.random.default_rng(12).uniform(-2e6, 2e6, 50)
    np.savez_compressed(os.path.join(OUT, 'fit_normal.npz'), MAP=sol.MAP, power_spectrum=sol.power_spectrum,
                        num_iterations=diag['num_iterations'], p_first=np.array(diag['power_spectrum'][:3]),
                        MAP_first=np.array(diag['MAP'][:3]), r=sol.r, q=sol.q, covariance=sol.covariance,
                        log_evidence=FF.log_evidence_laplace(), log_like=sol.log_likelihood(),
                        upred=uu, vpred=vv, Vpred=sol.predict(uu, vv),
                        Vpred_deproj=sol.predict_deprojected(sol.q),
                        self_noise=self_noise(lambda kw: FrankFitter(1.6, N, g, alpha=1.05, weights_smooth=1e-4,
                                                                     verbose=False, **kw), u, v, V, w, sol.MAP))
    # alpha / wsmooth variation (config 4 grid points)
    sweep = []
    for alpha, ws in [(1.3, 1e-2), (1.01, 1e-1)]:
        FF2 = FrankFitter(1.6, N, g, alpha=alpha, weights_smooth=ws, verbose=False,
                          store_iteration_diagnostics=True)
        s2 = FF2.fit(u, v, V, w)
        sn = self_noise(lambda kw: FrankFitter(1.6, N, g, alpha=alpha, weights_smooth=ws, verbose=False, **kw),
                        u, v, V, w, s2.MAP)
        sweep.append((alpha, ws, FF2.iteration_diagnostics['num_iterations'], s2.MAP, s2.power_spectrum, sn))
    np.savez_compressed(os.path.join(OUT, 'fit_sweep.npz'), alpha=[s[0] for s in sweep], ws=[s[1] for s in sweep],
                        num_iterations=[s[2] for s in sweep], MAP=np.array([s[3] for s in sweep]),
                        power_spectrum=np.array([s[4] for s in sweep]), self_noise=np.array([s[5] for s in sweep]))
    # non-parametric (no prior) fit
    FB = FourierBesselFitter(1.6, 20, g, verbose=False)
    sb = FB.fit(u, v, V, w)
    np.savez_compressed(os.path.join(OUT, 'fit_fourier_bessel.npz'), MAP=sb.MAP,
                        self_noise=self_noise(lambda kw: FourierBesselFitter(1.6, 20, g, verbose=False, **kw), u, v, V, w, sb.MAP))

    # ---- LogNormal fit --------------------------------------------------------------------
    ul, vl, Vl, wl = gen_lognormal(g)

    # ---- debris Normal fit ----------------------------------------------------------------
    FD = FrankDebrisFitter(1.6, 40, g, lambda r: 0.05 * r, alpha=1.3, weights_smooth=1e-2, verbose=False,
                           store_iteration_diagnostics=True)
    sd = FD.fit(ul, vl, Vl, wl)
    np.savez_compressed(os.path.join(OUT, 'fit_debris.npz'), N=40, MAP=sd.MAP, power_spectrum=sd.power_spectrum,
                        num_iterations=FD.iteration_diagnostics['num_iterations'],
                        self_noise=self_noise(lambda kw: FrankDebrisFitter(1.6, 40, g, lambda r: 0.05 * r, alpha=1.3,
                                                                           weights_smooth=1e-2, verbose=False, **kw), ul, vl, Vl, wl, sd.MAP))

    # ---- AS 209 subsample (196 visibilities), N=20 -------------------------------------
    ua, va, re, im, wa = np.genfromtxt(os.path.join(REF, 'docs/tutorials/test_datafile.txt')).T
    Va = re + 1j * im
    ga = FixedGeometry(34.97, 85.76, dRA=1.9e-3, dDec=-2.5e-3)      # AS 209 geometry, frank/tests.py:150-151
    FA = FrankFitter(1.6, 20, ga, alpha=1.05, weights_smooth=1e-2, verbose=False,
                     store_iteration_diagnostics=True, check_qbounds=False, convergence_failure='warn')
    sa = FA.fit(ua, va, Va, wa)
    np.savez_compressed(os.path.join(OUT, 'fit_as209sub.npz'), u=ua, v=va, V=Va, w=wa,
                        geom=np.array([ga.inc, ga.PA, ga.dRA, ga.dDec]), MAP=sa.MAP,
                        power_spectrum=sa.power_spectrum, M=FA._M, j=FA._j, H0=FA._H0,
                        num_iterations=FA.iteration_diagnostics['num_iterations'],
                        self_noise=self_noise(lambda kw: FrankFitter(1.6, 20, ga, alpha=1.05, weights_smooth=1e-2, verbose=False,
                                                                     check_qbounds=False, convergence_failure='warn', **kw),
                                              ua, va, Va, wa, sa.MAP))

    # ---- UV binner ------------------------------------------------------------------------
    uvd = out['q']
    for tag, width in [('a', 5e4), ('b', 1e3)]:
        b = UVDataBinner(uvd, out['Vp'], w, width)
        np.savez_compressed(os.path.join(OUT, f'uvbin_{tag}.npz'), uv_in=uvd, V_in=out['Vp'], w_in=w, width=width,
                            idx=b.determine_uv_bin(uvd), uv=b.uv.filled(0), V=b.V.filled(0),
                            weights=b.weights.filled(0), counts=b.bin_counts.filled(0),
                            error=b.error.filled(np.nan), mask=np.ma.getmaskarray(b.uv),
                            left=b.bin_edges[0].filled(0), right=b.bin_edges[1].filled(0))
    # real-valued V and boundary-exact points
    uvx = np.concatenate([np.arange(0, 11) * 1e3, [9999.999999999998, 1e4 * (1 - 1e-16), 3e3 + 1e-9]])
    Vx = np.linspace(-1, 1, len(uvx))
    wx = np.linspace(1, 2, len(uvx))
    b = UVDataBinner(uvx, Vx, wx, 1e3)
    np.savez_compressed(os.path.join(OUT, 'uvbin_edge.npz'), uv_in=uvx, V_in=Vx, w_in=wx, width=1e3,
                        idx=b.determine_uv_bin(uvx), uv=b.uv.filled(0), V=b.V.filled(0),
                        weights=b.weights.filled(0), counts=b.bin_counts.filled(0),
                        error=b.error.filled(np.nan), mask=np.ma.getmaskarray(b.uv))

    gen_estweights()
    gen_geomfit()

    # ---- analytic Gaussian pair (frank/tests.py:37-130) --------------------------------------
    def gauss_vis(q, inc):
        return np.cos(inc * deg_to_rad) * 2 * np.pi * np.exp(-0.5 * (2 * np.pi * q) ** 2)
    dht = DiscreteHankelTransform(5.0, 100)
    gk = FixedGeometry(60, 0)
    vm = VisibilityMapping(dht, gk, verbose=False)
    Ig = np.exp(-0.5 * (dht.r) ** 2)
    np.savez_compressed(os.path.join(OUT, 'gauss_kat.npz'), r=dht.r, q=dht.q, I=Ig,
                        V_model=vm.predict_visibilities(Ig, dht.q), V_exact=gauss_vis(dht.q, 60.))
    print('golden fixtures written to', OUT)
    for fn in sorted(os.listdir(OUT)):
        if fn.endswith('.npz'):
            print(f'  {fn:24s} {os.path.getsize(os.path.join(OUT, fn)) / 1024:8.1f} KB')


if __name__ == '__main__':
    main()
