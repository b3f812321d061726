"""Pin the CPU oracle (oracle/frank_oracle.py, oracle/cephes_j0.c) to outputs of the unmodified
reference (fixtures from tests/golden/make_golden.py).  CPU only."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import frank_oracle as fo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    return np.max(np.abs(a - b)) / np.max(np.abs(b))


@pytest.fixture(scope='module')
def j0lib():
    so = os.path.join(ROOT, 'oracle', '_build', 'liboracle_j0.so')
    subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle')])
    lib = ctypes.CDLL(so)
    return lib


def _call(lib, fn, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    out = np.empty_like(x)
    getattr(lib, fn)(x.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(x.size))
    return out


def test_c_j0_matches_scipy_golden(golden, j0lib):
    g = golden('j0_golden.npz')
    x, ref = g['x'], g['j0']
    got = _call(j0lib, 'oracle_j0_array', x)
    small = x <= 5.0
    # Cephes rational branch: bit-for-bit what SciPy returns
    assert np.array_equal(got[small], ref[small])
    # asymptotic branch: SciPy 1.18 evaluates the phase more accurately than classic Cephes
    # (no rounding of x - pi/4); the restatement carries the classic form, whose deviation is
    # bounded by ulp(x)/2 * |J1(x)|.
    big = ~small
    bound = 0.5 * np.spacing(x[big]) * np.sqrt(2 / (np.pi * x[big])) + 4e-16
    assert np.all(np.abs(got[big] - ref[big]) <= bound)
    # and SciPy's J0 sits within 3e-16 of the exactly rounded function (x87 j0l)
    exact = _call(j0lib, 'oracle_j0_exact_array', x)
    assert np.max(np.abs(exact - ref)) < 5e-16
    assert np.max(np.abs(exact[x > 30] - ref[x > 30])) < 1e-16


def test_dht_tables(golden):
    g = golden('dht.npz')
    d = fo.DHTTables(float(g['Rmax']), int(g['N']))
    for name, got in [('r', d.r), ('q', d.q), ('Ykm', d.Ykm), ('scale_factor', d.scale_factor),
                      ('coeff', d.coefficients()), ('coeff_qs', d.coefficients(g['qs'])),
                      ('Hf', d.transform(g['f']))]:
        assert np.array_equal(got, g[name]), name
    assert d.Qmax == float(g['Qmax'])
    assert np.array_equal(np.dot(d.coefficients(g['qs']), g['f']), g['Hf_qs'])


def test_gaussian_hankel_pair(golden):
    """Reference KAT frank/tests.py:97-130: optically thick Gaussian, inc = 60 deg."""
    g = golden('gauss_kat.npz')
    d = fo.DHTTables(5.0, 100)
    V = fo.predict_visibilities(d, g['I'], d.q, None, 'opt_thick', 60.)
    assert np.array_equal(V, g['V_model'])
    np.testing.assert_allclose(V, g['V_exact'], atol=1e-5, rtol=0)


def _map(g, **kw):
    d = fo.DHTTables(float(g['Rmax']) / fo.RAD_TO_ARCSEC, int(g['N']))
    inc, PA, dRA, dDec = g['geom']
    return d, fo.map_visibilities(d, g['u'], g['v'], g['V'], kw.pop('w', g['w']), inc, PA, dRA, dDec, **kw)


def test_geometry(golden):
    g = golden('mapping.npz')
    inc, PA, dRA, dDec = g['geom']
    up, vp, wp, Vp = fo.apply_correction(g['u'], g['v'], g['V'], inc, PA, dRA, dDec)
    for name, got in [('up', up), ('vp', vp), ('wp', wp), ('Vp', Vp)]:
        assert np.array_equal(got, g[name]), name
    assert np.array_equal(np.hypot(up, vp), g['q'])


@pytest.mark.parametrize('model', ['opt_thick', 'opt_thin'])
def test_mapping_thin_thick(golden, model):
    g = golden('mapping.npz')
    _, m = _map(g, vis_model=model)
    assert np.array_equal(m['M'], g[f'M_{model}'])
    assert np.array_equal(m['j'], g[f'j_{model}'])
    assert m['null_likelihood'] == float(g[f'H0_{model}'])


def test_mapping_debris_multi_scalar(golden):
    g = golden('mapping.npz')
    d = fo.DHTTables(float(g['Rmax']) / fo.RAD_TO_ARCSEC, int(g['N']))
    H2 = fo.debris_H2(d, lambda r: 0.05 * r)
    assert np.array_equal(H2, g['H2_debris'])
    _, m = _map(g, vis_model='debris', H2=H2)
    assert np.array_equal(m['M'], g['M_debris']) and np.array_equal(m['j'], g['j_debris'])
    _, m = _map(g, frequencies=g['freqs'])
    assert np.array_equal(m['M'], g['M_multi']) and np.array_equal(m['j'], g['j_multi'])
    assert np.array_equal(m['channels'], g['channels'])
    assert m['null_likelihood'] == float(g['H0_multi'])
    _, m = _map(g, w=2.5)
    assert np.array_equal(m['M'], g['M_scalar_w']) and m['null_likelihood'] == float(g['H0_scalar_w'])
    V = fo.predict_visibilities(d, g['I_pred'], g['q'], g['wp'], 'opt_thick', g['geom'][0])
    assert np.array_equal(V, g['V_pred'])


def test_qbounds_error(golden):
    """frank/tests.py:415-431 expects ValueError when data extend past the last collocation point."""
    g = golden('mapping.npz')
    d = fo.DHTTables(1.6 / fo.RAD_TO_ARCSEC, 20)
    inc, PA, dRA, dDec = g['geom']
    with pytest.raises(ValueError):
        fo.map_visibilities(d, g['u'], g['v'], g['V'], g['w'], inc, PA, dRA, dDec)


def test_normal_fit(golden):
    g, f = golden('mapping.npz'), golden('fit_normal.npz')
    d = fo.DHTTables(1.6 / fo.RAD_TO_ARCSEC, int(g['N']))
    r = fo.frank_fit(d, g['M_opt_thick'], g['j_opt_thick'], alpha=1.05, weights_smooth=1e-4, store=True)
    assert r['num_iterations'] == int(f['num_iterations'])
    assert np.array_equal(np.array(r['history']['power_spectrum'][:3]), f['p_first'])
    assert rel(r['MAP'], f['MAP']) < 1e-12
    assert np.max(np.abs(r['power_spectrum'] / f['power_spectrum'] - 1)) < 1e-9
    cov = r['fit'].Dsolve(np.eye(d.N))
    assert rel(cov, f['covariance']) < 1e-10


def test_sweep_points(golden):
    g, f = golden('mapping.npz'), golden('fit_sweep.npz')
    d = fo.DHTTables(1.6 / fo.RAD_TO_ARCSEC, int(g['N']))
    for i in range(len(f['alpha'])):
        r = fo.frank_fit(d, g['M_opt_thick'], g['j_opt_thick'], alpha=float(f['alpha'][i]),
                         weights_smooth=float(f['ws'][i]))
        assert r['num_iterations'] == int(f['num_iterations'][i])
        assert rel(r['MAP'], f['MAP'][i]) < 1e-12


def test_fourier_bessel_fit(golden):
    g, f = golden('mapping.npz'), golden('fit_fourier_bessel.npz')
    d = fo.DHTTables(1.6 / fo.RAD_TO_ARCSEC, 20)
    inc, PA, dRA, dDec = g['geom']
    m = fo.map_visibilities(d, g['u'], g['v'], g['V'], g['w'], inc, PA, dRA, dDec, check_qbounds=False)
    assert rel(fo.GaussianSolve(d, m['M'], m['j']).mu, f['MAP']) < 1e-13


def test_lognormal_fit(golden):
    f = golden('fit_lognormal.npz')
    d = fo.DHTTables(1.6 / fo.RAD_TO_ARCSEC, int(f['N']))
    r = fo.frank_fit(d, f['M'], f['j'], alpha=1.3, weights_smooth=1e-2, method='LogNormal', store=True)
    assert r['num_iterations'] == int(f['num_iterations'])
    assert rel(np.array(r['history']['MAP'][:3]), f['MAP_first']) < 1e-12
    assert rel(r['MAP'], f['MAP']) < 1e-9
    assert rel(r['fit'].MAP, f['s_MAP']) < 1e-9


def test_as209_subsample_fit(golden):
    f = golden('fit_as209sub.npz')
    d = fo.DHTTables(1.6 / fo.RAD_TO_ARCSEC, 20)
    inc, PA, dRA, dDec = f['geom']
    m = fo.map_visibilities(d, f['u'], f['v'], f['V'], f['w'], inc, PA, dRA, dDec, check_qbounds=False)
    assert np.array_equal(m['M'], f['M']) and np.array_equal(m['j'], f['j'])
    assert m['null_likelihood'] == float(f['H0'])
    r = fo.frank_fit(d, m['M'], m['j'], alpha=1.05, weights_smooth=1e-2)
    assert r['num_iterations'] == int(f['num_iterations'])
    assert rel(r['MAP'], f['MAP']) < 1e-12


@pytest.mark.parametrize('tag', ['a', 'b', 'edge'])
def test_uv_binner(golden, tag):
    g = golden(f'uvbin_{tag}.npz')
    b = fo.uv_bin(g['uv_in'], g['V_in'], g['w_in'], float(g['width']))
    assert np.array_equal(b['idx'], g['idx'])
    assert np.array_equal(b['counts'], g['counts'])
    assert np.array_equal(b['mask'], g['mask'])
    ok = ~g['mask']
    for name in ['uv', 'V', 'weights']:
        assert np.array_equal(b[name][ok], g[name][ok]), name
    assert np.array_equal(b['error'][ok], g['error'][ok], equal_nan=True)


def test_estimate_weights(golden):
    """estimate_weights (utilities.py:515-631) in its call forms, incl. the reference's real-part-only variance
    (np.iscomplex(dtype) quirk) and the neighbour fill of single-count bins."""
    g = golden('estweights.npz')
    u, v, V = g['u'], g['v'], g['V']
    for tag, kw in [('log', dict(nbins=300)), ('lin', dict(nbins=300, log=False)), ('fine', dict(nbins=8000)),
                    ('median', dict(nbins=300, use_median=True))]:
        got = fo.estimate_weights(u, v, V, **kw)
        assert np.allclose(got, g[tag], rtol=1e-12, atol=0), tag
    got = fo.estimate_weights(np.hypot(u, v), V.real, nbins=100)
    assert np.allclose(got, g['uV'], rtol=1e-12, atol=0)


def test_svd_fallback(golden):
    """GaussianModel's SVD pseudo-inverse branch (statistical_models.py:747-755) on the reference's own outputs:
    indefinite systems with and without a prior (entry by entry), and the rank-deficient mapping of 25 visibilities
    at N = 40 (the solution is round-off along the null space, so the residual and the singular values are compared)."""
    g = golden('svd_fallback.npz')
    dht = fo.DHTTables(1.6 / fo.RAD_TO_ARCSEC, int(g["N"]))
    for tag, p in [('a', None), ('b', g['p'])]:
        fit = fo.GaussianSolve(dht, g['M'], g['j'], p)
        assert fit.svd is not None and fit.chol is None
        assert np.max(np.abs(fit.mu - g['mu_' + tag])) <= 1e-12 * np.max(np.abs(g['mu_' + tag]))
        assert np.allclose(fit.svd[1], g['s1_' + tag], rtol=1e-12, atol=0)
    fit = fo.GaussianSolve(dht, g['M'], g['j'])
    assert np.max(np.abs(fit.Dsolve(g['j']) - g['Dj_a'])) <= 1e-12 * np.max(np.abs(g['Dj_a']))
    fc = fo.GaussianSolve(dht, g['Mc'], g['jc'])
    assert fc.svd is not None
    assert np.max(np.abs(g['Mc'] @ fc.mu - g['jc'])) <= 1e-10 * np.max(np.abs(g['jc']))
    assert np.allclose(np.sort(fc.svd[1])[:25], np.sort(g['s1_c'])[:25], rtol=1e-9, atol=0)
