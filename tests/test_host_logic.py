"""CPU-only tests of the host-side mirror: DHT tables, geometry helpers, smoothing-matrix bands, API surface."""
import os
import numpy as np
import pytest

from oracle import frank_oracle as fo
from frank_b200.constants import rad_to_arcsec
from frank_b200.filter import banded_ldl, smoothing_bands, spectral_smoothing_matrix
from frank_b200.geometry import FixedGeometry, apply_phase_shift, deproject
from frank_b200.hankel import DiscreteHankelTransform


def test_dht_matches_reference_golden(golden):
    g = golden('dht.npz')
    d = DiscreteHankelTransform(float(g['Rmax']), int(g['N']))
    for name, got in [('r', d.r), ('q', d.q), ('Ykm', d._Ykm), ('scale_factor', d._scale_factor),
                      ('coeff', d.coefficients()), ('coeff_qs', d.coefficients(g['qs'])), ('Hf', d.transform(g['f'])),
                      ('Hf_qs', d.transform(g['f'], g['qs']))]:
        assert np.array_equal(got, g[name]), name
    assert d.Qmax == float(g['Qmax']) and d.size == int(g['N']) and d.order == 0
    r, q = DiscreteHankelTransform.get_collocation_points(float(g['Rmax']), int(g['N']))
    assert np.array_equal(r, g['r']) and np.array_equal(q, g['q'])
    with pytest.raises(AttributeError):
        d.coefficients(direction='sideways')


def test_hankel_gauss_pair():
    """frank/tests.py:37-94: exp(-r^2/2) <-> 2 pi exp(-(2 pi q)^2/2)."""
    d = DiscreteHankelTransform(5.0, 100)
    f = np.exp(-0.5 * d.r ** 2)
    F = 2 * np.pi * np.exp(-0.5 * (2 * np.pi * d.q) ** 2)
    np.testing.assert_allclose(d.transform(f, direction='forward'), F, atol=1e-5, rtol=0)
    np.testing.assert_allclose(d.transform(F, direction='backward'), f, atol=1e-5, rtol=0)
    np.testing.assert_allclose(np.dot(d.coefficients(d.q), f), d.transform(f), atol=1e-12, rtol=0)


def test_geometry_matches_reference_golden(golden):
    g = golden('mapping.npz')
    geom = FixedGeometry(*[float(x) for x in g['geom']])
    up, vp, wp, Vp = geom.apply_correction(g['u'], g['v'], g['V'], use3D=True)
    for name, got in [('up', up), ('vp', vp), ('wp', wp), ('Vp', Vp)]:
        assert np.array_equal(got, g[name]), name
    u2, v2 = geom.reproject(*geom.deproject(g['u'], g['v']))
    np.testing.assert_allclose(u2, g['u'], rtol=1e-12)
    _, _, V2 = geom.undo_correction(up, vp, Vp)
    np.testing.assert_allclose(V2, g['V'], rtol=1e-10)
    s = geom.device_scalars()
    assert s.cos_inc == np.cos(float(g['geom'][0]) * np.pi / 180)
    assert geom.clone().inc == geom.inc and 'FixedGeometry' in repr(geom)


def test_smoothing_matrix_and_band_factorisation():
    for N, ws in [(20, 1e-4), (60, 1e-2), (300, 1e-1)]:
        d = DiscreteHankelTransform(1.6 / rad_to_arcsec, N)
        o = fo.DHTTables(1.6 / fo.RAD_TO_ARCSEC, N)
        Tref = fo.smoothing_matrix(o, ws).toarray()
        T = spectral_smoothing_matrix(d, ws).toarray()
        assert np.max(np.abs(T - Tref)) <= 1e-14 * np.max(np.abs(Tref))
        b = smoothing_bands(d, ws)
        b[2] += 1
        ldl = banded_ldl(b)
        L = np.eye(N) + np.diag(ldl[1][1:], -1) + np.diag(ldl[2][2:], -2)
        A = Tref + np.eye(N)
        assert np.max(np.abs(L @ np.diag(ldl[0]) @ L.T - A)) <= 1e-13 * np.max(np.abs(A))
        assert np.all(ldl[0] > 0)


def test_fitter_constructor_errors():
    from frank_b200.radial_fitters import FrankFitter, FourierBesselFitter
    from frank_b200.statistical_models import VisibilityMapping
    g = FixedGeometry(30, 40)
    with pytest.raises(ValueError):
        FrankFitter(1.6, 20, g, method='Poisson')
    with pytest.raises(ValueError):
        FrankFitter(1.6, 20, g, convergence_failure='explode')
    with pytest.raises(ValueError):
        FourierBesselFitter(1.6, 20, g, assume_optically_thick=True, scale_height=lambda r: r)
    with pytest.raises(ValueError):
        VisibilityMapping(DiscreteHankelTransform(1e-5, 20), g, vis_model='opaque')
    with pytest.raises(ValueError):
        VisibilityMapping(DiscreteHankelTransform(1e-5, 20), g, vis_model='debris')
    FF = FrankFitter(1.6, 20, g, verbose=False)
    assert FF.size == 20 and abs(FF.Rmax - 1.6) < 1e-12 and FF.fit_method() == 'FrankFitter: Normal method'
    vm = FF._vis_map
    assert vm.check_hash([False, FF._DHT, g, 'opt_thick', None])
    assert not vm.check_hash([True, FF._DHT, g, 'opt_thick', None])
    assert not vm.check_hash([False, FF._DHT, FixedGeometry(31, 40), 'opt_thick', None])


def test_j0_table_accuracy(tmp_path):
    """The overlapping degree-7 J0 table of the Gram kernel (frank_b200/csrc/fb_j0_table.h), built and evaluated on
    the host: nearest row (gather path, |t| <= 1/32) and the far neighbour (|t| up to 1/16, the worst case of the
    one-row-per-tile path) against glibc's 80-bit j0l.  SciPy's own J0 -- what the reference calls -- is within
    4.4e-16 (x <= 30) / 5.6e-17 (x > 30) of the exact function; the table has to stay inside that."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / 'j0_table_check')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-o', exe, os.path.join(root, 'tests', 'native', 'j0_table_check.cpp')])
    out = subprocess.check_output([exe, '6400', '3000000']).decode().split()
    near, far = [float(v) for v in out[:4]], [float(v) for v in out[4:]]
    # ranges: x < 5, < 30, < 200, < 6400 (N = 2000 reaches j_N = 6283)
    # (half an ulp of the value itself is 1.1e-16 / 5.6e-17 / 2.8e-17 / 1.4e-17 at the top of these ranges)
    assert near[0] <= 1.2e-16 and near[1] <= 7e-17 and near[2] <= 3.5e-17 and near[3] <= 1.5e-17
    assert far[0] <= 1.3e-16 and far[1] <= 7e-17 and far[2] <= 3.5e-17 and far[3] <= 1.5e-17


def test_j0_recentred_product_accuracy(tmp_path):
    """The J0 phase of k_gram forms a tile of the design matrix as a rank-8 product: the table row of a (column, tile) pair is
    shifted to the tile centre and rescaled to the tile variable, the visibilities contribute sqrt(w) times its powers
    (fb_gram.cu: select / prepare / j0_gemm).  tests/native/j0_recentre_check.cpp restates that arithmetic on the host, in the
    kernel's operation order, and measures G / sqrt(w) against glibc's 80-bit j0l(a j_k).  The bars sit where SciPy's own J0
    -- what the reference calls -- does: 4.4e-16 for x <= 30, 5.6e-17 above."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / 'j0_recentre_check')
    subprocess.check_call(['g++', '-O2', '-std=c++17', '-o', exe, os.path.join(root, 'tests', 'native', 'j0_recentre_check.cpp')])
    out = [float(v) for v in subprocess.check_output([exe, '6400', '40000']).decode().split()]
    worst, rejected = out[:4], out[4]
    # ranges: x < 5, < 30, < 200, < 6400; measured 4.0e-16 / 1.4e-16 / 6.6e-17 / 2.3e-17
    assert worst[0] <= 4.5e-16 and worst[1] <= 2.0e-16 and worst[2] <= 1.0e-16 and worst[3] <= 4.0e-17
    assert 0.0 < rejected < 0.2          # tiles wider than a row's validity window are rejected (per-visibility path), not mis-evaluated
