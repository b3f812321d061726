"""GPU parity tests of the visibility-mapping path (through the C ABI) against the CPU oracle and against
fixtures produced by the unmodified reference.  Run on a B200: pytest -m gpu."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import frank_oracle as fo

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPS = np.finfo(np.float64).eps


def gram_ok(M, Mref, rtol=1e-10, c_eps=16):
    """Per-entry criterion of the north star with the floor any float64 summation needs:
    |dM_kl| <= 1e-10 |M_kl| + c_eps * eps * sqrt(M_kk M_ll).
    The second term is the Cauchy-Schwarz scale of the k,l dot product; entries that cancel to 1e-7..1e-9 of it
    change by more than 1e-10 relative when the REFERENCE itself is re-run with its visibilities permuted or
    with another block_size ('Parity floor')."""
    d = np.sqrt(np.abs(np.diag(Mref)))
    tol = rtol * np.abs(Mref) + c_eps * EPS * np.outer(d, d)
    return float(np.max(np.abs(M - Mref) / tol))


def vec_ok(j, jref, M, H0scale, rtol=1e-10, c_eps=16):
    tol = rtol * np.abs(jref) + c_eps * EPS * np.sqrt(np.abs(np.diag(M))) * H0scale
    return float(np.max(np.abs(j - jref) / tol))


@pytest.fixture(scope='module')
def fb():
    import frank_b200  # noqa: F401
    from frank_b200 import _lib
    from frank_b200.geometry import FixedGeometry
    from frank_b200.hankel import DiscreteHankelTransform
    from frank_b200.statistical_models import VisibilityMapping
    from frank_b200.constants import rad_to_arcsec

    class NS:
        pass
    ns = NS()
    ns.lib, ns.FixedGeometry, ns.DHT, ns.VM, ns.r2a = _lib, FixedGeometry, DiscreteHankelTransform, VisibilityMapping, rad_to_arcsec
    return ns


def mapping_from_golden(fb, g, **kw):
    inc, PA, dRA, dDec = [float(x) for x in g['geom']]
    dht = fb.DHT(float(g['Rmax']) / fb.r2a, int(g['N']))
    vm = fb.VM(dht, fb.FixedGeometry(inc, PA, dRA, dDec), verbose=False, **kw)
    return dht, vm


def test_j0_device_accuracy(fb, golden):
    """The device J0 (Taylor table) against the exactly rounded function and against SciPy's J0 (reference)."""
    g = golden('j0_golden.npz')
    ctx = fb.lib.get_context()
    dht = fb.DHT(1.6 / fb.r2a, 2000)
    ctx.dht_setup(dht)
    x = g['x'][g['x'] < dht._j_nk[-1]]
    got = ctx.debug_j0(x)
    subprocess.check_call(['make', '-s', '-C', os.path.join(ROOT, 'oracle')])
    lib = ctypes.CDLL(os.path.join(ROOT, 'oracle', '_build', 'liboracle_j0.so'))
    exact = np.empty_like(x)
    lib.oracle_j0_exact_array(x.ctypes.data_as(ctypes.c_void_p), exact.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(x.size))
    assert np.max(np.abs(got - exact)) <= 1.2e-16           # half an ulp of values <= 1
    big = x > 30
    assert np.max(np.abs(got[big] - exact[big])) <= 3e-17
    ref = g['j0'][g['x'] < dht._j_nk[-1]]
    assert np.max(np.abs(got - ref)) <= 5e-16                # SciPy's own error for x <= 30
    assert np.max(np.abs(got[big] - ref[big])) <= 1e-16
    # the one-row-per-tile path may use a row up to 1/16 away from its centre (fb_j0_table.h): worst case
    far = ctx.debug_j0(x, far=True)
    assert np.max(np.abs(far - exact)) <= 1.3e-16
    assert np.max(np.abs(far[big] - exact[big])) <= 4e-17   # SciPy itself: 5.6e-17 for x > 30


def test_prepass_bits(fb, golden):
    """Deprojection / hypot / kz bit-identical to NumPy's; Re V' to 2 ulp (sin/cos differ by an ulp)."""
    g = golden('mapping.npz')
    dht, vm = mapping_from_golden(fb, g)
    vm.map_visibilities(g['u'], g['v'], g['V'], g['w'])
    n = len(g['u'])
    a, kz, Vre, perm = fb.lib.get_context().debug_prepped(n)
    assert np.array_equal(np.sort(perm), np.arange(n))
    assert np.array_equal(a, (g['q'] * (1. / dht.Qmax))[perm])
    assert np.array_equal(kz, g['wp'][perm])
    assert np.max(np.abs(Vre - g['Vp'].real[perm])) <= 4 * EPS * np.max(np.abs(g['Vp'].real))
    assert np.all(np.diff(a) >= -a.max() / 60000)          # sorted by baseline bin


@pytest.mark.parametrize('model', ['opt_thick', 'opt_thin'])
def test_mapping_vs_reference_golden(fb, golden, model):
    g = golden('mapping.npz')
    dht, vm = mapping_from_golden(fb, g, vis_model=model)
    m = vm.map_visibilities(g['u'], g['v'], g['V'], g['w'])
    assert m['mult_freq'] is False and m['channels'] is None and m['M'].shape == (dht.size, dht.size)
    assert gram_ok(m['M'], g[f'M_{model}']) <= 1.0
    assert np.max(np.abs(m['M'] - g[f'M_{model}'])) <= 1e-14 * np.max(np.abs(g[f'M_{model}']))
    assert np.max(np.abs(m['j'] - g[f'j_{model}'])) <= 1e-13 * np.max(np.abs(g[f'j_{model}']))
    assert abs(m['null_likelihood'] - float(g[f'H0_{model}'])) <= 1e-12 * abs(float(g[f'H0_{model}']))
    assert np.array_equal(m['M'], m['M'].T)


def test_mapping_debris_scalar_multi(fb, golden):
    g = golden('mapping.npz')
    dht, vm = mapping_from_golden(fb, g, vis_model='debris', scale_height=lambda r: 0.05 * r)
    assert np.array_equal(vm._H2, g['H2_debris'])
    m = vm.map_visibilities(g['u'], g['v'], g['V'], g['w'])
    assert gram_ok(m['M'], g['M_debris']) <= 1.0
    assert np.max(np.abs(m['j'] - g['j_debris'])) <= 1e-13 * np.max(np.abs(g['j_debris']))
    dht, vm = mapping_from_golden(fb, g)
    m = vm.map_visibilities(g['u'], g['v'], g['V'], 2.5)     # scalar weights (radial_fitters.py:544)
    assert gram_ok(m['M'], g['M_scalar_w']) <= 1.0
    assert abs(m['null_likelihood'] - float(g['H0_scalar_w'])) <= 1e-12 * abs(float(g['H0_scalar_w']))
    m = vm.map_visibilities(g['u'], g['v'], g['V'], g['w'], frequencies=g['freqs'])
    assert m['mult_freq'] is True and np.array_equal(m['channels'], g['channels'])
    for c in range(len(g['channels'])):
        assert gram_ok(m['M'][c], g['M_multi'][c]) <= 1.0
    assert np.max(np.abs(m['j'] - g['j_multi'])) <= 1e-13 * np.max(np.abs(g['j_multi']))
    assert abs(m['null_likelihood'] - float(g['H0_multi'])) <= 1e-12 * abs(float(g['H0_multi']))


def test_qbounds_error(fb, golden):
    """frank/tests.py:415-431: data beyond the last collocation point raise ValueError."""
    g = golden('mapping.npz')
    inc, PA, dRA, dDec = [float(x) for x in g['geom']]
    vm = fb.VM(fb.DHT(1.6 / fb.r2a, 20), fb.FixedGeometry(inc, PA, dRA, dDec), verbose=False)
    with pytest.raises(ValueError):
        vm.map_visibilities(g['u'], g['v'], g['V'], g['w'])
    vm.check_qbounds = False
    vm.map_visibilities(g['u'], g['v'], g['V'], g['w'])     # FourierBesselFitter path: no check


@pytest.mark.parametrize('n,N', [(1, 20), (63, 20), (64, 33), (65, 100), (5000, 300), (40000, 500), (3000, 1000)])
def test_mapping_vs_oracle_sizes(fb, n, N):
    """Ragged / tiny / multi-panel shapes against the CPU oracle on the same seeded inputs."""
    u, v, V, w, odht = fo.synthetic_disc(n, N, seed=100 + n)
    ref = fo.map_visibilities(odht, u, v, V, w, 30., 40., 1e-3, -2e-3)
    vm = fb.VM(fb.DHT(1.6 / fb.r2a, N), fb.FixedGeometry(30., 40., 1e-3, -2e-3), verbose=False)
    m = vm.map_visibilities(u, v, V, w)
    assert gram_ok(m['M'], ref['M']) <= 1.0
    assert np.max(np.abs(m['M'] - ref['M'])) <= 2e-14 * np.max(np.abs(ref['M']))
    assert np.max(np.abs(m['j'] - ref['j'])) <= 1e-12 * np.max(np.abs(ref['j']))
    assert abs(m['null_likelihood'] - ref['null_likelihood']) <= 1e-12 * abs(ref['null_likelihood'])


@pytest.mark.parametrize('N', [1, 2, 7, 8, 9, 31, 33, 63, 64, 65, 127, 128, 151, 152, 159, 160, 161, 255, 256, 257,
                               319, 320, 321, 479, 480, 481, 639, 640, 641, 800])
def test_block_plans(fb, N):
    """Every shape of the block decomposition (panel counts 1..6, full / half OFF blocks, odd and even DIAG
    panels, partly filled last tiles, 16- and 24-bit sort keys) against the CPU oracle; 700 visibilities are
    sparse enough that the upper modes also take the per-visibility gather path of the J0 phase."""
    n = 700
    u, v, V, w, odht = fo.synthetic_disc(n, N, seed=7 * N + 1)
    ref = fo.map_visibilities(odht, u, v, V, w, 30., 40., 1e-3, -2e-3)
    vm = fb.VM(fb.DHT(1.6 / fb.r2a, N), fb.FixedGeometry(30., 40., 1e-3, -2e-3), verbose=False)
    m = vm.map_visibilities(u, v, V, w)
    assert gram_ok(m['M'], ref['M']) <= 1.0
    assert np.max(np.abs(m['M'] - ref['M'])) <= 2e-14 * np.max(np.abs(ref['M']))
    assert np.max(np.abs(m['j'] - ref['j'])) <= 1e-12 * np.max(np.abs(ref['j']))
    assert np.array_equal(m['M'], m['M'].T)


def test_mapping_empty(fb):
    vm = fb.VM(fb.DHT(1.6 / fb.r2a, 40), fb.FixedGeometry(30., 40.), verbose=False, check_qbounds=False)
    m = vm.map_visibilities(np.zeros(0), np.zeros(0), np.zeros(0, dtype=complex), np.zeros(0))
    assert np.all(m['M'] == 0) and np.all(m['j'] == 0) and m['null_likelihood'] == 0


def test_device_and_host_entry_points_agree(fb):
    import torch
    u, v, V, w, odht = fo.synthetic_disc(20000, 120, seed=5)
    vm = fb.VM(fb.DHT(1.6 / fb.r2a, 120), fb.FixedGeometry(30., 40., 1e-3, -2e-3), verbose=False)
    mh = vm.map_visibilities(u, v, V, w)
    md = vm.map_visibilities(*[torch.from_numpy(x).cuda() for x in (u, v, V, w)])
    assert np.array_equal(mh['M'], md['M']) and np.array_equal(mh['j'], md['j'])
    assert mh['null_likelihood'] == md['null_likelihood']
    # deterministic: a second call gives the same bits (frank/tests.py:296-314 relies on this)
    m2 = vm.map_visibilities(u, v, V, w)
    assert np.array_equal(mh['M'], m2['M']) and np.array_equal(mh['j'], m2['j'])


def test_full_size_properties(fb):
    """BASELINE.json config 2 shape (1e7 visibilities, N = 300) through size-independent properties:
    linearity in the weights, additivity over a split of the visibilities, symmetry, positive diagonal,
    agreement of the (split) host entry point with the device entry point, and agreement of a 2e4-visibility slice
    with the oracle."""
    import torch
    n, N = 10_000_000, 300
    gen = torch.Generator(device='cuda').manual_seed(3)
    dht = fb.DHT(1.6 / fb.r2a, N)
    q = 0.98 * dht.q[-1] * torch.sqrt(torch.rand(n, device='cuda', dtype=torch.float64, generator=gen))
    th = 2 * np.pi * torch.rand(n, device='cuda', dtype=torch.float64, generator=gen)
    u, v = q * torch.cos(th), q * torch.sin(th)
    V = torch.complex(torch.exp(-(q / 1e6) ** 2), 0.1 * torch.sin(q / 3e5))
    w = 1e4 * (0.5 + 1.5 * torch.rand(n, device='cuda', dtype=torch.float64, generator=gen))
    vm = fb.VM(dht, fb.FixedGeometry(30., 40., 1e-3, -2e-3), verbose=False, check_qbounds=False)
    full = vm.map_visibilities(u, v, V, w)
    M = full['M']
    assert np.array_equal(M, M.T) and np.all(np.diag(M) > 0)
    half = n // 2 + 12345
    a = vm.map_visibilities(u[:half], v[:half], V[:half], w[:half])
    b = vm.map_visibilities(u[half:], v[half:], V[half:], w[half:])
    d = np.sqrt(np.diag(M))
    # float64 accumulation of ~7e4 DMMA steps per accumulator: random-walk round-off ~ eps * sqrt(7e4) / 3.5 = 75 eps
    # relative to the Cauchy-Schwarz scale of each entry (the reference's chunked dgemm accumulation is at ~50 eps)
    SUM_TOL = 256 * EPS
    assert np.max(np.abs(a['M'] + b['M'] - M) / np.outer(d, d)) < SUM_TOL
    assert np.max(np.abs(a['j'] + b['j'] - full['j'])) < 1e-12 * np.max(np.abs(full['j']))
    assert abs(a['null_likelihood'] + b['null_likelihood'] - full['null_likelihood']) < 1e-12 * abs(full['null_likelihood'])
    w2 = vm.map_visibilities(u, v, V, 2.0 * w)
    assert np.array_equal(w2['M'], 2.0 * M) or np.max(np.abs(w2['M'] - 2.0 * M) / np.outer(d, d)) < SUM_TOL
    # the host entry point runs this size as two overlapped halves (copy of the second under the kernels of the
    # first): same result up to the summation order, and the same bits on every call
    uh, vh, Vh, wh = [x.cpu().numpy() for x in (u, v, V, w)]
    host = vm.map_visibilities(uh, vh, Vh, wh)
    assert np.max(np.abs(host['M'] - M) / np.outer(d, d)) < SUM_TOL
    assert np.max(np.abs(host['j'] - full['j'])) < 1e-12 * np.max(np.abs(full['j']))
    assert abs(host['null_likelihood'] - full['null_likelihood']) < 1e-12 * abs(full['null_likelihood'])
    host2 = vm.map_visibilities(uh, vh, Vh, wh)
    assert np.array_equal(host['M'], host2['M']) and np.array_equal(host['j'], host2['j'])
    del uh, vh, Vh, wh
    k = 20000
    us, vs, Vs, ws = [x[:k].cpu().numpy() for x in (u, v, V, w)]
    ref = fo.map_visibilities(fo.DHTTables(1.6 / fo.RAD_TO_ARCSEC, N), us, vs, Vs, ws, 30., 40., 1e-3, -2e-3, check_qbounds=False)
    got = vm.map_visibilities(us, vs, Vs, ws)
    assert gram_ok(got['M'], ref['M']) <= 1.0


def test_predict_visibilities_vs_reference_golden(fb, golden):
    """statistical_models.py:279-329 and FrankRadialFit.predict (radial_fitters.py:56-98)."""
    g, f = golden('mapping.npz'), golden('fit_normal.npz')
    dht, vm = mapping_from_golden(fb, g)
    V = vm.predict_visibilities(g['I_pred'], g['q'], g['wp'])
    assert np.max(np.abs(V - g['V_pred'])) <= 1e-13 * np.max(np.abs(g['V_pred']))
    k = golden('gauss_kat.npz')                                   # frank/tests.py:97-130, inc = 60 deg
    vmk = fb.VM(fb.DHT(5.0, 100), fb.FixedGeometry(60, 0), verbose=False)
    Vk = vmk.predict_visibilities(k['I'], k['q'])
    assert np.max(np.abs(Vk - k['V_model'])) <= 1e-13 * np.max(np.abs(k['V_model']))
    np.testing.assert_allclose(Vk, k['V_exact'], atol=1e-5, rtol=0)
    from frank_b200.radial_fitters import FrankFitter
    FF = FrankFitter(1.6, int(g['N']), fb.FixedGeometry(*[float(x) for x in g['geom']]), verbose=False)
    sol = FF.fit(g['u'], g['v'], g['V'], g['w'])
    Vp = sol.predict(f['upred'], f['vpred'])
    assert np.max(np.abs(Vp - f['Vpred'])) <= 1e-7 * np.max(np.abs(f['Vpred']))
    assert np.max(np.abs(sol.predict_deprojected(sol.q) - f['Vpred_deproj'])) <= 1e-7 * np.max(np.abs(f['Vpred_deproj']))


# ---------------------------------------------------------------------------------------------------------------------
# Round 2: BASELINE-size fixtures, device-side multi-channel, the chunked host pipeline, the asynchronous entry point
# ---------------------------------------------------------------------------------------------------------------------
def strict_errors(M, Mref):
    """The north star's per-entry figure with NO floor, the same relative to the Cauchy-Schwarz scale, and the max-norm."""
    d = np.sqrt(np.diag(Mref))
    return (float(np.max(np.abs(M - Mref) / np.abs(Mref))), float(np.max(np.abs(M - Mref) / np.outer(d, d))),
            float(np.max(np.abs(M - Mref)) / np.max(np.abs(Mref))))


def test_config1_mapping_and_fit_vs_reference_golden(fb, golden):
    """BASELINE.json configs[0] at full size: 1e6 visibilities, N = 300, alpha = 1.05, wsmooth = 1e-4.  The inputs are
    regenerated bit for bit (checksums in the fixture); M, j, H0, the fitted profile, the power spectrum and the iteration
    count are the unmodified reference's (tests/golden/make_golden.py config1).

    M is held to the parity floor |dM_kl| <= 1e-10 |M_kl| + 16 eps sqrt(M_kk M_ll) and to 2e-14 in the max norm.  The strict
    per-entry figure (no floor) is printed next to the reference's own: entries that cancel to 1e-7 .. 1e-9 of
    sqrt(M_kk M_ll) cannot be reproduced to 1e-10 of THEMSELVES by anything but a bit-for-bit copy of SciPy's J0 -- the
    reference moves them by 1.8e-10 when its visibilities are merely permuted (same J0 values, another summation order),
    and a different J0 approximation (ours errs by <= 1.2e-16 of the exact function, SciPy's Cephes by <= 4.4e-16,
    tests/test_oracle_golden.py) shifts them by a few 1e-15 of sqrt(M_kk M_ll), i.e. ~1e-8 of the smallest entries."""
    g = golden('config1_normal_1e6_N300.npz')
    n, N = int(g['n_vis']), int(g['N'])
    u, v, V, w, _ = fo.synthetic_disc(n, N)
    chk = np.array([u.sum(), v.sum(), V.real.sum(), V.imag.sum(), w.sum(), u[123456], V[654321].real])
    assert np.array_equal(chk, g['in_check']), "regenerated inputs differ from the ones the reference was run on"
    from frank_b200.radial_fitters import FrankFitter
    geom = fb.FixedGeometry(*[float(x) for x in g['geom']])
    FF = FrankFitter(1.6, N, geom, alpha=float(g['alpha']), weights_smooth=float(g['wsmooth']), verbose=False,
                     store_iteration_diagnostics=True)
    sol = FF.fit(u, v, V, w)
    e_entry, e_cs, e_max = strict_errors(FF._M, g['M'])
    print(f"\nconfig1 M: strict per-entry {e_entry:.3e} (reference vs itself {float(g['self_noise_M_entry']):.3e}), "
          f"/sqrt(MkkMll) {e_cs:.3e} ({float(g['self_noise_M_cs']):.3e}), max-norm {e_max:.3e} ({float(g['self_noise_M_max']):.3e})")
    assert gram_ok(FF._M, g['M']) <= 1.0
    # closer to the reference than the reference is to itself under a permutation of its visibilities
    assert e_cs <= float(g['self_noise_M_cs']) and e_cs <= 64 * EPS and e_max <= 2e-14
    assert np.max(np.abs(FF._j - g['j'])) <= 1e-12 * np.max(np.abs(g['j']))
    assert abs(FF._H0 - float(g['H0'])) <= 1e-12 * abs(float(g['H0']))
    assert FF.iteration_diagnostics['num_iterations'] == int(g['num_iterations'])
    err = np.max(np.abs(sol.MAP - g['MAP'])) / np.max(np.abs(g['MAP']))
    perr = np.max(np.abs(sol.power_spectrum - g['power_spectrum']) / g['power_spectrum'])
    print(f"config1 fit: {FF.iteration_diagnostics['num_iterations']} iterations, profile error / peak {err:.3e} "
          f"(reference vs itself {float(g['self_noise']):.3e}), power spectrum rel {perr:.3e}")
    assert err <= max(1e-8, 4 * float(g['self_noise']))
    assert perr <= 1e-6
    # the solver alone, on the reference's own M and j: the north star's 1e-8 of peak
    FF2 = FrankFitter(1.6, N, geom, alpha=float(g['alpha']), weights_smooth=float(g['wsmooth']), verbose=False,
                      store_iteration_diagnostics=True)
    sol2 = FF2.fit_preprocessed({'M': g['M'], 'j': g['j'], 'null_likelihood': float(g['H0']),
                                 'hash': [False, FF2._DHT, geom, 'opt_thick', None]})
    err2 = np.max(np.abs(sol2.MAP - g['MAP'])) / np.max(np.abs(g['MAP']))
    print(f"config1 solver on the reference's M, j: profile error / peak {err2:.3e}")
    assert FF2.iteration_diagnostics['num_iterations'] == int(g['num_iterations'])
    assert err2 <= 1e-8


def _config5_inputs(g):
    n, N = int(g['n_vis']), int(g['N'])
    u, v, V, w, odht = fo.synthetic_disc(n, N, seed=int(g['seed']))
    freqs = np.random.default_rng(56).choice(np.array([2.1e11, 2.3e11, 3.3e11, 3.4e11]), n)
    chk = np.array([u.sum(), v.sum(), V.real.sum(), V.imag.sum(), w.sum(), freqs.sum()])
    assert np.array_equal(chk, g['in_check'])
    return u, v, V, w, freqs, odht


def test_config5_shape_vs_reference_golden(fb, golden):
    """BASELINE.json configs[4] shape -- N = 2000, 4 frequency channels, debris scale height -- in ONE device call
    (channel in the sort key, one Gram launch per channel), against the reference's j, H0, diagonal and 6000 sampled
    entries of every channel's M, and the whole of M against the oracle (itself checked against the same samples)."""
    import torch
    g = golden('config5_debris_N2000.npz')
    u, v, V, w, freqs, odht = _config5_inputs(g)
    N = int(g['N'])
    vm = fb.VM(fb.DHT(1.6 / fb.r2a, N), fb.FixedGeometry(30., 40., 1e-3, -2e-3), vis_model='debris',
               scale_height=lambda r: 0.05 * r, verbose=False)
    m = vm.map_visibilities(u, v, V, w, frequencies=freqs)
    assert m['mult_freq'] and np.array_equal(m['channels'], g['channels'])
    assert m['M'].shape == (4, N, N) and m['j'].shape == (4, N)
    rows, cols = g['rows'], g['cols']
    for c in range(4):
        d = np.sqrt(g['M_diag'][c])
        got, ref = m['M'][c][rows, cols], g['M_sample'][c]
        floor = 1e-10 * np.abs(ref) + 16 * EPS * d[rows] * d[cols]
        assert np.max(np.abs(got - ref) / floor) <= 1.0
        assert np.max(np.abs(np.diag(m['M'][c]) - g['M_diag'][c]) / g['M_diag'][c]) <= 1e-10
        assert np.max(np.abs(m['j'][c] - g['j'][c])) <= 1e-12 * np.max(np.abs(g['j'][c]))
        assert abs(m['M'][c].sum() - g['M_sum'][c]) <= 1e-9 * np.sum(np.abs(m['M'][c]))
    assert abs(m['null_likelihood'] - float(g['H0'])) <= 1e-12 * abs(float(g['H0']))
    # the whole matrices against the oracle
    ref = fo.map_visibilities(odht, u, v, V, w, 30., 40., 1e-3, -2e-3, vis_model='debris', H2=g['H2'], frequencies=freqs)
    for c in range(4):
        assert np.array_equal(ref['M'][c][rows, cols], g['M_sample'][c]) or \
            np.max(np.abs(ref['M'][c][rows, cols] - g['M_sample'][c]) / np.abs(g['M_sample'][c])) < 1e-9
        assert gram_ok(m['M'][c], ref['M'][c]) <= 1.0
        e = strict_errors(m['M'][c], ref['M'][c])
        print(f"\nconfig5 channel {c}: strict per-entry {e[0]:.3e}, /sqrt(MkkMll) {e[1]:.3e}, max-norm {e[2]:.3e}")
    # device-resident inputs give the same bits as host inputs (single chunk at this size)
    md = vm.map_visibilities(*[torch.from_numpy(x).cuda() for x in (u, v, V, w)], frequencies=torch.from_numpy(freqs).cuda())
    assert np.array_equal(md['M'], m['M']) and np.array_equal(md['j'], m['j']) and md['null_likelihood'] == m['null_likelihood']


def test_multichannel_matches_per_channel_calls(fb, golden):
    """One multi-channel call == separate calls on each channel's visibilities (the reference's loop,
    statistical_models.py:183-214) up to the summation order (the sort bins of a call follow the call's longest baseline),
    including a nearly empty channel and ragged channel sizes."""
    g = golden('mapping.npz')
    dht, vm = mapping_from_golden(fb, g)
    rng = np.random.default_rng(9)
    n = len(g['u'])
    freqs = rng.choice(np.array([1.0, 2.0, 5.0]), n, p=[0.7, 0.299, 0.001])
    m = vm.map_visibilities(g['u'], g['v'], g['V'], g['w'], frequencies=freqs)
    H0 = 0.0
    for c, f in enumerate(m['channels']):
        s = freqs == f
        one = vm.map_visibilities(g['u'][s], g['v'][s], g['V'][s], g['w'][s])
        d = np.sqrt(np.diag(one['M']))
        assert np.max(np.abs(one['M'] - m['M'][c]) / np.outer(d, d)) < 16 * EPS
        assert np.max(np.abs(one['j'] - m['j'][c])) <= 1e-14 * np.max(np.abs(one['j']))
        H0 += one['null_likelihood']
    assert m['M'].shape[0] == 3 and int(np.sum(freqs == 5.0)) < 20
    assert abs(H0 - m['null_likelihood']) <= 1e-13 * abs(H0)


def test_explicit_channel_list_with_an_empty_channel(fb, golden):
    """`channels=` fixes the channel list (what a rank of a sharded multi-frequency call must do): a channel without
    visibilities comes back as zeros, the others equal the call that only knows the channels present."""
    import torch
    g = golden('mapping.npz')
    dht, vm = mapping_from_golden(fb, g)
    rng = np.random.default_rng(11)
    n = len(g['u'])
    freqs = rng.choice(np.array([1.0, 5.0]), n, p=[0.6, 0.4])
    ref = vm.map_visibilities(g['u'], g['v'], g['V'], g['w'], frequencies=freqs)
    m = vm.map_visibilities(g['u'], g['v'], g['V'], g['w'], frequencies=freqs, channels=[1.0, 2.0, 5.0])
    assert m['M'].shape == (3,) + ref['M'].shape[1:] and list(m['channels']) == [1.0, 2.0, 5.0]
    assert not m['M'][1].any() and not m['j'][1].any()
    for c, r in ((0, 0), (2, 1)):
        d = np.sqrt(np.diag(ref['M'][r]))
        assert np.max(np.abs(m['M'][c] - ref['M'][r]) / np.outer(d, d)) < 16 * EPS
        assert np.max(np.abs(m['j'][c] - ref['j'][r])) <= 1e-14 * np.max(np.abs(ref['j'][r]))
    assert abs(m['null_likelihood'] - ref['null_likelihood']) <= 1e-13 * abs(ref['null_likelihood'])
    dev = [torch.as_tensor(np.ascontiguousarray(x)).cuda() for x in (g['u'], g['v'], g['V'], g['w'], freqs)]
    md = vm.map_visibilities(dev[0], dev[1], dev[2], dev[3], frequencies=dev[4], channels=[1.0, 2.0, 5.0])
    assert np.array_equal(md['M'], m['M']) and np.array_equal(md['j'], m['j'])
    with pytest.raises(ValueError):
        vm.map_visibilities(g['u'], g['v'], g['V'], g['w'], frequencies=freqs, channels=[1.0, 2.0])


@pytest.mark.parametrize('staging', ['pinned', 'pageable'])
def test_chunked_host_pipeline(fb, staging):
    """The K-deep copy / compute pipeline of the host entry point (several chunks over two lanes, growing chunk sizes,
    pinned inputs copied directly, pageable inputs gathered into the pinned staging ring by host threads): same result
    as the one-pass device entry point up to the summation order, bit-reproducible, multi-channel included."""
    import torch
    n, N = 300_000, 120
    u, v, V, w, odht = fo.synthetic_disc(n, N, seed=77)
    freqs = np.random.default_rng(1).choice(np.array([3., 1., 2.]), n)
    vm = fb.VM(fb.DHT(1.6 / fb.r2a, N), fb.FixedGeometry(30., 40., 1e-3, -2e-3), verbose=False)
    dev = vm.map_visibilities(*[torch.from_numpy(x).cuda() for x in (u, v, V, w)])
    devm = vm.map_visibilities(*[torch.from_numpy(x).cuda() for x in (u, v, V, w)], frequencies=torch.from_numpy(freqs).cuda())
    ctx = fb.lib.get_context()
    ctx.set_option('map_chunk', 20_000)
    ctx.set_option('map_kmax', 7)
    ctx.set_option('map_growth', 1.3)            # six chunks of growing size
    ctx.set_option('force_staging', 1 if staging == 'pageable' else 0)
    try:
        if staging == 'pinned':
            arrs = [torch.from_numpy(x).pin_memory() for x in (u, v, V, w)]
            hu, hv, hV, hw = [a.numpy() for a in arrs]
        else:
            hu, hv, hV, hw = u, v, V, w
        a = vm.map_visibilities(hu, hv, hV, hw)
        b = vm.map_visibilities(hu, hv, hV, hw)
        am = vm.map_visibilities(hu, hv, hV, hw, frequencies=freqs)
        sc = vm.map_visibilities(hu, hv, hV, 2.5)                     # broadcast scalar weight through the ring
    finally:
        ctx.set_option('map_chunk', 250_000)
        ctx.set_option("map_kmax", 8)
        ctx.set_option("map_growth", 1.5)
        ctx.set_option('force_staging', 0)
    assert np.array_equal(a['M'], b['M']) and np.array_equal(a['j'], b['j']) and a['null_likelihood'] == b['null_likelihood']
    d = np.sqrt(np.diag(dev['M']))
    assert np.max(np.abs(a['M'] - dev['M']) / np.outer(d, d)) < 64 * EPS
    assert np.max(np.abs(a['j'] - dev['j'])) <= 1e-13 * np.max(np.abs(dev['j']))
    assert abs(a['null_likelihood'] - dev['null_likelihood']) <= 1e-13 * abs(dev['null_likelihood'])
    for c in range(3):
        dc = np.sqrt(np.diag(devm['M'][c]))
        assert np.max(np.abs(am['M'][c] - devm['M'][c]) / np.outer(dc, dc)) < 64 * EPS
    ref = fo.map_visibilities(odht, u, v, V, w, 30., 40., 1e-3, -2e-3)
    assert gram_ok(a['M'], ref['M']) <= 1.0
    refs = fo.map_visibilities(odht, u, v, V, 2.5, 30., 40., 1e-3, -2e-3)
    assert gram_ok(sc['M'], refs['M']) <= 1.0 and abs(sc['null_likelihood'] - refs['null_likelihood']) <= 1e-12 * abs(refs['null_likelihood'])


def test_async_entry_point_and_device_side_checks(fb, golden):
    """fb_map_visibilities_dev_async + fb_map_sync: enqueue only; the q-range check and the J0-table check are made on the
    device and reported at the sync (statistical_models.py:512-535)."""
    import torch
    g = golden('mapping.npz')
    N = int(g['N'])
    dht, vm = mapping_from_golden(fb, g)
    ctx = fb.lib.get_context()
    ctx.dht_setup(dht)
    u, v, w = [torch.from_numpy(np.ascontiguousarray(g[k])).cuda() for k in ('u', 'v', 'w')]
    Vr = torch.view_as_real(torch.from_numpy(np.ascontiguousarray(g['V'])).cuda()).contiguous()
    out = torch.zeros(N * N + N + 1, dtype=torch.float64, device='cuda')
    geom = vm._geometry.device_scalars()
    args = (len(g['u']), u, v, Vr, w, 1, geom, 0, vm._model_scale(), None)
    torch.cuda.synchronize()
    ctx.map_visibilities_async(*args, True, float(dht.q[-1]), out[:N * N], out[N * N:N * N + N], out[N * N + N:])
    rc, qmin, qmax = ctx.map_sync()
    assert rc == 0
    ref = vm.map_visibilities(g['u'], g['v'], g['V'], g['w'])
    assert np.array_equal(out[:N * N].cpu().numpy().reshape(N, N), ref['M'])
    assert out[N * N + N].item() == ref['null_likelihood']
    assert qmax == np.max(g['q']) and qmin == np.min(g['q'])
    # out of range: flagged on the device, outputs untouched
    out.fill_(-7.0)
    ctx.map_visibilities_async(*args, True, 0.5 * float(np.max(g['q'])), out[:N * N], out[N * N:N * N + N], out[N * N + N:])
    rc, _, qmax2 = ctx.map_sync()
    assert rc == fb.lib.FB_E_QRANGE and qmax2 == qmax
    assert torch.all(out == -7.0).item()
    # data far beyond the collocation range without the check (FourierBesselFitter): the J0 table grows and the call is redone
    dht20 = fb.DHT(1.6 / fb.r2a, 20)
    vm20 = fb.VM(dht20, vm._geometry, verbose=False, check_qbounds=False)
    odht20 = fo.DHTTables(1.6 / fo.RAD_TO_ARCSEC, 20)
    got = vm20.map_visibilities(g['u'], g['v'], g['V'], g['w'])
    want = fo.map_visibilities(odht20, g['u'], g['v'], g['V'], g['w'], *[float(x) for x in g['geom']], check_qbounds=False)
    assert gram_ok(got['M'], want['M']) <= 1.0


def test_device_resident_predict(fb, golden):
    """fb_predict_visibilities_dev / fb_predict_sky_dev (FrankRadialFit.predict on device-resident sky baselines: deproject,
    H(q) I, undo_correction fused in one pass) against the reference fixtures and against the host entry point."""
    import torch
    from frank_b200.radial_fitters import FrankFitter
    g, f = golden('mapping.npz'), golden('fit_normal.npz')
    dht, vm = mapping_from_golden(fb, g)
    Vd = vm.predict_visibilities(g['I_pred'], torch.from_numpy(g['q']).cuda(), torch.from_numpy(g['wp']).cuda())
    assert Vd.is_cuda
    assert np.max(np.abs(Vd.cpu().numpy() - g['V_pred'])) <= 1e-13 * np.max(np.abs(g['V_pred']))
    FF = FrankFitter(1.6, int(g['N']), fb.FixedGeometry(*[float(x) for x in g['geom']]), verbose=False)
    sol = FF.fit(g['u'], g['v'], g['V'], g['w'])
    Vh = sol.predict(f['upred'], f['vpred'])
    Vs = sol.predict(torch.from_numpy(f['upred']).cuda(), torch.from_numpy(f['vpred']).cuda())
    assert Vs.is_cuda and Vs.dtype == torch.complex128
    assert np.max(np.abs(Vs.cpu().numpy() - Vh)) <= 1e-14 * np.max(np.abs(Vh))
    assert np.max(np.abs(Vs.cpu().numpy() - f['Vpred'])) <= 1e-7 * np.max(np.abs(f['Vpred']))
    # baselines beyond the J0 table: the table grows inside the call
    far = torch.from_numpy(np.array([3.0 * dht.q[-1], 10.0 * dht.q[-1]])).cuda()
    Vfar = vm.predict_visibilities(g['I_pred'], far, torch.zeros(2, dtype=torch.float64, device='cuda'))
    odht = fo.DHTTables(1.6 / fo.RAD_TO_ARCSEC, int(g['N']))
    want = fo.predict_visibilities(odht, g['I_pred'], far.cpu().numpy(), None, 'opt_thick', float(g['geom'][0]))
    assert np.max(np.abs(Vfar.cpu().numpy() - want)) <= 1e-12 * np.max(np.abs(g['V_pred']))
