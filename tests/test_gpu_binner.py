"""GPU parity tests of UVDataBinner (fb_uv_bin) against fixtures produced by the unmodified reference."""
import numpy as np
import pytest

from oracle import frank_oracle as fo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('tag', ['a', 'b', 'edge'])
def test_binner_vs_reference_golden(golden, tag):
    from frank_b200.utilities import UVDataBinner
    g = golden(f'uvbin_{tag}.npz')
    b = UVDataBinner(g['uv_in'], g['V_in'], g['w_in'], float(g['width']))
    assert np.array_equal(b._idx, g['idx'].astype(np.int32))                    # bit-exact indices
    assert np.array_equal(b.determine_uv_bin(g['uv_in']), g['idx'])
    assert np.array_equal(b.bin_counts.filled(0), g['counts'])                  # bit-exact counts
    assert np.array_equal(np.ma.getmaskarray(b.uv), g['mask'])
    ok = ~g['mask']
    for name, got in [('uv', b.uv), ('V', b.V), ('weights', b.weights)]:
        ref = g[name][ok]
        assert np.max(np.abs(got.filled(0)[ok] - ref)) <= 1e-13 * np.max(np.abs(ref)), name
    e, eref = b.error.filled(np.nan)[ok], g['error'][ok]
    assert np.array_equal(np.isnan(e.real), np.isnan(eref.real))                # single-count bins: nan, as the reference
    fin = ~np.isnan(eref.real)
    assert np.max(np.abs(e[fin] - eref[fin])) <= 1e-10 * np.max(np.abs(eref[fin]))
    if 'left' in g.files:
        assert np.array_equal(b.bin_edges[0].filled(0)[ok], g['left'][ok])
        assert np.array_equal(b.bin_edges[1].filled(0)[ok], g['right'][ok])


def test_binner_large_vs_oracle():
    from frank_b200.utilities import UVDataBinner
    rng = np.random.default_rng(9)
    n = 3_000_000
    uv = 2e6 * np.sqrt(rng.uniform(0, 1, n))
    V = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    w = rng.uniform(0.5, 2, n)
    ref = fo.uv_bin(uv, V, w, 1e3)
    b = UVDataBinner(uv, V, w, 1e3)
    assert np.array_equal(b._idx, ref['idx'])
    assert np.array_equal(b.bin_counts.filled(0), ref['counts'])
    ok = ~ref['mask']
    assert np.max(np.abs(b.V.filled(0)[ok] - ref['V'][ok])) <= 1e-12 * np.max(np.abs(ref['V'][ok]))
    assert np.max(np.abs(b.weights.filled(0)[ok] - ref['weights'][ok])) <= 1e-13 * np.max(ref['weights'][ok])
    # deterministic
    b2 = UVDataBinner(uv, V, w, 1e3)
    assert np.array_equal(b.V.filled(0), b2.V.filled(0)) and np.array_equal(b.error.filled(0), b2.error.filled(0), equal_nan=True)


@pytest.mark.parametrize('case', ['sorted_coarse', 'real_scalar_w', 'medium', 'wide_keys'])
def test_binner_device_entry_vs_oracle(case):
    """fb_uv_bin_dev (device-resident arrays) against the oracle on ragged shapes: already sorted baselines with a
    coarse binning (a block of warps per bin), real visibilities with one scalar weight, a medium binning (2 / 4 warps
    per bin), and more than 65536 bins (three radix passes); empty bins inside and at the end."""
    import torch
    from frank_b200 import _lib
    rng = np.random.default_rng(31)
    n, width = 400_003, 1e3
    uv = 2e6 * np.sqrt(rng.uniform(0, 1, n))
    V = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    w = rng.uniform(0.5, 2, n)
    if case == 'sorted_coarse':
        uv = np.sort(uv); width = 2.5e4                       # 80 bins
    elif case == 'real_scalar_w':
        V = V.real.copy(); w = np.array([1.7])
    elif case == 'medium':
        width = 1.2e3; uv[uv < 3e5] += 3e5                    # ~1667 bins, the first 250 empty
    elif case == 'wide_keys':
        width = 20.0; uv[:1000] = rng.uniform(0, 40.0, 1000)  # 1e5 bins, most of them with 0..8 points
    uv_max = uv.max()
    nbins = int(np.ceil(uv_max / width))
    if nbins * width < uv_max:
        nbins += 1
    ref = fo.uv_bin(uv, V, np.ones_like(uv) * w, width)
    ctx = _lib.get_context(0)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    idx, counts, sums, err = ctx.uv_bin_dev(d(uv), d(V), d(w), width, nbins)
    idx, counts, sums, err = idx.cpu().numpy(), counts.cpu().numpy(), sums.cpu().numpy(), err.cpu().numpy()
    assert np.array_equal(idx, ref['idx'])
    assert np.array_equal(counts, ref['counts'])
    assert counts.sum() == n
    ok = counts > 0
    assert np.allclose(sums[ok, 1], ref['weights'][ok], rtol=1e-13, atol=0)
    assert np.allclose(sums[ok, 0] / sums[ok, 1], ref['uv'][ok], rtol=1e-13, atol=0)
    Vb = (sums[ok, 2] + 1j * sums[ok, 3]) / sums[ok, 1]
    assert np.max(np.abs(Vb - ref['V'][ok])) <= 1e-12 * np.max(np.abs(ref['V'][ok]))
    assert np.all(sums[~ok] == 0) and np.all(err[~ok] == 0)
    # the host entry point goes through the same kernels: identical bits
    idx2, counts2, sums2, err2 = ctx.uv_bin(uv, V, np.ones_like(uv) * w if w.size > 1 else w, width, nbins)
    assert np.array_equal(idx2, idx) and np.array_equal(counts2, counts)
    assert np.array_equal(sums2, sums) and np.array_equal(err2, err)
    # variance sums against a direct evaluation with the device's own means
    many = counts > 1
    mu = np.zeros(nbins, dtype=complex); mu[ok] = Vb
    ww = (np.ones_like(uv) * w) ** 2
    e_re = np.bincount(idx, weights=ww * (np.real(V) - mu.real[idx]) ** 2, minlength=nbins)
    e_im = np.bincount(idx, weights=ww * (np.imag(V) - mu.imag[idx]) ** 2, minlength=nbins)
    assert np.allclose(err[many, 0], e_re[many], rtol=1e-10, atol=0)
    assert np.allclose(err[many, 1], e_im[many], rtol=1e-10, atol=1e-300)


def test_estimate_weights_vs_reference_golden(golden):
    """estimate_weights on top of the GPU binner against the unmodified reference (tests/golden/make_golden.py):
    bin membership is exact, so the weights agree to the round-off of the per-bin variance sums."""
    from frank_b200.utilities import estimate_weights
    g = golden('estweights.npz')
    u, v, V = g['u'], g['v'], g['V']
    for tag, kw in [('log', dict(nbins=300)), ('lin', dict(nbins=300, log=False)), ('fine', dict(nbins=8000)),
                    ('median', dict(nbins=300, use_median=True))]:
        got = np.ma.filled(estimate_weights(u, v, V, verbose=False, **kw), np.nan)
        assert got.shape == g[tag].shape
        assert np.allclose(got, g[tag], rtol=1e-11, atol=0), tag
    got = np.ma.filled(estimate_weights(np.hypot(u, v), V.real, nbins=100, verbose=False), np.nan)
    assert np.allclose(got, g['uV'], rtol=1e-11, atol=0)
    with pytest.raises(ValueError):
        estimate_weights(u)


def test_device_pipeline_correction_then_binning():
    """Device-resident pipeline of BASELINE.json configs[4]: SourceGeometry.apply_correction on CUDA tensors (phase
    shift + deprojection, geometry.py:202-236), q = hypot(u', v'), UVDataBinner on the deprojected baselines -- against
    the oracle's NumPy chain.  up, vp, wp and q are bit-equal to NumPy's; V' to a few ulp (sin / cos differ by an ulp);
    bin indices and counts exact."""
    import torch
    from frank_b200 import _lib
    from frank_b200.geometry import FixedGeometry
    from frank_b200.utilities import UVDataBinner
    n = 1_000_003
    u, v, V, w, dht = fo.synthetic_disc(n, 100, analytic=True)
    geom = FixedGeometry(30., 40., 1e-3, -2e-3)
    up, vp, wp, Vp = fo.apply_correction(u, v, V, 30., 40., 1e-3, -2e-3)
    q = np.hypot(up, vp)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    du, dv, dV, dw = d(u), d(v), d(V), d(w)
    gup, gvp, gwp, gVp = geom.apply_correction(du, dv, dV, use3D=True)
    assert np.array_equal(gup.cpu().numpy(), up) and np.array_equal(gvp.cpu().numpy(), vp) and np.array_equal(gwp.cpu().numpy(), wp)
    assert np.max(np.abs(gVp.cpu().numpy() - Vp)) <= 4 * np.finfo(float).eps * np.max(np.abs(Vp))
    g2 = geom.deproject(du, dv)
    assert len(g2) == 2 and np.array_equal(g2[0].cpu().numpy(), up)
    ctx = _lib.get_context(0)
    *_, gq = ctx.apply_correction_dev(du, dv, None, geom.device_scalars(), want_q=True)
    assert np.array_equal(gq.cpu().numpy(), q)                                   # glibc-exact hypot
    ref = fo.uv_bin(q, Vp, w, 1e3)
    b = UVDataBinner(gq, gVp, dw, 1e3)
    assert b._idx.is_cuda and np.array_equal(b._idx.cpu().numpy(), ref['idx'])
    assert np.array_equal(b.bin_counts.filled(0), ref['counts'])
    ok = ~ref['mask']
    assert np.max(np.abs(b.V.filled(0)[ok] - ref['V'][ok])) <= 1e-12 * np.max(np.abs(ref['V'][ok]))
    assert np.max(np.abs(b.uv.filled(0)[ok] - ref['uv'][ok])) <= 1e-13 * np.max(ref['uv'][ok])
    fin = ~np.isnan(ref['error'].real) & ok
    e = b.error.filled(np.nan)
    assert np.array_equal(np.isnan(e.real[ok]), np.isnan(ref['error'].real[ok]))
    assert np.max(np.abs(e[fin] - ref['error'][fin])) <= 1e-9 * np.max(np.abs(ref['error'][fin]))
    # host-array and device-tensor constructors run the same kernels on the same values
    bh = UVDataBinner(q, gVp.cpu().numpy(), w, 1e3)
    assert np.array_equal(bh.V.filled(0), b.V.filled(0)) and np.array_equal(bh.error.filled(0), b.error.filled(0), equal_nan=True)
