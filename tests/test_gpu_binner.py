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
