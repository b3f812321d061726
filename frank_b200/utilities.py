"""Data utilities on the hot path's edge: the uv-plane binner.

API mirror of frank.utilities.UVDataBinner (frank/utilities.py:180-400).  Bin indices, counts and the weighted
sums come from the GPU (frank_b200/csrc/fb_bin.cu: bit-exact index arithmetic, stable sort by bin, fixed-order
segmented reduction); this class only normalises the O(nbins) results and wraps them in masked arrays.
"""
import numpy as np

from frank_b200 import _lib

__all__ = ['UVDataBinner']


class UVDataBinner(object):
    r"""Average uv-data into bins of equal width (weighted means; empty bins masked).

    Parameters: uv (baselines / lambda), V (complex or real visibilities / Jy), weights (Jy^-2), bin_width
    (lambda) -- as frank/utilities.py:204.  `device` selects the GPU."""

    def __init__(self, uv, V, weights, bin_width, device=None):
        ctx = _lib.get_context(device)
        uv = np.ascontiguousarray(uv, dtype=np.float64)
        uv_max = ctx.uv_max(uv)
        nbins = np.ceil(uv_max / bin_width).astype('int')                  # utilities.py:205-208
        if nbins * bin_width < uv_max:
            nbins += 1
        nbins = int(nbins)
        bins = np.arange(nbins + 1, dtype='float64') * bin_width
        self._bins, self._nbins, self._norm = bins, nbins, 1 / bin_width
        self._ctx = ctx
        is_c = np.iscomplexobj(V)
        w = np.ones_like(uv) * weights
        idx, counts, sums, err = ctx.uv_bin(uv, V, w, bin_width, nbins)
        self._idx = idx
        bin_uv, bin_wgt = sums[:, 0].copy(), sums[:, 1].copy()
        bin_vis = (sums[:, 2] + 1j * sums[:, 3]) if is_c else sums[:, 2].copy()
        has = counts > 0
        bin_uv[has] /= bin_wgt[has]                                          # :220-224
        bin_vis[has] /= bin_wgt[has]
        mask = counts == 0
        self._uv = np.ma.masked_where(mask, bin_uv)
        self._V = np.ma.masked_where(mask, bin_vis)
        self._w = np.ma.masked_where(mask, bin_wgt)
        self._count = np.ma.masked_where(mask, counts)
        self._uv_left = np.ma.masked_where(mask, bins[:-1])
        self._uv_right = np.ma.masked_where(mask, bins[1:])
        # standard error of the mean (:236-264); single-count bins stay nan (+0j), as in the reference, whose
        # `bin_vis_err[idx1].real = ...` assigns into a temporary
        bin_err = np.full(nbins, np.nan, dtype=np.asarray(V).dtype if is_c else np.float64)
        many = counts > 1
        e = err.copy()
        e[many] /= (bin_wgt[many] ** 2 * (1 - 1 / counts[many]))[:, None]
        val = np.sqrt(e[many, 0])
        if is_c:
            val = val + 1.j * np.sqrt(e[many, 1])
        bin_err[many] = val
        bin_err[mask] = np.nan
        self._Verr = np.ma.masked_where(mask, bin_err)

    def determine_uv_bin(self, uv):
        r"""Bin index of each baseline; -1 outside the binned range (frank/utilities.py:267-298)."""
        bins, nbins = self._bins, self._nbins
        uv = np.asarray(uv, dtype=np.float64)
        idx = np.floor(uv * self._norm).astype('int32')
        ok = (idx >= 0) & (idx <= nbins)
        safe = np.where(ok, idx, 0)
        idx[ok & (uv < bins[safe])] -= 1
        idx[uv == bins[nbins]] -= 1
        too_high = idx >= nbins
        idx[too_high] = -1
        sel = ~too_high
        tmp = idx[sel]
        inc = (uv[sel] >= bins[tmp + 1]) & (tmp + 1 < nbins)
        tmp[inc] += 1
        idx[sel] = tmp
        return idx

    def __len__(self):
        return len(self._uv)

    uv = property(lambda self: self._uv, doc="Binned uv points, lambda")
    V = property(lambda self: self._V, doc="Binned visibility, Jy")
    weights = property(lambda self: self._w, doc="Binned weights, Jy^-2")
    error = property(lambda self: self._Verr, doc="Uncertainty on the binned visibilities, Jy")
    bin_counts = property(lambda self: self._count, doc="Number of points in each bin")
    bin_edges = property(lambda self: [self._uv_left, self._uv_right], doc="Edges of the histogram bins")
