"""Data utilities on the hot path's edge: the uv-plane binner and the weight estimate built on it.

API mirror of frank.utilities.UVDataBinner (frank/utilities.py:180-400).  Bin indices, counts and the weighted
sums come from the GPU (frank_b200/csrc/fb_bin.cu: bit-exact index arithmetic, stable sort by bin, fixed-order
segmented reduction); this class only normalises the O(nbins) results and wraps them in masked arrays.
"""
import numpy as np

from frank_b200 import _lib

import logging

__all__ = ['UVDataBinner', 'estimate_weights']


class UVDataBinner(object):
    r"""Average uv-data into bins of equal width (weighted means; empty bins masked).

    Parameters: uv (baselines / lambda), V (complex or real visibilities / Jy), weights (Jy^-2), bin_width
    (lambda) -- as frank/utilities.py:204.  `device` selects the GPU."""

    def __init__(self, uv, V, weights, bin_width, device=None):
        on_device = type(uv).__module__.startswith('torch') and getattr(uv, 'is_cuda', False)
        if on_device:
            # device-resident arrays (torch CUDA tensors): nothing but the O(nbins) results crosses to the host; the
            # per-visibility bin index stays on the device (`_idx` is an int32 CUDA tensor)
            import torch
            ctx = _lib.get_context(uv.device.index if device is None else device)
            uv = uv.contiguous()
            uv_max = float(uv.max().item())
        else:
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
        if on_device:
            is_c = V.is_complex()
            V = V.contiguous()
            if type(weights).__module__.startswith('torch'):
                w = weights.to(dtype=torch.float64, device=uv.device).reshape(-1).contiguous()
            else:
                w = torch.as_tensor(np.atleast_1d(np.asarray(weights, dtype=np.float64)), device=uv.device).contiguous()
            idx, counts, sums, err = ctx.uv_bin_dev(uv, V, w, bin_width, nbins)
            counts, sums, err = counts.cpu().numpy(), sums.cpu().numpy(), err.cpu().numpy()
            V = np.zeros(0, dtype=np.complex128 if is_c else np.float64)      # dtype carrier for the error array below
        else:
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


def estimate_weights(u, v=None, V=None, nbins=300, log=True, use_median=False, verbose=True, device=None):
    r"""Estimate the weights from the variance of the binned visibilities (frank/utilities.py:515-631).

    Same call forms as the reference: ``estimate_weights(u, v, V)``, ``estimate_weights(u, V)``,
    ``estimate_weights(u, V=V)``.  The binning (indices, counts, weighted sums, variance sums) runs on the GPU
    through :class:`UVDataBinner`; the O(nbins) post-processing below follows the reference line by line.

    Reference behaviour kept on purpose: the reference tests ``np.iscomplex(V.dtype)`` (utilities.py:598), which is
    False for every dtype object, so the variance of the REAL part alone is used even for complex visibilities
    (the docstring's "average of real and imaginary variance" branch is never taken).
    """
    if verbose:
        logging.info('  Estimating visibility weights')

    if V is None:                                                        # utilities.py:577-586
        if v is not None:
            V = v
            q = np.abs(u)
        else:
            raise ValueError("The visibilities, V, must be supplied")
    elif v is not None:
        q = np.hypot(u, v)
    else:
        q = np.abs(u)

    if log:                                                              # :588-590
        q = np.log(q)
        q -= q.min()

    bin_width = (q.max() - q.min()) / nbins                              # :592

    uvBin = UVDataBinner(q, V, np.ones_like(q), bin_width, device=device)

    if uvBin.bin_counts.max() == 1:                                      # :596-598
        raise ValueError("No bin contains more than one uv point, can't"
                         " estimate the variance. Use fewer bins.")

    var = uvBin.error.real ** 2 * uvBin.bin_counts                       # :600-603 (see the docstring)

    if use_median:                                                       # :605-609
        if verbose:
            logging.info('    Setting all weights as median binned visibility '
                         'variance')
        return np.full(len(u), 1 / np.ma.median(var[uvBin.bin_counts > 1]))
    else:
        if verbose:
            logging.info('    Setting weights according to baseline-dependent '
                         'binned visibility variance')
        # For bins with 1 uv point, use the average of the adjacent bins (:614-625)
        no_var = np.argwhere(uvBin.bin_counts == 1).reshape(-1)
        if len(no_var) > 0:
            good_var = np.argwhere(uvBin.bin_counts > 1).reshape(-1)
            loc = np.searchsorted(good_var, no_var, side='right')
            im = good_var[np.maximum(loc - 1, 0)]
            ip = good_var[np.minimum(loc, len(good_var) - 1)]
            var[no_var] = 0.5 * (var[im] + var[ip])

        bin_id = uvBin.determine_uv_bin(q)
        assert np.all(bin_id != -1), "Error in binning"  # Should never occur

        weights = 1 / var[bin_id]

        return weights
