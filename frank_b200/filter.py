"""Power-spectrum prior and its fixed-point optimiser.

API mirror of frank.filter (frank/filter.py).  The smoothing matrix is a small host-side set-up
(O(N)); the fixed-point update itself runs on the GPU (frank_b200/csrc/fb_solve.cu, k_ps_update).
"""
import numpy as np

from frank_b200 import _lib

__all__ = ['spectral_smoothing_matrix', 'CriticalFilter']


def smoothing_bands(DHT, weights):
    r"""The five diagonals of T_ij = w Delta^T diag(dc) Delta (frank/filter.py:41-62) as T[i, i + o] for
    o = -2..2 in rows 0..4 of a [5, N] array (zero where i + o falls outside).

    Delta is the second-difference operator in log q: rows 1..N-2 carry
    1/(dc_m de_{m-1}), -(1/de_m + 1/de_{m-1})/dc_m, 1/(dc_m de_m); rows 0 and N-1 are zero."""
    N = DHT.size
    lq = np.log(DHT.q)
    dc = (lq[2:] - lq[:-2]) / 2
    de = np.diff(lq)
    lo = np.zeros(N); mid = np.zeros(N); hi = np.zeros(N); dce = np.zeros(N)
    lo[1:-1] = 1 / (dc * de[:-1])                  # Delta[m, m-1]
    mid[1:-1] = -(1 / de[1:] + 1 / de[:-1]) / dc   # Delta[m, m]
    hi[1:-1] = 1 / (dc * de[1:])                   # Delta[m, m+1]
    dce[1:-1] = dc
    rows = {-1: lo, 0: mid, 1: hi}
    T = np.zeros([5, N])
    # T[i, k] = w * sum_m Delta[m, i] dce[m] Delta[m, k],  Delta[m, m + a] = rows[a][m]
    for a in (-1, 0, 1):
        for c in (-1, 0, 1):
            o = c - a                               # k - i
            for m in range(1, N - 1):
                i = m + a
                if 0 <= i < N and 0 <= i + o < N:
                    T[o + 2, i] += rows[a][m] * dce[m] * rows[c][m]
    return weights * T


def spectral_smoothing_matrix(DHT, weights):
    r"""Sparse spectral smoothing prior matrix T_ij (frank/filter.py:23-62)."""
    import scipy.sparse
    N = DHT.size
    T = smoothing_bands(DHT, weights)
    diags = [T[o + 2, max(0, -o):N - max(0, o)] for o in range(-2, 3)]
    return scipy.sparse.diags(diags, list(range(-2, 3)), shape=(N, N), format='csc')


def banded_ldl(bands):
    r"""L D L^T factorisation of the SPD pentadiagonal matrix whose diagonals are bands[o + 2, i] = A[i, i + o].
    Returns [3, N]: D, L[i, i-1], L[i, i-2] (unit lower-triangular L).  O(N), host side, once per filter."""
    N = bands.shape[1]
    d = np.zeros(N); l1 = np.zeros(N); l2 = np.zeros(N)
    for i in range(N):
        a0 = bands[2, i]
        a1 = bands[1, i] if i >= 1 else 0.0        # A[i, i-1]
        a2 = bands[0, i] if i >= 2 else 0.0        # A[i, i-2]
        if i >= 2:
            l2[i] = a2 / d[i - 2]
        if i >= 1:
            l1[i] = (a1 - (l2[i] * d[i - 2] * l1[i - 1] if i >= 2 else 0.0)) / d[i - 1]
        d[i] = a0 - (l1[i] ** 2 * d[i - 1] if i >= 1 else 0.0) - (l2[i] ** 2 * d[i - 2] if i >= 2 else 0.0)
    return np.stack([d, l1, l2])


def banded_inverse(ldl):
    r"""Dense inverse of A = L D L^T given banded_ldl(A): column c solves L D L^T x = e_c.  O(N^2)."""
    d, l1, l2 = ldl
    N = d.size
    X = np.eye(N)
    for i in range(N):                       # L y = e
        if i >= 1:
            X[i] -= l1[i] * X[i - 1]
        if i >= 2:
            X[i] -= l2[i] * X[i - 2]
    X /= d[:, None]
    for i in range(N - 1, -1, -1):           # L^T x = y
        if i + 1 < N:
            X[i] -= l1[i + 1] * X[i + 1]
        if i + 2 < N:
            X[i] -= l2[i + 2] * X[i + 2]
    return X


class CriticalFilter(object):
    """Optimiser for power-spectrum priors (frank/filter.py:64-263).

    Parameters: DHT, alpha (>= 1), p_0 (>= 0), weights_smooth (>= 0), tol."""

    def __init__(self, DHT, alpha, p_0, weights_smooth, tol=1e-3):
        self._DHT = DHT
        self._alpha = alpha
        self._p_0 = p_0
        self._rho = 1.0
        self._tol = tol
        self._weights_smooth = weights_smooth
        self._Tij = spectral_smoothing_matrix(DHT, weights_smooth)
        bands = smoothing_bands(DHT, weights_smooth)
        bands[2] += 1.0                                      # T + I  (filter.py:155)
        self._ldl = banded_ldl(bands)
        # dense inverse of the SPD pentadiagonal T + I through its banded factorisation (host, once per filter): the
        # device then solves (T + I) tau = rhs with one matrix-vector product per iteration
        self._Tinv = banded_inverse(self._ldl)

    def update_power_spectrum(self, fit, device=None):
        """One fixed-point update of the power spectrum for the current fit (frank/filter.py:154-177).

        Dispatches on the fit, as the reference's `fit.MAP` / `fit.Dsolve` interface does: a log-normal fit updates with
        its own Hessian factor, a Gaussian fit with a Cholesky factor runs one step of the device loop, and a Gaussian
        fit that went through the SVD pseudo-inverse branch (statistical_models.py:747-755) takes the reference's
        generic formula on its `Dsolve`."""
        if hasattr(fit, '_update_power_spectrum'):                       # LogNormalMAPModel
            return fit._update_power_spectrum(self._alpha, self._p_0, self._Tinv)
        if getattr(fit, '_Dsvd', None) is not None:
            Ykm = self._DHT.coefficients()
            Tr1 = np.dot(Ykm, fit.MAP) ** 2
            Tr2 = np.einsum('ij,ji->i', Ykm, fit.Dsolve(Ykm.T))
            pi = fit.power_spectrum
            beta = (self._p_0 + 0.5 * (Tr1 + Tr2)) / pi - (self._alpha - 1.0 + 0.5 * self._rho)
            return np.exp(np.dot(self._Tinv, beta + np.log(pi)))
        ctx = _lib.get_context(device)
        ctx.dht_setup(self._DHT)
        # one pass of the device loop: max_iter = 0 lets exactly one update through (count <= max_iter)
        out = ctx.frank_normal_loop(fit._M, fit._j, fit.power_spectrum, self._alpha, self._p_0, self._Tinv,
                                    self._tol, 0, want_chol=False)
        return out['p'][0]

    def check_convergence(self, pi_new, pi_old):
        return np.all(np.abs(pi_new - pi_old) <= self._tol * pi_new)

    def covariance_MAP(self, fit, ret_inv=False):
        """Covariance of the power spectrum at maximum likelihood (frank/filter.py:184-227); post-fit helper."""
        Ykm = self._DHT.coefficients()
        mq = np.dot(Ykm, fit.MAP)
        mqq = np.outer(mq, mq)
        Dqq = np.dot(Ykm, np.dot(fit.covariance, Ykm.T))
        p = fit.power_spectrum
        hess = np.diag(self._p_0 / p + 0.5 * (mq ** 2 + np.diag(Dqq)) / p) + self._Tij.toarray() \
            - 0.5 * np.outer(1 / p, 1 / p) * (2 * mqq + Dqq) * Dqq
        if ret_inv:
            return hess
        # inverse through a device Cholesky factorisation (fb_gaussian_fit without a prior) and N device solves
        ctx = _lib.get_context(getattr(fit, '_device', None))
        ctx.dht_setup(self._DHT)
        _, chol, _, rc = ctx.gaussian_fit(hess, np.zeros(self._DHT.size), None)
        if rc == _lib.FB_E_NOTPD:
            raise np.linalg.LinAlgError("covariance_MAP: the Hessian of the power-spectrum posterior is not positive definite")
        return ctx.chol_solve(np.triu(chol[0]), np.eye(self._DHT.size))

    def log_prior(self, p):
        """log P(p) up to a constant (frank/filter.py:229-263)."""
        xi = self._p_0 / p
        like = -np.sum(xi + (self._alpha - 1) * np.log(xi))
        tau = np.log(p)
        return like - 0.5 * np.dot(tau, self._Tij.dot(tau))
