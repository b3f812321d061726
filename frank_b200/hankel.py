"""Discrete Hankel transform tables (host side, O(N^2) set-up).

Same public interface as frank.hankel.DiscreteHankelTransform (frank/hankel.py:25-294): the
Baddour & Chouinard (2015) DHT with collocation points at the zeros of J_nu.  The tables are
small and built once per fit on the host; the per-visibility evaluation H(q) -- the hot part of
`coefficients(q)` -- is what the CUDA Gram kernel replaces (frank_b200/csrc/fb_gram.cu).
"""
import numpy as np
from scipy.special import j0, j1, jn_zeros, jv

__all__ = ['DiscreteHankelTransform']


def _bessel_pair(nu):
    """(J_nu, J_{nu+1}) evaluators; orders 0 and 1 use the dedicated routines (hankel.py:58-66)."""
    if nu == 0:
        return j0, j1
    if nu == 1:
        return j1, (lambda x: jv(2, x))
    return (lambda x: jv(nu, x)), (lambda x: jv(nu + 1, x))


class DiscreteHankelTransform(object):
    r"""DHT of order `nu` on [0, Rmax] with `N` collocation points.

        H[f](q) = \int_0^{Rmax} f(r) J_nu(2 pi q r) 2 pi r dr

    Parameters follow frank/hankel.py:55: Rmax (radians), N, nu=0.
    """

    def __init__(self, Rmax, N, nu=0):
        self._jnu0, self._jnup = _bessel_pair(nu)
        self._N, self._nu = N, nu
        self._Rmax = Rmax
        self._Rnk, self._Qnk, self._j_nk, self._j_nN = self._points(Rmax, N, nu)
        self._Qmax = self._j_nN / (2 * np.pi * Rmax)

        # Y_km = 2 / (j_N+1 J_{nu+1}(j_k)^2) J_nu(j_k j_m / j_N+1)        (hankel.py:84-87)
        Jp = np.outer(np.ones_like(self._j_nk), self._jnup(self._j_nk))
        arg = np.prod(np.meshgrid(self._j_nk, self._j_nk / self._j_nN), axis=0)
        self._Ykm = (2 / (self._j_nN * Jp * Jp)) * self._jnu0(arg)
        self._scale_factor = 1 / self._jnup(self._j_nk) ** 2               # hankel.py:89

    @staticmethod
    def _points(Rmax, N, nu):
        zeros = jn_zeros(nu, N + 1)
        j_nk, j_nN = zeros[:-1], zeros[-1]
        Qmax = j_nN / (2 * np.pi * Rmax)
        return Rmax * (j_nk / j_nN), Qmax * (j_nk / j_nN), j_nk, j_nN

    @classmethod
    def get_collocation_points(cls, Rmax, N, nu=0):
        """Radius and frequency collocation points (hankel.py:95-125)."""
        Rnk, Qnk, _, _ = cls._points(Rmax, N, nu)
        return Rnk, Qnk

    def _direction(self, direction):
        if direction == 'forward':
            return self._Rmax, self._Qmax
        if direction == 'backward':
            return self._Qmax, self._Rmax
        raise AttributeError("direction must be one of {}".format(['forward', 'backward']))

    def transform(self, f, q=None, direction='forward'):
        """Hankel transform of f sampled at the collocation points (hankel.py:127-165)."""
        if q is None:
            span, _ = self._direction(direction)
            return ((2 * np.pi * span ** 2) / self._j_nN) * np.dot(self._Ykm, f)
        return 1.0 * np.dot(self.coefficients(q, direction=direction), f)

    def coefficients(self, q=None, direction='forward'):
        """Transform matrix Y with H[f](q) = Y f (hankel.py:167-204)."""
        _, conj = self._direction(direction)
        norm = 1 / (np.pi * conj ** 2)
        k = 1. / conj
        if q is None:
            return 0.5 * self._j_nN * norm * self._Ykm
        return (norm * self._scale_factor) * self._jnu0(np.outer(k * q, self._j_nk))

    def interpolation_coefficients(self, q, space='Real'):
        """Fourier-Bessel interpolation matrix (hankel.py:206-236)."""
        if space == 'Real':
            x = np.atleast_1d(2 * np.pi * q * self._Qmax)
        elif space == 'Fourier':
            x = np.atleast_1d(2 * np.pi * q * self._Rmax)
        else:
            raise ValueError(f"Space must be one of 'Real' or 'Fourier', not {space}.")
        num = np.outer(np.where(x < self._j_nN, self._jnu0(x), 0), 2 * self._j_nk / self._jnup(self._j_nk))
        return num / (self._j_nk.reshape(1, -1) ** 2 - x.reshape(-1, 1) ** 2)

    def interpolate(self, f, q, space='Real'):
        """Interpolate f from the collocation points to q (hankel.py:238-264)."""
        return np.dot(self.interpolation_coefficients(q, space), f)

    r = property(lambda self: self._Rnk, doc="Radius points")
    Rmax = property(lambda self: self._Rmax, doc="Maximum radius")
    q = property(lambda self: self._Qnk, doc="Frequency points")
    Qmax = property(lambda self: self._Qmax, doc="Maximum frequency")
    size = property(lambda self: self._N, doc="Number of points used in the DHT")
    order = property(lambda self: self._nu, doc="Order of the Bessel function")
