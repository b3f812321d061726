"""Source geometry: phase-centre shift, deprojection, and the geometry fitters.

API mirror of frank.geometry (frank/geometry.py:41-763).  On the fit path the per-visibility arithmetic runs on the
GPU (frank_b200/csrc/fb_prep.cu); the NumPy helpers here serve the small host-side uses (re-projection of a handful
of points, `undo_correction` of predicted visibilities) and define the host scalars handed to the device
(`device_scalars`).  `FitGeometryFourierBessel` is a pure caller of the hot path: every
residual evaluation of its Levenberg-Marquardt search is one GPU mapping + solve + GPU prediction.
"""
import logging

import numpy as np
from scipy.optimize import least_squares

from frank_b200.constants import rad_to_arcsec, deg_to_rad

__all__ = ['apply_phase_shift', 'deproject', 'rescale_total_flux', 'SourceGeometry', 'FixedGeometry',
           'FitGeometryGaussian', 'FitGeometryFourierBessel']


def _is_cuda_tensor(x):
    return type(x).__module__.startswith('torch') and getattr(x, 'is_cuda', False)


def _fix_inc_and_PA_ranges(inc, PA):
    """Make sure the inclination and PA are in the ranges [0,90] and [0-180] (frank/geometry.py:33-39)."""
    inc = inc % 180
    PA = PA % 180
    if inc > 90:
        inc = 180 - inc
    return inc, PA


def apply_phase_shift(u, v, V, dRA, dDec, inverse=False):
    r"""Shift the phase centre by (dRA, dDec) arcsec (frank/geometry.py:41-79)."""
    phi = u * (dRA * (2. * np.pi / rad_to_arcsec)) + v * (dDec * (2. * np.pi / rad_to_arcsec))
    rot = np.cos(phi) + 1j * np.sin(phi)
    return V / rot if inverse else V * rot


def deproject(u, v, inc, PA, inverse=False):
    r"""Rotate by PA and (de)compress by cos(inc), degrees (frank/geometry.py:82-131)."""
    ci, si = np.cos(inc * deg_to_rad), np.sin(inc * deg_to_rad)
    ct, st = np.cos(PA * deg_to_rad), np.sin(PA * deg_to_rad)
    if inverse:
        st = st * -1
        u = u / ci
    up = u * ct - v * st
    vp = u * st + v * ct
    if inverse:
        return up, vp
    return up * ci, vp, up * si


def rescale_total_flux(V, weights, inc):
    r"""Optically-thick amplitude / weight rescaling (frank/geometry.py:133-170)."""
    c = np.cos(inc * deg_to_rad)
    return V.real / c, weights * c ** 2


class SourceGeometry(object):
    """Geometry correction with parameters inc, PA (deg), dRA, dDec (arcsec)
    (frank/geometry.py:173-370)."""

    def __init__(self, inc=None, PA=None, dRA=None, dDec=None):
        self._inc, self._PA, self._dRA, self._dDec = inc, PA, dRA, dDec

    def apply_correction(self, u, v, V, use3D=False):
        if _is_cuda_tensor(u):                      # device-resident arrays: one fused pass (fb_prep.cu, k_apply_correction)
            from frank_b200 import _lib
            up, vp, wp, Vp = _lib.get_context(u.device.index).apply_correction_dev(u, v, V, self.device_scalars())
            return (up, vp, wp, Vp) if use3D else (up, vp, Vp)
        Vp = apply_phase_shift(u, v, V, self._dRA, self._dDec, inverse=True)
        up, vp, wp = deproject(u, v, self._inc, self._PA)
        return (up, vp, wp, Vp) if use3D else (up, vp, Vp)

    def undo_correction(self, u, v, V):
        up, vp = self.reproject(u, v)
        return up, vp, apply_phase_shift(up, vp, V, self._dRA, self._dDec, inverse=False)

    def deproject(self, u, v, use3D=False):
        if _is_cuda_tensor(u):
            from frank_b200 import _lib
            up, vp, wp, _ = _lib.get_context(u.device.index).apply_correction_dev(u, v, None, self.device_scalars())
            return (up, vp, wp) if use3D else (up, vp)
        out = deproject(u, v, self._inc, self._PA)
        return out if use3D else out[:2]

    def reproject(self, u, v):
        return deproject(u, v, self._inc, self._PA, inverse=True)

    def rescale_total_flux(self, V, weights):
        return rescale_total_flux(V, weights, self._inc)

    def fit(self, u, v, V, weights):
        """Fixed geometries have nothing to fit (frank/geometry.py:319-335)."""
        return

    def clone(self):
        return FixedGeometry(self.inc, self.PA, self.dRA, self.dDec)

    def device_scalars(self):
        """Host-side scalars exactly as the reference forms them before touching the arrays
        (geometry.py:69-70, 111-115); passed by value to the CUDA pre-pass."""
        from frank_b200._lib import FBGeometry
        inc = self._inc * deg_to_rad
        PA = self._PA * deg_to_rad
        return FBGeometry(self._dRA * (2. * np.pi / rad_to_arcsec), self._dDec * (2. * np.pi / rad_to_arcsec),
                          float(np.cos(PA)), float(np.sin(PA)), float(np.cos(inc)), float(np.sin(inc)))

    dRA = property(lambda self: self._dRA, doc="Phase centre offset in right ascension, arcsec")
    dDec = property(lambda self: self._dDec, doc="Phase centre offset in declination, arcsec")
    PA = property(lambda self: self._PA, doc="Position angle, deg")
    inc = property(lambda self: self._inc, doc="Inclination, deg")

    @property
    def rescale_factor(self):
        return 1.0 / np.cos(self._inc * deg_to_rad)

    def __repr__(self):
        return "SourceGeometry(inc={}, PA={}, dRA={}, dDec={})".format(self.inc, self.PA, self.dRA, self.dDec)


class FixedGeometry(SourceGeometry):
    """Pre-determined geometry (frank/geometry.py:372-401)."""

    def __init__(self, inc, PA, dRA=0.0, dDec=0.0):
        super(FixedGeometry, self).__init__(inc, PA, dRA, dDec)

    def __repr__(self):
        return "FixedGeometry(inc={}, PA={}, dRA={}, dDEC={})".format(self.inc, self.PA, self.dRA, self.dDec)


class FitGeometryGaussian(SourceGeometry):
    """Determine the disc geometry by fitting a Gaussian in Fourier space (frank/geometry.py:404-620).

    Same parameters and behaviour as the reference: `inc_pa` / `phase_centre` fix part of the geometry, `guess` is
    [inc, PA, dRA, dDec].  Like the reference this is a host-side SciPy Levenberg-Marquardt fit of six parameters
    (not on the GPU path: SURVEY 8f keeps it out of the hot path)."""

    def __init__(self, inc_pa=None, phase_centre=None, guess=None):
        super(FitGeometryGaussian, self).__init__()
        self._inc_pa = inc_pa
        self._phase_centre = phase_centre
        if guess is None:
            guess = [10.0, 10.0, 0.0, 0.0, 1.0, 1.0]
        else:
            guess.extend([1.0, 1.0])
        if self._inc_pa is not None:
            guess[0], guess[1] = self._inc_pa
        if self._phase_centre is not None:
            guess[2], guess[3] = self._phase_centre
        self._guess = guess

    def fit(self, u, v, V, weights):
        if self._inc_pa and self._phase_centre:
            logging.info('    You requested a Gaussian fit to determine the geometry,'
                         ' but you provided values for inclination, PA, and the phase offset.'
                         ' --> Using your provided values (not fitting for the geometry)')
            self._inc, self._PA = self._inc_pa
            self._dRA, self._dDec = self._phase_centre
        else:
            logging.info('    Fitting Gaussian to determine geometry')
            inc, PA, dRA, dDec = _fit_geometry_gaussian(u, v, V, weights, guess=self._guess, inc_pa=self._inc_pa,
                                                        phase_centre=self._phase_centre)
            if not self._inc_pa:
                inc, PA = _fix_inc_and_PA_ranges(inc, PA)
            self._inc, self._PA, self._dRA, self._dDec = inc, PA, dRA, dDec


def _fit_geometry_gaussian(u, v, V, weights, guess, inc_pa=None, phase_centre=None):
    """Gaussian fit in uv-space by Levenberg-Marquardt (frank/geometry.py:497-626)."""
    fac = 2 * np.pi / rad_to_arcsec
    w = np.sqrt(weights)

    if inc_pa is not None:
        inc, PA = inc_pa
        inc *= deg_to_rad
        PA *= deg_to_rad

    if phase_centre is not None:
        dRA, dDec = phase_centre
        phi = dRA * fac * u + dDec * fac * v
        V = V * (np.cos(phi) - 1j * np.sin(phi))

    guess[0] *= deg_to_rad
    guess[1] *= deg_to_rad

    def wrap(fun):
        return np.concatenate([fun.real, fun.imag])

    def _gauss_fun(params):
        inc, PA, dRA, dDec, norm, scal = params
        if phase_centre is None:
            phi = dRA * fac * u + dDec * fac * v
            Vp = V * (np.cos(phi) - 1j * np.sin(phi))
        else:
            Vp = V
        c_t, s_t, c_i = np.cos(PA), np.sin(PA), np.cos(inc)
        up = (u * c_t - v * s_t) * c_i / (scal * rad_to_arcsec)
        vp = (u * s_t + v * c_t) / (scal * rad_to_arcsec)
        return wrap(w * (norm * np.exp(-0.5 * (up * up + vp * vp)) - Vp))

    def _gauss_jac(params):
        inc, PA, dRA, dDec, norm, scal = params
        jac = np.zeros([6, 2 * len(w)])
        if phase_centre is None:
            phi = dRA * fac * u + dDec * fac * v
            dVp = - w * V * (-np.sin(phi) - 1j * np.cos(phi)) * fac
            jac[2] = wrap(dVp * u)
            jac[3] = wrap(dVp * v)
        c_t, s_t, c_i, s_i = np.cos(PA), np.sin(PA), np.cos(inc), np.sin(inc)
        up = (u * c_t - v * s_t)
        vp = (u * s_t + v * c_t)
        uv = (up * up * c_i * c_i + vp * vp)
        G = w * np.exp(-0.5 * uv / (scal * rad_to_arcsec) ** 2)
        norm = norm / (scal * rad_to_arcsec) ** 2
        if inc_pa is None:
            jac[0] = wrap(norm * G * up * up * c_i * s_i)
            jac[1] = wrap(norm * G * up * vp * (c_i * c_i - 1) / 2)
        jac[4] = wrap(G)
        jac[5] = wrap(norm * G * uv / scal)
        return jac.T

    res = least_squares(_gauss_fun, guess, jac=_gauss_jac, method='lm')
    inc, PA, dRA, dDec, _, _ = res.x
    if inc_pa is not None:
        inc, PA = inc_pa
    else:
        inc /= deg_to_rad
        PA /= deg_to_rad
    if phase_centre is not None:
        dRA, dDec = phase_centre
    return inc, PA, dRA, dDec


class FitGeometryFourierBessel(SourceGeometry):
    """Determine the disc geometry by minimising the chi^2 of a non-parametric Fourier-Bessel fit
    (frank/geometry.py:623-763).

    Parameters as in the reference: Rmax (arcsec), N, inc_pa, phase_centre, guess, verbose.  Each residual
    evaluation builds a `FourierBesselFitter` for the trial geometry, fits (GPU mapping + N x N solve) and predicts
    the visibilities at the data's (u, v) (GPU); SciPy's Levenberg-Marquardt drives the search exactly as in the
    reference (finite-difference Jacobian, `method='lm'`)."""

    def __init__(self, Rmax, N, inc_pa=None, phase_centre=None, guess=None, verbose=False):
        super(FitGeometryFourierBessel, self).__init__()
        self._N = N
        self._R = Rmax
        self._inc_pa = inc_pa
        self._phase_centre = phase_centre
        if guess is None:
            guess = [10., 10., 0., 0.]
        if self._inc_pa is not None:
            guess[0], guess[1] = self._inc_pa
        if self._phase_centre is not None:
            guess[2], guess[3] = self._phase_centre
        self._guess = guess
        self._verbose = verbose

    def _residual(self, params, uvdata=None):
        from frank_b200.radial_fitters import FourierBesselFitter
        inc, pa, dRA, dDec = params
        if self._inc_pa is not None:
            inc, pa = self._inc_pa
        if self._phase_centre is not None:
            dRA, dDec = self._phase_centre
        geom = FixedGeometry(inc, pa, dRA, dDec)
        FBF = FourierBesselFitter(self._R, self._N, geom, verbose=False)
        u, v, vis, w_half = uvdata
        sol = FBF.fit(u, v, vis, w_half * w_half)
        error = w_half * (sol.predict(u, v) - vis)
        if self._verbose:
            Chi2 = 0.5 * np.sum(error.real ** 2 + error.imag ** 2) / len(w_half)
            print('\n      FitGeometryFourierBessel: Iteration {}, chi^2={:.8f}, inc={:.3f} PA={:.3f} dRA={:.5f} dDec={:.5f}'
                  ''.format(self._counter, Chi2, inc, pa, dRA, dDec), end='', flush=True)
            self._counter += 1
        return np.concatenate([error.real, error.imag])

    def fit(self, u, v, vis, w):
        if self._inc_pa and self._phase_centre:
            logging.info('    You requested a nonparametric fit to determine the geometry,'
                         ' but you provided values for inclination, PA, and the phase offset.'
                         ' --> Using your provided values (not fitting for the geometry)')
            self._inc, self._PA = self._inc_pa
            self._dRA, self._dDec = self._phase_centre
        else:
            logging.info('    Fitting nonparametric form to determine geometry')
            uvdata = [u, v, vis, w ** 0.5]
            self._counter = 0
            result = least_squares(self._residual, self._guess, kwargs={'uvdata': uvdata}, method='lm')
            if not result.success:
                raise RuntimeError("FitGeometryFourierBessel failed to converge")
            inc, pa, dRA, dDec = result.x
            if self._inc_pa:
                inc, pa = self._inc_pa
            if self._phase_centre:
                dRA, dDec = self._phase_centre
            if not self._inc_pa:
                inc, pa = _fix_inc_and_PA_ranges(inc, pa)
            self._inc, self._PA, self._dRA, self._dDec = inc, pa, dRA, dDec
            self._nfev = result.nfev
