"""Source geometry: phase-centre shift, deprojection, and the geometry fitters.

API mirror of frank.geometry (frank/geometry.py:41-763).  On the fit path the per-visibility arithmetic runs on the
GPU (frank_b200/csrc/fb_prep.cu); the NumPy helpers here serve the small host-side uses (re-projection of a handful
of points, `undo_correction` of predicted visibilities) and define the host scalars handed to the device
(`device_scalars`).  `FitGeometryFourierBessel` is a pure caller of the hot path: every
residual evaluation of its Levenberg-Marquardt search is one GPU mapping + solve + GPU prediction on visibilities that
stay resident on the device.
"""
import logging

import numpy as np
from scipy.optimize import least_squares

from frank_b200.constants import rad_to_arcsec, deg_to_rad

__all__ = ['apply_phase_shift', 'deproject', 'rescale_total_flux', 'SourceGeometry', 'FixedGeometry',
           'FitGeometryFourierBessel']


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


class _TrialGeometryResidual(object):
    """Weighted residual of the non-parametric fit under a trial geometry, with the visibilities resident on the GPU.

    The data are uploaded once; an evaluation is a mapping call on the resident arrays (fb_map_visibilities_dev), an
    N x N solve, one fused prediction pass (fb_predict_sky_dev: deproject, H(q) I, re-project, phase rotation) and one
    elementwise residual; only the 2 n residual values travel back for the optimiser."""

    def __init__(self, Rmax, N, u, v, vis, weights, device=None):
        import torch
        from frank_b200 import _lib
        ctx = _lib.get_context(device)
        dev = torch.device('cuda', ctx.device)
        self._Rmax, self._N, self._device = Rmax, N, ctx.device
        self._u = torch.as_tensor(np.ascontiguousarray(u, dtype=np.float64)).to(dev)
        self._v = torch.as_tensor(np.ascontiguousarray(v, dtype=np.float64)).to(dev)
        self._vis = torch.as_tensor(np.ascontiguousarray(vis, dtype=np.complex128)).to(dev)
        w = np.ones(len(u)) * np.asarray(weights, dtype=np.float64)
        self._root_w = torch.as_tensor(w ** 0.5).to(dev)
        self._w_fit = self._root_w * self._root_w           # the reference fits with sqrt(w)**2 (geometry.py:690)
        self._host = torch.empty(2 * len(u), dtype=torch.float64).pin_memory()
        self.evaluations = 0

    def on_device(self, inc, PA, dRA, dDec):
        """The residual as one float64 CUDA tensor [2 n] (real parts, then imaginary parts)."""
        import torch
        from frank_b200.radial_fitters import FourierBesselFitter
        trial = FixedGeometry(inc, PA, dRA, dDec)
        fitter = FourierBesselFitter(self._Rmax, self._N, trial, verbose=False, device=self._device)
        model = fitter.fit(self._u, self._v, self._vis, self._w_fit).predict(self._u, self._v)
        miss = torch.view_as_real(self._root_w * (model - self._vis))
        self.evaluations += 1
        return torch.cat([miss[:, 0], miss[:, 1]])

    def __call__(self, inc, PA, dRA, dDec):
        import torch
        from frank_b200.radial_fitters import FourierBesselFitter
        trial = FixedGeometry(inc, PA, dRA, dDec)
        fitter = FourierBesselFitter(self._Rmax, self._N, trial, verbose=False, device=self._device)
        model = fitter.fit(self._u, self._v, self._vis, self._w_fit).predict(self._u, self._v)
        miss = self._root_w * (model - self._vis)
        n = miss.numel()
        self._host[:n].copy_(miss.real)
        self._host[n:].copy_(miss.imag)
        torch.cuda.current_stream(miss.device).synchronize()
        self.evaluations += 1
        return self._host.numpy().copy()


def _levenberg_marquardt_device(residual, x0, free, ctx, ftol=1e-8, xtol=1e-8, gtol=1e-8, max_nfev=None):
    """Levenberg-Marquardt with every 2n-vector on the device.

    What scipy.optimize.least_squares(method='lm') (MINPACK lmdif) does for the reference (frank/geometry.py:745-746), with the
    same ingredients -- forward-difference Jacobian with step sqrt(eps) |x_j| (sqrt(eps) when x_j = 0), column scaling by the
    largest column norm seen, tolerances ftol = xtol = gtol = 1e-8 -- but on the normal equations: an iteration needs only
    J^T J, J^T r and r.r, which one fixed-order device reduction over the residual and the Jacobian columns delivers
    (fb_columns_gram_dev) instead of a 2n x 4 host Jacobian and its QR factorisation.  The damping parameter follows Nielsen's
    gain-ratio rule.  `free` lists the indices of x that are optimised.  Returns x, number of residual evaluations, converged."""
    x = np.array(x0, dtype=np.float64)
    k = len(free)
    if max_nfev is None:
        max_nfev = 100 * (len(x) + 1) * 4
    eps = np.sqrt(np.finfo(np.float64).eps)
    nfev = 0

    def res(xv):
        return residual.on_device(*xv)

    r = res(x)
    nfev += 1
    scale = np.zeros(k)
    lam, nu = None, 2.0
    while nfev < max_nfev:
        cols = []
        for j in free:                                     # forward differences (MINPACK fdjac2)
            h = eps * abs(x[j]) if x[j] != 0 else eps
            xp = x.copy()
            xp[j] += h
            cols.append((res(xp) - r) / h)
            nfev += 1
        G = ctx.columns_gram_dev(cols + [r])               # [J | r]^T [J | r]
        A, g, f = G[:k, :k], G[:k, k], G[k, k]
        norms = np.sqrt(np.diag(A))
        if f == 0 or np.max(np.abs(g) / np.where(norms > 0, norms, 1.0)) / np.sqrt(f) <= gtol:
            return x, nfev, True                           # gradient orthogonal to the residual (MINPACK info = 4)
        scale = np.maximum(scale, np.where(norms > 0, norms, 1.0))
        if lam is None:
            lam = 1e-3
        while True:
            step = np.linalg.solve(A + lam * np.diag(scale * scale), -g)
            xn = x.copy()
            xn[list(free)] += step
            rn = res(xn)
            nfev += 1
            fn = float(ctx.columns_gram_dev([rn])[0, 0])
            predicted = -(2.0 * g @ step + step @ A @ step)
            gain = (f - fn) / predicted if predicted > 0 else -1.0
            small_step = np.linalg.norm(scale * step) <= xtol * np.linalg.norm(scale * x[list(free)])
            if gain > 1e-4:
                converged = small_step or ((f - fn) <= ftol * f and predicted <= ftol * f)
                x, r = xn, rn
                lam *= max(1.0 / 3.0, 1.0 - (2.0 * gain - 1.0) ** 3)
                nu = 2.0
                if converged:
                    return x, nfev, True
                break
            if small_step:
                return x, nfev, True
            lam *= nu
            nu *= 2.0
            if nfev >= max_nfev or not np.isfinite(lam):
                return x, nfev, False
    return x, nfev, False


class FitGeometryFourierBessel(SourceGeometry):
    """Determine the disc geometry by minimising the weighted chi^2 of a non-parametric Fourier-Bessel fit
    (frank/geometry.py:623-763).

    Parameters as in the reference: Rmax (arcsec), N, inc_pa, phase_centre, guess = [inc, PA, dRA, dDec], verbose;
    `device` selects the GPU.  Levenberg-Marquardt with a finite-difference Jacobian drives the search as in the reference
    (`solver='scipy'`: scipy.optimize.least_squares(method='lm') itself, on residual vectors brought back to the host;
    `solver='device'`: the same algorithm with the residual, the Jacobian columns and their reductions on the device;
    'auto' switches at 2e5 visibilities); every residual evaluation runs on GPU-resident visibilities
    (_TrialGeometryResidual).  Parameters fixed through `inc_pa` / `phase_centre` are held at the given values inside every
    evaluation and reported back unchanged.

    (frank's FitGeometryGaussian, a 6-parameter uv-plane Gaussian fit, is outside the hot path and is not provided: use
    frank's own and pass the result in as a FixedGeometry.)"""

    def __init__(self, Rmax, N, inc_pa=None, phase_centre=None, guess=None, verbose=False, device=None, solver='auto'):
        super(FitGeometryFourierBessel, self).__init__()
        if solver not in ('auto', 'scipy', 'device'):
            raise ValueError("solver must be 'auto', 'scipy' or 'device'")
        self._solver = solver
        self._Rmax_fit, self._N_fit = Rmax, N
        self._fixed_inc_pa = None if inc_pa is None else tuple(inc_pa)
        self._fixed_centre = None if phase_centre is None else tuple(phase_centre)
        start = [10., 10., 0., 0.] if guess is None else list(guess)
        if self._fixed_inc_pa is not None:
            start[0:2] = self._fixed_inc_pa
        if self._fixed_centre is not None:
            start[2:4] = self._fixed_centre
        self._start = start
        self._verbose = verbose
        self._device = device
        self._nfev = 0

    def _pin(self, params):
        """Trial parameters with the user-fixed ones substituted (geometry.py:679-683)."""
        inc, PA, dRA, dDec = params
        if self._fixed_inc_pa is not None:
            inc, PA = self._fixed_inc_pa
        if self._fixed_centre is not None:
            dRA, dDec = self._fixed_centre
        return inc, PA, dRA, dDec

    def fit(self, u, v, vis, w):
        if self._fixed_inc_pa and self._fixed_centre:
            logging.info('    You requested a nonparametric fit to determine the geometry,'
                         ' but you provided values for inclination, PA, and the phase offset.'
                         ' --> Using your provided values (not fitting for the geometry)')
            self._inc, self._PA = self._fixed_inc_pa
            self._dRA, self._dDec = self._fixed_centre
            return
        logging.info('    Fitting nonparametric form to determine geometry')
        residual = _TrialGeometryResidual(self._Rmax_fit, self._N_fit, u, v, vis, w, device=self._device)

        def objective(params):
            trial = self._pin(params)
            r = residual(*trial)
            if self._verbose:
                print('\n      FitGeometryFourierBessel: Iteration {}, chi^2={:.8f}, inc={:.3f} PA={:.3f} dRA={:.5f} dDec={:.5f}'
                      ''.format(residual.evaluations - 1, 0.5 * float(np.dot(r, r)) / (len(r) // 2), *trial), end='', flush=True)
            return r

        # 'scipy': the reference's driver on host residual vectors (what the small reference fixtures pin); 'device': the same
        # algorithm on device-resident vectors, for data sets whose 2n x 4 host Jacobian is the bottleneck ('auto': from 2e5
        # visibilities on)
        solver = self._solver if self._solver != 'auto' else ('device' if len(u) >= 200_000 else 'scipy')
        if solver == 'scipy':
            result = least_squares(objective, self._start, method='lm')
            if not result.success:
                raise RuntimeError("FitGeometryFourierBessel failed to converge")
            best, self._nfev = result.x, result.nfev
        else:
            from frank_b200 import _lib
            free = [j for j in range(4) if not ((j < 2 and self._fixed_inc_pa is not None) or (j >= 2 and self._fixed_centre is not None))]
            pinned_residual = residual
            if len(free) < 4:                          # fixed parameters are substituted inside every evaluation
                class _Pinned(object):
                    evaluations = property(lambda s_: residual.evaluations)

                    def on_device(s_, *params):
                        return residual.on_device(*self._pin(params))
                pinned_residual = _Pinned()
            best, nfev, ok = _levenberg_marquardt_device(pinned_residual, self._start, free, _lib.get_context(self._device))
            if not ok:
                raise RuntimeError("FitGeometryFourierBessel failed to converge")
            self._nfev = nfev
        inc, PA, dRA, dDec = self._pin(best)
        if self._fixed_inc_pa is None:
            inc, PA = _fix_inc_and_PA_ranges(inc, PA)
        self._inc, self._PA, self._dRA, self._dDec = inc, PA, dRA, dDec
