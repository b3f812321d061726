"""Radial-profile fitters and their result objects.

API mirror of frank.radial_fitters (frank/radial_fitters.py): FourierBesselFitter, FrankFitter,
FrankRadialFit, FrankGaussianFit, FrankLogNormalFit -- same constructor arguments, methods, properties,
exceptions and messages.  The data-dependent work (visibility mapping, Cholesky solves, the power-spectrum
fixed-point loop) runs in hand-written CUDA (libfrankb200); this module only sequences it.
"""
import abc
import logging
from collections import defaultdict

import numpy as np

from frank_b200 import _lib
from frank_b200.constants import rad_to_arcsec
from frank_b200.filter import CriticalFilter
from frank_b200.hankel import DiscreteHankelTransform
from frank_b200.statistical_models import GaussianModel, LogNormalMAPModel, VisibilityMapping

__all__ = ['FrankRadialFit', 'FrankGaussianFit', 'FrankLogNormalFit', 'FourierBesselFitter', 'FrankFitter']


class FrankRadialFit(metaclass=abc.ABCMeta):
    """Base class for results of frank fits (frank/radial_fitters.py:35-219)."""

    def __init__(self, vis_map, info, geometry):
        self._vis_map = vis_map
        self._geometry = geometry
        self._info = info

    def predict(self, u, v, I=None, geometry=None):
        r"""Predict the visibilities in the sky-plane (frank/radial_fitters.py:56-98)."""
        if geometry is None:
            geometry = self._geometry
        if I is None:
            I = self.I
        if geometry is not None and hasattr(u, 'is_cuda') and u.is_cuda:
            # device-resident baselines: deprojection, H(q) I and undo_correction fused in one pass (fb_predict_sky_dev)
            return self._vis_map.predict_sky(I, u, v, geometry)
        if geometry is not None:
            u, v, wz = geometry.deproject(u, v, use3D=True)
        else:
            wz = np.zeros_like(u)
        q = np.hypot(u, v)
        V = self._vis_map.predict_visibilities(I, q, wz, geometry=geometry)
        if geometry is not None:
            _, _, V = geometry.undo_correction(u, v, V)
        return V

    def predict_deprojected(self, q=None, I=None, geometry=None, block_size=10 ** 5, assume_optically_thick=True):
        r"""Predict the visibilities in the deprojected plane (frank/radial_fitters.py:100-144)."""
        if geometry is None:
            geometry = self._geometry
        if I is None:
            I = self.I
        if q is None:
            q = self.q
        return self._vis_map.predict_visibilities(I, q, q * 0, geometry=geometry)

    def interpolate_brightness(self, Rpts, I=None):
        r"""Fourier-Bessel interpolation of the profile to Rpts / arcsec (frank/radial_fitters.py:146-176)."""
        if I is None:
            I = self.I
        return self._vis_map.interpolate(I, np.array(Rpts), space='Real')

    @abc.abstractproperty
    def MAP(self):
        pass

    I = property(lambda self: self.MAP)
    r = property(lambda self: self._vis_map.r, doc="Radius points, arcsec")
    Rmax = property(lambda self: self._vis_map.Rmax, doc="Maximum radius, arcsec")
    q = property(lambda self: self._vis_map.q, doc="Frequency points, lambda")
    Qmax = property(lambda self: self._vis_map.Qmax, doc="Maximum frequency, lambda")
    size = property(lambda self: self._vis_map.size, doc="Number of points in reconstruction")
    geometry = property(lambda self: self._geometry, doc="SourceGeometry object")
    info = property(lambda self: self._info, doc="Fit quantities for reference")


class FrankGaussianFit(FrankRadialFit):
    """Result of a fit with a Gaussian brightness model (frank/radial_fitters.py:222-323)."""

    def __init__(self, DHT, fit, info={}, geometry=None):
        FrankRadialFit.__init__(self, DHT, info, geometry)
        self._fit = fit

    def draw(self, N):
        return np.random.multivariate_normal(self.mean, self.covariance, N)

    def log_likelihood(self, I=None):
        return self._fit.log_likelihood(I)

    def solve_non_negative(self):
        return self._fit.solve_non_negative()

    mean = property(lambda self: self._fit.mean, doc="Posterior mean, Jy / sr")
    MAP = property(lambda self: self._fit.mean, doc="Posterior maximum, Jy / sr")
    covariance = property(lambda self: self._fit.covariance, doc="Posterior covariance, (Jy / sr)**2")
    power_spectrum = property(lambda self: self._fit.power_spectrum, doc="Power spectrum coefficients")


class FrankLogNormalFit(FrankRadialFit):
    """Result of a fit with a log-normal brightness model (frank/radial_fitters.py:326-402)."""

    def __init__(self, DHT, fit, info={}, geometry=None):
        FrankRadialFit.__init__(self, DHT, info, geometry)
        self._fit = fit

    def log_likelihood(self, I=None):
        return self._fit.log_likelihood(None if I is None else np.log(I))

    @property
    def MAP(self):
        return np.exp((self._fit.MAP + self._fit.s_0) * self._fit.scale)

    covariance = property(lambda self: self._fit.covariance)
    power_spectrum = property(lambda self: self._fit.power_spectrum)


class FourierBesselFitter(object):
    """Fourier-Bessel series model for fitting visibilities, no prior (frank/radial_fitters.py:405-613).

    Parameters: Rmax (arcsec), N, geometry, nu=0, block_data=True, assume_optically_thick=True,
    scale_height=None, block_size=10**5, verbose=True  -- as in the reference; `device` selects the GPU."""

    def __init__(self, Rmax, N, geometry, nu=0, block_data=True, assume_optically_thick=True, scale_height=None,
                 block_size=10 ** 5, verbose=True, device=None):
        Rmax /= rad_to_arcsec
        self._geometry = geometry
        self._device = device
        self._DHT = DiscreteHankelTransform(Rmax, N, nu)
        if assume_optically_thick:
            if scale_height is not None:
                raise ValueError("Optically thick models must have zero scale-height")
            model = 'opt_thick'
        elif scale_height is not None:
            model = 'debris'
        else:
            model = 'opt_thin'
        self._vis_map = VisibilityMapping(self._DHT, geometry, model, scale_height=scale_height, block_data=block_data,
                                          block_size=block_size, check_qbounds=False, verbose=verbose, device=device)
        self._info = {'Rmax': self._DHT.Rmax * rad_to_arcsec, 'N': self._DHT.size}
        self._verbose = verbose

    def preprocess_visibilities(self, u, v, V, weights=1):
        r"""Map the visibilities onto the normal equations once (frank/radial_fitters.py:468-498)."""
        return self._vis_map.map_visibilities(u, v, V, weights)

    def _build_matrices(self, mapping):
        self._vis_map.check_hash(mapping['hash'])            # result unused, as in the reference (:509)
        self._M = mapping['M']
        self._j = mapping['j']
        self._H0 = mapping['null_likelihood']

    def fit_method(self):
        return type(self).__name__

    def fit_preprocessed(self, preproc_vis):
        if self._verbose:
            logging.info('  Fitting pre-processed visibilities for brightness'
                         ' profile using {}'.format(self.fit_method()))
        self._build_matrices(preproc_vis)
        return self._fit()

    def fit(self, u, v, V, weights=1):
        r"""Fit the visibilities (frank/radial_fitters.py:544-572)."""
        if self._verbose:
            logging.info('  Fitting for brightness profile using {}'.format(self.fit_method()))
        self._geometry.fit(u, v, V, weights)
        mapping = self.preprocess_visibilities(u, v, V, weights)
        self._build_matrices(mapping)
        return self._fit()

    def _fit(self):
        fit = GaussianModel(self._DHT, self._M, self._j, noise_likelihood=self._H0, device=self._device)
        self._sol = FrankGaussianFit(self._vis_map, fit, self._info, geometry=self._geometry.clone())
        return self._sol

    r = property(lambda self: self._DHT.r * rad_to_arcsec, doc="Radius points, arcsec")
    Rmax = property(lambda self: self._DHT.Rmax * rad_to_arcsec, doc="Maximum radius, arcsec")
    q = property(lambda self: self._DHT.q, doc="Frequency points, lambda")
    Qmax = property(lambda self: self._DHT.Qmax, doc="Maximum frequency, lambda")
    size = property(lambda self: self._DHT.size, doc="Number of points in reconstruction")
    geometry = property(lambda self: self._geometry, doc="Geometry object")


class FrankFitter(FourierBesselFitter):
    """Gaussian-process fit with the DHT of Baddour & Chouinard (2015) and the critical-filter power-spectrum
    prior of Oppermann et al. (2013) (frank/radial_fitters.py:616-991).

    Same parameters and defaults as the reference:
    Rmax, N, geometry, nu=0, block_data=True, block_size=10**5, alpha=1.05, p_0=None, weights_smooth=1e-4,
    tol=1e-3, method='Normal', I_scale=1e5, max_iter=2000, check_qbounds=True,
    store_iteration_diagnostics=False, assume_optically_thick=True, scale_height=None, verbose=True,
    convergence_failure='raise'.
    """

    def __init__(self, Rmax, N, geometry, nu=0, block_data=True, block_size=10 ** 5, alpha=1.05, p_0=None,
                 weights_smooth=1e-4, tol=1e-3, method='Normal', I_scale=1e5, max_iter=2000, check_qbounds=True,
                 store_iteration_diagnostics=False, assume_optically_thick=True, scale_height=None, verbose=True,
                 convergence_failure='raise', device=None):
        if method not in {'Normal', 'LogNormal'}:
            raise ValueError('FrankFitter supports following mehods:\n\t{ "Normal", "LogNormal"}"')
        self._method = method
        super(FrankFitter, self).__init__(Rmax, N, geometry, nu, block_data, assume_optically_thick, scale_height,
                                          block_size, verbose, device=device)
        self._vis_map.check_qbounds = check_qbounds          # FourierBesselFitter does not check bounds (:707)
        if p_0 is None:
            p_0 = 1e-15 if method == 'Normal' else 1e-35
        self._s_scale = np.log(I_scale)
        self._tol = tol
        self._filter = CriticalFilter(self._DHT, alpha, p_0, weights_smooth, tol)
        self._max_iter = max_iter
        self._store_iteration_diagnostics = store_iteration_diagnostics
        self._info.update({'alpha': alpha, 'wsmooth': weights_smooth, 'p0': p_0, 'method': method})
        if convergence_failure not in {'raise', 'warn', 'ignore'}:
            raise ValueError("convergence_failure must be one of 'raise',"
                             f"'warn', or 'ignore', nor {convergence_failure}")
        self._convergence_failure = convergence_failure

    def fit_method(self):
        return '{}: {} method'.format(type(self).__name__, self._method)

    def _starting_spectrum(self):
        """The two warm-up Normal fits that seed the iteration (frank/radial_fitters.py:744-752)."""
        pI = np.ones([self.size])
        fit = self._perform_fit(pI, guess=np.ones_like(pI), fit_method='Normal')
        pI = np.max(self._DHT.transform(fit.MAP) ** 2)
        pI = pI * (self.q / self.q[0]) ** -2
        return pI

    def _fit(self):
        """Power-spectrum fixed-point iteration (frank/radial_fitters.py:737-832), device resident."""
        if self._store_iteration_diagnostics:
            self._iteration_diagnostics = defaultdict(list)
        if self._method == 'LogNormal':
            return self._fit_lognormal()
        pI = self._starting_spectrum()
        ctx = _lib.get_context(self._device)
        ctx.dht_setup(self._DHT)
        hist_cap = self._max_iter + 2 if self._store_iteration_diagnostics else 0
        out = ctx.frank_normal_loop(self._M, self._j, pI, self._filter._alpha, self._filter._p_0, self._filter._Tinv,
                                    self._tol, self._max_iter, want_chol=True, hist_cap=hist_cap)
        if out['status'] == _lib.FB_E_NOTPD:
            # a Cholesky pivot failed inside the device loop.  The reference's GaussianModel would have switched to its SVD
            # pseudo-inverse for that solve and carried on (statistical_models.py:747-755): redo the iteration sequenced
            # from the host, one GaussianModel (device Cholesky, device SVD on failure) per step
            return self._fit_sequenced(pI)
        count = int(out['niter'][0])
        pI = out['p'][0]
        fit = GaussianModel(self._DHT, self._M, self._j, pI, noise_likelihood=self._H0, device=self._device,
                            _solution=(out['mu'][0], np.triu(out['chol'][0])))
        if self._store_iteration_diagnostics:
            self._iteration_diagnostics['power_spectrum'] = [out['hist_p'][0, i].copy() for i in range(count)]
            self._iteration_diagnostics['MAP'] = [out['hist_mu'][0, i].copy() for i in range(count)]
        self._report_convergence(count)
        if self._store_iteration_diagnostics:
            self._iteration_diagnostics['num_iterations'] = count
        self._sol = FrankGaussianFit(self._vis_map, fit, self._info, geometry=self._geometry.clone())
        self._ps = pI
        self._ps_cov = None
        return self._sol

    def _sequenced_iteration(self, pI, filt, store=False):
        """The Normal iteration of frank/radial_fitters.py:765-785 sequenced from the host, one GaussianModel (device Cholesky,
        device SVD on failure, statistical_models.py:747-755) per step: the path taken when a factorisation failed inside the
        device-resident loop, so that the reference's SVD fallback of individual solves is honoured.  Returns fit, p, count."""
        fit = self._perform_fit(pI, fit_method='Normal')
        count, pi_old = 0, 0
        while (not filt.check_convergence(pI, pi_old)) and count <= self._max_iter:
            pi_old = pI.copy()
            pI = filt.update_power_spectrum(fit, device=self._device)
            fit = self._perform_fit(pI, guess=fit.MAP, fit_method='Normal')
            if store:
                self._iteration_diagnostics['power_spectrum'].append(pI)
                self._iteration_diagnostics['MAP'].append(fit.MAP)
            count += 1
        return fit, pI, count

    def _fit_sequenced(self, pI):
        fit, pI, count = self._sequenced_iteration(pI, self._filter, store=self._store_iteration_diagnostics)
        self._report_convergence(count)
        if self._store_iteration_diagnostics:
            self._iteration_diagnostics['num_iterations'] = count
        self._sol = FrankGaussianFit(self._vis_map, fit, self._info, geometry=self._geometry.clone())
        self._ps = pI
        self._ps_cov = None
        return self._sol

    # -- hyper-parameter sweeps (BASELINE config 4) -------------------------------------------------------------------
    def fit_sweep(self, u, v, V, weights=1, alphas=(1.05,), weights_smooths=(1e-4,), group=None, on_cholesky_failure='svd'):
        r"""Fit the same visibilities for every (alpha, w_smooth) pair of a grid.

        The reference's `run_multiple_fits` (frank/fit.py:493-563) loops `perform_fit` over the pairs and re-maps the
        visibilities every time although M and j do not depend on the hyper-parameters.  Here the data are mapped once
        and all grid points run as ONE batched device loop (`fb_frank_normal_loop` with B = number of points); with
        torch.distributed initialised the points are sharded over the ranks and the results all-gathered
        (frank_b200.distributed.sweep_sharded), so every rank returns the full grid.

        Returns a list of FrankGaussianFit in the reference's loop order (alpha outer, w_smooth inner);
        `self.sweep_diagnostics` holds alpha, wsmooth, num_iterations and converged per point.

        A grid point whose Cholesky factorisation fails inside the batched loop (alpha -> 1 with little smoothing) is, by
        default, redone sequenced from the host with the reference's SVD fallback (`on_cholesky_failure='svd'`: thousands of
        host-sequenced iterations, seconds per point); `'flag'` leaves it marked as not converged instead."""
        self._geometry.fit(u, v, V, weights)
        mapping = self.preprocess_visibilities(u, v, V, weights)
        return self.fit_sweep_preprocessed(mapping, alphas, weights_smooths, group=group, on_cholesky_failure=on_cholesky_failure)

    def fit_sweep_preprocessed(self, preproc_vis, alphas=(1.05,), weights_smooths=(1e-4,), group=None, on_cholesky_failure='svd'):
        if on_cholesky_failure not in ('svd', 'flag'):
            raise ValueError("on_cholesky_failure must be 'svd' or 'flag'")
        if self._method != 'Normal':
            raise ValueError("fit_sweep batches the Normal method; loop FrankFitter(method='LogNormal') over the grid instead")
        from frank_b200 import distributed
        self._build_matrices(preproc_vis)
        pI = self._starting_spectrum()                        # the warm-up fits do not depend on (alpha, w_smooth)
        self._sweep_fallbacks = []
        grid = [(float(a), float(w)) for a in np.atleast_1d(alphas) for w in np.atleast_1d(weights_smooths)]
        ctx = _lib.get_context(self._device)
        ctx.dht_setup(self._DHT)
        N = self.size

        def solve_points(idx):
            filters = [CriticalFilter(self._DHT, grid[i][0], self._filter._p_0, grid[i][1], self._tol) for i in idx]
            out = ctx.frank_normal_loop(self._M, self._j, np.tile(pI, (len(idx), 1)), [f._alpha for f in filters],
                                        [f._p_0 for f in filters], np.stack([f._Tinv for f in filters]), self._tol,
                                        self._max_iter, want_chol=False)
            res = {'p': out['p'], 'mu': out['mu'], 'niter': out['niter'], 'converged': out['converged']}
            for k in np.nonzero(out['info'])[0]:
                if on_cholesky_failure == 'flag':
                    res['converged'][k] = 0
                    self._sweep_fallbacks.append(int(idx[k]))
                    continue
                # a Cholesky pivot failed for this point: redo it sequenced from the host, where a failed factorisation
                # falls back to the SVD pseudo-inverse like the reference's GaussianModel (statistical_models.py:747-755)
                fit, pk, count = self._sequenced_iteration(pI.copy(), filters[k])
                res['p'][k], res['mu'][k], res['niter'][k] = pk, fit.MAP, count
                res['converged'][k] = int(count <= self._max_iter)
                self._sweep_fallbacks.append(int(idx[k]))
            return res

        res = distributed.sweep_sharded(solve_points, len(grid), N, group=group, ctx=ctx)
        self.sweep_diagnostics = {'alpha': [g[0] for g in grid], 'wsmooth': [g[1] for g in grid],
                                  'num_iterations': res['niter'].tolist(), 'converged': res['converged'].tolist(),
                                  'cholesky_failed_points_this_rank': list(self._sweep_fallbacks)}
        sols = []
        for k, (a, w) in enumerate(grid):
            fit = GaussianModel(self._DHT, self._M, self._j, res['p'][k], noise_likelihood=self._H0, device=self._device,
                                _solution=(res['mu'][k], None))      # the factor is recomputed on demand (Dsolve, covariance)
            info = dict(self._info, alpha=a, wsmooth=w)
            sols.append(FrankGaussianFit(self._vis_map, fit, info, geometry=self._geometry.clone()))
        return sols

    def _fit_lognormal(self):
        """LogNormal branch of frank/radial_fitters.py:754-785: log-space start, then the fixed-point loop with
        LogNormalMAPModel solves, device resident (fb_frank_lognormal_loop)."""
        pI = self._starting_spectrum()
        fit = self._perform_fit(pI, fit_method='Normal')
        s = np.log(np.maximum(fit.MAP, 1e-3 * fit.MAP.max()))
        s -= self._s_scale
        pI = np.max(self._DHT.transform(s) ** 2)
        pI = pI * (self.q / self.q[0]) ** -4
        # first log-normal fit and the whole fixed-point iteration in one device-resident call
        ctx = _lib.get_context(self._device)
        ctx.dht_setup(self._DHT)
        hist_cap = self._max_iter + 2 if self._store_iteration_diagnostics else 0
        out = ctx.frank_lognormal_loop(self._M, self._j, pI, s, self._s_scale, alpha=self._filter._alpha, p0=self._filter._p_0,
                                       Tinv=self._filter._Tinv, tol=self._tol, max_iter=self._max_iter, hist_cap=hist_cap)
        LogNormalMAPModel._raise_for_status(out['status'])
        count, pI = out['niter'], out['p']
        fit = LogNormalMAPModel(self._DHT, self._M, self._j, pI, guess=out['s'], s0=self._s_scale, noise_likelihood=self._H0,
                                device=self._device, _solution=(out['s'], np.triu(out['chol']), out['newton']))
        if self._store_iteration_diagnostics:
            self._iteration_diagnostics['power_spectrum'] = [out['hist_p'][i].copy() for i in range(count)]
            self._iteration_diagnostics['MAP'] = [out['hist_s'][i].copy() for i in range(count)]
        self._report_convergence(count)
        if self._store_iteration_diagnostics:
            self._iteration_diagnostics['num_iterations'] = count
        self._sol = FrankLogNormalFit(self._vis_map, fit, self._info, geometry=self._geometry.clone())
        self._ps = pI
        self._ps_cov = None
        return self._sol

    def _report_convergence(self, count):
        """Convergence policy of frank/radial_fitters.py:787-815."""
        show = self._verbose and logging.getLogger().isEnabledFor(logging.INFO)
        if count < self._max_iter:
            if show:
                logging.info('    Convergence criterion met at iteration {}'.format(count - 1))
            return
        if show:
            logging.info('    Convergence criterion not met; fit stopped at'
                         ' max_iter specified in your parameter file,'
                         ' {}'.format(self._max_iter))
        msg = f'Convergence not met within {self._max_iter} '
        msg += 'iterations.\nTry increasing max_iter, or '
        msg += 'try increasing alpha since convergence can '
        msg += 'be very slow for alpha close to 1.'
        if self._convergence_failure == 'raise':
            msg += '\nAlternatively set convergence_failure to'
            msg += "'warn' or 'ignore' to continue despite the"
            msg += 'failure.'
            raise RuntimeError(msg)
        if self._convergence_failure == 'warn':
            if logging.getLogger().isEnabledFor(logging.INFO):
                logging.info(msg)
            else:
                print(msg)

    def _perform_fit(self, p, guess=None, fit_method=None):
        """Posterior for a given power spectrum (frank/radial_fitters.py:858-890)."""
        if fit_method is None:
            fit_method = self._method
        if fit_method == 'Normal':
            return GaussianModel(self._DHT, self._M, self._j, p, guess=guess, noise_likelihood=self._H0,
                                 device=self._device)
        if fit_method == 'LogNormal':
            return LogNormalMAPModel(self._DHT, self._M, self._j, p, guess=guess, s0=self._s_scale,
                                     noise_likelihood=self._H0, device=self._device)
        raise ValueError('fit_method must be one of the following:\n\t{"Normal", "LogNormal"}')

    def draw_powerspectrum(self, Ndraw=1):
        log_p = np.random.multivariate_normal(np.log(self._ps), self.MAP_spectrum_covariance, Ndraw)
        return np.exp(log_p)

    def log_prior(self, p=None):
        return self._filter.log_prior(self._ps if p is None else p)

    def log_likelihood(self, sol=None):
        if sol is None:
            sol = self.MAP_solution
        return self.log_prior(sol.power_spectrum) + sol.log_likelihood()

    def log_evidence_laplace(self):
        r"""Laplace-approximated evidence of the best-fit model (frank/radial_fitters.py:951-967)."""
        Sigma_inv = self._filter.covariance_MAP(self._sol, ret_inv=True)
        sign, logdet = np.linalg.slogdet(Sigma_inv / (2 * np.pi))
        return self.log_likelihood() - 0.5 * logdet

    MAP_solution = property(lambda self: self._sol, doc="Reconstruction for the maximum a posteriori power spectrum")
    MAP_spectrum = property(lambda self: self._ps, doc="Maximum a posteriori power spectrum")
    iteration_diagnostics = property(lambda self: self._iteration_diagnostics)

    @property
    def MAP_spectrum_covariance(self):
        if self._ps_cov is None:
            self._ps_cov = self._filter.covariance_MAP(self._sol)
        return self._ps_cov
