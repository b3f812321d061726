"""Statistical models of the fit: visibility mapping (data -> normal equations) and the
Gaussian / log-normal posterior solves.

API mirror of frank.statistical_models (frank/statistical_models.py).  The arithmetic runs in
libfrankb200 (hand-written sm_100a CUDA, reached through ctypes); NumPy only carries the O(N^2)
results.  There is no CPU path: constructing these objects without a CUDA device raises.
"""
import logging
import os

import numpy as np

from frank_b200 import _lib
from frank_b200.constants import rad_to_arcsec, deg_to_rad

__all__ = ['VisibilityMapping', 'GaussianModel', 'LogNormalMAPModel']


# Dsolve through fb_chol_solve (device; N <= 512) rather than SciPy on the GPU-computed factor; FRANK_B200_DEVICE_DSOLVE=0
# selects the SciPy helper
_DEVICE_DSOLVE = os.environ.get('FRANK_B200_DEVICE_DSOLVE', '1') == '1'


def _is_cuda_tensor(x):
    return hasattr(x, 'data_ptr') and hasattr(x, 'is_cuda') and x.is_cuda


class VisibilityMapping(object):
    r"""Mapping between the visibility plane and the brightness at the DHT collocation points.

    Same constructor and results as frank.statistical_models.VisibilityMapping
    (frank/statistical_models.py:29-107).  ``block_data`` / ``block_size`` are accepted for
    compatibility; the GPU kernel tiles the visibility axis itself (64 visibilities per
    shared-memory tile), so they do not change the result beyond round-off.

    Parameters
    ----------
    DHT : DiscreteHankelTransform
    geometry : SourceGeometry
    vis_model : {'opt_thick', 'opt_thin', 'debris'}
    scale_height : callable H(R) in arcsec, required for 'debris'
    block_data, block_size : ignored (see above)
    check_qbounds : bool
        Raise when the data extend beyond the last collocation point.
    verbose : bool
    device : int, optional
        CUDA device ordinal (default: LOCAL_RANK, else 0).
    """

    def __init__(self, DHT, geometry, vis_model='opt_thick', scale_height=None, block_data=True,
                 block_size=10 ** 5, check_qbounds=True, verbose=True, device=None):
        models = ['opt_thick', 'opt_thin', 'debris']
        if vis_model not in models:
            raise ValueError(f"vis_model must be one of {models}")
        self._vis_model = vis_model
        self.check_qbounds = check_qbounds
        self._verbose = verbose
        self._chunking = block_data
        self._chunk_size = block_size
        self._DHT = DHT
        self._geometry = geometry
        self._device = device
        self._scale_height = None
        self._H2 = None
        if vis_model == 'debris':
            if scale_height is None:
                raise ValueError('You requested a model with a non-zero scale height'
                                 ' but did not specify H(R) (scale_height=None)')
            self._scale_height = scale_height(self.r)
            self._H2 = 0.5 * (2 * np.pi * self._scale_height / rad_to_arcsec) ** 2     # :101-102
        if verbose:
            logging.info('  Visibility model: %s', vis_model)
        self._timing = None

    # -- the hot path ----------------------------------------------------------------------------
    def _context(self):
        ctx = _lib.get_context(self._device)
        ctx.dht_setup(self._DHT)
        return ctx

    def _model_scale(self, geometry=None):
        if self._vis_model == 'opt_thick':
            g = self._geometry if geometry is None else geometry
            return float(np.cos(g.inc * deg_to_rad))                                   # :486-490
        return 1.0

    def _map_one(self, ctx, u, v, V, w, w_stride):
        """One channel through fb_map_visibilities_{host,dev}: returns M, j, H0, qmin, qmax."""
        import torch
        N = self.size
        n = int(u.shape[0])
        geom = self._geometry.device_scalars()
        q_last = float(self.q[-1])
        on_device = _is_cuda_tensor(u)
        if on_device:
            dev = u.device
            out = torch.empty(N * N + N + 1, dtype=torch.float64, device=dev)
            M, j, H0 = out[:N * N], out[N * N:N * N + N], out[N * N + N:]
            Vr = torch.view_as_real(V) if V.is_complex() else torch.stack([V, torch.zeros_like(V)], dim=-1)
            Vr = Vr.contiguous()
            rc, qmin, qmax = ctx.map_visibilities(n, u.contiguous(), v.contiguous(), Vr, w, w_stride, geom,
                                                  _lib.MODEL_CODE[self._vis_model], self._model_scale(), self._H2,
                                                  self.check_qbounds, q_last, M, j, H0, host=False)
            if rc == 0:
                host = out.cpu().numpy()
        else:
            host = np.empty(N * N + N + 1)
            M, j, H0 = host[:N * N], host[N * N:N * N + N], host[N * N + N:]
            rc, qmin, qmax = ctx.map_visibilities(n, u, v, V, w, w_stride, geom, _lib.MODEL_CODE[self._vis_model],
                                                  self._model_scale(), self._H2, self.check_qbounds, q_last,
                                                  M, j, H0, host=True)
        self._timing = ctx.last_map_timing()
        if rc == _lib.FB_E_QRANGE:
            self._raise_qrange(qmax)
        return host[:N * N].reshape(N, N).copy(), host[N * N:N * N + N].copy(), float(host[N * N + N]), qmin, qmax

    def map_visibilities(self, u, v, V, weights, frequencies=None, geometry=None):
        r"""Compute M = H^T w H, j = H^T w V and the null likelihood H0 from the visibilities
        (frank/statistical_models.py:109-237).

        u, v, V, weights may be NumPy arrays (host memory; copied to the GPU inside the call) or
        torch CUDA tensors (used in place).  Returns the reference's dict:
        ``mult_freq, channels, M, j, null_likelihood, hash``.

        As in the reference the deprojection always uses the geometry given at construction;
        a `geometry` argument only lands in the returned hash (statistical_models.py:158-165, 227).
        """
        if geometry is None:
            geometry = self._geometry
        if self._verbose:
            logging.info('    Building visibility matrices M and j')
        ctx = self._context()
        on_device = _is_cuda_tensor(u)
        if on_device:
            import torch
            w_stride = 1
            if not hasattr(weights, 'data_ptr'):
                weights = torch.full((1,), float(weights), dtype=torch.float64, device=u.device)
                w_stride = 0
            elif weights.numel() == 1:
                weights = weights.reshape(1).to(torch.float64)
                w_stride = 0
            V = V if V.is_complex() else V.to(torch.float64)
        else:
            u = np.ascontiguousarray(u, dtype=np.float64)
            v = np.ascontiguousarray(v, dtype=np.float64)
            V = np.ascontiguousarray(V, dtype=np.complex128)
            weights = np.asarray(weights, dtype=np.float64)
            w_stride = 1
            if weights.ndim == 0 or weights.size == 1:
                weights = weights.reshape(1).copy()
                w_stride = 0
            else:
                weights = np.ascontiguousarray(weights)

        multi_freq = frequencies is not None
        if not multi_freq:
            M, j, H0, qmin, qmax = self._map_one(ctx, u, v, V, weights, w_stride)
            self._warn_qmin(qmin)
            return {'mult_freq': False, 'channels': None, 'M': M, 'j': j, 'null_likelihood': H0,
                    'hash': [False, self._DHT, geometry, self._vis_model, self._scale_height]}

        # multi-frequency: one Gram per channel (statistical_models.py:180-214); H0 over all data
        if on_device:
            import torch
            channels = torch.unique(frequencies)
            sel = [frequencies == f for f in channels]
            channels = channels.cpu().numpy()
        else:
            frequencies = np.asarray(frequencies)
            channels = np.unique(frequencies)
            sel = [frequencies == f for f in channels]
        N = self.size
        Ms, js, H0 = np.zeros([len(channels), N, N]), np.zeros([len(channels), N]), 0.0
        qlo, qhi = np.inf, -np.inf
        for c, idx in enumerate(sel):
            wc = weights if w_stride == 0 else weights[idx]
            Ms[c], js[c], h0c, qmin, qmax = self._map_one(ctx, u[idx], v[idx], V[idx], wc, w_stride)
            H0 += h0c
            qlo, qhi = min(qlo, qmin), max(qhi, qmax)
        self._warn_qmin(qlo)
        return {'mult_freq': True, 'channels': channels, 'M': Ms, 'j': js, 'null_likelihood': H0,
                'hash': [True, self._DHT, geometry, self._vis_model, self._scale_height]}

    def _warn_qmin(self, qmin):
        if self.check_qbounds and self.q[0] < qmin:                                     # :519-525
            logging.warning(r"WARNING: First collocation point, q[0] = {:.3e} \lambda,"
                            " is at a baseline shorter than the"
                            " shortest deprojected baseline in the dataset,"
                            r" min(uv) = {:.3e} \lambda. For q[0] << min(uv),"
                            " the fit's total flux may be biased"
                            " low.".format(self.q[0], qmin))

    def _raise_qrange(self, qmax):
        raise ValueError(r"ERROR: Last collocation point, {:.3e} \lambda, is at"                # :526-535
                         " a shorter baseline than the longest deprojected"
                         r" baseline in the dataset, {:.3e} \lambda. Please"
                         " increase N (this is `hyperparameters: n` if you're using a parameter"
                         " file). Or if you'd like to fit to shorter maximum baseline,"
                         " cut the (u, v) distribution before fitting"
                         " (`modify_data: baseline_range` in the"
                         " parameter file).".format(self.q[-1], qmax))

    @property
    def last_timing(self):
        """CUDA-event timings (ms) of the most recent map_visibilities call."""
        return self._timing

    def check_hash(self, hash, multi_freq=False, geometry=None):
        """Compatibility test of mapped visibilities with this mapping (statistical_models.py:239-276)."""
        if geometry is None:
            geometry = self._geometry
        same = (multi_freq == hash[0]
                and all(getattr(self._DHT, k) == getattr(hash[1], k) for k in ('Rmax', 'size', 'order'))
                and all(getattr(geometry, k) == getattr(hash[2], k) for k in ('inc', 'PA', 'dRA', 'dDec'))
                and self._vis_model == hash[3])
        if not same:
            return False
        if self._scale_height is None:
            return hash[4] is None
        return False if hash[4] is None else bool(np.all(self._scale_height == hash[4]))

    def predict_visibilities(self, I, q, k=None, geometry=None):
        r"""Predicted (deprojected-plane) visibilities of the profile I at baselines q
        (frank/statistical_models.py:279-329); k is the vertical uv-distance, needed by the debris model."""
        q = np.asarray(q, dtype=np.float64)
        ctx = self._context()
        if self._vis_model == 'debris':
            if k is None:
                raise ValueError("the debris model needs the vertical uv-distance k")
            return ctx.predict_visibilities(q.reshape(-1), np.asarray(k).reshape(-1), I, 2, 1.0, self._H2).reshape(q.shape)
        return ctx.predict_visibilities(q.reshape(-1), None, I, _lib.MODEL_CODE[self._vis_model],
                                        self._model_scale(geometry), None).reshape(q.shape)

    def invert_visibilities(self, V, R, geometry=None):
        r"""Brightness at radii R / arcsec from visibilities at the collocation frequencies
        (frank/statistical_models.py:331-384); a handful of points, evaluated on the host."""
        R = np.atleast_1d(R)
        if self._vis_model == 'debris':
            scale = np.ones(self.size)
        else:
            scale = np.atleast_1d(self._model_scale(geometry))
        H = self._DHT.coefficients(R / rad_to_arcsec, direction='backward') * (1 / scale).reshape(1, -1)
        return np.dot(H, V)[R < self.Rmax]

    # -- small host-side helpers kept for API compatibility -----------------------------------
    def transform(self, f, q=None, direction='forward'):
        if direction == 'backward' and q is not None:
            q = q / rad_to_arcsec
        return self._DHT.transform(f, q, direction)

    def DHT_coefficients(self, direction='forward'):
        return self._DHT.coefficients(direction=direction)

    def interpolate(self, f, r, space='Real'):
        if space == 'Real':
            r = r / rad_to_arcsec
        r = np.array(r)
        return self._DHT.interpolate(f, r.reshape(-1), space).reshape(*r.shape)

    r = property(lambda self: self._DHT.r * rad_to_arcsec, doc="Radius points, arcsec")
    Rmax = property(lambda self: self._DHT.Rmax * rad_to_arcsec, doc="Maximum radius, arcsec")
    q = property(lambda self: self._DHT.q, doc="Frequency points, lambda")
    Qmax = property(lambda self: self._DHT.Qmax, doc="Maximum frequency, lambda")
    size = property(lambda self: self._DHT.size, doc="Number of points in reconstruction")

    @property
    def scale_height(self):
        return self._scale_height


class GaussianModel(object):
    r"""Posterior of the linear (Gaussian) brightness model: D^-1 = M + S(p)^-1, mu = D j.

    API mirror of frank.statistical_models.GaussianModel (frank/statistical_models.py:571-904) for one
    field (Nfields = 1).  The Cholesky factorisation and the solve run on the GPU
    (frank_b200/csrc/fb_solve.cu); the O(N^2) results are held as NumPy arrays so the object pickles like
    the reference's.

    Parameters
    ----------
    DHT : DiscreteHankelTransform
    M : array (N, N) or (Nf, N, N) ; j : array (N,) or (Nf, N)
    p : array (N,), optional power spectrum
    scale : optional per-channel scale factors (unit scale if None)
    guess : accepted for signature compatibility (the reference only uses it to infer Nfields)
    noise_likelihood : float
    """

    def __init__(self, DHT, M, j, p=None, scale=None, guess=None, Nfields=None, noise_likelihood=0, device=None,
                 _solution=None):
        self._DHT = DHT
        M = np.asarray(M)
        j = np.asarray(j)
        if M.ndim == 2:
            M = M.reshape(1, *M.shape)
        if j.ndim == 1:
            j = j.reshape(1, *j.shape)
        Nf, Nr = j.shape
        if Nfields is None:
            Nfields = 1 if guess is None else np.asarray(guess).reshape(-1, Nr).shape[0]
        if Nfields != 1:
            raise NotImplementedError("frank_b200 solves single-field models (Nfields = 1)")
        self._Nfields = 1
        if p is not None:
            p = np.asarray(p, dtype=np.float64).reshape(-1, Nr)
            if np.any(p <= 0) or np.any(np.isnan(p)):                                  # :688-698
                raise ValueError("Bad value in power spectrum. The power"
                                 " spectrum must be positive and not contain"
                                 " any NaN values. This is likely due to"
                                 " your UVtable (incorrect units or weights), "
                                 " or the deprojection being applied (incorrect"
                                 " geometry and/or phase center). Else you may"
                                 " want to adjust `rout` (ensure it is larger than"
                                 " the source) or `n` (up to ~300).")
        self._p = p
        s = np.ones([Nf, 1]) if scale is None else np.asarray(scale, dtype=np.float64).reshape(Nf, -1)
        # channel sum with scale factors (statistical_models.py:711-726)
        self._M = np.zeros([Nr, Nr])
        self._j = np.zeros(Nr)
        for si, Mi, ji in zip(s, M, j):
            self._j += si[0] * ji
            self._M += si[0] * si[0] * Mi
        self._like_noise = noise_likelihood
        self._device = device
        self._cov = None
        self._Sinv_cache = None
        if _solution is not None:
            self._mu, self._U = _solution
        else:
            self._fit()

    def _fit(self):
        ctx = _lib.get_context(self._device)
        ctx.dht_setup(self._DHT)
        mu, chol, info, rc = ctx.gaussian_fit(self._M, self._j, None if self._p is None else self._p[0])
        self._Dsvd = None
        if rc == _lib.FB_E_NOTPD:
            # SVD pseudo-inverse fallback (statistical_models.py:747-755): U, s, V of D^-1 from the device (one-sided
            # Jacobi, fb_gaussian_svd); s1 and mu formed as in the reference
            U, s, V, _ = ctx.gaussian_svd(self._M, None if self._p is None else self._p[0])
            s1 = np.where(s > 0, 1. / np.where(s > 0, s, 1.), 0)
            self._Dsvd = U, s1, V
            self._U = None
            self._mu = np.dot(V.T, np.multiply(np.dot(U.T, self._j), s1))
            return
        self._mu = mu[0]
        self._U = np.triu(chol[0])

    @property
    def _Sinv(self):
        if self._p is None:
            return None
        if self._Sinv_cache is None:
            Y = self._DHT.coefficients()
            self._Sinv_cache = np.dot(Y.T * (1 / self._p[0]), Y)
        return self._Sinv_cache

    def Dsolve(self, b):
        r"""D b through the GPU-computed Cholesky factor (post-fit helper, statistical_models.py:762-781)."""
        if getattr(self, '_Dsvd', None) is not None:                                   # :778-781
            U, s1, V = self._Dsvd
            b = np.asarray(b)
            return np.dot(V.T, (np.dot(U.T, b).T * s1).T)
        if _DEVICE_DSOLVE and self._DHT.size <= 512:
            ctx = _lib.get_context(self._device)
            ctx.dht_setup(self._DHT)
            return ctx.chol_solve(self._U, b)
        import scipy.linalg
        return scipy.linalg.cho_solve((self._U, False), b)

    def log_likelihood(self, I=None):
        r"""log P(V|p) (I is None) or log P(I, V|p), statistical_models.py:790-856."""
        if I is None:
            like = 0.5 * np.sum(self._j * self._mu)
            if self._p is not None:
                like += 0.5 * np.linalg.slogdet(self.Dsolve(self._Sinv))[1]
        else:
            Sinv = 0 if self._p is None else self._Sinv
            like = np.sum(self._j * I) - 0.5 * np.dot(I, np.dot(self._M + Sinv, I))
            if self._p is not None:
                like += 0.5 * np.linalg.slogdet(2 * np.pi * Sinv)[1]
        return like + self._like_noise

    def solve_non_negative(self):
        import scipy.optimize
        Sinv = 0 if self._p is None else self._Sinv
        return scipy.optimize.nnls(self._M + Sinv, self._j, maxiter=100 * len(self._j))[0]

    def draw(self, N):
        return np.random.multivariate_normal(self.mean.reshape(-1), self.covariance, N)

    mean = property(lambda self: self._mu, doc="Posterior mean, Jy / sr")
    MAP = property(lambda self: self._mu, doc="Posterior maximum, Jy / sr")
    s_0 = property(lambda self: 0)
    num_fields = property(lambda self: 1)
    size = property(lambda self: self._DHT.size)

    @property
    def covariance(self):
        if self._cov is None:
            self._cov = self.Dsolve(np.eye(self.size))
        return self._cov

    @property
    def power_spectrum(self):
        return None if self._p is None else self._p.reshape(self.size)


class LogNormalMAPModel(object):
    r"""Maximum a posteriori field of the log-normal model, P(s|V,p) ∝ G(H exp(s + s0) - V, M) G(s, S(p)).

    API mirror of frank.statistical_models.LogNormalMAPModel (frank/statistical_models.py:907-1295) for one
    channel, one field and unit scale (what FrankFitter uses).  The objective, gradient, Hessian, its
    factorisation and the Newton solves run on the GPU (fb_ln_* in include/frankb200.h); the Newton / line-search
    decisions are taken on the host (frank_b200/minimizer.py).
    """

    def __init__(self, DHT, M, j, p=None, scale=None, s0=None, guess=None, Nfields=None, full_hessian=1,
                 noise_likelihood=0, device=None):
        from frank_b200.minimizer import LineSearch, MinimizeNewton
        self._DHT = DHT
        self._full_hess = full_hessian
        M = np.asarray(M)
        j = np.asarray(j)
        if M.ndim == 3:
            if M.shape[0] != 1:
                raise NotImplementedError("frank_b200 solves single-channel log-normal models")
            M, j = M[0], j[0]
        Nr = j.shape[0]
        if (Nfields or 1) != 1 or scale is not None:
            raise NotImplementedError("frank_b200 solves single-field, unit-scale log-normal models")
        if s0 is None or guess is None or p is None:
            raise ValueError("LogNormalMAPModel needs p, s0 and an initial guess")
        p = np.asarray(p, dtype=np.float64).reshape(Nr)
        if np.any(p <= 0) or np.any(np.isnan(p)):                                     # :1053-1062
            raise ValueError("Bad value in power spectrum. The power"
                             " spectrum must be positive and not contain"
                             " any NaN values. This is likely due to"
                             " your UVtable (incorrect units or weights), "
                             " or the deprojection being applied (incorrect"
                             " geometry and/or phase center). Else you may"
                             " want to increase `rout` by 10-20% or `n` so"
                             " that it is large, >~300.")
        self._M, self._j, self._p = M, j, p
        self._s0 = float(np.atleast_1d(s0)[0])
        self._like_noise = noise_likelihood
        self._device = device
        self._cov = None

        ctx = _lib.get_context(device)
        ctx.dht_setup(DHT)
        ctx.ln_setup(M, j, self._s0, full_hessian)
        ctx.ln_set_spectrum(p)
        self._ctx = ctx

        def limit_step(dx, x):                                                        # :1136-1140
            return min(1.1 * np.min(np.abs(x / dx)), 1) * dx

        def newton_dir(x, refactor):
            g, dx, rc = ctx.ln_newton_direction(x, refactor)
            if rc == _lib.FB_E_NOTPD:
                # The reference factorises with LU and would still get a (not necessarily descending) direction;
                # an indefinite Hessian here makes the step fall back to gradient descent (minimizer.py:250-253).
                # Not met on any fit path tested (0 of 1.2e5 Hessians in the reference runs).
                return g, None
            return g, dx

        search = LineSearch(reduce_step=limit_step)
        x0 = np.array(guess, dtype=np.float64).reshape(Nr)
        s, self._status = MinimizeNewton(lambda x: ctx.ln_eval(x), lambda x: ctx.ln_eval(x, True)[1], newton_dir, x0,
                                         search, tol=1e-7)
        self._s_MAP = s
        chol, _, rc = ctx.ln_posterior(s, p)                                          # cho_factor(hess(s)), :1148-1150
        self._U = np.triu(chol)

    def _update_power_spectrum(self, alpha, p0, Tinv):
        """CriticalFilter.update_power_spectrum with this model's Hessian factor (device)."""
        self._ctx.ln_setup(self._M, self._j, self._s0, self._full_hess)
        self._ctx.ln_set_spectrum(self._p)
        _, p_new, _ = self._ctx.ln_posterior(self._s_MAP, self._p, alpha, p0, Tinv, want_chol=False)
        return p_new

    def Dsolve(self, b):
        import scipy.linalg
        return scipy.linalg.cho_solve((self._U, False), b)

    def log_likelihood(self, s=None):
        r"""log P(I, V|p) as the reference evaluates it (statistical_models.py:1196-1245)."""
        if s is None:
            s = self._s_MAP
        Y = self._DHT.coefficients()
        Sinv = np.dot(Y.T * (1 / self._p), Y)
        I = np.exp(s)
        like = -0.5 * np.dot(s - self._s0, np.dot(Sinv, s - self._s0))
        like -= 0.5 * np.dot(I, np.dot(self._M, I))
        like += np.sum(I * self._j)
        like += 0.5 * np.linalg.slogdet(2 * np.pi * Sinv)[1]
        return like + self._like_noise

    def solve_non_negative(self):
        return self.MAP

    def draw(self, N):
        return np.random.multivariate_normal(self.MAP.reshape(-1), self.covariance, N)

    MAP = property(lambda self: self._s_MAP, doc="Posterior maximum of s = log I - s0")
    power_spectrum = property(lambda self: self._p)
    scale = property(lambda self: np.ones(1))
    s_0 = property(lambda self: np.array([self._s0]))
    num_fields = property(lambda self: 1)
    size = property(lambda self: self._DHT.size)

    @property
    def covariance(self):
        if self._cov is None:
            self._cov = self.Dsolve(np.eye(self.size))
        return self._cov
