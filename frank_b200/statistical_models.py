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


# Dsolve through fb_chol_solve (device) rather than SciPy on the GPU-computed factor; FRANK_B200_DEVICE_DSOLVE=0 selects the
# SciPy helper (diagnostics)
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

    def _map(self, ctx, u, v, V, w, w_stride, chan=None, nchan=1):
        """All channels through one fb_map_visibilities_{host,dev} call: returns M [nchan, N, N], j [nchan, N], H0,
        qmin, qmax."""
        import torch
        N = self.size
        n = int(u.shape[0])
        geom = self._geometry.device_scalars()
        q_last = float(self.q[-1])
        nM, nj = nchan * N * N, nchan * N
        if _is_cuda_tensor(u):
            out = torch.empty(nM + nj + 1, dtype=torch.float64, device=u.device)
            M, j, H0 = out[:nM], out[nM:nM + nj], out[nM + nj:]
            # the library runs on its own stream: everything the caller (or the conversions above) enqueued on torch's
            # current stream must have finished before the kernels read it
            torch.cuda.current_stream(u.device).synchronize()
            rc, qmin, qmax = ctx.map_visibilities(n, u, v, V, w, w_stride, geom, _lib.MODEL_CODE[self._vis_model],
                                                  self._model_scale(), self._H2, self.check_qbounds, q_last, M, j, H0,
                                                  host=False, chan=chan, nchan=nchan)
            host = out.cpu().numpy() if rc == 0 else None
        else:
            host = np.empty(nM + nj + 1)
            M, j, H0 = host[:nM], host[nM:nM + nj], host[nM + nj:]
            rc, qmin, qmax = ctx.map_visibilities(n, u, v, V, w, w_stride, geom, _lib.MODEL_CODE[self._vis_model],
                                                  self._model_scale(), self._H2, self.check_qbounds, q_last,
                                                  M, j, H0, host=True, chan=chan, nchan=nchan)
        self._timing = ctx.last_map_timing()
        if rc == _lib.FB_E_QRANGE:
            self._raise_qrange(qmax)
        return (host[:nM].reshape(nchan, N, N).copy(), host[nM:nM + nj].reshape(nchan, N).copy(), float(host[nM + nj]),
                qmin, qmax)

    @staticmethod
    def _device_inputs(u, v, V, weights):
        """Validate / convert CUDA-tensor inputs: contiguous float64 u, v, weights and complex128 V on one device.
        Returns u, v, V as a real [n, 2] view, weights, w_stride."""
        import torch
        dev = u.device
        n = u.numel()

        def f64(t, name):
            if not (hasattr(t, 'is_cuda') and t.is_cuda and t.device == dev):
                raise ValueError(f"map_visibilities: {name} must be a CUDA tensor on {dev} like u")
            if t.numel() != n:
                raise ValueError(f"map_visibilities: {name} has {t.numel()} elements, u has {n}")
            return t.reshape(-1).to(torch.float64).contiguous()
        u, v = f64(u, 'u'), f64(v, 'v')
        if not (hasattr(V, 'is_cuda') and V.is_cuda and V.device == dev and V.numel() == n):
            raise ValueError("map_visibilities: V must be a CUDA tensor of the same length and device as u")
        V = V.reshape(-1)
        if V.is_complex():
            Vr = torch.view_as_real(V.to(torch.complex128).contiguous())
        else:
            Vr = torch.stack([V.to(torch.float64), torch.zeros_like(V, dtype=torch.float64)], dim=-1).contiguous()
        w_stride = 1
        if not hasattr(weights, 'data_ptr'):
            weights = torch.full((1,), float(weights), dtype=torch.float64, device=dev)
            w_stride = 0
        elif weights.numel() == 1:
            weights = weights.reshape(1).to(device=dev, dtype=torch.float64).contiguous()
            w_stride = 0
        else:
            weights = f64(weights, 'weights')
        return u, v, Vr, weights, w_stride

    def map_visibilities(self, u, v, V, weights, frequencies=None, geometry=None, channels=None):
        r"""Compute M = H^T w H, j = H^T w V and the null likelihood H0 from the visibilities
        (frank/statistical_models.py:109-237).

        u, v, V, weights may be NumPy arrays (host memory, pinned or pageable; streamed to the GPU inside the call) or
        torch CUDA tensors (used in place).  Returns the reference's dict:
        ``mult_freq, channels, M, j, null_likelihood, hash``.

        With `frequencies` the Gram matrices of all channels come out of ONE device call (the channel index is the
        high part of the kernel's sort key, statistical_models.py:175-214); nothing is split on the host.
        `channels` (not in the reference; sorted, unique) fixes the channel list instead of np.unique(frequencies): a rank
        of a sharded call must use the GLOBAL list even if its slice misses a channel (that channel's M, j are zero).

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
            u, v, V, weights, w_stride = self._device_inputs(u, v, V, weights)
        else:
            u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
            v = np.ascontiguousarray(v, dtype=np.float64).reshape(-1)
            V = np.ascontiguousarray(V, dtype=np.complex128).reshape(-1)
            if v.size != u.size or V.size != u.size:
                raise ValueError("map_visibilities: u, v and V must have the same length")
            weights = np.asarray(weights, dtype=np.float64)
            w_stride = 1
            if weights.ndim == 0 or weights.size == 1:
                weights = weights.reshape(1).copy()
                w_stride = 0
            else:
                weights = np.ascontiguousarray(weights).reshape(-1)
                if weights.size != u.size:
                    raise ValueError("map_visibilities: weights must be a scalar or have the length of u")

        multi_freq = frequencies is not None
        if not multi_freq:
            M, j, H0, qmin, qmax = self._map(ctx, u, v, V, weights, w_stride)
            self._warn_qmin(qmin)
            return {'mult_freq': False, 'channels': None, 'M': M[0], 'j': j[0], 'null_likelihood': H0,
                    'hash': [False, self._DHT, geometry, self._vis_model, self._scale_height]}

        # multi-frequency: channel index = position in np.unique(frequencies) (statistical_models.py:180-189)
        if channels is not None:
            channels = np.asarray(channels, dtype=np.float64).reshape(-1)
            if channels.size == 0 or np.any(np.diff(channels) <= 0):
                raise ValueError("map_visibilities: channels must be sorted and unique")
        if on_device:
            import torch
            f = frequencies.reshape(-1)
            if channels is None:
                channels, chan = torch.unique(f, return_inverse=True)
                channels = channels.cpu().numpy()
            else:
                ch = torch.as_tensor(channels, dtype=f.dtype, device=f.device)
                chan = torch.searchsorted(ch, f.contiguous()).clamp_(max=len(channels) - 1)
                if f.numel() and not bool(torch.all(ch[chan] == f)):
                    raise ValueError("map_visibilities: a frequency is not in `channels`")
            chan = chan.to(torch.int32).contiguous()
        else:
            f = np.asarray(frequencies).reshape(-1)
            if channels is None:
                channels, chan = np.unique(f, return_inverse=True)
            else:
                chan = np.minimum(np.searchsorted(channels, f), len(channels) - 1)
                if f.size and not np.array_equal(channels[chan], f):
                    raise ValueError("map_visibilities: a frequency is not in `channels`")
            chan = np.ascontiguousarray(chan, dtype=np.int32)
        if chan.shape[0] != u.shape[0]:
            raise ValueError("map_visibilities: frequencies must have the length of u")
        if len(channels) > 64:
            raise ValueError("frank_b200 maps at most 64 frequency channels per call")
        Ms, js, H0, qlo, qhi = self._map(ctx, u, v, V, weights, w_stride, chan=chan, nchan=len(channels))
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
        ctx = self._context()
        if _is_cuda_tensor(q):                      # device-resident baselines: the result stays on the device
            if self._vis_model == 'debris' and k is None:
                raise ValueError("the debris model needs the vertical uv-distance k")
            return ctx.predict_visibilities_dev(q, k if self._vis_model == 'debris' else None, I, _lib.MODEL_CODE[self._vis_model],
                                                1.0 if self._vis_model == 'debris' else self._model_scale(geometry),
                                                self._H2).reshape(q.shape)
        q = np.asarray(q, dtype=np.float64)
        if self._vis_model == 'debris':
            if k is None:
                raise ValueError("the debris model needs the vertical uv-distance k")
            return ctx.predict_visibilities(q.reshape(-1), np.asarray(k).reshape(-1), I, 2, 1.0, self._H2).reshape(q.shape)
        return ctx.predict_visibilities(q.reshape(-1), None, I, _lib.MODEL_CODE[self._vis_model],
                                        self._model_scale(geometry), None).reshape(q.shape)

    def predict_sky(self, I, u, v, geometry):
        r"""Sky-plane visibilities of the profile I at device-resident sky baselines u, v (CUDA tensors): what
        FrankRadialFit.predict computes (frank/radial_fitters.py:56-98), in one fused device pass."""
        ctx = self._context()
        return ctx.predict_sky_dev(u, v, geometry.device_scalars(), I, _lib.MODEL_CODE[self._vis_model],
                                   1.0 if self._vis_model == 'debris' else self._model_scale(geometry), self._H2).reshape(u.shape)

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
            self._mu, self._U = _solution                 # U may be None: factorised on first use (see _factor)
            self._Dsvd = None
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

    def _factor(self):
        """The upper Cholesky factor of D^-1; computed on the device on first use for fits that arrived without one
        (batched sweeps return only p and mu)."""
        if self._U is None and getattr(self, '_Dsvd', None) is None:
            mu = self._mu
            self._fit()
            self._mu = mu
        return self._U

    def Dsolve(self, b):
        r"""D b through the GPU-computed Cholesky factor (post-fit helper, statistical_models.py:762-781)."""
        self._factor()
        if getattr(self, '_Dsvd', None) is not None:                                   # :778-781
            U, s1, V = self._Dsvd
            b = np.asarray(b)
            return np.dot(V.T, (np.dot(U.T, b).T * s1).T)
        if _DEVICE_DSOLVE:
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
    channel, one field and unit scale (what FrankFitter uses).  The whole fit -- objective, gradient, Hessian, its
    factorisation, the Newton solves and the line search -- is one device-resident call (fb_frank_lognormal_loop in
    include/frankb200.h; the host thread of the call only reads the scalars the line-search decisions need).
    """

    def __init__(self, DHT, M, j, p=None, scale=None, s0=None, guess=None, Nfields=None, full_hessian=1,
                 noise_likelihood=0, device=None, _solution=None):
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
        self._ctx = ctx
        if _solution is not None:                  # produced by the device-resident loop of FrankFitter
            self._s_MAP, self._U, self._status = _solution
            return
        # MinimizeNewton + LineSearch + the factor of the Hessian at the MAP point, device resident (fb_frank_lognormal_loop
        # with max_iter < 0; statistical_models.py:1073-1160, minimizer.py:74-283)
        out = ctx.frank_lognormal_loop(M, j, p, guess, self._s0, full_hessian=full_hessian, max_iter=-1)
        self._raise_for_status(out['status'])
        self._s_MAP = out['s']
        self._status = out['newton']
        self._U = np.triu(out['chol'])

    @staticmethod
    def _raise_for_status(rc):
        if rc == _lib.FB_E_SLOPE:
            raise ValueError("Round off in slope calculation")                        # minimizer.py:130-133
        if rc == _lib.FB_E_BADP:
            raise ValueError("Bad value in power spectrum. The power"
                             " spectrum must be positive and not contain"
                             " any NaN values. This is likely due to"
                             " your UVtable (incorrect units or weights), "
                             " or the deprojection being applied (incorrect"
                             " geometry and/or phase center). Else you may"
                             " want to increase `rout` by 10-20% or `n` so"
                             " that it is large, >~300.")
        if rc == _lib.FB_E_NOTPD:
            # the reference switches to an SVD pseudo-inverse of the Hessian here (:1152-1158); on every case tried its
            # next power-spectrum update then aborts with "Bad value in power spectrum" (tests/golden/make_golden.py,
            # gen_config3): report the loss of definiteness instead of carrying an unusable factor
            raise np.linalg.LinAlgError("LogNormalMAPModel: the Hessian at the MAP point is not positive definite")

    def _update_power_spectrum(self, alpha, p0, Tinv):
        """CriticalFilter.update_power_spectrum with this model's Hessian factor (device)."""
        self._ctx.dht_setup(self._DHT)            # the process-wide context may have served another DHT since
        self._ctx.ln_setup(self._M, self._j, self._s0, self._full_hess)
        self._ctx.ln_set_spectrum(self._p)
        _, p_new, rc = self._ctx.ln_posterior(self._s_MAP, self._p, alpha, p0, Tinv, want_chol=False)
        self._raise_for_status(rc)
        return p_new

    def Dsolve(self, b):
        r"""(Hessian at the MAP)^-1 b through its GPU-computed Cholesky factor (statistical_models.py:1160-1170)."""
        if _DEVICE_DSOLVE:
            self._ctx.dht_setup(self._DHT)
            return self._ctx.chol_solve(self._U, b)
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
