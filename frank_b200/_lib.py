"""ctypes binding of libfrankb200.so (the C ABI in include/frankb200.h).

There is no CPU fallback: importing this module without the compiled library, or creating a
context without a CUDA device, raises.
"""
import ctypes
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('FRANK_B200_LIB') or os.path.join(_HERE, 'lib', 'libfrankb200.so')     # (override: kernel A/B builds)

FB_E_QRANGE, FB_E_NOTPD, FB_E_BADP, FB_E_NOCONV, FB_E_RETRY, FB_E_SLOPE = 1, 2, 3, 4, 5, 6
MODEL_CODE = {'opt_thick': 0, 'opt_thin': 1, 'debris': 2}


class FBGeometry(ctypes.Structure):
    _fields_ = [('a_ra', ctypes.c_double), ('a_dec', ctypes.c_double),
                ('cos_pa', ctypes.c_double), ('sin_pa', ctypes.c_double),
                ('cos_inc', ctypes.c_double), ('sin_inc', ctypes.c_double)]


_c_p = ctypes.c_void_p
_c_i = ctypes.c_int
_c_l = ctypes.c_int64
_c_d = ctypes.c_double

_SIGNATURES = {
    'fb_version': ([], _c_i),
    'fb_ctx_create': ([ctypes.POINTER(_c_p), _c_i], _c_i),
    'fb_ctx_destroy': ([_c_p], _c_i),
    'fb_last_error': ([_c_p], ctypes.c_char_p),
    'fb_set_option': ([_c_p, ctypes.c_char_p, _c_d], _c_i),
    'fb_dht_setup': ([_c_p, _c_i, _c_d, _c_p, _c_p, _c_p, _c_d], _c_i),
    'fb_map_visibilities_dev': ([_c_p, _c_l, _c_p, _c_p, _c_p, _c_p, _c_i, _c_p, _c_i, ctypes.POINTER(FBGeometry),
                                 _c_i, _c_d, _c_p, _c_i, _c_d, _c_p, _c_p, _c_p, _c_p], _c_i),
    'fb_map_visibilities_host': ([_c_p, _c_l, _c_p, _c_p, _c_p, _c_p, _c_i, _c_p, _c_i, ctypes.POINTER(FBGeometry),
                                  _c_i, _c_d, _c_p, _c_i, _c_d, _c_p, _c_p, _c_p, _c_p], _c_i),
    'fb_map_visibilities_dev_async': ([_c_p, _c_l, _c_p, _c_p, _c_p, _c_p, _c_i, _c_p, _c_i, ctypes.POINTER(FBGeometry),
                                       _c_i, _c_d, _c_p, _c_i, _c_d, _c_p, _c_p, _c_p], _c_i),
    'fb_map_sync': ([_c_p, _c_p], _c_i),
    'fb_comm_unique_id': ([_c_p], _c_i),
    'fb_comm_init': ([_c_p, _c_i, _c_i, _c_p], _c_i),
    'fb_comm_destroy': ([_c_p], _c_i),
    'fb_comm_info': ([_c_p, _c_p, _c_p], _c_i),
    'fb_comm_allgather': ([_c_p, _c_p, _c_l, _c_p], _c_i),
    'fb_comm_allreduce_sum_dev': ([_c_p, _c_p, _c_l], _c_i),
    'fb_last_map_timing': ([_c_p, _c_p], _c_i),
    'fb_timer_start': ([_c_p], _c_i),
    'fb_timer_stop': ([_c_p, _c_p], _c_i),
    'fb_debug_prepped': ([_c_p, _c_l, _c_p, _c_p, _c_p, _c_p], _c_i),
    'fb_debug_j0': ([_c_p, _c_l, _c_p, _c_p], _c_i),
    'fb_debug_j0_far': ([_c_p, _c_l, _c_p, _c_p], _c_i),
    'fb_gaussian_fit': ([_c_p, _c_i, _c_p, _c_p, _c_p, _c_i, _c_p, _c_p, _c_p], _c_i),
    'fb_chol_solve': ([_c_p, _c_p, _c_i, _c_p, _c_p], _c_i),
    'fb_gaussian_svd': ([_c_p, _c_p, _c_p, _c_i, _c_p, _c_p, _c_p, _c_p], _c_i),
    'fb_predict_visibilities': ([_c_p, _c_l, _c_p, _c_p, _c_p, _c_i, _c_d, _c_p, _c_p], _c_i),
    'fb_predict_visibilities_dev': ([_c_p, _c_l, _c_p, _c_p, _c_p, _c_i, _c_d, _c_p, _c_p], _c_i),
    'fb_predict_sky_dev': ([_c_p, _c_l, _c_p, _c_p, ctypes.POINTER(FBGeometry), _c_p, _c_i, _c_d, _c_p, _c_p], _c_i),
    'fb_columns_gram_dev': ([_c_p, _c_l, _c_i, _c_p, _c_p], _c_i),
    'fb_apply_correction_dev': ([_c_p, _c_l, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p], _c_i),
    'fb_uv_max': ([_c_p, _c_l, _c_p, _c_p], _c_i),
    'fb_uv_bin': ([_c_p, _c_l, _c_p, _c_p, _c_i, _c_p, _c_i, _c_d, _c_i, _c_p, _c_p, _c_p, _c_p], _c_i),
    'fb_uv_bin_dev': ([_c_p, _c_l, _c_p, _c_p, _c_i, _c_p, _c_i, _c_d, _c_i, _c_p, _c_p, _c_p, _c_p], _c_i),
    'fb_ln_setup': ([_c_p, _c_p, _c_p, _c_d, _c_d], _c_i),
    'fb_ln_set_spectrum': ([_c_p, _c_p], _c_i),
    'fb_ln_eval': ([_c_p, _c_p, _c_p, _c_p], _c_i),
    'fb_ln_newton_direction': ([_c_p, _c_p, _c_i, _c_p, _c_p, _c_p], _c_i),
    'fb_ln_posterior': ([_c_p, _c_p, _c_p, _c_d, _c_d, _c_p, _c_p, _c_p, _c_p], _c_i),
    'fb_frank_lognormal_loop': ([_c_p, _c_p, _c_p, _c_p, _c_p, _c_d, _c_d, _c_d, _c_d, _c_p, _c_d, _c_i, _c_d, _c_p, _c_p, _c_p,
                                 _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_i], _c_i),
    'fb_frank_normal_loop': ([_c_p, _c_i, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_d, _c_i, _c_p, _c_p, _c_p, _c_p,
                              _c_p, _c_p, _c_p, _c_p, _c_i], _c_i),
}

_lib = None
_lock = threading.Lock()


def load():
    """Load libfrankb200.so (built by `make -C frank_b200/csrc` / __graft_entry__.build())."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise RuntimeError(
                    f"frank_b200: compiled CUDA library not found at {LIB_PATH}. Build it with "
                    "`make -C frank_b200/csrc` (needs nvcc, sm_100a). There is no CPU fallback.")
            lib = ctypes.CDLL(LIB_PATH)
            for name, (args, res) in _SIGNATURES.items():
                fn = getattr(lib, name)     # AttributeError if the ABI and the header drifted apart
                fn.argtypes = args
                fn.restype = res
            _lib = lib
    return _lib


def exported_symbols():
    return sorted(_SIGNATURES)


def _ptr(x):
    """Raw pointer of a torch tensor / numpy array / None."""
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data_as(_c_p)
    return _c_p(x.data_ptr())          # torch.Tensor


class Context(object):
    """One fb_ctx: a CUDA stream, workspaces and DHT tables on one device."""

    def __init__(self, device=0):
        self._lib = load()
        h = _c_p()
        rc = self._lib.fb_ctx_create(ctypes.byref(h), int(device))
        if rc != 0:
            raise RuntimeError(f"frank_b200: fb_ctx_create(device={device}) failed with status {rc} "
                               "(no CUDA device, or not an sm_100 part). There is no CPU fallback.")
        self._h = h
        self.device = int(device)
        self._dht_key = None

    def close(self):
        if getattr(self, '_h', None):
            self._lib.fb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def error(self):
        return self._lib.fb_last_error(self._h).decode()

    def set_option(self, name, value):
        self.check(self._lib.fb_set_option(self._h, name.encode(), float(value)), 'fb_set_option')

    def check(self, rc, what):
        if rc < 0:
            raise RuntimeError(f"frank_b200: {what} failed ({rc}): {self.error()}")
        return rc

    # -- DHT ---------------------------------------------------------------------------------
    def dht_setup(self, dht, x_max=0.0):
        if dht.order != 0:
            # the device kernels evaluate J0 only (fb_j0_table.h); an order-nu transform would silently mix J_nu tables
            # with J0 design rows
            raise NotImplementedError("frank_b200: the CUDA path implements the order-0 transform only (nu = 0)")
        key = (dht.Rmax, dht.size, dht.order)
        if self._dht_key == key:
            return
        j_nk = np.ascontiguousarray(dht._j_nk, dtype=np.float64)
        coef = np.ascontiguousarray((1 / (np.pi * dht.Qmax ** 2)) * dht._scale_factor, dtype=np.float64)
        Y = np.ascontiguousarray(dht.coefficients(), dtype=np.float64)
        self.check(self._lib.fb_dht_setup(self._h, dht.size, float(dht.Qmax), _ptr(j_nk), _ptr(coef), _ptr(Y),
                                          float(x_max)), 'fb_dht_setup')
        self._dht_key = key

    # -- mapping -----------------------------------------------------------------------------
    def map_visibilities(self, n, u, v, V, w, w_stride, geom, vis_model, model_scale, H2, check_qbounds,
                         q_last, M, j, H0, host=False, chan=None, nchan=1):
        """Thin call into fb_map_visibilities_{dev,host}.  Returns (status, qmin, qmax).  `chan` (int32 [n], values in
        [0, nchan)) selects the multi-frequency path: M [nchan, N, N], j [nchan, N]."""
        qmm = np.zeros(2)
        fn = self._lib.fb_map_visibilities_host if host else self._lib.fb_map_visibilities_dev
        H2p = None if H2 is None else np.ascontiguousarray(H2, dtype=np.float64)
        rc = fn(self._h, int(n), _ptr(u), _ptr(v), _ptr(V), _ptr(w), int(w_stride), _ptr(chan), int(nchan), ctypes.byref(geom),
                int(vis_model), float(model_scale), _ptr(H2p), int(bool(check_qbounds)), float(q_last),
                _ptr(M), _ptr(j), _ptr(H0), _ptr(qmm))
        self.check(rc, 'fb_map_visibilities')
        return rc, qmm[0], qmm[1]

    def map_visibilities_async(self, n, u, v, V, w, w_stride, geom, vis_model, model_scale, H2, check_qbounds,
                               q_last, M, j, H0, chan=None, nchan=1):
        """fb_map_visibilities_dev_async: enqueue only (device tensors); pair with map_sync()."""
        H2p = None if H2 is None else np.ascontiguousarray(H2, dtype=np.float64)
        rc = self._lib.fb_map_visibilities_dev_async(self._h, int(n), _ptr(u), _ptr(v), _ptr(V), _ptr(w), int(w_stride), _ptr(chan),
                                                     int(nchan), ctypes.byref(geom), int(vis_model), float(model_scale), _ptr(H2p),
                                                     int(bool(check_qbounds)), float(q_last), _ptr(M), _ptr(j), _ptr(H0))
        return self.check(rc, 'fb_map_visibilities_dev_async')

    def map_sync(self):
        """fb_map_sync: wait for the asynchronous mapping call; returns (status, qmin, qmax)."""
        qmm = np.zeros(2)
        rc = self.check(self._lib.fb_map_sync(self._h, _ptr(qmm)), 'fb_map_sync')
        return rc, qmm[0], qmm[1]

    # -- multi-GPU ---------------------------------------------------------------------------
    def comm_unique_id(self):
        buf = np.zeros(128, dtype=np.uint8)
        rc = self._lib.fb_comm_unique_id(_ptr(buf))
        if rc != 0:
            raise RuntimeError(f"frank_b200: fb_comm_unique_id failed ({rc}): is NCCL (libnccl.so.2) available?")
        return buf

    def comm_init(self, nranks, rank, unique_id):
        uid = np.ascontiguousarray(unique_id, dtype=np.uint8)
        assert uid.size == 128
        self.check(self._lib.fb_comm_init(self._h, int(nranks), int(rank), _ptr(uid)), 'fb_comm_init')

    def comm_destroy(self):
        self._lib.fb_comm_destroy(self._h)

    def comm_info(self):
        """(attached, rank, nranks) of the context's communicator."""
        r, n = ctypes.c_int(0), ctypes.c_int(1)
        has = self._lib.fb_comm_info(self._h, ctypes.byref(r), ctypes.byref(n))
        return bool(has), r.value, n.value

    def comm_allgather(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        _, _, n = self.comm_info()
        out = np.empty((n, x.size))
        self.check(self._lib.fb_comm_allgather(self._h, _ptr(x), x.size, _ptr(out)), 'fb_comm_allgather')
        return out

    def comm_allreduce_sum_dev(self, t):
        """In-place sum of a float64 CUDA tensor over the communicator, asynchronous on the library stream."""
        self.check(self._lib.fb_comm_allreduce_sum_dev(self._h, _ptr(t), t.numel()), 'fb_comm_allreduce_sum_dev')

    def apply_correction_dev(self, u, v, V, geom, want_q=False):
        """fb_apply_correction_dev on torch CUDA tensors: returns (up, vp, wp, Vp[, q]) as device tensors (Vp None when V
        is None)."""
        import torch
        for t in (u, v):
            if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float64):
                raise ValueError("apply_correction_dev: contiguous float64 CUDA tensors expected")
        n = u.numel()
        Vr = Vp = None
        if V is not None:
            if not (V.is_cuda and V.is_contiguous() and V.dtype == torch.complex128):
                raise ValueError("apply_correction_dev: V must be a contiguous complex128 CUDA tensor")
            Vr = torch.view_as_real(V)
            Vp = torch.empty_like(V)
        up, vp, wp = torch.empty_like(u), torch.empty_like(u), torch.empty_like(u)
        q = torch.empty_like(u) if want_q else None
        torch.cuda.current_stream(u.device).synchronize()
        self.check(self._lib.fb_apply_correction_dev(self._h, n, _ptr(u), _ptr(v), _ptr(Vr), ctypes.byref(geom), _ptr(up), _ptr(vp),
                                                     _ptr(wp), None if Vp is None else _ptr(torch.view_as_real(Vp)), _ptr(q)),
                   'fb_apply_correction_dev')
        return (up, vp, wp, Vp, q) if want_q else (up, vp, wp, Vp)

    def timer_start(self):
        self.check(self._lib.fb_timer_start(self._h), 'fb_timer_start')

    def timer_stop(self):
        """Device milliseconds on the library stream since timer_start (CUDA events)."""
        t = np.zeros(1)
        self.check(self._lib.fb_timer_stop(self._h, _ptr(t)), 'fb_timer_stop')
        return float(t[0])

    def last_map_timing(self):
        t = np.zeros(4)
        self._lib.fb_last_map_timing(self._h, _ptr(t))
        return {'prep_ms': t[0], 'gram_ms': t[1], 'finalize_ms': t[2], 'copy_ms': t[3]}

    def debug_prepped(self, n):
        a, kz, Vre, perm = np.empty(n), np.empty(n), np.empty(n), np.empty(n, dtype=np.uint32)
        self.check(self._lib.fb_debug_prepped(self._h, int(n), _ptr(a), _ptr(kz), _ptr(Vre), _ptr(perm)),
                   'fb_debug_prepped')
        return a, kz, Vre, perm

    # -- solver ------------------------------------------------------------------------------
    def gaussian_fit(self, M, j, p=None, want_chol=True):
        """fb_gaussian_fit for a batch of power spectra p [B, N] (or no prior).  Returns mu [B, N],
        chol [B, N, N] (upper factor in the upper triangle), info [B], status."""
        N = M.shape[0]
        M = np.ascontiguousarray(M, dtype=np.float64)
        j = np.ascontiguousarray(j, dtype=np.float64)
        if p is None:
            B, pp = 1, None
        else:
            pp = np.ascontiguousarray(np.atleast_2d(p), dtype=np.float64)
            B = pp.shape[0]
        mu = np.empty((B, N))
        chol = np.empty((B, N, N)) if want_chol else None
        info = np.zeros(B, dtype=np.int32)
        rc = self._lib.fb_gaussian_fit(self._h, B, _ptr(M), _ptr(j), _ptr(pp), int(p is not None), _ptr(mu), _ptr(chol),
                                       _ptr(info))
        self.check(rc, 'fb_gaussian_fit')
        return mu, chol, info, rc

    def chol_solve(self, U, b):
        """fb_chol_solve: (U^T U)^-1 b for a vector b [N] or a matrix b [N, k] (columns = right-hand sides)."""
        U = np.ascontiguousarray(U, dtype=np.float64)
        b = np.asarray(b, dtype=np.float64)
        B = np.ascontiguousarray(b.reshape(b.shape[0], -1).T)          # one right-hand side per row
        X = np.empty_like(B)
        self.check(self._lib.fb_chol_solve(self._h, _ptr(U), B.shape[0], _ptr(B), _ptr(X)), 'fb_chol_solve')
        return np.ascontiguousarray(X.T).reshape(b.shape)

    def gaussian_svd(self, M, p=None):
        """fb_gaussian_svd: U, s, Vt of D^-1 = M (+ Y^T diag(1/p) Y) as scipy.linalg.svd would return them."""
        N = M.shape[0]
        M = np.ascontiguousarray(M, dtype=np.float64)
        pp = None if p is None else np.ascontiguousarray(p, dtype=np.float64).reshape(N)
        U, s, Vt = np.empty((N, N)), np.empty(N), np.empty((N, N))
        sweeps = np.zeros(1, dtype=np.int32)
        rc = self.check(self._lib.fb_gaussian_svd(self._h, _ptr(M), _ptr(pp), int(p is not None), _ptr(U), _ptr(s), _ptr(Vt),
                                                  _ptr(sweeps)), 'fb_gaussian_svd')
        if rc == FB_E_NOCONV:
            raise np.linalg.LinAlgError("SVD did not converge")
        return U, s, Vt, int(sweeps[0])

    def frank_normal_loop(self, M, j, p_init, alpha, p0, Tinv, tol, max_iter, want_chol=True, hist_cap=0):
        """fb_frank_normal_loop for B hyper-parameter points.  Returns a dict."""
        N = M.shape[0]
        M = np.ascontiguousarray(M, dtype=np.float64)
        j = np.ascontiguousarray(j, dtype=np.float64)
        p_init = np.ascontiguousarray(np.atleast_2d(p_init), dtype=np.float64)
        B = p_init.shape[0]
        alpha = np.ascontiguousarray(np.broadcast_to(np.asarray(alpha, dtype=np.float64), (B,)))
        p0 = np.ascontiguousarray(np.broadcast_to(np.asarray(p0, dtype=np.float64), (B,)))
        Tinv = np.ascontiguousarray(np.broadcast_to(np.asarray(Tinv, dtype=np.float64), (B, N, N)))
        p = np.empty((B, N)); mu = np.empty((B, N))
        chol = np.empty((B, N, N)) if want_chol else None
        niter = np.zeros(B, dtype=np.int32); conv = np.zeros(B, dtype=np.int32); info = np.zeros(B, dtype=np.int32)
        hp = hm = None
        if hist_cap > 0:
            # history buffers are kept across calls (callers copy the rows they keep): a fresh multi-megabyte
            # allocation per fit costs page faults that, on a busy host, showed up as 0.1-0.8 s of jitter
            if getattr(self, '_hist_shape', None) != (B, hist_cap, N):
                self._hist = (np.zeros((B, hist_cap, N)), np.zeros((B, hist_cap, N)))
                self._hist_shape = (B, hist_cap, N)
            hp, hm = self._hist
        rc = self._lib.fb_frank_normal_loop(self._h, B, _ptr(M), _ptr(j), _ptr(p_init), _ptr(alpha), _ptr(p0), _ptr(Tinv),
                                            float(tol), int(max_iter), _ptr(p), _ptr(mu), _ptr(chol), _ptr(niter),
                                            _ptr(conv), _ptr(info), _ptr(hp), _ptr(hm), int(hist_cap))
        self.check(rc, 'fb_frank_normal_loop')
        return {'p': p, 'mu': mu, 'chol': chol, 'niter': niter, 'converged': conv, 'info': info, 'status': rc,
                'hist_p': hp, 'hist_mu': hm}

    # -- LogNormal model ------------------------------------------------------------------------
    def frank_lognormal_loop(self, M, j, p_init, guess, s0, alpha=0.0, p0=0.0, Tinv=None, tol=1e-3, max_iter=-1,
                             full_hessian=1.0, newton_tol=1e-7, want_chol=True, hist_cap=0):
        """fb_frank_lognormal_loop: the log-normal MAP fit (max_iter < 0) or the whole power-spectrum iteration around
        it, device resident.  Returns a dict; 'status' is 0, FB_E_BADP, FB_E_NOTPD or FB_E_SLOPE."""
        N = M.shape[0]
        M = np.ascontiguousarray(M, dtype=np.float64); j = np.ascontiguousarray(j, dtype=np.float64)
        p_init = np.ascontiguousarray(p_init, dtype=np.float64).reshape(N)
        guess = np.ascontiguousarray(guess, dtype=np.float64).reshape(N)
        Tc = None if Tinv is None else np.ascontiguousarray(Tinv, dtype=np.float64)
        s, p = np.empty(N), np.empty(N)
        chol = np.empty((N, N)) if want_chol else None
        niter, conv, info = np.zeros(1, dtype=np.int32), np.zeros(1, dtype=np.int32), np.zeros(1, dtype=np.int32)
        stats = np.zeros(7, dtype=np.int64)
        hp = hs = None
        if hist_cap > 0:
            hp, hs = np.zeros((hist_cap, N)), np.zeros((hist_cap, N))
        rc = self._lib.fb_frank_lognormal_loop(self._h, _ptr(M), _ptr(j), _ptr(p_init), _ptr(guess), float(s0), float(full_hessian),
                                               float(alpha), float(p0), _ptr(Tc), float(tol), int(max_iter), float(newton_tol),
                                               _ptr(s), _ptr(p), _ptr(chol), _ptr(niter), _ptr(conv), _ptr(info), _ptr(stats),
                                               _ptr(hp), _ptr(hs), int(hist_cap))
        self.check(rc, 'fb_frank_lognormal_loop')
        return {'s': s, 'p': p, 'chol': chol, 'niter': int(niter[0]), 'converged': bool(conv[0]), 'info': int(info[0]),
                'status': rc, 'hist_p': hp, 'hist_s': hs,
                'newton': {'steps': int(stats[0]), 'evaluations': int(stats[1]), 'hessians': int(stats[2]),
                           'status_counts': stats[3:7].tolist()}}

    def ln_setup(self, M, j, s0, full_hessian=1.0):
        M = np.ascontiguousarray(M, dtype=np.float64); j = np.ascontiguousarray(j, dtype=np.float64)
        self.check(self._lib.fb_ln_setup(self._h, _ptr(M), _ptr(j), float(s0), float(full_hessian)), 'fb_ln_setup')

    def ln_set_spectrum(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64)
        return self.check(self._lib.fb_ln_set_spectrum(self._h, _ptr(p)), 'fb_ln_set_spectrum')

    def ln_eval(self, s, want_grad=False):
        s = np.ascontiguousarray(s, dtype=np.float64)
        f = np.zeros(1)
        g = np.empty_like(s) if want_grad else None
        self.check(self._lib.fb_ln_eval(self._h, _ptr(s), _ptr(f), _ptr(g)), 'fb_ln_eval')
        return (float(f[0]), g) if want_grad else float(f[0])

    def ln_newton_direction(self, s, refactor):
        s = np.ascontiguousarray(s, dtype=np.float64)
        g, dx = np.empty_like(s), np.empty_like(s)
        info = np.zeros(1, dtype=np.int32)
        rc = self.check(self._lib.fb_ln_newton_direction(self._h, _ptr(s), int(refactor), _ptr(g), _ptr(dx), _ptr(info)),
                        'fb_ln_newton_direction')
        return g, dx, rc

    def ln_posterior(self, s, p, alpha=None, p0=None, Tinv=None, want_chol=True):
        s = np.ascontiguousarray(s, dtype=np.float64); p = np.ascontiguousarray(p, dtype=np.float64)
        N = s.size
        chol = np.empty((N, N)) if want_chol else None
        p_new = np.empty(N) if Tinv is not None else None
        ldl_c = None if Tinv is None else np.ascontiguousarray(Tinv, dtype=np.float64)
        info = np.zeros(1, dtype=np.int32)
        rc = self.check(self._lib.fb_ln_posterior(self._h, _ptr(s), _ptr(p), float(alpha or 0.0), float(p0 or 0.0), _ptr(ldl_c),
                                                  _ptr(chol), _ptr(p_new), _ptr(info)), 'fb_ln_posterior')
        return chol, p_new, rc

    def predict_visibilities(self, q, kz, I, vis_model, model_scale, H2):
        q = np.ascontiguousarray(q, dtype=np.float64)
        I = np.ascontiguousarray(I, dtype=np.float64)
        kzc = None if kz is None else np.ascontiguousarray(np.broadcast_to(kz, q.shape), dtype=np.float64)
        H2c = None if H2 is None else np.ascontiguousarray(H2, dtype=np.float64)
        V = np.empty_like(q)
        self.check(self._lib.fb_predict_visibilities(self._h, q.size, _ptr(q), _ptr(kzc), _ptr(I), int(vis_model), float(model_scale),
                                                     _ptr(H2c), _ptr(V)), 'fb_predict_visibilities')
        return V

    def predict_visibilities_dev(self, q, kz, I, vis_model, model_scale, H2):
        """fb_predict_visibilities_dev on float64 CUDA tensors q [n] (and kz [n] for the debris model); returns V [n] on the device."""
        import torch
        I = np.ascontiguousarray(I, dtype=np.float64)
        H2c = None if H2 is None else np.ascontiguousarray(H2, dtype=np.float64)
        q = q.reshape(-1).to(torch.float64).contiguous()
        kzc = None if kz is None else kz.reshape(-1).to(torch.float64).contiguous()
        V = torch.empty_like(q)
        torch.cuda.current_stream(q.device).synchronize()
        self.check(self._lib.fb_predict_visibilities_dev(self._h, q.numel(), _ptr(q), _ptr(kzc), _ptr(I), int(vis_model), float(model_scale),
                                                         _ptr(H2c), _ptr(V)), 'fb_predict_visibilities_dev')
        return V

    def predict_sky_dev(self, u, v, geom, I, vis_model, model_scale, H2):
        """fb_predict_sky_dev: sky-plane prediction for float64 CUDA tensors u, v [n]; returns a complex128 CUDA tensor [n]."""
        import torch
        I = np.ascontiguousarray(I, dtype=np.float64)
        H2c = None if H2 is None else np.ascontiguousarray(H2, dtype=np.float64)
        u = u.reshape(-1).to(torch.float64).contiguous()
        v = v.reshape(-1).to(torch.float64).contiguous()
        V = torch.empty(u.numel(), dtype=torch.complex128, device=u.device)
        torch.cuda.current_stream(u.device).synchronize()
        self.check(self._lib.fb_predict_sky_dev(self._h, u.numel(), _ptr(u), _ptr(v), ctypes.byref(geom), _ptr(I), int(vis_model),
                                                float(model_scale), _ptr(H2c), _ptr(torch.view_as_real(V))), 'fb_predict_sky_dev')
        return V

    def columns_gram_dev(self, cols):
        """fb_columns_gram_dev: X^T X of up to 8 float64 CUDA vectors of equal length, summed in a fixed order; returns a
        NumPy [K, K] array."""
        import torch
        K, n = len(cols), cols[0].numel()
        cols = [c.reshape(-1).contiguous() for c in cols]
        ptrs = (ctypes.c_void_p * K)(*[c.data_ptr() for c in cols])
        G = np.empty((K, K))
        torch.cuda.current_stream(cols[0].device).synchronize()
        self.check(self._lib.fb_columns_gram_dev(self._h, n, K, ctypes.cast(ptrs, _c_p), _ptr(G)), 'fb_columns_gram_dev')
        return G

    # -- uv binning -----------------------------------------------------------------------------
    def uv_max(self, uv):
        uv = np.ascontiguousarray(uv, dtype=np.float64)
        m = np.zeros(1)
        self.check(self._lib.fb_uv_max(self._h, uv.size, _ptr(uv), _ptr(m)), 'fb_uv_max')
        return float(m[0])

    def uv_bin(self, uv, V, w, bin_width, nbins):
        uv = np.ascontiguousarray(uv, dtype=np.float64)
        n = uv.size
        is_c = np.iscomplexobj(V)
        V = np.ascontiguousarray(V, dtype=np.complex128 if is_c else np.float64)
        w = np.ascontiguousarray(np.atleast_1d(w), dtype=np.float64)
        idx = np.empty(n, dtype=np.int32)
        counts = np.empty(nbins, dtype=np.int64)
        sums = np.empty((nbins, 4)); err = np.empty((nbins, 2))
        self.check(self._lib.fb_uv_bin(self._h, n, _ptr(uv), _ptr(V), int(is_c), _ptr(w), int(w.size > 1), float(bin_width),
                                       int(nbins), _ptr(idx), _ptr(counts), _ptr(sums), _ptr(err)), 'fb_uv_bin')
        return idx, counts, sums, err

    def uv_bin_dev(self, uv, V, w, bin_width, nbins):
        """fb_uv_bin_dev on torch CUDA tensors (uv, w float64 [n] or w [1]; V complex128 or float64 [n]).  Returns
        device tensors idx (int32 [n]), counts (int64 [nbins]), sums (float64 [nbins, 4]), err (float64 [nbins, 2])."""
        import torch
        n = uv.numel()
        is_c = V.is_complex()
        Vr = torch.view_as_real(V) if is_c else V
        for t in (uv, Vr, w):
            if not (t.is_cuda and t.is_contiguous() and t.dtype == torch.float64):
                raise ValueError("uv_bin_dev: contiguous float64 / complex128 CUDA tensors expected")
        idx = torch.empty(n, dtype=torch.int32, device=uv.device)
        counts = torch.empty(nbins, dtype=torch.int64, device=uv.device)
        sums = torch.empty((nbins, 4), dtype=torch.float64, device=uv.device)
        err = torch.empty((nbins, 2), dtype=torch.float64, device=uv.device)
        torch.cuda.current_stream(uv.device).synchronize()       # inputs were produced on torch's stream
        self.check(self._lib.fb_uv_bin_dev(self._h, n, _ptr(uv), _ptr(Vr), int(is_c), _ptr(w), int(w.numel() > 1),
                                           float(bin_width), int(nbins), _ptr(idx), _ptr(counts), _ptr(sums), _ptr(err)),
                   'fb_uv_bin_dev')
        return idx, counts, sums, err

    def debug_j0(self, x, far=False):
        x = np.ascontiguousarray(x, dtype=np.float64)
        out = np.empty_like(x)
        fn = self._lib.fb_debug_j0_far if far else self._lib.fb_debug_j0
        self.check(fn(self._h, x.size, _ptr(x), _ptr(out)), 'fb_debug_j0')
        return out


_contexts = {}


def get_context(device=None):
    """Process-wide context per device (one process per GPU is the deployment model)."""
    if device is None:
        device = int(os.environ.get('LOCAL_RANK', '0')) if 'FRANK_B200_DEVICE' not in os.environ \
            else int(os.environ['FRANK_B200_DEVICE'])
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]
