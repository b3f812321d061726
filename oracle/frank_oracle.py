"""
oracle/frank_oracle.py -- TEST INFRASTRUCTURE ONLY.

CPU restatement (NumPy/SciPy, float64) of the arithmetic of discsim/frank's
visibility -> Gaussian-process normal-equations path and of the dense solves that
consume it.  It exists to CHECK the CUDA path; the product (`frank_b200/`) never
imports it.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s CPU-baseline /
`--impl reference` legs may use it.

Every function names the reference lines (`/root/reference/frank/...`) it restates.
The third-party numerics the reference leans on are used through the same public
entry points (scipy.special.j0/j1/jn_zeros, numpy.dot -> BLAS, scipy.linalg
cho_factor/cho_solve/lu_factor/lu_solve, scipy.sparse.linalg.spsolve); a standalone C
restatement of J0 lives in oracle/cephes_j0.c.

Parity pinning: tests/test_oracle_golden.py compares every function here against
fixtures produced by running the UNMODIFIED reference (frank 1.2.3 imported from
/root/reference) on seeded inputs -- tests/golden/make_golden.py is the generator --
plus the reference's own analytic known-answer test (Gaussian Hankel pair,
frank/tests.py:37-94).  The reference's AS209 golden vectors cannot be replayed (the
data blob is absent from the checkout).
"""
import numpy as np
import scipy.linalg
import scipy.sparse
import scipy.sparse.linalg
from scipy.special import j0, j1, jn_zeros

# frank/constants.py:23-25
RAD_TO_ARCSEC = 3600 * 180 / np.pi
DEG_TO_RAD = np.pi / 180


# --------------------------------------------------------------------------------------
# Discrete Hankel transform tables            frank/hankel.py:55-93, 127-204
# --------------------------------------------------------------------------------------
class DHTTables(object):
    """Baddour-Chouinard DHT tables of order 0 (frank/hankel.py:55-93).

    Rmax is in radians (FourierBesselFitter divides arcsec by rad_to_arcsec first,
    frank/radial_fitters.py:441)."""

    def __init__(self, Rmax, N):
        zeros = jn_zeros(0, N + 1)                       # hankel.py:72
        self.j_nk, self.j_nN = zeros[:-1], zeros[-1]      # hankel.py:73
        self.N = N
        self.Rmax = Rmax
        self.Qmax = self.j_nN / (2 * np.pi * Rmax)        # hankel.py:75
        self.r = Rmax * (self.j_nk / self.j_nN)           # hankel.py:77
        self.q = self.Qmax * (self.j_nk / self.j_nN)      # hankel.py:78
        J1row = np.outer(np.ones_like(self.j_nk), j1(self.j_nk))          # hankel.py:84
        self.Ykm = (2 / (self.j_nN * J1row * J1row)) * \
            j0(np.prod(np.meshgrid(self.j_nk, self.j_nk / self.j_nN), axis=0))  # hankel.py:86-87
        self.scale_factor = 1 / j1(self.j_nk) ** 2        # hankel.py:89

    def coefficients(self, q=None):
        """Forward transform matrix (hankel.py:187-204)."""
        norm = 1 / (np.pi * self.Qmax ** 2)
        k = 1. / self.Qmax
        if q is None:
            return 0.5 * self.j_nN * norm * self.Ykm      # hankel.py:198-199
        return (norm * self.scale_factor) * j0(np.outer(k * q, self.j_nk))  # hankel.py:201-202

    def transform(self, f):
        """Forward DHT at the collocation points (hankel.py:150-165)."""
        norm = (2 * np.pi * self.Rmax ** 2) / self.j_nN
        return norm * np.dot(self.Ykm, f)


# --------------------------------------------------------------------------------------
# Geometry                                      frank/geometry.py:41-131, 202-236
# --------------------------------------------------------------------------------------
def phase_shift(u, v, V, dRA, dDec, inverse=False):
    """geometry.py:69-79.  dRA/dDec in arcsec."""
    a = dRA * (2. * np.pi / RAD_TO_ARCSEC)
    b = dDec * (2. * np.pi / RAD_TO_ARCSEC)
    phi = u * a + v * b
    rot = np.cos(phi) + 1j * np.sin(phi)
    return V / rot if inverse else V * rot


def deproject(u, v, inc, PA, inverse=False):
    """geometry.py:111-131.  inc/PA in degrees."""
    inc = inc * DEG_TO_RAD
    PA = PA * DEG_TO_RAD
    cos_t, sin_t = np.cos(PA), np.sin(PA)
    if inverse:
        sin_t = sin_t * -1
        u = u / np.cos(inc)
    up = u * cos_t - v * sin_t
    vp = u * sin_t + v * cos_t
    if inverse:
        return up, vp
    wp = up * np.sin(inc)
    up = up * np.cos(inc)
    return up, vp, wp


def apply_correction(u, v, V, inc, PA, dRA, dDec):
    """SourceGeometry.apply_correction(use3D=True), geometry.py:230-236."""
    Vp = phase_shift(u, v, V, dRA, dDec, inverse=True)
    up, vp, wp = deproject(u, v, inc, PA)
    return up, vp, wp, Vp


def undo_correction(u, v, V, inc, PA, dRA, dDec):
    """SourceGeometry.undo_correction, geometry.py:260-265."""
    up, vp = deproject(u, v, inc, PA, inverse=True)
    return up, vp, phase_shift(up, vp, V, dRA, dDec, inverse=False)


# --------------------------------------------------------------------------------------
# Visibility mapping                             frank/statistical_models.py:109-237, 483-535
# --------------------------------------------------------------------------------------
def debris_H2(dht, scale_height):
    """statistical_models.py:101-102; scale_height is a function of r in arcsec."""
    h = scale_height(dht.r * RAD_TO_ARCSEC)
    return 0.5 * (2 * np.pi * h / RAD_TO_ARCSEC) ** 2


def mapping_coefficients(dht, qs, ks, vis_model, inc, H2=None):
    """_get_mapping_coefficients (forward), statistical_models.py:486-509."""
    if vis_model == 'opt_thick':
        scale = np.cos(inc * DEG_TO_RAD)
    elif vis_model == 'opt_thin':
        scale = 1
    elif vis_model == 'debris':
        scale = np.exp(-np.outer(ks * ks, H2))
    else:
        raise ValueError("vis_model must be one of ['opt_thick', 'opt_thin', 'debris']")
    return dht.coefficients(qs) * scale


def map_visibilities(dht, u, v, V, weights, inc, PA, dRA, dDec, vis_model='opt_thick',
                     H2=None, block_size=10 ** 5, frequencies=None, check_qbounds=True):
    """VisibilityMapping.map_visibilities, statistical_models.py:165-237.

    Returns dict(M, j, null_likelihood, channels, qmin, qmax).  Raises ValueError when the
    data reach beyond the last collocation point (statistical_models.py:526-535)."""
    up, vp, k, Vp = apply_correction(u, v, V, inc, PA, dRA, dDec)
    q = np.hypot(up, vp)                                   # :166
    if check_qbounds and dht.q[-1] < q.max():              # :526
        raise ValueError("Last collocation point is at a shorter baseline than the longest "
                         "deprojected baseline in the dataset")
    Vre = Vp.real                                          # :172
    w = np.ones_like(Vre) * weights                        # :173
    multi = frequencies is not None
    if not multi:
        frequencies = np.ones_like(Vre)
    channels = np.unique(frequencies)                      # :180
    N = dht.N
    Ms = np.zeros([len(channels), N, N])
    js = np.zeros([len(channels), N])
    Nstep = int(block_size / N + 1)                        # :194
    for c, f in enumerate(channels):
        sel = frequencies == f
        qi, ki, wi, Vi = q[sel], k[sel], w[sel], Vre[sel]
        for start in range(0, len(Vi), Nstep):             # :200-214
            sl = slice(start, start + Nstep)
            X = mapping_coefficients(dht, qi[sl], ki[sl], vis_model, inc, H2)
            wXT = np.array(X.T * wi[sl], order='C')        # :208
            Ms[c] += np.dot(wXT, X)                        # :210
            js[c] += np.dot(wXT, Vi[sl])                   # :211
    H0 = 0.5 * np.sum(np.log(w / (2 * np.pi)) - Vre * w * Vre)   # :218
    out = {'channels': channels if multi else None, 'null_likelihood': H0,
           'qmin': q.min(), 'qmax': q.max(), 'q': q, 'k': k, 'Vre': Vre, 'w': w}
    out['M'], out['j'] = (Ms, js) if multi else (Ms[0], js[0])
    return out


def predict_visibilities(dht, I, q, k, vis_model, inc, H2=None, block_size=10 ** 5):
    """VisibilityMapping.predict_visibilities, statistical_models.py:306-329."""
    Ni = int(block_size / dht.N + 1)
    out = []
    for start in range(0, len(q), Ni):
        sl = slice(start, start + Ni)
        H = mapping_coefficients(dht, q[sl], None if k is None else k[sl], vis_model, inc, H2)
        out.append(np.dot(H, I))
    return np.concatenate(out)


# --------------------------------------------------------------------------------------
# Gaussian model                                 frank/statistical_models.py:700-781
# --------------------------------------------------------------------------------------
class GaussianSolve(object):
    """GaussianModel for Nfields=1 (statistical_models.py:650-760): D^-1 = M + S(p)^-1."""

    def __init__(self, dht, M, j, p=None):
        if M.ndim == 3:                                    # multi-channel, unit scale (:713-726)
            M, j = M.sum(axis=0), j.sum(axis=0)
        self.p = p
        if p is not None:
            if np.any(p <= 0) or np.any(np.isnan(p)):      # :688
                raise ValueError("Bad value in power spectrum")
            Y = dht.coefficients()
            self.Sinv = np.einsum('ji,j,jk->ik', Y, 1 / p, Y)     # :701
            Dinv = M + self.Sinv
        else:
            self.Sinv = None
            Dinv = M + 0
        self.Dinv = Dinv
        self.j = j
        try:
            self.chol = scipy.linalg.cho_factor(Dinv)      # :742 (upper)
            self.svd = None
            self.mu = scipy.linalg.cho_solve(self.chol, j)
        except np.linalg.LinAlgError:                      # :747-755
            U, s, Vt = scipy.linalg.svd(Dinv, full_matrices=False)
            s1 = np.where(s > 0, 1. / s, 0)
            self.chol, self.svd = None, (U, s1, Vt)
            self.mu = np.dot(Vt.T, np.multiply(np.dot(U.T, j), s1))

    def Dsolve(self, b):                                   # :777-781
        if self.chol is not None:
            return scipy.linalg.cho_solve(self.chol, b)
        U, s1, Vt = self.svd
        return np.dot(Vt.T, np.multiply(np.dot(U.T, b), s1))

    @property
    def MAP(self):
        return self.mu

    @property
    def power_spectrum(self):
        return self.p


# --------------------------------------------------------------------------------------
# Critical filter                                frank/filter.py:23-62, 154-181
# --------------------------------------------------------------------------------------
def smoothing_matrix(dht, weights_smooth):
    """spectral_smoothing_matrix, filter.py:41-62 (returns a scipy sparse matrix)."""
    log_q = np.log(dht.q)
    dc = (log_q[2:] - log_q[:-2]) / 2
    de = np.diff(log_q)
    N = dht.N
    D = np.zeros([3, N])
    D[0, :-2] = 1 / (dc * de[:-1])
    D[1, 1:-1] = -(1 / de[1:] + 1 / de[:-1]) / dc
    D[2, 2:] = 1 / (dc * de[1:])
    Delta = scipy.sparse.dia_matrix((D, [-1, 0, 1]), shape=(N, N))
    dce = np.zeros_like(log_q)
    dce[1:-1] = dc
    dce = scipy.sparse.dia_matrix((dce.reshape(1, -1), 0), shape=(N, N))
    return weights_smooth * Delta.T.dot(dce.dot(Delta))


def smoothing_bands(dht, weights_smooth):
    """The five diagonals (offsets -2..2) of T + I as a dense [5, N] array; band[o+2, i] = (T+I)[i, i+o].
    Derived from filter.py:41-62; this is what the device band solver consumes."""
    T = (smoothing_matrix(dht, weights_smooth) + scipy.sparse.identity(dht.N)).toarray()
    N = dht.N
    bands = np.zeros([5, N])
    for o in range(-2, 3):
        for i in range(N):
            if 0 <= i + o < N:
                bands[o + 2, i] = T[i, i + o]
    return bands


def update_power_spectrum(dht, fit, Tij, alpha, p0):
    """CriticalFilter.update_power_spectrum, filter.py:156-177 (rho = 1)."""
    Y = dht.coefficients()
    TpI = Tij + scipy.sparse.identity(dht.N)
    Tr1 = np.dot(Y, fit.MAP) ** 2
    Tr2 = np.einsum('ij,ji->i', Y, fit.Dsolve(Y.T))
    p = fit.power_spectrum
    beta = (p0 + 0.5 * (Tr1 + Tr2)) / p - (alpha - 1.0 + 0.5 * 1.0)
    tau = scipy.sparse.linalg.spsolve(scipy.sparse.csc_matrix(TpI), beta + np.log(p))
    return np.exp(tau)


def check_convergence(p_new, p_old, tol):
    """filter.py:179-181."""
    return np.all(np.abs(p_new - p_old) <= tol * p_new)


# --------------------------------------------------------------------------------------
# Log-normal MAP model                           frank/statistical_models.py:1073-1160
# Newton minimiser / line search                 frank/minimizer.py:74-283
# --------------------------------------------------------------------------------------
class _Backtrack(object):
    """LineSearch (minimizer.py:44-184) for the scalar-objective (root=False) use."""

    def __init__(self, reduce_step, armijo=1e-4, min_step_frac=0.1):
        self.reduce_step = reduce_step
        self.armijo = armijo
        self.l_min = min_step_frac
        self.reduction = None

    def __call__(self, func, grad, x0, p, f0):
        nfev = 0
        cost = f0
        p = self.reduce_step(p, x0)                        # :120
        delta_f = np.dot(grad, p)                          # :127
        if delta_f > 0:
            raise ValueError("Round off in slope calculation")
        lam = 1.0
        cost_save = lam_save = None
        while True:
            x_new = x0 + lam * p
            if np.all(x_new == x0):                        # :139
                return x0, f0, nfev, True
            cost_new = func(x_new)
            nfev += 1
            if cost_new <= (cost + self.armijo * lam * delta_f):   # :146
                self.reduction = lam
                return x_new, cost_new, nfev, False
            if lam == 1.0:                                 # :151-154
                lam_new = -0.5 * delta_f / (cost_new - cost - delta_f)
            else:                                          # :156-173
                r1 = (cost_new - cost - lam * delta_f) / (lam * lam)
                r2 = (cost_save - cost - lam_save * delta_f) / (lam_save * lam_save)
                a = (r1 - r2) / (lam - lam_save)
                b = (lam * r2 - lam_save * r1) / (lam - lam_save)
                if a == 0:
                    lam_new = -0.5 * delta_f / b
                else:
                    d = b * b - 3 * a * delta_f
                    if d < 0:
                        lam_new = 0.5 * lam
                    elif b <= 0:
                        lam_new = (-b + np.sqrt(d)) / (3 * a)
                    else:
                        lam_new = -1 * delta_f / (b + np.sqrt(d))
                    lam_new = min(0.5 * lam, lam_new)
            if np.isnan(lam_new):                          # :175-177
                lam_new = self.l_min * lam
            lam_save, cost_save = lam, cost_new
            lam = max(lam_new, self.l_min * lam)           # :181


def minimize_newton(fun, jac, hess, guess, search, max_step=10 ** 5, max_hev=1000, tol=1e-5):
    """MinimizeNewton, minimizer.py:228-283.  Returns x, (status, nstep, nfev, nhess)."""
    need_hess = True
    nfev, nhess = 1, 0
    x = guess
    fx = fun(x)
    lu = None
    for nstep in range(max_step):
        if need_hess:
            if nhess == max_hev:
                return x, (3, nstep, nfev, nhess)
            lu = scipy.linalg.lu_factor(hess(x))           # :238
            nhess += 1
        jx = jac(x)
        dx = scipy.linalg.lu_solve(lu, -jx)
        if np.dot(jx, dx) < 0:                             # :244
            x, fx, fev, failed = search(fun, jx, x, dx, fx)
            nfev += fev
        else:
            failed = True
        if failed:                                         # :250-274
            x, fx, fev, failed_descent = search(fun, jx, x, -jx, fx)
            nfev += fev
            if failed_descent:
                dx = search.reduce_step(-jx, x)
                for _ in range(10):
                    xn = x + dx
                    fn = fun(xn)
                    nfev += 1
                    if fn < fx:
                        break
                    dx = dx * 2 ** -4
                else:
                    return x, (1, nstep, nfev, nhess)
                fx, x = fn, xn
        need_hess = failed or (search.reduction != 1.0)    # :276
        if (np.abs(jac(x)) * np.abs(x)).max() < tol * max(np.abs(fx), 1):   # :281
            return x, (0, nstep, nfev, nhess)
    return x, (2, max_step, nfev, nhess)


class LogNormalSolve(object):
    """LogNormalMAPModel, one channel / one field / unit scale
    (statistical_models.py:1064-1158)."""

    def __init__(self, dht, M, j, p, guess, s0, full_hessian=1):
        if np.any(p <= 0) or np.any(np.isnan(p)):
            raise ValueError("Bad value in power spectrum")
        Y = dht.coefficients()
        Sinv = np.einsum('ji,j,jk->ik', Y, 1 / p, Y)       # :1065
        self.p, self.s0, self.Sinv = p, s0, Sinv

        # Same einsum contractions as the reference (Nf = Ns = 1) so that round-off, and with it
        # the Armijo accept/reject decisions of the line search, follow the same trajectory.
        M3, j2, S3 = M.reshape(1, *M.shape), j.reshape(1, -1), Sinv.reshape(1, *Sinv.shape)
        scale = np.ones([1, 1])
        s0c = np.atleast_1d(s0).reshape(1, 1)
        Nr = len(j)

        def f(s):                                          # :1088-1098
            s = s.reshape(1, Nr)
            I = np.exp(np.dot(scale, s + s0c))
            val = 0.5 * np.einsum('ij,ijk,ik', s, S3, s)
            val += 0.5 * np.einsum('ij,ijk,ik', I, M3, I)
            val -= np.sum(I * j2)
            return val

        def g(s):                                          # :1100-1111
            s = s.reshape(1, Nr)
            I = np.exp(np.dot(scale, s + s0c))
            sI = np.einsum('is,ij->isj', scale, I)
            S1_s = np.einsum('sjk,sk->sj', S3, s)
            MI = np.einsum('isj,ijk,ik->sj', sI, M3, I)
            jI = np.einsum('isj,ij->sj', sI, j2)
            return (S1_s + (MI - jI)).reshape(Nr)

        def h(s):                                          # :1113-1132
            s = s.reshape(1, Nr)
            I = np.exp(np.dot(scale, s + s0c))
            sI = np.einsum('is,ij->isj', scale, I)
            Mjk = np.einsum('isj,ijk,itk->sjtk', sI, M3, sI)
            resid = 0
            if full_hessian > 0:
                MI = Mjk.sum(3)
                jI = np.einsum('is,itj,ij->sjt', scale, sI, j2)
                resid = np.einsum('sjt,jk->sjtk', MI - jI, np.eye(Nr)).reshape(Nr, Nr)
                if full_hessian < 1:
                    resid *= full_hessian
            return Mjk.reshape(Nr, Nr) + resid + Sinv

        def limit_step(dx, x):                             # :1136-1140
            return min(1.1 * np.min(np.abs(x / dx)), 1) * dx

        self.f, self.g, self.h = f, g, h
        s, self.status = minimize_newton(f, g, h, guess.copy(), _Backtrack(limit_step), tol=1e-7)
        self.s_MAP = s
        Dinv = h(s)
        try:
            self.chol, self.svd = scipy.linalg.cho_factor(Dinv), None          # :1150
        except np.linalg.LinAlgError:
            U, sv, Vt = scipy.linalg.svd(Dinv, full_matrices=False)
            self.chol, self.svd = None, (U, np.where(sv > 0, 1. / sv, 0), Vt)

    def Dsolve(self, b):
        if self.chol is not None:
            return scipy.linalg.cho_solve(self.chol, b)
        U, s1, Vt = self.svd
        return np.dot(Vt.T, np.multiply(np.dot(U.T, b), s1))

    @property
    def MAP(self):
        return self.s_MAP

    @property
    def power_spectrum(self):
        return self.p


# --------------------------------------------------------------------------------------
# FrankFitter power-spectrum loop                frank/radial_fitters.py:737-832
# --------------------------------------------------------------------------------------
def frank_fit(dht, M, j, alpha=1.05, p0=None, weights_smooth=1e-4, tol=1e-3, method='Normal',
              I_scale=1e5, max_iter=2000, store=False):
    """FrankFitter._fit.  Returns dict(MAP (brightness), power_spectrum, num_iterations,
    converged, [history])."""
    if p0 is None:
        p0 = 1e-15 if method == 'Normal' else 1e-35        # :709-713
    s_scale = np.log(I_scale)
    Tij = smoothing_matrix(dht, weights_smooth)

    def solve(p, guess=None, how=method):
        if how == 'Normal':
            return GaussianSolve(dht, M, j, p)
        return LogNormalSolve(dht, M, j, p, guess, s_scale)

    pI = np.ones([dht.N])
    fit = solve(pI, how='Normal')                          # :747
    pI = np.max(dht.transform(fit.MAP) ** 2)               # :749
    pI = pI * (dht.q / dht.q[0]) ** -2
    fit = solve(pI, how='Normal')                          # :752
    if method == 'LogNormal':                              # :756-763
        s = np.log(np.maximum(fit.MAP, 1e-3 * fit.MAP.max()))
        s -= s_scale
        pI = np.max(dht.transform(s) ** 2)
        pI = pI * (dht.q / dht.q[0]) ** -4
        fit = solve(pI, guess=s)
    count, p_old = 0, 0
    hist = {'power_spectrum': [], 'MAP': []}
    while (not check_convergence(pI, p_old, tol)) and count <= max_iter:   # :769-770
        p_old = pI.copy()
        pI = update_power_spectrum(dht, fit, Tij, alpha, p0)
        fit = solve(pI, guess=fit.MAP)
        if store:
            hist['power_spectrum'].append(pI)
            hist['MAP'].append(fit.MAP)
        count += 1
    MAP = fit.MAP if method == 'Normal' else np.exp(fit.MAP + s_scale)     # :390-392
    out = {'MAP': MAP, 'power_spectrum': pI, 'num_iterations': count,
           'converged': count < max_iter, 'fit': fit}
    if store:
        out['history'] = hist
    return out


# --------------------------------------------------------------------------------------
# UVDataBinner                                   frank/utilities.py:204-367
# --------------------------------------------------------------------------------------
def uv_bin_index(uv, bin_width, nbins=None):
    """Bin index arithmetic of bin_quantities (utilities.py:333-347) with the bin count of
    __init__ (utilities.py:205-213)."""
    if nbins is None:
        nbins = np.ceil(uv.max() / bin_width).astype('int')
        if nbins * bin_width < uv.max():
            nbins += 1
    bins = np.arange(nbins + 1, dtype='float64') * bin_width
    norm = 1 / bin_width
    idx = np.floor(uv * norm).astype('int32')
    idx[uv < bins[idx]] -= 1
    idx[idx == nbins] -= 1
    inc = (uv >= bins[idx + 1]) & (idx + 1 != nbins)
    idx[inc] += 1
    return idx, int(nbins), bins


def uv_bin(uv, V, weights, bin_width):
    """UVDataBinner.__init__ (utilities.py:204-264): weighted means per bin, counts, and the
    standard error of the mean (nan+0j for single-count bins).
    Empty bins carry zeros in sums and nan in error; `mask` flags them."""
    idx, nbins, bins = uv_bin_index(uv, bin_width)
    w = np.ones_like(uv) * weights
    BLOCK = 65536

    def accumulate(wts, qty):
        res = np.zeros(nbins, dtype=qty.dtype)
        for i in range(0, len(uv), BLOCK):                 # :333-361
            t = wts[i:i + BLOCK] * qty[i:i + BLOCK]
            ii = idx[i:i + BLOCK]
            if np.iscomplexobj(qty):
                res.real += np.bincount(ii, weights=t.real, minlength=nbins)
                res.imag += np.bincount(ii, weights=t.imag, minlength=nbins)
            else:
                res += np.bincount(ii, weights=t, minlength=nbins)
        return res

    counts = np.zeros(nbins, dtype='int64')
    for i in range(0, len(uv), BLOCK):
        counts += np.bincount(idx[i:i + BLOCK], minlength=nbins)
    bin_uv = accumulate(w, uv)
    bin_wgt = accumulate(w, np.ones_like(uv))
    bin_vis = accumulate(w, V)
    has = counts > 0
    bin_uv[has] /= bin_wgt[has]
    bin_vis[has] /= bin_wgt[has]
    mu = bin_vis[idx]                                      # :240
    qty = (V - mu).real ** 2
    if np.iscomplexobj(V):
        qty = qty + 1j * (V - mu).imag ** 2
    err = accumulate(w ** 2, qty)
    many = counts > 1
    err[many] /= bin_wgt[many] ** 2 * (1 - 1 / counts[many])
    bin_err = np.full(nbins, np.nan, dtype=V.dtype)
    e = np.sqrt(err.real[many])
    if np.iscomplexobj(V):
        e = e + 1.j * np.sqrt(err.imag[many])
    bin_err[many] = e
    bin_err[~has] = np.nan
    return {'idx': idx, 'nbins': nbins, 'bins': bins, 'uv': bin_uv, 'V': bin_vis, 'weights': bin_wgt,
            'counts': counts, 'error': bin_err, 'mask': ~has}


def uv_determine_bin(uv, bins, nbins, bin_width):
    """UVDataBinner.determine_uv_bin (utilities.py:267-298): bin index, -1 beyond the last edge."""
    norm = 1 / bin_width
    idx = np.floor(uv * norm).astype('int32')
    idx[uv < bins[np.clip(idx, 0, nbins)]] -= 1
    idx[uv == bins[nbins]] -= 1
    too_high = idx >= nbins
    idx[too_high] = -1
    tmp = idx[~too_high]
    inc = (uv[~too_high] >= bins[tmp + 1]) & (tmp + 1 != nbins)
    tmp[inc] += 1
    idx[~too_high] = tmp
    return idx


def estimate_weights(u, v=None, V=None, nbins=300, log=True, use_median=False):
    """estimate_weights (utilities.py:515-631) on top of uv_bin.  The reference's `np.iscomplex(V.dtype)`
    (:598) is False for any dtype object, so only the variance of the real part is ever used."""
    if V is None:
        if v is None:
            raise ValueError("The visibilities, V, must be supplied")
        V = v
        q = np.abs(u)
    elif v is not None:
        q = np.hypot(u, v)
    else:
        q = np.abs(u)
    if log:
        q = np.log(q)
        q -= q.min()
    bin_width = (q.max() - q.min()) / nbins
    b = uv_bin(q, V, np.ones_like(q), bin_width)
    counts = b['counts']
    if counts.max() == 1:
        raise ValueError("No bin contains more than one uv point, can't estimate the variance. Use fewer bins.")
    var = b['error'].real ** 2 * counts                                  # nan where counts <= 1
    if use_median:
        return np.full(len(u), 1 / np.median(var[counts > 1]))
    no_var = np.argwhere(counts == 1).reshape(-1)
    if len(no_var) > 0:
        good_var = np.argwhere(counts > 1).reshape(-1)
        loc = np.searchsorted(good_var, no_var, side='right')
        im = good_var[np.maximum(loc - 1, 0)]
        ip = good_var[np.minimum(loc, len(good_var) - 1)]
        var[no_var] = 0.5 * (var[im] + var[ip])
    bin_id = uv_determine_bin(q, b['bins'], b['nbins'], bin_width)
    assert np.all(bin_id != -1)
    return 1 / var[bin_id]


# --------------------------------------------------------------------------------------
# Synthetic workload of SURVEY.md section 8(d) / BASELINE.md section 3
# --------------------------------------------------------------------------------------
def synthetic_disc(n_vis, N, Rmax_arcsec=1.6, seed=12345, inc=30., PA=40., dRA=1e-3, dDec=-2e-3,
                   noise=True, analytic=False):
    """Gaussian-ring disc visibilities on random baselines.
    Returns u, v, V (complex), w and the DHT tables."""
    rng = np.random.default_rng(seed)
    dht = DHTTables(Rmax_arcsec / RAD_TO_ARCSEC, N)
    q = 0.98 * dht.q[-1] * np.sqrt(rng.uniform(1e-5, 1, n_vis))
    th = rng.uniform(0, 2 * np.pi, n_vis)
    ud, vd = q * np.cos(th), q * np.sin(th)
    u, v = deproject(ud, vd, inc, PA, inverse=True)
    r_as = dht.r * RAD_TO_ARCSEC
    if analytic:
        # closed-form Hankel pair A exp(-r^2/2s^2) <-> 2 pi s^2 A exp(-2 pi^2 s^2 q^2) (cost is value independent)
        s = 0.3 / RAD_TO_ARCSEC
        Vd = np.cos(inc * DEG_TO_RAD) * 2 * np.pi * s * s * 3e9 * np.exp(-2 * np.pi ** 2 * s * s * q * q)
    else:
        I = 1e10 * np.exp(-0.5 * ((r_as - 0.6) / 0.08) ** 2) + 3e9 * np.exp(-0.5 * (r_as / 0.3) ** 2)
        Vd = predict_visibilities(dht, I, q, None, 'opt_thick', inc)
    _, _, V = undo_correction(ud, vd, Vd.astype(complex), inc, PA, dRA, dDec)
    w = 1e4 * rng.uniform(0.5, 2, n_vis)
    if noise:
        V = V + (rng.standard_normal(n_vis) + 1j * rng.standard_normal(n_vis)) / np.sqrt(w)
    return u, v, V, w, dht
