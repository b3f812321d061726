"""Back-tracking line search and Newton iteration used by the log-normal MAP model.

Same decision logic as frank.minimizer (frank/minimizer.py:24-283; the line search is the Numerical-Recipes
scheme): the control flow runs on the host on O(N) vectors, the objective / gradient / Newton direction are
supplied by callables that evaluate them on the GPU.
"""
import numpy as np

__all__ = ['LineSearch', 'MinimizeNewton']


class LineSearch(object):
    """Back-tracking line search for scalar objectives (frank/minimizer.py:44-184, root=False use).

    armijo: sufficient-decrease coefficient; min_step_frac: floor of the step reduction per trial;
    reduce_step(dx, x) -> dx limits the step before the search."""

    def __init__(self, armijo=1e-4, min_step_frac=0.1, reduce_step=None):
        self.reduction = None
        self.reduce_step = reduce_step if reduce_step is not None else (lambda dx, _: dx)
        self._armijo = armijo
        self._l_min = min_step_frac

    def __call__(self, func, grad, x0, p, f0):
        """Returns x_new, f_new, number of evaluations, failed."""
        nfev = 0
        cost = f0
        p = self.reduce_step(p, x0)
        slope = np.dot(grad, p)                              # expected first-order change
        if slope > 0:
            raise ValueError("Round off in slope calculation")
        lam, lam_prev, cost_prev = 1.0, None, None
        while True:
            x_new = x0 + lam * p
            if np.all(x_new == x0):
                return x0, f0, nfev, True
            cost_new = func(x_new)
            nfev += 1
            if cost_new <= cost + self._armijo * lam * slope:
                self.reduction = lam
                return x_new, cost_new, nfev, False
            if lam == 1.0:
                lam_new = -0.5 * slope / (cost_new - cost - slope)           # quadratic model
            else:                                                            # cubic model through the last two trials
                r1 = (cost_new - cost - lam * slope) / (lam * lam)
                r2 = (cost_prev - cost - lam_prev * slope) / (lam_prev * lam_prev)
                a = (r1 - r2) / (lam - lam_prev)
                b = (lam * r2 - lam_prev * r1) / (lam - lam_prev)
                if a == 0:
                    lam_new = -0.5 * slope / b
                else:
                    disc = b * b - 3 * a * slope
                    if disc < 0:
                        lam_new = 0.5 * lam
                    elif b <= 0:
                        lam_new = (-b + np.sqrt(disc)) / (3 * a)
                    else:
                        lam_new = -1 * slope / (b + np.sqrt(disc))
                    lam_new = min(0.5 * lam, lam_new)
            if np.isnan(lam_new):
                lam_new = self._l_min * lam
            lam_prev, cost_prev = lam, cost_new
            lam = max(lam_new, self._l_min * lam)


def MinimizeNewton(fun, jac, newton_dir, guess, line_search, max_step=10 ** 5, max_hev=1000, tol=1e-5):
    """Newton's method with back-tracking and a gradient-descent fallback (frank/minimizer.py:187-283).

    fun(x) -> f ; jac(x) -> g ; newton_dir(x, refactor) -> (g(x), -H^-1 g(x)) where H is re-evaluated and
    re-factorised at x when `refactor` is true and re-used otherwise (the reference keeps its LU factors while
    full steps are accepted, minimizer.py:276).  Returns x, (status, nstep, nfev, nhess) with status
    0 success, 1 failed to improve, 2 too many iterations, 3 too many Hessian evaluations."""
    need_hess = True
    nfev, nhess = 1, 0
    x = guess
    fx = fun(x)
    for nstep in range(max_step):
        if need_hess:
            if nhess == max_hev:
                return x, (3, nstep, nfev, nhess)
            nhess += 1
        jx, dx = newton_dir(x, need_hess)
        if dx is not None and np.dot(jx, dx) < 0:
            x, fx, fev, failed = line_search(fun, jx, x, dx, fx)
            nfev += fev
        else:
            failed = True
        if failed:                                                           # gradient descent instead
            x, fx, fev, failed_descent = line_search(fun, jx, x, -jx, fx)
            nfev += fev
            if failed_descent:                                               # last resort: shrinking steps
                dx = line_search.reduce_step(-jx, x)
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
        need_hess = failed or (line_search.reduction != 1.0)
        if (np.abs(jac(x)) * np.abs(x)).max() < tol * max(np.abs(fx), 1):
            return x, (0, nstep, nfev, nhess)
    return x, (2, max_step, nfev, nhess)
