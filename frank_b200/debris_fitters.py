"""Fitters for optically thin discs with a Gaussian vertical structure (frank/debris_fitters.py)."""
from frank_b200.radial_fitters import FourierBesselFitter, FrankFitter

__all__ = ['FourierBesselDebrisFitter', 'FrankDebrisFitter']


class FourierBesselDebrisFitter(FourierBesselFitter):
    """I(R, z) = I(R) exp(-z^2 / 2 H(R)^2) with known scale height H(R) / arcsec (debris_fitters.py:26-66)."""

    def __init__(self, Rmax, N, geometry, scale_height, nu=0, block_data=True, block_size=10 ** 5, verbose=True,
                 device=None):
        super(FourierBesselDebrisFitter, self).__init__(Rmax, N, geometry, nu=nu, block_data=block_data,
                                                        assume_optically_thick=False, scale_height=scale_height,
                                                        block_size=block_size, verbose=verbose, device=device)


class FrankDebrisFitter(FrankFitter):
    """FrankFitter for the vertically extended optically thin model (debris_fitters.py:68-156)."""

    def __init__(self, Rmax, N, geometry, scale_height, nu=0, block_data=True, block_size=10 ** 5, alpha=1.05,
                 p_0=None, weights_smooth=1e-4, tol=1e-3, method='Normal', I_scale=1e5, max_iter=2000,
                 check_qbounds=True, store_iteration_diagnostics=False, verbose=True, convergence_failure='raise',
                 device=None):
        super(FrankDebrisFitter, self).__init__(
            Rmax, N, geometry, nu=nu, block_data=block_data, block_size=block_size, alpha=alpha, p_0=p_0,
            weights_smooth=weights_smooth, tol=tol, method=method, I_scale=I_scale, max_iter=max_iter,
            check_qbounds=check_qbounds, store_iteration_diagnostics=store_iteration_diagnostics,
            assume_optically_thick=False, scale_height=scale_height, verbose=verbose,
            convergence_failure=convergence_failure, device=device)
