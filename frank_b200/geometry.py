"""Source geometry: phase-centre shift and deprojection.

API mirror of frank.geometry (frank/geometry.py:41-401) for the fixed-geometry classes.  On the fit
path the per-visibility arithmetic runs on the GPU (frank_b200/csrc/fb_prep.cu); the NumPy helpers
here serve the small host-side uses (re-projection of a handful of points, `undo_correction` of
predicted visibilities) and define the host scalars handed to the device (`device_scalars`).
"""
import numpy as np

from frank_b200.constants import rad_to_arcsec, deg_to_rad

__all__ = ['apply_phase_shift', 'deproject', 'rescale_total_flux', 'SourceGeometry', 'FixedGeometry']


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
        Vp = apply_phase_shift(u, v, V, self._dRA, self._dDec, inverse=True)
        up, vp, wp = deproject(u, v, self._inc, self._PA)
        return (up, vp, wp, Vp) if use3D else (up, vp, Vp)

    def undo_correction(self, u, v, V):
        up, vp = self.reproject(u, v)
        return up, vp, apply_phase_shift(up, vp, V, self._dRA, self._dDec, inverse=False)

    def deproject(self, u, v, use3D=False):
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
