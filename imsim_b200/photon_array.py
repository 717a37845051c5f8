"""Minimal ``galsim.PhotonArray`` stand-in (host SoA, float64).

The B200 photon ops are duck-typed: they work on a real ``galsim.PhotonArray``
(when GalSim is installed beside this package) or on this class, which carries
the same attribute names and ``hasAllocated*`` predicates the reference uses
(imsim/photon_ops.py:139-140, imsim/photon_pooling.py:186-191).
"""
from __future__ import annotations

import numpy as np

_LAZY = ("dxdz", "dydz", "wavelength", "pupil_u", "pupil_v", "time")


class PhotonArray:
    def __init__(self, N, x=None, y=None, flux=None, dxdz=None, dydz=None, wavelength=None, pupil_u=None,
                 pupil_v=None, time=None):
        self._N = int(N)
        self._x = np.zeros(self._N) if x is None else self._chk(x)
        self._y = np.zeros(self._N) if y is None else self._chk(y)
        self._flux = np.zeros(self._N) if flux is None else self._chk(flux)
        self._dxdz = None if dxdz is None else self._chk(dxdz)
        self._dydz = None if dydz is None else self._chk(dydz)
        self._wavelength = None if wavelength is None else self._chk(wavelength)
        self._pupil_u = None if pupil_u is None else self._chk(pupil_u)
        self._pupil_v = None if pupil_v is None else self._chk(pupil_v)
        self._time = None if time is None else self._chk(time)

    def _chk(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        if a.shape != (self._N,):
            raise ValueError("photon array field has wrong shape %r" % (a.shape,))
        return a

    def size(self):
        return self._N

    def __len__(self):
        return self._N

    # always-allocated fields
    x = property(lambda s: s._x, lambda s, v: s._x.__setitem__(slice(None), v))
    y = property(lambda s: s._y, lambda s, v: s._y.__setitem__(slice(None), v))
    flux = property(lambda s: s._flux, lambda s, v: s._flux.__setitem__(slice(None), v))

    def _lazy(name):  # noqa: N805
        priv = "_" + name

        def get(self):
            if getattr(self, priv) is None:
                setattr(self, priv, np.zeros(self._N))
            return getattr(self, priv)

        def set_(self, v):
            get(self)[:] = v

        return property(get, set_)

    dxdz = _lazy("dxdz")
    dydz = _lazy("dydz")
    wavelength = _lazy("wavelength")
    pupil_u = _lazy("pupil_u")
    pupil_v = _lazy("pupil_v")
    time = _lazy("time")
    del _lazy

    def hasAllocatedAngles(self):
        return self._dxdz is not None and self._dydz is not None

    def hasAllocatedWavelengths(self):
        return self._wavelength is not None

    def hasAllocatedPupil(self):
        return self._pupil_u is not None and self._pupil_v is not None

    def hasAllocatedTimes(self):
        return self._time is not None

    def copyFrom(self, rhs, target_indices=slice(None), source_indices=slice(None)):
        """galsim.PhotonArray.copyFrom: copy every field ``rhs`` has allocated."""
        self._x[target_indices] = rhs.x[source_indices]
        self._y[target_indices] = rhs.y[source_indices]
        self._flux[target_indices] = rhs.flux[source_indices]
        if rhs.hasAllocatedAngles():
            self.dxdz[target_indices] = rhs.dxdz[source_indices]
            self.dydz[target_indices] = rhs.dydz[source_indices]
        if rhs.hasAllocatedWavelengths():
            self.wavelength[target_indices] = rhs.wavelength[source_indices]
        if rhs.hasAllocatedPupil():
            self.pupil_u[target_indices] = rhs.pupil_u[source_indices]
            self.pupil_v[target_indices] = rhs.pupil_v[source_indices]
        if rhs.hasAllocatedTimes():
            self.time[target_indices] = rhs.time[source_indices]
        return self


def field(pa, name):
    """Contiguous float64 view of a PhotonArray field (GalSim's or ours)."""
    a = getattr(pa, name)
    if not (isinstance(a, np.ndarray) and a.dtype == np.float64 and a.flags.c_contiguous):
        raise TypeError("PhotonArray.%s must be a contiguous float64 array" % name)
    return a
