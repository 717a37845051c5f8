"""Minimal stand-in for the parts of GalSim that imsim_b200.galsim_plugin touches -- TEST INFRASTRUCTURE.
It records registrations so that tests can check the plugin's wiring without GalSim installed."""


class Angle:
    def __init__(self, rad):
        self.rad = rad


class CelestialCoord:
    def __init__(self, ra=0.0, dec=0.0):
        self.ra, self.dec = ra, dec


class PositionD:
    def __init__(self, x=0.0, y=0.0):
        self.x, self.y = x, y


class BaseDeviate:
    def __init__(self, seed=0):
        self._seed = int(seed)

    def raw(self):
        return self._seed


class UniformDeviate:
    def __init__(self, seed=0):
        import numpy as np

        self._g = np.random.default_rng(seed.raw() if hasattr(seed, "raw") else int(seed))

    def __call__(self):
        return float(self._g.random())


class SiliconSensor:  # the plugin's sensor subclasses this for the isinstance test of photon_pooling.py:209
    def __init__(self, *a, **k):
        raise AssertionError("galsim.SiliconSensor.__init__ must not run for the B200 sensor")


from . import config  # noqa: E402,F401


# ---- what the pooled image builder of the plugin touches (round 2) -------------------------------------------------
import numpy as _np  # noqa: E402

from imsim_b200.photon_array import PhotonArray  # noqa: E402,F401  (same attribute names as galsim.PhotonArray)


class BoundsI:
    def __init__(self, xmin, xmax, ymin, ymax):
        self.xmin, self.xmax, self.ymin, self.ymax = int(xmin), int(xmax), int(ymin), int(ymax)

    def isDefined(self):
        return self.xmax >= self.xmin and self.ymax >= self.ymin

    def __and__(self, o):
        return BoundsI(max(self.xmin, o.xmin), min(self.xmax, o.xmax), max(self.ymin, o.ymin), min(self.ymax, o.ymax))

    def withBorder(self, n):
        return BoundsI(self.xmin - n, self.xmax + n, self.ymin - n, self.ymax + n)


class JacobianWCS:
    def __init__(self, dudx, dudy, dvdx, dvdy):
        self.m = _np.array([[dudx, dudy], [dvdx, dvdy]], float)

    def getMatrix(self):
        return self.m


class PixelScale:
    def __init__(self, scale):
        self.scale = float(scale)

    def local(self, pos=None):
        return JacobianWCS(self.scale, 0.0, 0.0, self.scale)

    def makeSkyImage(self, image, sky_level):
        image.array[:, :] = sky_level * self.scale ** 2


class ImageF:
    """array + integer bounds + wcs, the part of galsim.Image the builders use"""

    def __init__(self, ncol, nrow=None, xmin=1, ymin=1, wcs=None, dtype=_np.float32):
        if isinstance(ncol, BoundsI):  # galsim.ImageF(bounds, wcs=...)
            b = ncol
            ncol, nrow, xmin, ymin = b.xmax - b.xmin + 1, b.ymax - b.ymin + 1, b.xmin, b.ymin
        self.array = _np.zeros((nrow, ncol), dtype=dtype)
        self.bounds = BoundsI(xmin, xmin + ncol - 1, ymin, ymin + nrow - 1)
        self.wcs = wcs
        self.photons = None

    @property
    def dtype(self):
        return self.array.dtype

    @property
    def true_center(self):
        b = self.bounds
        return PositionD((b.xmin + b.xmax) / 2.0, (b.ymin + b.ymax) / 2.0)


class Sensor:
    """galsim.Sensor: photons binned by nominal pixel"""

    def accumulate(self, photons, image, orig_center=None, resume=False):
        b = image.bounds
        ix = _np.floor(photons.x + 0.5).astype(int) - b.xmin
        iy = _np.floor(photons.y + 0.5).astype(int) - b.ymin
        ok = (ix >= 0) & (ix <= b.xmax - b.xmin) & (iy >= 0) & (iy <= b.ymax - b.ymin)
        _np.add.at(image.array, (iy[ok], ix[ok]), photons.flux[ok])
        return float(photons.flux[ok].sum())

    def updateRNG(self, rng):
        pass


class _NumpyRng:
    @staticmethod
    def of(rng):
        seed = rng.raw() if hasattr(rng, "raw") else (0 if rng is None else int(rng))
        return _np.random.default_rng(seed)


class TimeSampler:
    def __init__(self, t0=0.0, exptime=0.0):
        self.t0, self.exptime = float(t0), float(exptime)

    def applyTo(self, photon_array, local_wcs=None, rng=None):
        photon_array.time = self.t0 + self.exptime * _NumpyRng.of(rng).random(len(photon_array))


class PupilAnnulusSampler:
    def __init__(self, R_outer, R_inner=0.0):
        self.R_outer, self.R_inner = float(R_outer), float(R_inner)

    def applyTo(self, photon_array, local_wcs=None, rng=None):
        g = _NumpyRng.of(rng)
        n = len(photon_array)
        r = _np.sqrt(g.uniform(self.R_inner**2, self.R_outer**2, n))
        phi = g.uniform(0, 2 * _np.pi, n)
        photon_array.pupil_u = r * _np.cos(phi)
        photon_array.pupil_v = r * _np.sin(phi)


class FocusDepth:
    def __init__(self, depth):
        self.depth = float(depth)

    def applyTo(self, photon_array, local_wcs=None, rng=None):
        photon_array.x = photon_array.x + photon_array.dxdz * self.depth
        photon_array.y = photon_array.y + photon_array.dydz * self.depth


class Refraction:
    def __init__(self, index_ratio):
        self.index_ratio = float(index_ratio)

    def applyTo(self, photon_array, local_wcs=None, rng=None):
        n2 = self.index_ratio**2
        f = 1.0 / _np.sqrt(n2 + (n2 - 1.0) * (photon_array.dxdz**2 + photon_array.dydz**2))
        photon_array.dxdz = photon_array.dxdz * f
        photon_array.dydz = photon_array.dydz * f
