"""PhotonOp classes for the B200 path (mirror of imsim/photon_ops.py).

Same class names, constructor arguments, ``_req_params`` / ``_opt_params`` and
``applyTo(photon_array, local_wcs=None, rng=None)`` contract as the reference,
so a GalSim config that names ``RubinOptics`` / ``RubinDiffraction`` /
``RubinDiffractionOptics`` runs unchanged once this module is listed in
``modules:`` instead of (or after) ``imsim`` -- see INTEGRATION.md.

The per-photon chain  xy -> v -> n_air -> spider kick -> ray trace -> pixels
(imsim/photon_ops.py:81-127,136-148,274-302,469-503) is ONE kernel launch
(``k_rubin_optics`` in csrc/optics.cu).  Host code only validates, uploads the
per-detector description once, and passes array pointers.

``telescope`` / ``img_wcs`` / ``icrf_to_field`` / ``camera[det_name]`` may be
the live reference objects (batoid.Optic, galsim.GSFitsWCS,
lsst.afw.cameraGeom.Detector) or this package's flat descriptions
(:class:`Telescope`, :class:`TanSipWCS`, :class:`DetectorGeometry`).
"""
from __future__ import annotations

from typing import Optional

import numpy as np

from . import _abi, _lib
from .context import OpticsContext
from .detector import DetectorGeometry
from .diffraction import RUBIN_LATITUDE, RUBIN_SPIDER_GEOMETRY, diffraction_config
from .photon_array import field
from .telescope import Telescope
from .wcs import TanSipWCS


def _as_telescope(t):
    if isinstance(t, Telescope):
        return t
    from .extract import telescope_from_batoid  # live batoid.Optic

    return telescope_from_batoid(t)


def _as_wcs(w):
    if isinstance(w, TanSipWCS):
        return w
    from .extract import tansip_from_galsim

    return tansip_from_galsim(w)


def _as_detector(d):
    if isinstance(d, DetectorGeometry):
        return d
    from .extract import detector_from_lsst

    return detector_from_lsst(d)


def _xy(p):
    if p is None:
        return None
    return (float(p.x), float(p.y)) if hasattr(p, "x") else (float(p[0]), float(p[1]))


def _seed_offset(rng):
    """64-bit Philox key from a GalSim deviate / int / None (None: fixed key 0)."""
    if rng is None:
        return 0
    if isinstance(rng, (int, np.integer)):
        return int(rng) & 0xFFFFFFFFFFFFFFFF
    if hasattr(rng, "raw"):
        return ((int(rng.raw()) << 32) | int(rng.raw())) & 0xFFFFFFFFFFFFFFFF
    raise TypeError("rng must be None, an int or a galsim.BaseDeviate")


#: device index -> the context shared by the photon ops of this process on that device
_SHARED_CONTEXTS: dict = {}


class _DeviceOp:
    """Shared plumbing: lazily creates the per-detector device context."""

    device = 0

    def _context(self) -> OpticsContext:
        """The device context of this op's GPU, holding this op's telescope / WCS / detector / spider set-up.

        One context per (process, device) is shared by all ops -- GalSim rebuilds the op objects for every image
        (and every stamp in the classic pipeline), and a context of its own for each would also force the sensor,
        which must live on the optics' context for the fused pooled step, to re-create its 2.5 GB of boundary
        arrays per image.  The context remembers which op's description it holds; another op re-uploads its own
        (a few kB of structs, no device work)."""
        ctx = _SHARED_CONTEXTS.get(self.device)
        if ctx is None:
            ctx = _SHARED_CONTEXTS[self.device] = OpticsContext(device=self.device)
        if getattr(ctx, "_described_by", None) is not self:
            ctx.set_telescope(_as_telescope(self.telescope))
            ctx.set_wcs(_as_wcs(self.img_wcs), _as_wcs(self.icrf_to_field))
            det = getattr(self, "detector", None)
            if det is not None:
                ctx.set_detector(_as_detector(det))
            ctx.set_diffraction(self._diffraction_pod())
            ctx._described_by = self
        self._ctx = ctx
        return ctx

    def _diffraction_pod(self):
        return None

    def _options(self, rng) -> _abi.B2OpticsOptions:
        opt = _abi.B2OpticsOptions()
        sc = _xy(self.stamp_center)
        opt.shift_in = int(bool(self.shift_photons) and sc is not None)
        opt.shift_out = int(sc is not None)
        if sc is not None:
            opt.stamp_center[0], opt.stamp_center[1] = sc
        opt.seed = _seed_offset(rng)
        opt.photon_offset = 0
        return opt


class RubinOptics(_DeviceOp):
    """Ray-trace photons through the Rubin optics (imsim/photon_ops.py:24-133).

    Parameters
    ----------
    telescope : batoid.Optic or imsim_b200.Telescope
    boresight : galsim.CelestialCoord (kept for interface parity; unused like in the reference)
    img_wcs : galsim.BaseWCS or TanSipWCS
    stamp_center : galsim.PositionD, (x, y) or None
    icrf_to_field : galsim.GSFitsWCS or TanSipWCS
    det_name : str
    camera : mapping det_name -> detector (lsst camera or dict of DetectorGeometry)
    shift_photons : bool, whether to shift photons at start. [default: False]
    """

    # type names mirror imsim/photon_ops.py:43-51; resolved against galsim in galsim_plugin.py
    _req_params = {"boresight": "CelestialCoord", "camera": str, "det_name": str}
    _opt_params = {"shift_photons": bool}

    def __init__(self, telescope, boresight, img_wcs, stamp_center, icrf_to_field, det_name, camera,
                 shift_photons=False, device: int = 0):
        self.telescope = telescope
        self.detector = camera[det_name]
        self.boresight = boresight
        self.img_wcs = img_wcs
        self.stamp_center = stamp_center
        self.icrf_to_field = icrf_to_field
        self.shift_photons = shift_photons
        self.device = device
        self.last_stats: Optional[_abi.B2OpticsStats] = None

    def photon_velocity(self, photon_array, rng=None) -> np.ndarray:
        """Velocity of the photons entering the pupil, shape (n, 3)
        (imsim/photon_ops.py:72-79,136-148)."""
        return photon_velocity(photon_array, XyToV(self.icrf_to_field, self.img_wcs, _ctx=self._context()),
                               self._get_n)

    def _get_n(self, wavelength_m):
        return _as_telescope(self.telescope).in_medium.n(wavelength_m)

    def applyTo(self, photon_array, local_wcs=None, rng=None, gauss=None):
        """Apply the photon operator to a PhotonArray, in place
        (imsim/photon_ops.py:81-127).  Pupil positions and arrival times must
        already be sampled.  ``gauss`` (testing) injects the per-photon standard
        normal draws the diffraction kick consumes."""
        assert photon_array.hasAllocatedPupil()
        assert photon_array.hasAllocatedTimes()
        ctx = self._context()
        opt = self._options(rng)
        x, y, flux = field(photon_array, "x"), field(photon_array, "y"), field(photon_array, "flux")
        dxdz, dydz = field(photon_array, "dxdz"), field(photon_array, "dydz")
        if gauss is not None:
            gauss = np.ascontiguousarray(gauss, dtype=np.float64)
        stats = ctx.rubin_optics(x, y, dxdz, dydz, flux, field(photon_array, "wavelength"),
                                 field(photon_array, "pupil_u"), field(photon_array, "pupil_v"),
                                 field(photon_array, "time"), gauss=gauss, options=opt)
        self.last_stats = stats
        # imsim/photon_ops.py:493-494 asserts this for all non-vignetted rays
        assert stats.n_offdetector_z == 0, "%d rays did not end on the detector plane" % stats.n_offdetector_z

    def __str__(self):
        return f"imsim.{type(self).__name__}()"

    def __repr__(self):
        return str(self)


def photon_velocity(photon_array, xy_to_v: "XyToV", get_n) -> np.ndarray:
    """Velocity of a photon array (imsim/photon_ops.py:136-148)."""
    assert photon_array.hasAllocatedPupil()
    assert photon_array.hasAllocatedTimes()
    v = xy_to_v(photon_array.x, photon_array.y)
    wavelength = photon_array.wavelength * 1e-9
    n = get_n(wavelength)
    v /= n[:, None]
    return v


class RubinDiffraction(_DeviceOp):
    """Statistical diffraction by the Rubin spider (imsim/photon_ops.py:211-358).

    Parameters: telescope, latitude, altitude, azimuth [rad], img_wcs, icrf_to_field,
    disable_field_rotation, stamp_center, shift_photons -- as in the reference.
    """

    _req_params = {"altitude": "Angle", "azimuth": "Angle", "latitude": "Angle"}
    _opt_params = {"disable_field_rotation": bool, "stamp_center": "PositionD", "shift_photons": bool}

    def __init__(self, telescope, latitude, altitude, azimuth, img_wcs, icrf_to_field,
                 disable_field_rotation: bool = False, stamp_center=None, shift_photons=False, device: int = 0):
        self.telescope = telescope
        self.img_wcs = img_wcs
        self.icrf_to_field = icrf_to_field
        self.stamp_center = stamp_center
        self.shift_photons = shift_photons
        self.latitude = _rad(latitude)
        self.altitude = _rad(altitude)
        self.azimuth = _rad(azimuth)
        self.disable_field_rotation = bool(disable_field_rotation)
        self.device = device

    def _diffraction_pod(self):
        return diffraction_config(latitude=self.latitude, altitude=self.altitude, azimuth=self.azimuth,
                                  disable_field_rotation=self.disable_field_rotation,
                                  geometry=RUBIN_SPIDER_GEOMETRY)

    def applyTo(self, photon_array, local_wcs=None, rng=None, gauss=None):
        """Kick the photons' incoming directions and map them back to pixel
        positions, in place (imsim/photon_ops.py:304-352)."""
        assert photon_array.hasAllocatedPupil()
        assert photon_array.hasAllocatedTimes()
        ctx = self._context()
        opt = self._options(rng)
        # RubinDiffraction shifts symmetrically (photon_ops.py:323-325,350-352)
        opt.shift_out = opt.shift_in
        if gauss is not None:
            gauss = np.ascontiguousarray(gauss, dtype=np.float64)
        ctx.rubin_diffraction(field(photon_array, "x"), field(photon_array, "y"), field(photon_array, "wavelength"),
                              field(photon_array, "pupil_u"), field(photon_array, "pupil_v"),
                              field(photon_array, "time"), gauss=gauss, options=opt)

    def __str__(self):
        return f"imsim.{type(self).__name__}()"

    def __repr__(self):
        return str(self)


class RubinDiffractionOptics(RubinOptics):
    """RubinDiffraction followed by RubinOptics without undoing the xy -> v
    transform in between (imsim/photon_ops.py:151-208)."""

    # verbatim from imsim/photon_ops.py:173-186 (altitude/azimuth listed twice there too)
    _req_params = {"boresight": "CelestialCoord", "camera": str, "det_name": str, "altitude": "Angle",
                   "azimuth": "Angle"}
    _opt_params = {"altitude": "Angle", "azimuth": "Angle", "latitude": "Angle", "disable_field_rotation": bool,
                   "shift_photons": bool}

    def __init__(self, telescope, boresight, stamp_center, det_name, camera, rubin_diffraction: RubinDiffraction,
                 shift_photons=False, device: int = 0):
        super().__init__(telescope, boresight, rubin_diffraction.img_wcs, stamp_center,
                         rubin_diffraction.icrf_to_field, det_name, camera, shift_photons, device=device)
        self.rubin_diffraction = rubin_diffraction

    def _diffraction_pod(self):
        return self.rubin_diffraction._diffraction_pod()


def _rad(angle) -> float:
    """radians of a float or a coord.Angle"""
    return float(angle.rad) if hasattr(angle, "rad") else float(angle)


class XyToV:
    """Maps image coordinates (x, y) to the 3d direction of the photons before they
    enter the telescope: (x,y) -> (ra,dec) -> (thx,thy) -> v
    (imsim/photon_ops.py:454-483).  Takes 2 vectors of shape (n,), returns (n, 3)."""

    def __init__(self, icrf_to_field, img_wcs, _ctx: OpticsContext = None, device: int = 0):
        self.icrf_to_field = icrf_to_field
        self.img_wcs = img_wcs
        if _ctx is None:
            _ctx = OpticsContext(device=device)
            _ctx.set_wcs(_as_wcs(img_wcs), _as_wcs(icrf_to_field))
        self._ctx = _ctx

    def __call__(self, x: np.ndarray, y: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.ascontiguousarray(y, dtype=np.float64)
        vx, vy, vz = self._ctx.xy_to_v(x, y)
        return np.array([vx, vy, vz]).T

    def inverse(self, v_photon: np.ndarray) -> tuple:
        v = np.asarray(v_photon, dtype=np.float64)
        return self._ctx.v_to_xy(np.ascontiguousarray(v[:, 0]), np.ascontiguousarray(v[:, 1]),
                                 np.ascontiguousarray(v[:, 2]))


def ray_vector_to_photon_array(ray_vector, detector, out):
    """Convert traced rays (``x, y, z, vx, vy, vz, vignetted`` attributes, detector
    frame, metres) to pixel coordinates and slopes, stored into ``out``
    (imsim/photon_ops.py:486-503).  Host version of the kernel epilogue, kept for
    interface parity and the reference's golden-vector test
    (tests/test_photon_ops.py:668-691); O(n) set-up arithmetic on already
    traced rays, not on the bench path."""
    det = _as_detector(detector)
    w = ~np.asarray(ray_vector.vignetted, dtype=bool)
    assert all(np.abs(np.asarray(ray_vector.z)[w]) < 1.0e-15)
    out.x, out.y = det.focal_to_pixel(np.asarray(ray_vector.y) * 1e3, np.asarray(ray_vector.x) * 1e3)
    jac = det.jhat()
    d = jac @ np.array([ray_vector.vx, ray_vector.vy]) / np.asarray(ray_vector.vz)
    out.dxdz, out.dydz = d[0].ravel(), d[1].ravel()
    out.flux[np.asarray(ray_vector.vignetted, dtype=bool)] = 0.0
    return out


class BandpassRatio:
    """Reweight photon fluxes from an initial to a target bandpass
    (imsim/photon_ops.py:506-521).  ``ratio`` is any callable of wavelength [nm]
    (``target / initial`` for GalSim Bandpasses)."""

    def __init__(self, target_bandpass, initial_bandpass):
        self.target = target_bandpass
        self.initial = initial_bandpass
        self.ratio = self.target / self.initial

    def applyTo(self, photon_array, local_wcs=None, rng=None):
        photon_array.flux *= self.ratio(photon_array.wavelength)


def make_rubin_diffraction_optics(telescope, boresight, img_wcs, icrf_to_field, det_name, camera, altitude, azimuth,
                                  latitude=RUBIN_LATITUDE, disable_field_rotation=False, stamp_center=None,
                                  shift_photons=False, device=0) -> RubinDiffractionOptics:
    """What ``deserialize_rubin_diffraction_optics`` builds (imsim/photon_ops.py:414-433)."""
    rd = RubinDiffraction(telescope=telescope, latitude=latitude, altitude=altitude, azimuth=azimuth,
                          img_wcs=img_wcs, icrf_to_field=icrf_to_field,
                          disable_field_rotation=disable_field_rotation, device=device)
    return RubinDiffractionOptics(telescope=telescope, boresight=boresight, stamp_center=stamp_center,
                                  det_name=det_name, camera=camera, rubin_diffraction=rd,
                                  shift_photons=shift_photons, device=device)


def air_refractive_index_minus_one(wave_nm, pressure=69.328, temperature=293.15, H2O_pressure=1.067):
    """galsim.dcr.air_refractive_index_minus_one (Filippenko 1982); host helper for set-up."""
    P = pressure * 7.50061683
    T = temperature - 273.15
    W = H2O_pressure * 7.50061683
    sigma_squared = 1.0 / (np.asarray(wave_nm, float) * 1.e-3) ** 2.0
    n_minus_one = (64.328 + (29498.1 / (146.0 - sigma_squared)) + (255.4 / (41.0 - sigma_squared))) * 1.e-6
    n_minus_one *= P * (1.0 + (1.049 - 0.0157 * T) * 1.e-6 * P) / (720.883 * (1.0 + 0.003661 * T))
    n_minus_one -= (0.0624 - 0.000680 * sigma_squared) / (1.0 + 0.003661 * T) * W * 1.e-6
    return n_minus_one


def get_refraction(wave_nm, zenith_angle, **kwargs):
    """galsim.dcr.get_refraction [radians]."""
    nm1 = air_refractive_index_minus_one(wave_nm, **kwargs)
    r0 = nm1 * (nm1 + 2) / 2.0 / (nm1**2 + 2 * nm1 + 1)
    return r0 * np.tan(zenith_angle)


def set_dcr_options(opt: _abi.B2OpticsOptions, base_wavelength, zenith_angle, parallactic_angle, jacobian,
                    center=(0.0, 0.0), alpha=0.0, scale_unit_rad=np.pi / (180.0 * 3600.0), pressure=69.328,
                    temperature=293.15, H2O_pressure=1.067):
    """Arm the fused ``galsim.PhotonDCR`` prologue of ``b2_rubin_optics``
    (config/imsim-config.yaml:290-296).  ``jacobian``: local WCS [[dudx, dudy], [dvdx, dvdy]]
    in ``scale_unit`` (arcsec) per pixel; the op moves photons by J^-1 (-s sin q, s cos q) with
    s = refraction(w) - refraction(base) expressed in ``scale_unit``."""
    J = np.asarray(jacobian, float).reshape(2, 2)
    Ji = np.linalg.inv(J)
    d = np.array([-np.sin(parallactic_angle), np.cos(parallactic_angle)]) / scale_unit_rad
    m = Ji @ d
    opt.do_dcr = 1
    opt.dcr_base_wavelength = float(base_wavelength)
    opt.dcr_alpha = float(alpha)
    opt.dcr_center[0], opt.dcr_center[1] = float(center[0]), float(center[1])
    opt.dcr_base_refraction = float(get_refraction(base_wavelength, zenith_angle, pressure=pressure,
                                                   temperature=temperature, H2O_pressure=H2O_pressure))
    opt.dcr_tanz = float(np.tan(zenith_angle))
    opt.dcr_pth[0], opt.dcr_pth[1], opt.dcr_pth[2] = pressure, temperature, H2O_pressure
    opt.dcr_m[0], opt.dcr_m[1] = float(m[0]), float(m[1])
    return opt
