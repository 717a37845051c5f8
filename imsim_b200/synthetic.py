"""Synthetic per-detector set-ups (telescope + WCS pair + detector geometry) for
benchmarks, smoke tests and parity tests: the stand-in for what
imsim/telescope_loader.py and imsim/batoid_wcs.py hand to the photon ops at run
time.  SYNTHETIC DATA around the LSST v3.3 design of ``telescope.lsst_v33``.

The image WCS is *fitted to chief-ray traces through the same telescope*, as the
reference does (imsim/batoid_wcs.py:408-453), so photons land within a few
pixels of where the WCS says their sky position is.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Optional

import numpy as np

from .detector import DetectorGeometry, lsstcam_like
from .telescope import UNIT_AIR, Telescope, rubin_like
from .wcs import TanSipWCS, field_wcs, fit_tan_sip, tan_deproject

# tracer(telescope, thx, thy, wavelength_m) -> (x, y) on the detector plane [m],
# for chief rays (through the centre of the stop) at field angles thx, thy [rad]
Tracer = Callable[[Telescope, np.ndarray, np.ndarray, float], tuple]


@dataclass
class DetectorSetup:
    det_name: str
    telescope: Telescope
    detector: DetectorGeometry
    img_wcs: TanSipWCS
    icrf_to_field: TanSipWCS
    boresight: tuple
    wavelength_nm: float


def chief_ray_inputs(thx, thy, wavelength_m, medium=None):
    """Stop-plane rays for ``batoid.RayVector.fromFieldAngles``-like chief rays; ``medium`` is the
    telescope's ``in_medium`` (|v| = 1/n, imsim/photon_ops.py:144-147)."""
    thx, thy = np.asarray(thx, float), np.asarray(thy, float)
    n = thx.size
    g = 1.0 / np.sqrt(1.0 + thx * thx + thy * thy)
    nair = float((UNIT_AIR if medium is None else medium).n(wavelength_m))
    z = np.zeros(n)
    return (z.copy(), z.copy(), z.copy(), thx * g / nair, thy * g / nair, -g / nair, z.copy(),
            np.full(n, wavelength_m))


def gpu_tracer(ctx) -> Tracer:
    """Chief-ray tracer running on the device context (product path)."""

    def trace(tel, thx, thy, wl):
        ctx.set_telescope(tel)
        x, y, z, vx, vy, vz, t, w = (np.ascontiguousarray(a) for a in chief_ray_inputs(thx, thy, wl, tel.in_medium))
        vig = np.zeros(x.size, np.uint8)
        fail = np.zeros(x.size, np.uint8)
        ctx.trace_rays(x, y, z, vx, vy, vz, t, w, vig, fail)
        return x, y

    return trace


def _field_of_focal(tracer, tel, fpx_mm, fpy_mm, wl):
    """Invert field angle -> focal plane for one point by Newton iterations
    (role of BatoidWCSFactory._focal_to_field, imsim/batoid_wcs.py:374-398)."""
    th = np.zeros(2)
    h = 1e-5
    for _ in range(8):
        thx = np.array([th[0], th[0] + h, th[0]])
        thy = np.array([th[1], th[1], th[1] + h])
        x, y = tracer(tel, thx, thy, wl)
        fx, fy = y * 1e3, x * 1e3  # EDCS -> DVCS transpose
        r = np.array([fx[0] - fpx_mm, fy[0] - fpy_mm])
        J = np.array([[(fx[1] - fx[0]) / h, (fx[2] - fx[0]) / h], [(fy[1] - fy[0]) / h, (fy[2] - fy[0]) / h]])
        step = np.linalg.solve(J, r)
        th = th - step
        if np.abs(step).max() < 1e-13:
            break
    return th


def make_detector_setup(tracer: Tracer, det_name: str = "R22_S11", band: str = "r", rot_tel_pos: float = 0.0,
                        boresight=(0.3, -0.5), rot_sky_pos: float = 0.4, wavelength_nm: float = 622.0,
                        field_distortion: float = 1e-4, order: int = 3,
                        detector: Optional[DetectorGeometry] = None) -> DetectorSetup:
    det = detector if detector is not None else lsstcam_like(det_name)
    tel = rubin_like(band, rot_tel_pos=rot_tel_pos, detector_z_offset=det.z_offset)
    wl = wavelength_nm * 1e-9
    fwcs = field_wcs(boresight, rot_sky_pos, distortion=field_distortion, seed=17)
    # field angle of the detector centre, then a hexapolar grid of 0.16 deg radius around it
    cfx, cfy = det.center_focal()
    th0 = _field_of_focal(tracer, tel, cfx, cfy, wl)
    rs, ths = [0.0], [0.0]
    nrings = 5
    for r in np.linspace(0.01, 0.16, nrings):
        nth = (int(r / 0.16 * 6 * nrings) // 6 + 1) * 6
        rs.extend([r] * nth)
        ths.extend([i / nth * 2 * np.pi for i in range(nth)])
    thxs = th0[0] + np.deg2rad(np.array(rs) * np.cos(ths))
    thys = th0[1] + np.deg2rad(np.array(rs) * np.sin(ths))
    x, y = tracer(tel, thxs, thys, wl)
    px, py = det.focal_to_pixel(y * 1e3, x * 1e3)
    # field -> ICRF through the field WCS (forward TAN-SIP evaluated on the host: set-up only)
    ra, dec = _tansip_forward_host(fwcs, thxs, thys)
    iwcs = fit_tan_sip(px, py, ra, dec, order=order)
    return DetectorSetup(det_name, tel, det, iwcs, fwcs, tuple(boresight), wavelength_nm)


def _tansip_forward_host(w: TanSipWCS, x, y):
    u, v = x - w.crpix[0], y - w.crpix[1]
    if w.order > 0:
        f = np.zeros_like(u)
        g = np.zeros_like(u)
        for i in range(w.order + 1):
            for j in range(w.order + 1 - i):
                f = f + w.ab[0, i, j] * u**i * v**j
                g = g + w.ab[1, i, j] * u**i * v**j
    else:
        f, g = u, v
    xi = np.radians(w.cd[0, 0] * f + w.cd[0, 1] * g)
    eta = np.radians(w.cd[1, 0] * f + w.cd[1, 1] * g)
    return tan_deproject(xi, eta, *w.center)


def synthetic_photons(n: int, nx: int = 4096, ny: int = 4004, seed: int = 0, kind: str = "uniform",
                      wavelength=(550.0, 690.0), n_stars: int = 1000, sigma_px: float = 1.5, xmin=0, ymin=0):
    """Host arrays (x, y, wavelength, flux) of a synthetic pool in pixel coordinates.

    kind='uniform': flat illumination;  kind='stars': ``n_stars`` Gaussian spots of
    ``sigma_px`` with a bright-end weighted flux function (bright-star dominated)."""
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        x = rng.uniform(xmin - 0.5, xmin + nx - 0.5, n)
        y = rng.uniform(ymin - 0.5, ymin + ny - 0.5, n)
    elif kind == "stars":
        cx = rng.uniform(xmin + 50, xmin + nx - 50, n_stars)
        cy = rng.uniform(ymin + 50, ymin + ny - 50, n_stars)
        w = 10 ** (0.4 * rng.uniform(0, 5, n_stars))  # 5 magnitudes of dynamic range
        idx = rng.choice(n_stars, size=n, p=w / w.sum())
        x = cx[idx] + sigma_px * rng.standard_normal(n)
        y = cy[idx] + sigma_px * rng.standard_normal(n)
    else:
        raise ValueError(kind)
    wl = rng.uniform(wavelength[0], wavelength[1], n)
    return x, y, wl, np.ones(n)
