"""Host-side set-up of the spider-diffraction kick (imsim/diffraction.py).

Only the per-exposure scalars are computed here (geometry table, pointing and
zenith vectors of imsim/diffraction.py:284-415); the per-photon arithmetic
(directed_dist, phi_star, field rotation, apply_delta_v) runs in the
``diffraction_kick`` device function of csrc/optics.cu.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _abi

# Earth rotation rate [rad/s] (imsim/diffraction.py:280)
OMEGA_EARTH = 7.292115826090781e-05
# Simonyi telescope latitude used as default by the reference
# (imsim/photon_ops.py:420, lsst.obs.lsst SIMONYI_LOCATION): -30.24463 degrees
RUBIN_LATITUDE = np.radians(-30.24463)


@dataclass
class Geometry:
    """2D geometry of thick lines [nx, ny, d, thickness] and circles [x, y, r]
    (imsim/diffraction.py:15-29)."""

    thick_lines: np.ndarray
    circles: np.ndarray


# imsim/diffraction.py:32-42
RUBIN_SPIDER_GEOMETRY = Geometry(
    thick_lines=np.array(
        [
            [1 / np.sqrt(2.0), 1 / np.sqrt(2.0), -0.4, 0.025],
            [-1 / np.sqrt(2.0), 1 / np.sqrt(2.0), -0.4, 0.025],
            [1 / np.sqrt(2.0), 1 / np.sqrt(2.0), 0.4, 0.025],
            [-1 / np.sqrt(2.0), 1 / np.sqrt(2.0), 0.4, 0.025],
        ]
    ),
    circles=np.array([[0.0, 0.0, 2.558], [0.0, 0.0, 4.18]]),
)


def e_equatorial(latitude: float, altitude: float, azimuth: float) -> np.ndarray:
    """Pointing in the equatorial frame (imsim/diffraction.py:387-415)."""
    e_zenith = np.array([np.cos(latitude), 0.0, np.sin(latitude)])
    e_east = np.array([0.0, 1.0, 0.0])
    e_north = np.array([-e_zenith[2], 0.0, e_zenith[0]])
    return (e_east * np.cos(altitude) * np.sin(azimuth) + e_north * np.cos(altitude) * np.cos(azimuth)
            + e_zenith * np.sin(altitude))


def diffraction_config(latitude=None, altitude=None, azimuth=None, disable_field_rotation=False,
                       geometry: Geometry = RUBIN_SPIDER_GEOMETRY, enabled=True) -> _abi.B2Diffraction:
    """POD for ``b2_diffraction_config`` (RubinDiffraction.__init__, imsim/photon_ops.py:233-262)."""
    c = _abi.B2Diffraction()
    c.enabled = int(enabled)
    c.field_rotation = int(enabled and not disable_field_rotation)
    lines = np.asarray(geometry.thick_lines, float)
    circles = np.asarray(geometry.circles, float)
    if len(lines) > 8 or len(circles) > 4:
        raise ValueError("geometry too large")
    c.n_lines, c.n_circles = len(lines), len(circles)
    for k, row in enumerate(lines):
        for j in range(4):
            c.lines[k][j] = row[j]
    for k, row in enumerate(circles):
        for j in range(3):
            c.circles[k][j] = row[j]
    c.omega = OMEGA_EARTH
    if c.field_rotation:
        if latitude is None or altitude is None or azimuth is None:
            raise ValueError("latitude, altitude and azimuth are required for field rotation")
        lat = float(latitude)
        c.cos_lat, c.sin_lat = np.cos(lat), np.sin(lat)
        e0 = np.array([np.cos(lat), 0.0, np.sin(lat)])  # prepare_e_z (:284-304)
        ef = e_equatorial(latitude=lat, altitude=float(altitude), azimuth=float(azimuth))
        for k in range(3):
            c.e_z_0[k] = e0[k]
            c.e_focal[k] = ef[k]
    return c
