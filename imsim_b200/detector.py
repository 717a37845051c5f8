"""Detector geometry: focal plane [mm, DVCS] -> pixel affine map.

The reference asks ``lsst.afw.cameraGeom`` for this per call
(imsim/utils.py:42-78); it is constant per detector, so the B200 path carries it
as a 2x3 affine plus the normalised Jacobian of imsim/photon_ops.py:497-499.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from . import _abi


@dataclass
class DetectorGeometry:
    name: str
    A: np.ndarray  # (2,2) pixels per mm
    b: np.ndarray  # (2,) pixel coordinates of the focal-plane origin
    nx: int = 4096
    ny: int = 4004
    z_offset: float = 0.0  # detector height offset in metres (telescope_loader.py:399-405)
    xmin: int = 0
    ymin: int = 0

    def jhat(self) -> np.ndarray:
        """imsim/photon_ops.py:497-499: M @ J normalised by sqrt|det|."""
        M = np.array([[0.0, 1.0e3], [1.0e3, 0.0]])
        jac = M @ np.asarray(self.A, float)
        return jac / np.sqrt(np.abs(np.linalg.det(jac)))

    def focal_to_pixel(self, fpx, fpy):
        A = np.asarray(self.A, float)
        return A[0, 0] * fpx + A[0, 1] * fpy + self.b[0], A[1, 0] * fpx + A[1, 1] * fpy + self.b[1]

    def pixel_to_focal(self, x, y):
        Ai = np.linalg.inv(np.asarray(self.A, float))
        dx, dy = x - self.b[0], y - self.b[1]
        return Ai[0, 0] * dx + Ai[0, 1] * dy, Ai[1, 0] * dx + Ai[1, 1] * dy

    def center_focal(self):
        return self.pixel_to_focal(self.xmin + (self.nx - 1) / 2.0, self.ymin + (self.ny - 1) / 2.0)

    def to_pod(self) -> _abi.B2Detector:
        d = _abi.B2Detector()
        for k, v in enumerate(np.asarray(self.A, float).ravel()):
            d.A[k] = v
        d.b[0], d.b[1] = float(self.b[0]), float(self.b[1])
        for k, v in enumerate(self.jhat().ravel()):
            d.Jhat[k] = v
        return d


#: ITL rafts of LSSTCam (the other science rafts carry e2v CCDs)
ITL_RAFTS = {"R01", "R02", "R03", "R10", "R20", "R41", "R42", "R43"}


def lsstcam_like(det_name: str = "R22_S11") -> DetectorGeometry:
    """Synthetic LSSTCam-like science CCD: 10 micron pixels, 42.25 mm CCD pitch,
    127 mm raft pitch.  R22_S11 reproduces the reference's golden vector
    (tests/test_photon_ops.py:668-691): x = 100 fpx + 2047.5, y = 100 fpy + 2001.5."""
    rx, ry, sx, sy = int(det_name[1]), int(det_name[2]), int(det_name[5]), int(det_name[6])
    cx = (rx - 2) * 127.0 + (sx - 1) * 42.25
    cy = (ry - 2) * 127.0 + (sy - 1) * 42.25
    # e2v CCDs have 4096 x 4004 imaging pixels (16 segments of 512 x 2002), ITL ones 4072 x 4000 (509 x 2000)
    nx, ny = (4072, 4000) if det_name[:3] in ITL_RAFTS else (4096, 4004)
    A = np.array([[100.0, 0.0], [0.0, 100.0]])
    b = np.array([(nx - 1) / 2.0 - 100.0 * cx, (ny - 1) / 2.0 - 100.0 * cy])
    return DetectorGeometry(det_name, A, b, nx=nx, ny=ny)


def lsstcam_science_detectors():
    """The 189 science CCD names R10_S00 ... R34_S22 (corner rafts excluded)."""
    names = []
    for rx in range(5):
        for ry in range(5):
            if (rx, ry) in ((0, 0), (0, 4), (4, 0), (4, 4)):
                continue
            for sx in range(3):
                for sy in range(3):
                    names.append("R%d%d_S%d%d" % (rx, ry, sx, sy))
    return names
