"""Optical path differences of the ray trace on a pupil grid, and their annular Zernike fit.

The numerical content of imSim's ``opd`` extra output (/root/reference/imsim/opd.py:138-194, which calls
``batoid.analysis.wavefront`` / ``zernike``): a square grid of parallel rays over the entrance pupil is traced
to the detector, carried on to a reference sphere centred on the chief ray's image point, and the optical path
of every ray is compared with the chief ray's.  The reference's tests use it to pin the telescope model against
a Zemax wavefront (tests/test_opd.py:16-95); here it pins the trace kernel the same way.

The tracer is a callable ``trace(tel, x, y, z, vx, vy, vz, t, wavelength_m) -> (x, y, z, vx, vy, vz, t,
vignetted, failed)`` with positions relative to the stop -- ``OpticsContext.trace_rays`` on the device.
"""
from __future__ import annotations

from typing import Callable, Tuple

import numpy as np
from numpy.polynomial import polynomial as _P

from .telescope import LSST_PUPIL_OBSCURATION, LSST_PUPIL_SIZE, LSST_SPHERE_RADIUS, Telescope


def field_to_dircos(thx: float, thy: float, projection: str = "postel") -> np.ndarray:
    """Direction of travel of the incoming beam for a field angle (batoid ``fieldToDirCos`` with the
    sign of a ray heading down to the primary).  ``zemax``: tangents add; ``gnomonic``: tangent plane;
    ``postel``: azimuthal equidistant."""
    if projection == "zemax":
        tx, ty = np.tan(thx), np.tan(thy)
        n = np.sqrt(1.0 + tx * tx + ty * ty)
        return np.array([tx / n, ty / n, -1.0 / n])
    if projection == "gnomonic":
        n = np.sqrt(1.0 + thx * thx + thy * thy)
        return np.array([thx / n, thy / n, -1.0 / n])
    if projection == "postel":
        rho = np.hypot(thx, thy)
        sinc = np.sin(rho) / rho if rho > 0 else 1.0
        return np.array([thx * sinc, thy * sinc, -np.cos(rho)])
    raise ValueError("unknown projection %r" % projection)


def pupil_grid(nx: int, pupil_size: float) -> np.ndarray:
    """1-d sample positions of the pupil grid: ``nx`` odd puts both edges on the grid, ``nx`` even keeps
    the sample at 0 (the layout whose Fourier transform needs no phase ramp)."""
    if nx % 2:
        return np.linspace(-pupil_size / 2, pupil_size / 2, nx)
    d = pupil_size / (nx - 2)
    return (np.arange(nx) - nx // 2) * d


def wavefront(trace: Callable, tel: Telescope, thx: float, thy: float, wavelength_m: float, nx: int = 255,
              projection: str = "postel", sphere_radius: float = LSST_SPHERE_RADIUS,
              pupil_size: float = LSST_PUPIL_SIZE) -> Tuple[np.ndarray, np.ndarray]:
    """OPD map [waves] referenced to the chief ray, NaN where vignetted, and the grid coordinates [m].

    Row index = pupil y, column index = pupil x (stop coordinates)."""
    d = field_to_dircos(thx, thy, projection)
    n_in = float(tel.in_medium.n(wavelength_m))
    g = pupil_grid(nx, pupil_size)
    X, Y = np.meshgrid(g, g)
    x, y = X.ravel().copy(), Y.ravel().copy()
    n = x.size
    z = np.zeros(n)
    # all rays start on one wavefront: a ray through (x, y) of the stop plane is ahead by n (r . d)
    t = n_in * (x * d[0] + y * d[1])
    v = [np.full(n, d[k] / n_in) for k in range(3)]
    xo, yo, zo, vx, vy, vz, to, vig, fail = trace(tel, x, y, z, v[0], v[1], v[2], t, wavelength_m)
    chief = (nx // 2) * nx + nx // 2
    r = np.stack([xo - xo[chief], yo - yo[chief], zo - zo[chief]], axis=1)
    vel = np.stack([vx, vy, vz], axis=1)
    speed = np.linalg.norm(vel, axis=1)
    u = vel / speed[:, None]
    # forward intersection with the sphere of radius sphere_radius about the chief ray's image point
    b = np.einsum("ij,ij->i", r, u)
    c = np.einsum("ij,ij->i", r, r) - sphere_radius * sphere_radius
    s = -b + np.sqrt(b * b - c)
    tt = to + s / speed  # optical path: distance times the index, |v| = 1/n
    w = (tt[chief] - tt) / wavelength_m
    w[(vig != 0) | (fail != 0)] = np.nan
    return w.reshape(nx, nx), g


def noll_to_nm(j: int) -> Tuple[int, int]:
    """Noll index -> (n, m); m > 0 for the cosine term (even j), m < 0 for the sine term (odd j)."""
    n = 0
    while (n + 1) * (n + 2) // 2 < j:
        n += 1
    k = j - n * (n + 1) // 2
    ms = []
    for m in range(n % 2, n + 1, 2):
        ms += [0] if m == 0 else [m, m]
    m = ms[k - 1]
    return (n, 0) if m == 0 else (n, m if j % 2 == 0 else -m)


def _annular_radial(eps: float, nmax: int):
    """Radial polynomials orthogonal over eps <= rho <= 1 with weight rho (Gram-Schmidt on rho^m, rho^(m+2), ...),
    normalised so that sqrt(n+1) R (times sqrt(2) cos / sin) has unit mean square over the annulus."""

    def inner(a, b):
        prim = _P.polyint(_P.polymul(_P.polymul(a, b), [0.0, 1.0]))
        return _P.polyval(1.0, prim) - _P.polyval(eps, prim)

    out = {}
    for m in range(nmax + 1):
        basis = []
        for n in range(m, nmax + 1, 2):
            p = np.zeros(n + 1)
            p[n] = 1.0
            for q in basis:
                p = _P.polysub(p, inner(p, q) / inner(q, q) * q)
            p = p / np.sqrt(inner(p, p) * 2.0 / (1.0 - eps * eps) * (n + 1))
            if _P.polyval(1.0, p) < 0:
                p = -p
            basis.append(p)
            out[(n, m)] = p
    return out


def annular_zernike_basis(jmax: int, x, y, r_outer: float, r_inner: float) -> np.ndarray:
    """Noll-indexed annular Zernike polynomials Z_1..Z_jmax at (x, y); row 0 is unused (zeros)."""
    x, y = np.asarray(x, float), np.asarray(y, float)
    rho, th = np.hypot(x, y) / r_outer, np.arctan2(y, x)
    rad = _annular_radial(r_inner / r_outer, noll_to_nm(jmax)[0])
    B = np.zeros((jmax + 1,) + x.shape)
    for j in range(1, jmax + 1):
        n, m = noll_to_nm(j)
        R = _P.polyval(rho, rad[(n, abs(m))])
        if m == 0:
            B[j] = np.sqrt(n + 1.0) * R
        elif m > 0:
            B[j] = np.sqrt(2.0 * (n + 1)) * R * np.cos(m * th)
        else:
            B[j] = np.sqrt(2.0 * (n + 1)) * R * np.sin(-m * th)
    return B


def annular_zernikes(opd: np.ndarray, grid: np.ndarray, jmax: int = 28, pupil_size: float = LSST_PUPIL_SIZE,
                     eps: float = LSST_PUPIL_OBSCURATION) -> np.ndarray:
    """Least-squares annular Zernike coefficients (index 1..jmax; element 0 is 0) of an OPD map over its
    unvignetted points (the ``AZ_nnn`` header values of imsim/opd.py:178-190)."""
    X, Y = np.meshgrid(grid, grid)
    ok = ~np.isnan(opd)
    B = annular_zernike_basis(jmax, X[ok], Y[ok], pupil_size / 2, eps * pupil_size / 2)
    c, *_ = np.linalg.lstsq(B[1:].T, opd[ok], rcond=None)
    return np.concatenate([[0.0], c])
