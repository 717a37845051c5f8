"""TAN-SIP WCS description consumed by the trace kernel.

The reference passes two ``galsim.GSFitsWCS`` objects to the photon ops
(``base['current_image'].wcs`` and ``base['_icrf_to_field']``,
imsim/photon_ops.py:407-408), both order-3 ``FittedSIPWCS`` fits made by
imsim/batoid_wcs.py:429-453,499-505.  ``TanSipWCS`` is the flat equivalent;
``fit_tan_sip`` builds one from matched (x, y) <-> (ra, dec) samples the same
way for synthetic runs.  Host-side set-up only; per-photon evaluation happens
on the GPU.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import _abi


@dataclass
class TanSipWCS:
    crpix: np.ndarray  # (2,)
    cd: np.ndarray  # (2,2) degrees per pixel unit
    center: tuple  # (ra0, dec0) radians
    ab: np.ndarray = field(default_factory=lambda: np.zeros((2, 4, 4)))  # identity folded in
    order: int = 0

    def to_pod(self) -> _abi.B2TanSip:
        w = _abi.B2TanSip()
        w.crpix[0], w.crpix[1] = float(self.crpix[0]), float(self.crpix[1])
        cd = np.asarray(self.cd, float).reshape(2, 2)
        for k, v in enumerate(cd.ravel()):
            w.cd[k] = v
        ab = np.zeros((2, 4, 4))
        a = np.asarray(self.ab, float)
        if self.order > 0:
            if a.shape[1] > 4 or a.shape[2] > 4:
                raise ValueError("SIP order > 3 is not supported")
            ab[:, : a.shape[1], : a.shape[2]] = a
        for k in range(2):
            for i in range(4):
                for j in range(4):
                    w.ab[k][i][j] = ab[k, i, j]
        w.ra0, w.dec0 = float(self.center[0]), float(self.center[1])
        w.order = int(self.order)
        return w

    def intermediate(self, x, y):
        """Tangent-plane coordinates (xi, eta) in degrees of pixel (x, y): SIP polynomial, then CD."""
        u, v = np.asarray(x, float) - self.crpix[0], np.asarray(y, float) - self.crpix[1]
        f, g = u, v
        if self.order > 0:
            a = np.asarray(self.ab, float)
            f = sum(a[0, i, j] * u**i * v**j for i in range(a.shape[1]) for j in range(a.shape[2]) if a[0, i, j] != 0.0)
            g = sum(a[1, i, j] * u**i * v**j for i in range(a.shape[1]) for j in range(a.shape[2]) if a[1, i, j] != 0.0)
        cd = np.asarray(self.cd, float).reshape(2, 2)
        return cd[0, 0] * f + cd[0, 1] * g, cd[1, 0] * f + cd[1, 1] * g

    def local(self, image_pos=None):
        """``galsim.BaseWCS.local``: the Jacobian [[dudx, dudy], [dvdx, dvdy]] in arcsec per pixel at ``image_pos``
        (u towards west like GalSim: u = -xi), as an object with ``getMatrix()`` -- what ``PhotonDCR`` uses."""
        x0, y0 = (self.crpix if image_pos is None else
                  ((image_pos.x, image_pos.y) if hasattr(image_pos, "x") else image_pos))
        h = 0.5
        xi = [self.intermediate(x0 + dx, y0 + dy) for dx, dy in ((h, 0), (-h, 0), (0, h), (0, -h))]
        dxi_dx, deta_dx = (xi[0][0] - xi[1][0]) / (2 * h), (xi[0][1] - xi[1][1]) / (2 * h)
        dxi_dy, deta_dy = (xi[2][0] - xi[3][0]) / (2 * h), (xi[2][1] - xi[3][1]) / (2 * h)
        m = 3600.0 * np.array([[-dxi_dx, -dxi_dy], [deta_dx, deta_dy]])

        class _Jacobian:
            def getMatrix(self_inner):
                return m

        return _Jacobian()


def tan_project(ra, dec, ra0, dec0):
    """Gnomonic projection, FITS convention: (xi, eta) radians, xi east, eta north."""
    dra = ra - ra0
    cosc = np.sin(dec0) * np.sin(dec) + np.cos(dec0) * np.cos(dec) * np.cos(dra)
    xi = np.cos(dec) * np.sin(dra) / cosc
    eta = (np.cos(dec0) * np.sin(dec) - np.sin(dec0) * np.cos(dec) * np.cos(dra)) / cosc
    return xi, eta


def tan_deproject(xi, eta, ra0, dec0):
    c = 1.0 / np.sqrt(1.0 + xi * xi + eta * eta)
    dec = np.arcsin(c * (np.sin(dec0) + eta * np.cos(dec0)))
    ra = ra0 + np.arctan2(xi * c, c * (np.cos(dec0) - eta * np.sin(dec0)))
    return ra, dec


def fit_tan_sip(x, y, ra, dec, order=3, center=None) -> TanSipWCS:
    """Least-squares TAN-SIP fit of pixel <-> sky samples (the role of
    ``galsim.FittedSIPWCS`` in imsim/batoid_wcs.py:453).  Host set-up code."""
    x, y, ra, dec = (np.asarray(a, float) for a in (x, y, ra, dec))
    if center is None:
        # mean direction
        v = np.array([np.mean(np.cos(dec) * np.cos(ra)), np.mean(np.cos(dec) * np.sin(ra)), np.mean(np.sin(dec))])
        v /= np.linalg.norm(v)
        center = (float(np.arctan2(v[1], v[0])), float(np.arcsin(v[2])))
    xi, eta = tan_project(ra, dec, *center)
    xi, eta = np.degrees(xi), np.degrees(eta)
    # affine first: where does (xi, eta) = 0 land in pixels -> crpix
    A = np.column_stack([x, y, np.ones_like(x)])
    cx, *_ = np.linalg.lstsq(A, xi, rcond=None)
    cy, *_ = np.linalg.lstsq(A, eta, rcond=None)
    M = np.array([[cx[0], cx[1]], [cy[0], cy[1]]])
    crpix = -np.linalg.solve(M, np.array([cx[2], cy[2]]))
    if order <= 1:
        return TanSipWCS(crpix=crpix, cd=M, center=center, order=0)
    # refine crpix: it is where the full polynomial (not its affine part) vanishes, so that the
    # final fit needs no constant term
    all_terms = [(i, j) for i in range(order + 1) for j in range(order + 1 - i)]
    for _ in range(6):
        u, v = x - crpix[0], y - crpix[1]
        s = max(np.abs(u).max(), np.abs(v).max())
        D = np.column_stack([(u / s) ** i * (v / s) ** j for i, j in all_terms])
        px, *_ = np.linalg.lstsq(D, xi, rcond=None)
        py, *_ = np.linalg.lstsq(D, eta, rcond=None)
        c = {t: (a, b_) for t, a, b_ in zip(all_terms, px, py)}
        J = np.array([[c[(1, 0)][0], c[(0, 1)][0]], [c[(1, 0)][1], c[(0, 1)][1]]]) / s
        step = -np.linalg.solve(J, np.array(c[(0, 0)]))
        crpix = crpix + step
        if np.abs(step).max() < 1e-13 * s:
            break
    u, v = x - crpix[0], y - crpix[1]
    # polynomial without constant term, scaled for conditioning
    s = max(np.abs(u).max(), np.abs(v).max())
    us, vs = u / s, v / s
    terms = [(i, j) for i in range(order + 1) for j in range(order + 1 - i) if i + j >= 1]
    D = np.column_stack([us**i * vs**j for i, j in terms])
    px, *_ = np.linalg.lstsq(D, xi, rcond=None)
    py, *_ = np.linalg.lstsq(D, eta, rcond=None)
    P = np.zeros((2, 4, 4))
    for (i, j), a, b in zip(terms, px, py):
        P[0, i, j] = a / s ** (i + j)
        P[1, i, j] = b / s ** (i + j)
    cd = np.array([[P[0, 1, 0], P[0, 0, 1]], [P[1, 1, 0], P[1, 0, 1]]])
    cdinv = np.linalg.inv(cd)
    ab = np.einsum("kl,lij->kij", cdinv, P)
    # exact identity for the linear part
    ab[0, 1, 0], ab[0, 0, 1], ab[1, 1, 0], ab[1, 0, 1] = 1.0, 0.0, 0.0, 1.0
    return TanSipWCS(crpix=crpix, cd=cd, center=center, ab=ab, order=order)


def field_wcs(boresight, rot_sky_pos=0.0, distortion=0.0, seed=0) -> TanSipWCS:
    """Synthetic ICRF <-> field-angle WCS: field angles (radians) play the role of
    pixels, as in imsim/batoid_wcs.py:499-505.  ``distortion`` adds random
    quadratic / cubic SIP terms of that relative size at 2 degrees off axis so the
    Newton inversion is exercised."""
    c, s = np.cos(rot_sky_pos), np.sin(rot_sky_pos)
    cd = np.degrees(1.0) * np.array([[-c, s], [s, c]])
    ab = np.zeros((2, 4, 4))
    ab[0, 1, 0] = 1.0
    ab[1, 0, 1] = 1.0
    order = 0
    if distortion:
        rng = np.random.default_rng(seed)
        r = np.radians(2.0)
        for i in range(4):
            for j in range(4 - i):
                if i + j >= 2:
                    ab[:, i, j] = rng.uniform(-1, 1, 2) * distortion / r ** (i + j - 1)
        order = 3
    return TanSipWCS(crpix=np.zeros(2), cd=cd, center=tuple(boresight), ab=ab, order=order)
