"""Postage-stamp sizes of the classic per-object pipeline (mirror of ``imsim/stamp_utils.py`` and the
PSF proxies of ``imsim/psf_utils.py:7-91``).

The reference asks GalSim for ``getGoodImageSize`` of cheap proxy profiles -- Kolmogorov (x) Gaussian for
stars, with the folding threshold lowered to ``noise_var / flux`` for bright ones
(stamp_utils.py:79-156); the object (x) double Gaussian for galaxies, grown until the edge surface
brightness drops below ``sqrt(noise_var) / 8`` for bright ones (stamp_utils.py:159-330).  GalSim is not
vendored by the reference, so ``getGoodImageSize`` is restated from its definition: ``N = 2 R / scale``
rounded up to an even integer, ``R = pi / stepk`` the radius enclosing ``1 - folding_threshold`` of the
flux (never below ``stepk_minimum_hlr = 5`` half-light radii), radii of a convolution added in
quadrature.  The control flow and constants of the reference functions are followed line by line;
sizes can differ from GalSim's by its internal rounding of ``stepk`` (statistical parity only).
"""
from __future__ import annotations

from functools import lru_cache

import numpy as np

from . import _abi

_trapz = getattr(np, "trapezoid", None) or np.trapz  # NumPy >= 2 renamed trapz
from .atmosphere import WLEN_EFF

FT_DEFAULT = 5.0e-3          # galsim.GSParams().folding_threshold
STEPK_MINIMUM_HLR = 5.0      # galsim.GSParams().stepk_minimum_hlr
NMAX = 4096                  # LSST_SiliconBuilder._Nmax (stamp.py:104)
TINY_FLUX = 10               # LSST_SiliconBuilder._tiny_flux
PIXEL_SCALE = 0.2


def _good_size(radius_arcsec: float, pixel_scale: float) -> int:
    """GSObject.getGoodImageSize: ceil(2 pi / (stepk * scale)) with stepk = pi / R, rounded up to even."""
    n = int(np.ceil(2.0 * radius_arcsec / pixel_scale))
    return n + (n % 2)


@lru_cache(maxsize=64)
def _kolmogorov_ee():
    """Encircled energy of the Kolmogorov long-exposure PSF in units of lam / r0: theta(u) table."""
    from scipy.special import j1

    # MTF exp(-3.442 (rho / r0)^(5/3)) on baselines rho [r0]; E(theta) = 2 pi theta int tau(rho) J1(2 pi rho theta) drho
    rho = np.linspace(0.0, 6.0, 6001)
    tau = np.exp(-3.442 * rho ** (5.0 / 3.0))
    theta = np.concatenate([[0.0], np.logspace(-2, 2.5, 1200)])
    E = 2 * np.pi * theta * _trapz(tau[None, :] * j1(2 * np.pi * rho[None, :] * theta[:, None]), rho, axis=1)
    E = np.maximum.accumulate(np.clip(E, 0.0, 1.0))
    return theta, E


def kolmogorov_radius(fwhm_arcsec: float, enclosed: float) -> float:
    """Radius [arcsec] enclosing ``enclosed`` of a Kolmogorov profile of the given FWHM (0.9759 lam / r0);
    beyond the table the analytic wing E = 1 - c theta^(-5/3) is extrapolated."""
    theta, E = _kolmogorov_ee()
    unit = fwhm_arcsec / 0.9758634299  # lam / r0 in arcsec (galsim Kolmogorov._fwhm_factor)
    if enclosed <= E[-2]:
        k = int(np.searchsorted(E, enclosed))
        return float(theta[max(k, 1)] * unit)
    c = (1.0 - E[-200]) * theta[-200] ** (5.0 / 3.0)
    return float((c / max(1.0 - enclosed, 1e-300)) ** 0.6 * unit)


def gaussian_radius(sigma: float, ft: float) -> float:
    return max(np.sqrt(-2.0 * np.log(ft)), STEPK_MINIMUM_HLR * 1.1774100225154747) * sigma


def get_star_stamp_size(nominal_flux, noise_var, Nmax=NMAX, pixel_scale=PIXEL_SCALE, airmass=None, rawSeeing=None,
                        band=None):
    """stamp_utils.py:79-156 with make_kolmogorov_and_gaussian_psf (psf_utils.py:42-91)."""
    folding_threshold = noise_var / nominal_flux
    if folding_threshold >= FT_DEFAULT or folding_threshold == 0:
        ft = FT_DEFAULT
    else:
        ft = float(np.exp(np.floor(np.log(folding_threshold))))  # rounded down to an e-folding
    airmass = 1.2 if airmass is None else airmass
    rawSeeing = 0.7 if rawSeeing is None else rawSeeing
    band = "r" if band is None else band
    fwhm_atm = rawSeeing * (WLEN_EFF[band] / 500.0) ** -0.3 * airmass ** 0.6
    fwhm_sys = np.sqrt(0.25 ** 2 + 0.3 ** 2 + 0.08 ** 2) * airmass ** 0.6
    r_atm = max(kolmogorov_radius(fwhm_atm, 1.0 - ft), STEPK_MINIMUM_HLR * 0.5548 * fwhm_atm)
    r_sys = gaussian_radius(fwhm_sys / 2.3548200450309493, ft)
    return min(_good_size(np.hypot(r_atm, r_sys), pixel_scale), Nmax)


def _double_gaussian_radius(ft=FT_DEFAULT, fwhm1=0.6, fwhm2=0.12):
    """make_double_gaussian (psf_utils.py:7-38): a Sum takes the smallest stepk = the larger radius."""
    return max(gaussian_radius(fwhm1 / 2.355, ft), gaussian_radius(fwhm2 / 2.355, ft))


def _sersic_sb(n, hlr, r):
    """Surface brightness [1 / arcsec^2] of a unit-flux circular Sersic at radius r."""
    from scipy.special import gamma, gammaincinv

    b = gammaincinv(2.0 * n, 0.5)
    norm = b ** (2.0 * n) / (2.0 * np.pi * n * gamma(2.0 * n) * hlr * hlr)
    return norm * np.exp(-b * (r / hlr) ** (1.0 / n))


def get_gal_stamp_size(row, nominal_flux, noise_var, radial_tables=None, sersic_n=None, Nmax=NMAX,
                       pixel_scale=PIXEL_SCALE, arcsec_to_pix=None):
    """stamp_utils.py:159-219 for one object row (``B2Object``): the matrix ``row['m']`` holds size, shear and
    lensing in pixels; its larger singular value bounds the radius along the major axis."""
    a2p = np.eye(2) / pixel_scale if arcsec_to_pix is None else np.asarray(arcsec_to_pix, float)
    M = np.linalg.solve(a2p, np.asarray(row["m"], float).reshape(2, 2))  # back to arcsec
    sv = np.linalg.svd(M, compute_uv=False)
    kind = int(row["kind"])
    if kind == _abi.PROF_RADIAL:
        n = sersic_n[int(row["lut"])]
        from .stage1 import RADIAL_TMAX

        tab = radial_tables[int(row["lut"])]
        t = np.linspace(0.0, RADIAL_TMAX, tab.size)
        r_unit = max(float(np.interp(-np.log(FT_DEFAULT), t, tab)), STEPK_MINIMUM_HLR)
    elif kind in (_abi.PROF_KNOTS, _abi.PROF_GAUSSIAN):
        n = 0.5  # RandomKnots._profile is a Gaussian of the same half-light radius (stamp_utils.py:296-312)
        scale = 1.0 / 1.1774100225154747 if kind == _abi.PROF_KNOTS else 1.0
        r_unit = gaussian_radius(scale, FT_DEFAULT)
    elif kind == _abi.PROF_BOX:
        n = None
        r_unit = 0.5 * float(np.hypot(row["p0"], row["p1"]))
    else:
        n = None
        r_unit = 0.0
    r_psf = _double_gaussian_radius()
    stamp_size = _good_size(np.hypot(r_unit * sv[0], r_psf), pixel_scale)
    if (nominal_flux > 10 * stamp_size ** 2) or (stamp_size > Nmax):
        keep_sb_level = np.sqrt(noise_var) / 8.0

        def grow(level):
            # get_good_phot_stamp_size1 (stamp_utils.py:262-330) per component, sizes added in quadrature
            def one(sb_at, N0):
                N = N0
                while N < Nmax:
                    h = N / 2 * pixel_scale
                    if sb_at(h) * nominal_flux * pixel_scale ** 2 <= level:
                        break
                    N = int(1.1 * N) + 1
                return min(N, Nmax) if N >= Nmax else N

            if n is not None:
                # unit profile (hlr_unit) stretched by the matrix: along the major axis the surface
                # brightness is sb_unit(h / sv0) / (sv0 sv1)
                hlr_unit = 1.1774100225154747 if kind == _abi.PROF_GAUSSIAN else 1.0
                n_gal = one(lambda h: _sersic_sb(n, hlr_unit, h / sv[0]) / (sv[0] * sv[1]),
                            _good_size(r_unit * sv[0], pixel_scale))
            else:
                n_gal = _good_size(r_unit * sv[0], pixel_scale)
            s1 = 0.6 / 2.355
            n_psf = one(lambda h: np.exp(-0.5 * (h / s1) ** 2) / (2 * np.pi * s1 * s1), _good_size(r_psf, pixel_scale))
            return int(np.sqrt(n_gal ** 2 + n_psf ** 2))

        stamp_size = grow(keep_sb_level)
        if stamp_size > Nmax:
            stamp_size = min(grow(3 * keep_sb_level), Nmax)
    return stamp_size


def get_stamp_size(row, nominal_flux, noise_var, Nmax=NMAX, pixel_scale=PIXEL_SCALE, airmass=None, rawSeeing=None,
                   band=None, radial_tables=None, sersic_n=None, arcsec_to_pix=None):
    """stamp_utils.py:9-76 plus the two shortcuts of LSST_SiliconBuilder.setup (stamp.py:208-214)."""
    if nominal_flux < TINY_FLUX:
        return 32
    if int(row["kind"]) == _abi.PROF_DELTA:
        return get_star_stamp_size(nominal_flux, noise_var, Nmax, pixel_scale, airmass, rawSeeing, band)
    return get_gal_stamp_size(row, nominal_flux, noise_var, radial_tables, sersic_n, Nmax, pixel_scale, arcsec_to_pix)


def get_stamp_sizes(rows, nominal_flux, noise_var, Nmax=NMAX, pixel_scale=PIXEL_SCALE, airmass=None, rawSeeing=None,
                    band=None, radial_tables=None, sersic_n=None, arcsec_to_pix=None):
    """``get_stamp_size`` for a whole catalogue at once (same sizes, object for object): stars share the few
    e-folding levels of the folding threshold, galaxies take the batched singular values of their matrices,
    and only the bright ones (the growing loop of stamp_utils.py:262-330) go through the per-object code."""
    from .stage1 import RADIAL_TMAX

    rows = np.asarray(rows)
    flux = np.asarray(nominal_flux, dtype=np.float64)
    n = rows.size
    out = np.full(n, 32, dtype=np.int64)
    kind = rows["kind"].astype(np.int64)
    live = flux >= TINY_FLUX
    star = live & (kind == _abi.PROF_DELTA)
    if star.any():
        with np.errstate(divide="ignore"):
            ft = noise_var / flux[star]
        level = np.where((ft >= FT_DEFAULT) | (ft == 0), np.inf, np.floor(np.log(np.where(ft > 0, ft, 1.0))))
        sizes = np.empty(level.size, dtype=np.int64)
        for lv in np.unique(level):
            pick = level == lv
            # any flux of the level gives the level's size
            f = float(flux[star][pick][0])
            sizes[pick] = get_star_stamp_size(f, noise_var, Nmax, pixel_scale, airmass, rawSeeing, band)
        out[star] = sizes
    gal = np.flatnonzero(live & ~star)
    if gal.size:
        a2p = np.eye(2) / pixel_scale if arcsec_to_pix is None else np.asarray(arcsec_to_pix, float)
        M = np.linalg.solve(a2p, np.asarray(rows["m"][gal], float).reshape(-1, 2, 2))
        sv0 = np.linalg.svd(M, compute_uv=False)[:, 0]
        k = kind[gal]
        r_unit = np.zeros(gal.size)
        rad = k == _abi.PROF_RADIAL
        if rad.any():
            lut = rows["lut"][gal][rad].astype(np.int64)
            per_lut = {}
            for u in np.unique(lut):
                tab = radial_tables[int(u)]
                t = np.linspace(0.0, RADIAL_TMAX, tab.size)
                per_lut[int(u)] = max(float(np.interp(-np.log(FT_DEFAULT), t, tab)), STEPK_MINIMUM_HLR)
            r_unit[rad] = [per_lut[int(u)] for u in lut]
        r_unit[k == _abi.PROF_KNOTS] = gaussian_radius(1.0 / 1.1774100225154747, FT_DEFAULT)
        r_unit[k == _abi.PROF_GAUSSIAN] = gaussian_radius(1.0, FT_DEFAULT)
        box = k == _abi.PROF_BOX
        if box.any():
            r_unit[box] = 0.5 * np.hypot(rows["p0"][gal][box], rows["p1"][gal][box])
        nn = np.ceil(2.0 * np.hypot(r_unit * sv0, _double_gaussian_radius()) / pixel_scale).astype(np.int64)
        size = nn + (nn % 2)
        bright = (flux[gal] > 10 * size ** 2) | (size > Nmax)
        for i in np.flatnonzero(bright):
            j = gal[i]
            size[i] = get_gal_stamp_size(rows[j], float(flux[j]), noise_var, radial_tables, sersic_n, Nmax,
                                         pixel_scale, arcsec_to_pix)
        out[gal] = size
    return out
