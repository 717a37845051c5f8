"""Atmospheric PSF as a device photon kick (mirror of ``imsim.atmPSF.AtmosphericPSF``, imsim/atmPSF.py:84-336).

The reference builds ``galsim.Atmosphere`` -- six frozen-flow von Karman phase screens (Ellerbroek
altitudes / weights, random wind, outer scale drawn from a truncated log-normal, ``r0_500`` chosen by
bisection so that the Tokovinin von Karman FWHM equals ``rawSeeing * airmass^0.6 * (lam_eff/500)^-0.3``,
atmPSF.py:211-296) -- truncates the screens at ``kmax = kcrit / r0`` (atmPSF.py:173-189) and hands the
high-k remainder to ``galsim.SecondKick`` (atmPSF.py:195-202).  ``getPSF`` (atmPSF.py:298-336) returns
``Convolve(ChromaticAtmosphere(PhaseScreenPSF, alpha=-0.3), SecondKick)``, which photon shooting turns
into: bilinear wavefront-gradient lookup per screen at the photon's (u, v, t, theta), summed and scaled
by ``(lam / lam_eff)^alpha``, plus a radial draw from the second-kick profile.

The screen synthesis and the second-kick profile live in GalSim (``galsim/phase_screens.py``,
``SBSecondKick.cpp``), which the reference does not vendor: both are restated here from the published
definitions (von Karman phase PSD ``0.0228 r0^-5/3 (f^2 + L0^-2)^-11/6``; second kick = Airy MTF x
``exp(-D_hk / 2)`` with ``D_hk`` the structure function of the modes above ``kcrit / r0``), PARITY
UNPINNED against GalSim's realisations; the tests check the physics (structure function, FWHM).
The parameter draws of ``_getAtmKwargs`` and the seeing relations are the reference's own code and are
followed line by line (with a numpy Generator in place of galsim's deviates); they are PINNED: the reference's
source of those methods, executed on recorded deviates, gives ``tests/golden/atmosphere.npz``.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _abi, _lib

ARCSEC = 206264.80624709636
WLEN_EFF = dict(u=365.49, g=480.03, r=622.20, i=754.06, z=868.21, y=991.66)  # atmPSF.py:125 (LSE-40 table 2)


#: galsim/kolmogorov.py ``Kolmogorov._fwhm_factor`` (FWHM = factor x lam / r0; recalled, GalSim is not in the tree)
KOLMOGOROV_FWHM_FACTOR = 0.9758634299


def kolmogorov_fwhm(r0_500: float, lam_nm: float) -> float:
    """``galsim.Kolmogorov(r0_500=, lam=).fwhm`` [arcsec]: 0.9759 lam / r0(lam)."""
    r0 = r0_500 * (lam_nm / 500.0) ** 1.2
    return KOLMOGOROV_FWHM_FACTOR * lam_nm * 1e-9 / r0 * ARCSEC


def vk_seeing(r0_500: float, wavelength: float, L0: float) -> float:
    """atmPSF.py:211-221 (Tokovinin 2002 eq. 19)."""
    r0 = r0_500 * (wavelength / 500.0) ** 1.2
    arg = 1.0 - 2.183 * (r0 / L0) ** 0.356
    return kolmogorov_fwhm(r0_500, wavelength) * (np.sqrt(arg) if arg > 0.0 else 0.0)


def r0_500_for_seeing(wavelength: float, L0: float, target: float) -> float:
    """atmPSF.py:227-237."""
    from scipy.optimize import bisect

    r0_max = min(1.0, L0 * (1.0 / 2.183) ** (-0.356) * (wavelength / 500.0) ** 1.2)
    return bisect(lambda r: vk_seeing(r, wavelength, L0) - target, 0.01, r0_max)


def von_karman_screen(npix: int, scale: float, r0_500: float, L0: float, rng: np.random.Generator,
                      kmax: Optional[float] = None, device=None, dtype=np.float32):
    """One periodic phase screen [nm of optical path] of ``npix^2`` cells of ``scale`` metres with the von
    Karman spectrum for ``r0_500``, keeping only modes with ``2 pi |f| <= kmax`` (``galsim.AtmosphericScreen``
    after ``instantiate(kmax=)``).  ``device``: torch device for the FFT (None = numpy on the host)."""
    size = npix * scale
    if device is None:
        f = np.fft.fftfreq(npix, scale)
        f2 = f[:, None] ** 2 + f[None, :] ** 2
        amp = np.sqrt(0.0228) * r0_500 ** (-5.0 / 6.0) * (f2 + L0 ** -2.0) ** (-11.0 / 12.0) / size
        amp[0, 0] = 0.0
        if kmax is not None:
            amp[(2 * np.pi) ** 2 * f2 > kmax * kmax] = 0.0
        z = (rng.standard_normal((npix, npix)) + 1j * rng.standard_normal((npix, npix))) / np.sqrt(2.0)
        # phi = sum_f c_f exp(2 pi i f x): numpy's ifft2 carries 1/npix^2
        phi = np.fft.ifft2(amp * z) * (npix * npix)
        screen = np.sqrt(2.0) * phi.real * (500.0 / (2 * np.pi))  # rad at 500 nm -> nm
        return screen.astype(dtype)
    import torch

    tdt = torch.float32 if dtype == np.float32 else torch.float64
    gen = torch.Generator(device=device)
    gen.manual_seed(int(rng.integers(1 << 62)))
    f = torch.fft.fftfreq(npix, scale, device=device, dtype=torch.float64)
    f2 = f[:, None] ** 2 + f[None, :] ** 2
    amp = (0.0228 ** 0.5) * r0_500 ** (-5.0 / 6.0) * (f2 + L0 ** -2.0) ** (-11.0 / 12.0) / size
    amp[0, 0] = 0.0
    if kmax is not None:
        amp = torch.where((2 * np.pi) ** 2 * f2 > kmax * kmax, torch.zeros_like(amp), amp)
    zr = torch.randn((npix, npix), generator=gen, device=device, dtype=torch.float64)
    zi = torch.randn((npix, npix), generator=gen, device=device, dtype=torch.float64)
    phi = torch.fft.ifft2(torch.complex(amp * zr, amp * zi) / np.sqrt(2.0)) * (npix * npix)
    return (np.sqrt(2.0) * (500.0 / (2 * np.pi)) * phi.real).to(tdt).contiguous()


def _g_tail(xc: np.ndarray) -> np.ndarray:
    """G(xc) = int_xc^inf x^(-8/3) (1 - J0(x)) dx, the high-k share of the Kolmogorov structure function."""
    from scipy.special import j0

    # log grid to x = 1e4, analytic tail beyond (J0 averages out): int x^-8/3 = (3/5) x^-5/3
    x = np.logspace(-6, 4, 200001)
    fx = x ** (-8.0 / 3.0) * (1.0 - j0(x))
    seg = 0.5 * (fx[1:] + fx[:-1]) * np.diff(x)
    cum_from_top = np.concatenate([np.cumsum(seg[::-1])[::-1], [0.0]]) + 0.6 * x[-1] ** (-5.0 / 3.0)
    # below the grid (1 - J0 ~ x^2/4): int_0^x0 x^-2/3 / 4 dx = (3/4) x0^(1/3)
    total0 = cum_from_top[0] + 0.75 * x[0] ** (1.0 / 3.0)
    xc = np.asarray(xc, float)
    out = np.interp(xc, x, cum_from_top)
    small = xc < x[0]
    out[small] = total0 - 0.75 * xc[small] ** (1.0 / 3.0)
    return out


def second_kick_table(lam_nm: float, r0: float, diam: float, obscuration: float, kcrit: float, entries: int = 2048,
                      tmax: float = 12.0, theta_max_arcsec: float = 60.0):
    """Radial sampling table of ``galsim.SecondKick(lam, r0, diam, obscuration, kcrit)``: the expectation of the
    high-k turbulence PSF times the annular-aperture diffraction pattern.

    tau(rho) = MTF_annulus(rho) exp(-D_hk(rho) / 2) on baselines rho in [0, diam];
    D_hk(rho) = 6.8839 (rho / r0)^(5/3) G(kcrit rho / r0) / G(0);
    encircled energy E(theta) = (2 pi theta / lam) int_0^diam tau(rho) J1(2 pi rho theta / lam) drho.

    Returns ``(theta[entries] in arcsec on t = -log(1 - u) in [0, tmax], delta_prob = 0)``: the unscattered
    core is diffraction broadened, so it is part of the table rather than a delta function.  Energy beyond
    ``theta_max_arcsec`` is dropped (renormalised)."""
    from scipy.special import j1

    lam = lam_nm * 1e-9
    rho = np.linspace(0.0, diam, 4001)
    # annular aperture MTF = autocorrelation of the annulus (circle overlaps)
    def overlap(d, r1, r2):
        d = np.maximum(d, 1e-12)
        out = np.zeros_like(d)
        inside = d <= abs(r1 - r2)
        out[inside] = np.pi * min(r1, r2) ** 2
        part = (~inside) & (d < r1 + r2)
        dp = d[part]
        a1 = np.arccos(np.clip((dp * dp + r1 * r1 - r2 * r2) / (2 * dp * r1), -1, 1))
        a2 = np.arccos(np.clip((dp * dp + r2 * r2 - r1 * r1) / (2 * dp * r2), -1, 1))
        out[part] = r1 * r1 * a1 + r2 * r2 * a2 - 0.5 * np.sqrt(
            np.maximum((-dp + r1 + r2) * (dp + r1 - r2) * (dp - r1 + r2) * (dp + r1 + r2), 0.0))
        return out

    R, Ri = diam / 2.0, obscuration * diam / 2.0
    mtf = overlap(rho, R, R) - 2.0 * overlap(rho, R, Ri) + overlap(rho, Ri, Ri)
    mtf /= np.pi * (R * R - Ri * Ri)
    g0 = _g_tail(np.array([0.0]))[0]
    dhk = 6.8839 * (rho / r0) ** (5.0 / 3.0) * _g_tail(kcrit * rho / r0) / g0
    tau = mtf * np.exp(-0.5 * dhk)
    theta = np.concatenate([[0.0], np.logspace(-4, np.log10(theta_max_arcsec), 1500)]) / ARCSEC
    arg = 2 * np.pi * rho[None, :] * theta[:, None] / lam
    integrand = tau[None, :] * j1(arg)
    E = (2 * np.pi * theta / lam) * np.trapezoid(integrand, rho, axis=1)
    # the far (theta^-11/3) wings beyond the grid are truncated and the rest renormalised, as GalSim
    # truncates its profiles at the folding threshold
    E = np.maximum.accumulate(np.clip(E, 0.0, None))
    E /= E[-1]
    t = np.linspace(0.0, tmax, entries)
    u = -np.expm1(-t)
    table = np.interp(u, E, theta * ARCSEC)
    return np.ascontiguousarray(table), 0.0, tmax


class AtmosphericPSF:
    """Same constructor as ``imsim.atmPSF.AtmosphericPSF`` (atmPSF.py:110-147); ``rng`` is a numpy
    Generator / seed; ``doOpt`` (the mock optics screens) is not supported -- the optics are ray traced.
    ``npix`` overrides ``screen_size / screen_scale`` (tests use small screens)."""

    def __init__(self, airmass, rawSeeing, band, boresight=None, rng=None, t0=0.0, exptime=30.0, kcrit=0.2,
                 screen_size=819.2, screen_scale=0.1, doOpt=False, exponent=-0.3, logger=None, nproc=None,
                 save_file=None, _no2k=False, device=None, dtype=np.float32, gauss_fwhm=0.3):
        if doOpt:
            raise NotImplementedError("doOpt=True duplicates the ray-traced optics (atmPSF.py:389-403)")
        self.airmass, self.rawSeeing, self.band, self.boresight = airmass, rawSeeing, band, boresight
        self.wlen_eff = WLEN_EFF[band]
        self.targetFWHM = rawSeeing * airmass ** 0.6 * (self.wlen_eff / 500.0) ** (-0.3)  # atmPSF.py:128
        self.rng = rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)
        self.t0, self.exptime, self.kcrit = t0, exptime, kcrit
        self.screen_size, self.screen_scale, self.exponent = screen_size, screen_scale, exponent
        self.gauss_fwhm = gauss_fwhm  # config/imsim-config.yaml:252-256
        self.kw = self._getAtmKwargs()
        self.npix = int(round(screen_size / screen_scale))
        # galsim.Atmosphere: r0_500 of layer i = r0_500 * weight_i^(-3/5); r0_500_effective recombines them
        w = np.asarray(self.kw["r0_weights"])
        self.r0_500_layers = self.kw["r0_500"] * w ** (-3.0 / 5.0)
        self.r0_500_effective = float(np.sum(self.r0_500_layers ** (-5.0 / 3.0)) ** (-3.0 / 5.0))
        self.r0 = self.r0_500_effective * (self.wlen_eff / 500.0) ** 1.2  # atmPSF.py:176-177
        self.kmax = kcrit / self.r0
        self.screens = [von_karman_screen(self.npix, screen_scale, r, L, self.rng, kmax=self.kmax, device=device,
                                          dtype=dtype)
                        for r, L in zip(self.r0_500_layers, self.kw["L0"])]
        self.second_kick = None if _no2k else second_kick_table(self.wlen_eff, self.r0, 8.36, 0.61, kcrit)
        self._dtype = dtype

    def _getAtmKwargs(self):
        """atmPSF.py:239-296, draw for draw."""
        gd, ud = self.rng.standard_normal, self.rng.random
        altitudes = [0.2, 2.58, 5.16, 7.73, 12.89, 15.46]
        weights = [0.652, 0.172, 0.055, 0.025, 0.074, 0.022]
        weights = [np.abs(w * (1.0 + 0.1 * gd())) for w in weights]
        weights = np.clip(weights, 0.01, 0.8)
        weights /= np.sum(weights)
        L0 = 0
        while L0 < 10.0 or L0 > 100:
            L0 = np.exp(gd() * 0.6 + np.log(25.0))
        r0_500 = r0_500_for_seeing(self.wlen_eff, L0, self.targetFWHM)
        speeds = [ud() * 20.0 for _ in range(6)]
        directions = [ud() * 2 * np.pi for _ in range(6)]
        return dict(r0_500=r0_500, L0=[L0] * 6, speed=speeds, direction=directions, altitude=altitudes,
                    r0_weights=weights, screen_size=self.screen_size, screen_scale=self.screen_scale)

    # ------------------------------------------------------------------
    def to_pod(self, arcsec_to_pix=None) -> _abi.B2Psf:
        p = _abi.B2Psf()
        p.n_screens = len(self.screens)
        p.npix = self.npix
        p.screen_f32 = 1 if self._dtype == np.float32 else 0
        p.screen_scale = self.screen_scale
        for l in range(p.n_screens):
            p.altitude[l] = self.kw["altitude"][l] * 1000.0
            p.vx[l] = self.kw["speed"][l] * np.cos(self.kw["direction"][l])
            p.vy[l] = self.kw["speed"][l] * np.sin(self.kw["direction"][l])
        p.t0, p.exptime = self.t0, self.exptime
        p.r_outer, p.r_inner = 8.36 / 2.0, 0.61 * 8.36 / 2.0
        p.base_wavelength, p.exponent = self.wlen_eff, self.exponent
        if self.second_kick is not None:
            p.n_kick = self.second_kick[0].size
            p.kick_delta_prob, p.kick_tmax = self.second_kick[1], self.second_kick[2]
        p.gauss_sigma = self.gauss_fwhm / 2.3548200450309493
        a = np.eye(2) / 0.2 if arcsec_to_pix is None else np.asarray(arcsec_to_pix, float)
        for k in range(4):
            p.arcsec_to_pix[k] = a.ravel()[k]
        return p

    def upload(self, ctx, arcsec_to_pix=None, packed: bool = True):
        """Bind the screens (moved to the context's GPU once) and tables to ``ctx`` (``b2_psf_upload``).
        ``packed`` (float32 screens): store each cell with its three periodic neighbours as one float4, so the
        bilinear gradient costs one 16-byte gather per screen instead of four 4-byte ones (4x the memory:
        6.4 GB for six 8192^2 screens)."""
        import torch

        dev = "cuda:%d" % ctx.device
        pod = self.to_pod(arcsec_to_pix)
        key = (dev, bool(packed))
        if getattr(self, "_dev_key", None) != key:
            plain = [s.to(dev) if hasattr(s, "to") else torch.as_tensor(np.ascontiguousarray(s), device=dev)
                     for s in self.screens]
            if packed and self._dtype == np.float32:
                plain = [torch.stack([s, torch.roll(s, -1, 1), torch.roll(s, -1, 0), torch.roll(s, (-1, -1), (0, 1))],
                                     dim=-1).contiguous() for s in plain]
            self._dev_screens, self._dev_key = plain, key
        if packed and self._dtype == np.float32:
            pod.screen_f32 = 2
        ptrs = (C.c_void_p * len(self._dev_screens))(*[s.data_ptr() for s in self._dev_screens])
        kick = self.second_kick[0] if self.second_kick is not None else None
        # same PSF, same local WCS as the context already holds (the next detector of a visit): nothing to upload
        key = (id(self), id(kick), bytes(pod), tuple(s.data_ptr() for s in self._dev_screens))
        if getattr(ctx, "_psf_key", None) != key:
            _lib.check(_lib.load().b2_psf_upload(ctx.handle, C.byref(pod), ptrs,
                                                 kick.ctypes.data if kick is not None else None))
            ctx._psf_key = key
        return pod


class GaussianPSF:
    """``psf: {type: Gaussian, fwhm: ...}``: no atmosphere, just the Gaussian kick."""

    def __init__(self, fwhm=0.7):
        self.fwhm = fwhm

    def to_pod(self, arcsec_to_pix=None):
        p = _abi.B2Psf()
        p.gauss_sigma = self.fwhm / 2.3548200450309493
        a = np.eye(2) / 0.2 if arcsec_to_pix is None else np.asarray(arcsec_to_pix, float)
        for k in range(4):
            p.arcsec_to_pix[k] = a.ravel()[k]
        return p

    def upload(self, ctx, arcsec_to_pix=None):
        pod = self.to_pod(arcsec_to_pix)
        key = ("gaussian", bytes(pod))
        if getattr(ctx, "_psf_key", None) != key:
            _lib.check(_lib.load().b2_psf_upload(ctx.handle, C.byref(pod), None, None))
            ctx._psf_key = key
        return pod
