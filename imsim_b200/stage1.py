"""Stage 1 on the device: catalogue objects -> pooled photons.

The reference draws every object with GalSim on the host: ``InstCatalog.getObj``
(imsim/instcat.py:465-561) builds ``DeltaFunction`` / ``Sersic(n, hlr).shear(q, beta).lens(g1, g2, mu)`` /
``RandomKnots`` / ``Box``; ``LSST_PhotonsBuilder.draw`` (imsim/stamp.py:599-744) shoots it with the PSF as
the first photon op and ``merge_photon_arrays`` (imsim/photon_pooling.py:177-192) concatenates the stamps.
Here the same object descriptions become rows of a device table (``B2Object``) and ``b2_stage1_photons``
writes the pooled SoA directly in HBM.

GalSim's own random streams cannot be reproduced (its samplers are not in the reference tree), so parity
with the reference is statistical for this stage: profile moments, half-light radii, wavelength
distributions.  The test oracle restates the per-photon arithmetic for injected uniforms.
"""
from __future__ import annotations

import ctypes as C
import gzip
from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

from . import _abi, _lib

#: t = -log(1 - u) range of the radial tables: photons beyond 1 - exp(-TMAX) of the flux are clamped
RADIAL_TMAX = 14.0
RADIAL_ENTRIES = 4096


def shear_matrix(q: float = None, beta: float = 0.0, g1: float = None, g2: float = None) -> np.ndarray:
    """``galsim.Shear(q=, beta=)`` or ``Shear(g1=, g2=)`` as the 2x2 matrix ``getMatrix()`` returns:
    ``[[1 + g1, g2], [g2, 1 - g1]] / sqrt(1 - |g|^2)`` (area preserving)."""
    if q is not None:
        g = (1.0 - q) / (1.0 + q)
        g1, g2 = g * np.cos(2.0 * beta), g * np.sin(2.0 * beta)
    gsq = g1 * g1 + g2 * g2
    return np.array([[1.0 + g1, g2], [g2, 1.0 - g1]]) / np.sqrt(1.0 - gsq)


def lens_matrix(g1: float, g2: float, mu: float) -> np.ndarray:
    """``GSObject._lens(g1, g2, mu)``: reduced shear then magnification, ``Shear(g1, g2).getMatrix() * sqrt(mu)``."""
    return shear_matrix(g1=g1, g2=g2) * np.sqrt(mu)


def lens_params(gamma1: float, gamma2: float, kappa: float):
    """imsim/instcat.py:438-444."""
    g1 = gamma1 / (1.0 - kappa)
    g2 = gamma2 / (1.0 - kappa)
    mu = 1.0 / ((1.0 - kappa) ** 2 - (gamma1 ** 2 + gamma2 ** 2))
    return g1, g2, mu


def sersic_radial_table(n: float, entries: int = RADIAL_ENTRIES, tmax: float = RADIAL_TMAX) -> np.ndarray:
    """Radius (in half-light radii) enclosing the flux fraction ``u = 1 - exp(-t)`` of a Sersic-``n``
    profile, on a uniform grid of ``t``: ``F(r) = P(2n, b (r/re)^(1/n))``, ``P(2n, b) = 1/2``."""
    from scipy.special import gammaincinv

    t = np.linspace(0.0, tmax, entries)
    u = -np.expm1(-t)
    b = gammaincinv(2.0 * n, 0.5)
    return (gammaincinv(2.0 * n, u) / b) ** n


@dataclass
class ObjectTable:
    """Rows of ``B2Object`` plus the per-object photon fluxes; profile sizes are in arcsec and are mapped to
    pixels by ``arcsec_to_pix`` (the inverse local WCS Jacobian) when the table is built."""

    arcsec_to_pix: np.ndarray = field(default_factory=lambda: np.eye(2) / 0.2)
    rows: List[np.ndarray] = field(default_factory=list)
    flux: List[np.ndarray] = field(default_factory=list)
    sersic_n: List[float] = field(default_factory=list)  # distinct indices -> rows of the radial table

    def _new(self, n):
        r = np.zeros(n, dtype=_abi.OBJECT_DTYPE)
        r["m"] = np.eye(2).ravel()
        return r

    def _lut_row(self, n: float) -> int:
        n = round(float(n) * 20.0) / 20.0  # instcat.py:504-511: quantised at 0.05
        if n not in self.sersic_n:
            self.sersic_n.append(n)
        return self.sersic_n.index(n)

    def _push(self, r, flux, x, y, sed, tanx, tany):
        r["x"], r["y"], r["sed"], r["tanx"], r["tany"] = x, y, sed, tanx, tany
        self.rows.append(r)
        self.flux.append(np.broadcast_to(np.asarray(flux, dtype=np.float64), r.shape).copy())

    def add_points(self, x, y, flux, sed=0, tanx=0.0, tany=0.0):
        """``galsim.DeltaFunction`` (stars)."""
        x = np.atleast_1d(np.asarray(x, float))
        r = self._new(x.size)
        r["kind"] = _abi.PROF_DELTA
        self._push(r, flux, x, y, sed, tanx, tany)

    def add_gaussians(self, x, y, flux, sigma_arcsec, sed=0, tanx=0.0, tany=0.0):
        x = np.atleast_1d(np.asarray(x, float))
        r = self._new(x.size)
        r["kind"] = _abi.PROF_GAUSSIAN
        r["m"] = (np.asarray(sigma_arcsec, float).reshape(-1, 1, 1) * self.arcsec_to_pix).reshape(-1, 4)
        self._push(r, flux, x, y, sed, tanx, tany)

    def _galaxy_matrix(self, hlr, q, beta, g1, g2, mu):
        """arcsec_to_pix @ lens(g1, g2, mu) @ shear(q, beta) * hlr, vectorised over objects -> (N, 4)."""
        hlr, q, beta, g1, g2, mu = np.broadcast_arrays(*(np.atleast_1d(np.asarray(a, float))
                                                         for a in (hlr, q, beta, g1, g2, mu)))
        g = (1.0 - q) / (1.0 + q)
        s1, s2 = g * np.cos(2.0 * beta), g * np.sin(2.0 * beta)
        S = np.array([[1.0 + s1, s2], [s2, 1.0 - s1]]) / np.sqrt(1.0 - (s1 * s1 + s2 * s2))  # (2, 2, N)
        L = np.array([[1.0 + g1, g2], [g2, 1.0 - g1]]) * (np.sqrt(mu) / np.sqrt(1.0 - (g1 * g1 + g2 * g2)))
        M = np.einsum("ij,jkn,kln->nil", self.arcsec_to_pix, L, S) * hlr[:, None, None]
        return M.reshape(-1, 4)

    def add_sersic(self, x, y, flux, hlr_arcsec, n, q=1.0, beta=0.0, g1=0.0, g2=0.0, mu=1.0, sed=0, tanx=0.0, tany=0.0):
        """``Sersic(n, half_light_radius)._shear(Shear(q, beta))._lens(g1, g2, mu)`` (instcat.py:496-520);
        scalars or arrays (one Sersic index per call)."""
        m = self._galaxy_matrix(hlr_arcsec, q, beta, g1, g2, mu)
        r = self._new(max(m.shape[0], np.atleast_1d(x).size))
        r["kind"] = _abi.PROF_RADIAL
        r["lut"] = self._lut_row(n)
        r["m"] = m
        self._push(r, flux, x, y, sed, tanx, tany)

    def add_knots(self, x, y, flux, hlr_arcsec, npoints, q=1.0, beta=0.0, g1=0.0, g2=0.0, mu=1.0, sed=0, seed=0,
                  tanx=0.0, tany=0.0):
        """``RandomKnots(npoints, half_light_radius)`` sheared and lensed (instcat.py:522-545); scalars or arrays."""
        m = self._galaxy_matrix(hlr_arcsec, q, beta, g1, g2, mu)
        r = self._new(max(m.shape[0], np.atleast_1d(x).size))
        r["kind"] = _abi.PROF_KNOTS
        r["n_knots"] = npoints
        r["knot_seed"] = np.asarray(seed, dtype=np.uint64)
        r["m"] = m
        self._push(r, flux, x, y, sed, tanx, tany)

    def add_streak(self, x, y, flux, length_arcsec, width_arcsec, position_angle=0.0, sed=0, tanx=0.0, tany=0.0):
        """``Box(length, width).rotate(position_angle)`` (instcat.py:486-494)."""
        r = self._new(1)
        r["kind"] = _abi.PROF_BOX
        r["p0"], r["p1"] = length_arcsec, width_arcsec
        c, s = np.cos(position_angle), np.sin(position_angle)
        r["m"] = (self.arcsec_to_pix @ np.array([[c, -s], [s, c]])).ravel()
        self._push(r, flux, x, y, sed, tanx, tany)

    # ------------------------------------------------------------------
    def build(self):
        """``(objects structured array, flux array)``."""
        if not self.rows:
            return np.zeros(0, dtype=_abi.OBJECT_DTYPE), np.zeros(0)
        return np.concatenate(self.rows), np.concatenate(self.flux)

    def radial_tables(self) -> Optional[np.ndarray]:
        if not self.sersic_n:
            return None
        return np.ascontiguousarray(np.stack([sersic_radial_table(n) for n in self.sersic_n]))

    def __len__(self):
        return sum(r.size for r in self.rows)


# ---------------------------------------------------------------------- instance catalogues
@dataclass
class InstCatObject:
    objid: str
    ra_deg: float
    dec_deg: float
    magnorm: float
    sed_name: str
    redshift: float
    lens: tuple
    objinfo: List[str]
    dust: List[str]


def read_instcat_objects(file_name: str, flip_g2: bool = True, skip_invalid: bool = True) -> List[InstCatObject]:
    """The ``object`` lines of a phoSim instance catalogue, tokenised like ``InstCatalog._load``
    (imsim/instcat.py:208-300): columns 7-9 are (gamma1, gamma2, kappa), columns 12.. the source type and its
    spatial parameters up to the dust block, whose start depends on the type."""
    dust_index = {"point": 13, "sersic2d": 17, "knots": 17, "streak": 16}
    g2_sign = -1.0 if flip_g2 else 1.0
    out = []
    opener = gzip.open if file_name.endswith(".gz") else open
    with opener(file_name, "rt") as fh:
        for line in fh:
            if " inf " in line or not line.startswith("object"):
                continue
            tok = line.strip().split()
            di = dust_index.get(tok[12].lower(), 15)
            objinfo, dust = tok[12:di], tok[di:]
            magnorm = float(tok[4])
            if skip_invalid:
                ok = magnorm < 50.0
                if objinfo[0].lower() == "sersic2d" and float(objinfo[1]) < float(objinfo[2]):
                    ok = False
                if objinfo[0].lower() == "knots" and (float(objinfo[1]) < float(objinfo[2]) or int(objinfo[4]) <= 0):
                    ok = False
                if not ok:
                    continue
            out.append(InstCatObject(tok[1], float(tok[2]), float(tok[3]), magnorm, tok[5], float(tok[6]),
                                     (float(tok[7]), g2_sign * float(tok[8]), float(tok[9])), objinfo, dust))
    return out


def add_instcat_object(table: ObjectTable, obj: InstCatObject, x: float, y: float, flux: float, sed: int = 0,
                       flip_g2: bool = True, knot_seed: int = 0, tanx: float = 0.0, tany: float = 0.0) -> bool:
    """``InstCatalog.getObj`` (imsim/instcat.py:465-561) for one catalogue entry placed at image position
    (x, y) with ``flux`` photons.  Returns False for entries the reference skips (magnorm >= 50) and raises
    for FITS-image sources (``InterpolatedImage``), which stage 1 does not generate on the device."""
    p = obj.objinfo
    kind = p[0].lower()
    if obj.magnorm >= 50:
        return False
    if kind == "point":
        table.add_points(x, y, flux, sed=sed, tanx=tanx, tany=tany)
    elif kind == "streak":
        table.add_streak(x, y, flux, float(p[1]), float(p[2]), np.radians(float(p[3])), sed=sed, tanx=tanx, tany=tany)
    elif kind in ("sersic2d", "knots"):
        a, b, pa = float(p[1]), float(p[2]), float(p[3])
        assert a >= b
        beta = np.radians(90.0 - pa) if flip_g2 else np.radians(90.0 + pa)
        hlr = (a * b) ** 0.5
        g1, g2, mu = lens_params(*obj.lens)
        if kind == "sersic2d":
            table.add_sersic(x, y, flux, hlr, float(p[4]), q=b / a, beta=beta, g1=g1, g2=g2, mu=mu, sed=sed, tanx=tanx,
                             tany=tany)
        else:
            table.add_knots(x, y, flux, hlr, int(p[4]), q=b / a, beta=beta, g1=g1, g2=g2, mu=mu, sed=sed, seed=knot_seed,
                            tanx=tanx, tany=tany)
    else:
        raise RuntimeError("Do not know how to handle object type on the device: %s" % p[0])
    return True


# ---------------------------------------------------------------------- device side
class Stage1:
    """Device tables of one detector's objects and the call that fills a photon batch."""

    def __init__(self, ctx, objects: np.ndarray, sed_cdf: Optional[np.ndarray] = None,
                 sed_wave: Optional[np.ndarray] = None, radial_tables: Optional[np.ndarray] = None, psf=None):
        import torch

        self.ctx = ctx
        self.torch = torch
        dev = "cuda:%d" % ctx.device
        self.device = dev
        self.n_obj = int(objects.size)
        raw = np.ascontiguousarray(objects).view(np.uint8).reshape(-1)
        self.objects = _lib.h2d_async(raw, dev)  # no stream synchronisation: callers queue work ahead of the GPU
        if sed_cdf is not None:
            sed_cdf = np.atleast_2d(np.asarray(sed_cdf, np.float64))
            sed_wave = np.atleast_2d(np.asarray(sed_wave, np.float64))
            if sed_wave.shape[0] == 1 and sed_cdf.shape[0] > 1:
                sed_wave = np.repeat(sed_wave, sed_cdf.shape[0], axis=0)
            assert sed_cdf.shape == sed_wave.shape
            assert objects.size == 0 or int(objects["sed"].max()) < sed_cdf.shape[0]
            self.cdf = _lib.h2d_async(sed_cdf, dev)
            self.cdf_wave = _lib.h2d_async(sed_wave, dev)
        else:
            self.cdf = self.cdf_wave = None
        if radial_tables is not None:
            t = np.ascontiguousarray(radial_tables, dtype=np.float64)
            # the upload replaces a device allocation (it waits for the stream): skip it when the context already
            # holds these very tables, as it does from detector to detector of a visit
            key = (t.shape, hash(t.tobytes()))
            if getattr(ctx, "_radial_luts_key", None) != key:
                _lib.check(_lib.load().b2_radial_luts_upload(ctx.handle, t.ctypes.data, t.shape[0], t.shape[1],
                                                             RADIAL_TMAX))
                ctx._radial_luts_key = key
        elif objects.size and np.any(objects["kind"] == _abi.PROF_RADIAL):
            raise ValueError("objects with radial profiles need radial_tables")
        if psf is not None:
            psf.upload(ctx)

    def shoot(self, dp, counts: np.ndarray, seed: int, photon_offset: int = 0, select: Optional[np.ndarray] = None,
              rand=None):
        """Fill ``dp`` (DevicePhotons with x, y, flux, wavelength) with ``counts[j]`` photons of object j
        (``select``: the object rows ``counts`` refers to; default all).  Nothing here waits for the GPU: the
        small index tables go up through pinned staging, so batches can be queued ahead of the device."""
        counts = np.asarray(counts, dtype=np.int64)
        n = int(counts.sum())
        assert dp.n == n
        cum_h = np.zeros(counts.size + 1, dtype=np.int64)
        np.cumsum(counts, out=cum_h[1:])
        cum = _lib.h2d_async(cum_h, self.device)
        if select is None:
            objs = self.objects
            nobj = self.n_obj
        else:
            sel = _lib.h2d_async(np.asarray(select, dtype=np.int64), self.device)
            objs = self.objects.view(self.n_obj, -1)[sel].contiguous().view(-1)
            nobj = int(sel.shape[0])
        assert counts.size == nobj
        n_sed = 0 if self.cdf is None else int(self.cdf.shape[0])
        ncdf = 0 if self.cdf is None else int(self.cdf.shape[1])
        _lib.check(_lib.load().b2_stage1_photons(
            self.ctx.handle, n, _lib.ptr(dp.x), _lib.ptr(dp.y), _lib.ptr(dp.flux),
            _lib.ptr(dp.wavelength) if self.cdf is not None else None, C.c_void_p(objs.data_ptr()),
            C.c_void_p(cum.data_ptr()), nobj, _lib.ptr(self.cdf), _lib.ptr(self.cdf_wave), n_sed, ncdf,
            _lib.ptr(rand) if rand is not None else None, int(seed) & 0xFFFFFFFFFFFFFFFF, int(photon_offset)))
        self._keep = (cum, objs)  # alive until the stream has consumed them
        return n
