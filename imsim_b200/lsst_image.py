"""The classic per-object pipeline on the device: ``LSST_Image`` +
``LSST_Silicon`` stamps (imsim/lsst_image.py:266-395, imsim/stamp.py:251-572).

Reference sequence per object: stamp size from the flux (stamp.py:188-221, stamp_utils.py); the profile is
drawn with ``drawImage(method='phot', maxN=1e6, sensor=sensor, photon_ops=psfs + photon_ops)`` in chunks of
``maxN`` photons -- each chunk is shot, passed through the photon ops (``RubinOptics`` shifts the photons from
stamp to image coordinates and back, lsst_image.py:324-335) and accumulated on the object's *own stamp*
with ``resume=(chunk > 0)``, so brighter-fatter sees only this object's charge and the pixel
boundaries are recomputed every ``nrecalc`` electrons; objects below ``max_flux_simple`` get neither
optics nor silicon (stamp.py:534-537,555-556); finally ``full_image[bounds] += stamp[bounds]``
(lsst_image.py:359-368).

Here the same sequence runs on HBM-resident data, for all objects at once (``build``): stage 1 generates the
photons of every object in one launch (``b2_stage1_photons``), the samplers + fused optics kernel run on the
photons of all non-faint objects, and ``b2_sensor_accumulate_stamps`` gives every object its own zero stamp
with fresh boundaries, accumulates its photons at the ``nrecalc`` cadence inside the stamp -- the chunk loop
runs inside the kernel, one thread block per stamp -- and adds the stamp to the device-resident full image.
``build_per_object`` is the host-driven loop (one bind / optics / accumulate per object) the batched form
replaced; both draw identical stamps from identical photons (tests/test_gpu_stamps.py).  Photons stay in
full-image coordinates throughout, which is what the reference's shift / unshift pair amounts to.  FFT-rendered objects (stamp.py:467-513) are out of
scope: objects brighter than ``fft_flux_limit`` are still photon shot.
"""
from __future__ import annotations

import os
import time
from typing import Optional

import numpy as np

from . import _lib
from .photon_pooling import DevicePhotons, PhotonPool
from .sensor import Image
from .stage1 import Stage1
from .sensor import STAMP_JOB_DTYPE
from .stamp_utils import get_stamp_size, get_stamp_sizes


class ClassicImageBuilder:
    """``nbatch`` (checkpoint cadence) only groups objects in the reference; images do not depend on it."""

    def __init__(self, ctx, sensor, objects, radial_tables=None, sersic_n=None, sed_cdf=None, sed_wave=None,
                 psf=None, arcsec_to_pix=None, noise_var: float = 800.0, exptime: float = 30.0, band: str = "r",
                 airmass: Optional[float] = None, rawSeeing: Optional[float] = None, maxN: int = int(1e6),
                 max_flux_simple: float = 100.0, focus_depth: float = 0.0, seed: int = 1):
        self.ctx, self.sensor = ctx, sensor
        self.rows = objects
        self.radial_tables, self.sersic_n = radial_tables, sersic_n
        self.arcsec_to_pix = arcsec_to_pix
        self.noise_var, self.band, self.airmass, self.rawSeeing = noise_var, band, airmass, rawSeeing
        self.maxN, self.max_flux_simple = int(maxN), max_flux_simple
        self.stage1 = Stage1(ctx, objects, sed_cdf, sed_wave, radial_tables)
        if psf is not None:
            psf.upload(ctx, arcsec_to_pix)
        self.pool = PhotonPool(ctx, sensor, exptime=exptime, focus_depth=focus_depth, seed=seed)
        self.seed = seed
        from .atmosphere import WLEN_EFF

        self.wlen_eff = WLEN_EFF[band]
        self.stats = {}

    def stamp_bounds(self, j: int, nominal_flux: float):
        """(xmin, ymin, size): stamp centred on the pixel containing the object (galsim stamps are centred at
        ``floor(pos + 0.5)``; even sizes put the centre at (min + max + 1) / 2)."""
        r = self.rows[j]
        size = get_stamp_size(r, nominal_flux, self.noise_var, airmass=self.airmass, rawSeeing=self.rawSeeing,
                              band=self.band, radial_tables=self.radial_tables, sersic_n=self.sersic_n,
                              arcsec_to_pix=self.arcsec_to_pix)
        icx, icy = int(np.floor(r["x"] + 0.5)), int(np.floor(r["y"] + 0.5))
        return icx - size // 2, icy - size // 2, size

    def _to_device(self, arr, dev):
        """The caller's (pageable) full image as a device tensor, through the pinned ring."""
        import torch

        full = torch.empty(arr.shape, dtype=torch.float32 if arr.dtype == np.float32 else torch.float64, device=dev)
        if arr.flags.c_contiguous and arr.nbytes % 8 == 0 and arr.dtype in (np.float32, np.float64):
            _lib.check(_lib.load().b2_copy_through_ring(self.ctx.handle, arr.ctypes.data, full.data_ptr(), arr.nbytes, 1))
        else:
            full.copy_(torch.as_tensor(np.ascontiguousarray(arr)))
        return full

    def _to_host(self, full, arr):
        if arr.flags.c_contiguous and arr.nbytes % 8 == 0 and arr.dtype in (np.float32, np.float64):
            _lib.check(_lib.load().b2_copy_through_ring(self.ctx.handle, arr.ctypes.data, full.data_ptr(), arr.nbytes, 0))
        else:
            arr[:, :] = full.cpu().numpy()

    def all_stamp_bounds(self, nominal_flux):
        """``stamp_bounds`` of every catalogue row at once: arrays (xmin, ymin, size)."""
        size = get_stamp_sizes(self.rows, nominal_flux, self.noise_var, airmass=self.airmass,
                               rawSeeing=self.rawSeeing, band=self.band, radial_tables=self.radial_tables,
                               sersic_n=self.sersic_n, arcsec_to_pix=self.arcsec_to_pix)
        icx = np.floor(self.rows["x"] + 0.5).astype(np.int64)
        icy = np.floor(self.rows["y"] + 0.5).astype(np.int64)
        return icx - size // 2, icy - size // 2, size

    #: photons per group of objects handed to the device at once (9 float64 arrays each)
    GROUP_PHOTONS = 1 << 27

    def build(self, image: Image, nominal_flux, phot_flux=None, rng=None):
        """Draw every object onto ``image`` (added in place), all objects of a group in three launches.
        ``phot_flux``: Poisson realisation of the fluxes (stamp.py:194-196), drawn here from ``rng`` if not
        given.  Returns a record of counts."""
        import torch

        ctx, sensor = self.ctx, self.sensor
        dev = "cuda:%d" % ctx.device
        gen = rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)
        nominal_flux = np.asarray(nominal_flux, dtype=np.float64)
        if phot_flux is None:
            phot_flux = gen.poisson(nominal_flux)
        phot_flux = np.asarray(phot_flux, dtype=np.int64)
        arr = image.array
        X0, Y0 = image.xmin, image.ymin
        ny, nx = arr.shape
        t0 = time.perf_counter()
        # host: which objects are drawn, on which stamps (SkipThisObject / off-image stamps, lsst_image.py:361-366)
        xmin, ymin, size = self.all_stamp_bounds(nominal_flux)
        on_image = (np.maximum(xmin, X0) < np.minimum(xmin + size, X0 + nx)) & \
                   (np.maximum(ymin, Y0) < np.minimum(ymin + size, Y0 + ny))
        drawn = np.flatnonzero((phot_flux > 0) & on_image)
        is_faint = nominal_flux[drawn] < self.max_flux_simple
        n_skipped = int(self.rows.size - drawn.size)
        n_faint = int(is_faint.sum())
        t_host = time.perf_counter() - t0
        phases = {}
        profile = bool(os.environ.get("B2_CLASSIC_PROFILE"))  # synchronises after every phase

        def lap(name, t_start):
            if profile:
                torch.cuda.synchronize()
                phases[name] = phases.get(name, 0.0) + time.perf_counter() - t_start
            return time.perf_counter()

        t = time.perf_counter()
        full = self._to_device(arr, dev)
        t = lap("upload", t)
        n_photons = 0
        # groups of consecutive objects; non-faint objects first inside a group, so that the optics run on one
        # contiguous range
        cum = np.concatenate([[0], np.cumsum(phot_flux[drawn])])
        k = 0
        while k < drawn.size:
            k1 = max(int(np.searchsorted(cum, cum[k] + self.GROUP_PHOTONS, side="right")) - 1, k + 1)
            order = np.argsort(is_faint[k:k1], kind="stable") + k
            sel = drawn[order]
            counts = phot_flux[sel]
            tot = int(counts.sum())
            n_opt = int(counts[~is_faint[order]].sum())
            dp = DevicePhotons(tot, device=dev)
            self.stage1.shoot(dp, counts, seed=self.seed, photon_offset=n_photons, select=sel)
            if self.stage1.cdf is None:
                dp.wavelength.fill_(self.wlen_eff)  # monochromatic at the band's effective wavelength
            t = lap("stage1", t)
            if n_opt:
                self.pool.trace(dp, n_opt)
            t = lap("optics", t)
            dp._has.update(dxdz=True, dydz=True)
            jobs = np.zeros(sel.size, dtype=STAMP_JOB_DTYPE)
            jobs["p0"][1:] = np.cumsum(counts)[:-1]
            jobs["n"] = counts
            jobs["xmin"], jobs["ymin"] = xmin[sel], ymin[sel]
            jobs["nx"] = jobs["ny"] = size[sel]
            jobs["plain"] = is_faint[order]
            if n_opt < tot:  # faint objects carry no slopes: the plain jobs do not read them
                dp.dxdz[n_opt:].zero_()
                dp.dydz[n_opt:].zero_()
            sensor.accumulate_stamps(jobs, dp, full, X0, Y0, want_stats=False)
            t = lap("stamps", t)
            n_photons += tot
            k = k1
        self._to_host(full, arr)
        lap("read_back", t)
        self.stats = {"phot": int(drawn.size) - n_faint, "faint": n_faint, "skipped": n_skipped,
                      "photons": n_photons, "seconds": time.perf_counter() - t0, "host_setup_seconds": t_host}
        if profile:
            self.stats["phases"] = phases
        return self.stats

    def build_per_object(self, image: Image, nominal_flux, phot_flux=None, rng=None):
        """The same drawing, driven object by object from the host (one bind / optics / accumulate per object)."""
        import torch

        ctx, sensor = self.ctx, self.sensor
        dev = "cuda:%d" % ctx.device
        gen = rng if isinstance(rng, np.random.Generator) else np.random.default_rng(rng)
        nominal_flux = np.asarray(nominal_flux, dtype=np.float64)
        if phot_flux is None:
            phot_flux = gen.poisson(nominal_flux)
        phot_flux = np.asarray(phot_flux, dtype=np.int64)
        arr = image.array
        tdt = torch.float32 if arr.dtype == np.float32 else torch.float64
        full = torch.as_tensor(arr, device=dev).clone()
        X0, Y0 = image.xmin, image.ymin
        ny, nx = arr.shape
        token = object()
        n_phot = n_faint = n_skipped = n_photons = 0
        t0 = time.perf_counter()
        for j in range(self.rows.size):
            n = int(phot_flux[j])
            if n == 0:
                n_skipped += 1  # SkipThisObject('phot_flux=0'), stamp.py:205-208
                continue
            xmin, ymin, size = self.stamp_bounds(j, float(nominal_flux[j]))
            # overlap with the full image; stamps entirely off the image are skipped (lsst_image.py:361-366)
            ox0, ox1 = max(xmin, X0), min(xmin + size, X0 + nx)
            oy0, oy1 = max(ymin, Y0), min(ymin + size, Y0 + ny)
            if ox0 >= ox1 or oy0 >= oy1:
                n_skipped += 1
                continue
            sensor.bind_stamp(xmin, ymin, size, size, dtype=arr.dtype)
            faint = nominal_flux[j] < self.max_flux_simple
            sel = np.array([j])
            done = 0
            while done < n:
                m = min(self.maxN, n - done)
                dp = DevicePhotons(m, device=dev)
                self.stage1.shoot(dp, np.array([m]), seed=self.seed + 31 * j, photon_offset=done, select=sel)
                if self.stage1.cdf is None:
                    dp.wavelength.fill_(self.wlen_eff)  # monochromatic at the band's effective wavelength
                if faint:
                    sensor.plain_accumulate_bound(dp)
                else:
                    self.pool.process(dp, token, resume=(done > 0), recalc=False, fused=False, prebound=True)
                done += m
            stamp = torch.empty((size, size), dtype=tdt, device=dev)
            sensor.snapshot_image(stamp)
            full[oy0 - Y0:oy1 - Y0, ox0 - X0:ox1 - X0] += stamp[oy0 - ymin:oy1 - ymin, ox0 - xmin:ox1 - xmin]
            n_photons += n
            if faint:
                n_faint += 1
            else:
                n_phot += 1
        arr[:, :] = full.cpu().numpy()
        self.stats = {"phot": n_phot, "faint": n_faint, "skipped": n_skipped, "photons": n_photons,
                      "seconds": time.perf_counter() - t0}
        return self.stats
