"""Flat-field production on the B200 (mirror of ``LSST_FlatBuilder.addNoise``,
imsim/flat.py:133-281).

The reference builds the CCD in ``nx x ny`` sections with a ``buffer_size``
border, ``niter = ceil(counts / max_counts_per_iter)`` iterations per section:

* no ``sed`` (examples/flat.yaml): ``area = sensor.calculate_pixel_areas(section)``,
  ``temp = base * area / mean(area)``, Poisson realisation, ``section += temp``
  (flat.py:220-237) -- the areas come from ``k_pixel_areas`` and the Poisson realisation from the exact
  counter-based sampler of ``k_add_sky``, the section never leaves HBM (``fused=False``: numpy Poisson on the
  host, the first version);
* ``sed`` given (examples/flat_with_sed.yaml): Poisson number of photons uniform
  over the bordered section, wavelengths from SED x bandpass,
  ``sensor.accumulate(photons, section, resume=(it > 0))`` with
  ``nrecalc = 1e4 * xsize * ysize / (nx * ny)`` (flat.py:96-101,239-264) -- photons
  are generated in HBM (``k_flat_photons``) and never touch the host.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np

import ctypes as C

from . import _lib
from .sensor import Image, SiliconSensor


def flat_nrecalc(xsize, ysize, nx, ny) -> float:
    """imsim/flat.py:96-101: recalc every 10,000 electrons per pixel of a section."""
    return 10_000 * xsize * ysize / (nx * ny)


def wavelength_cdf(wave_nm, weight):
    """Tabulated CDF of ``sed(w) * bandpass(w)`` on ``wave_nm`` (trapezoid), for ``k_flat_photons``."""
    wave_nm = np.ascontiguousarray(wave_nm, dtype=np.float64)
    w = np.asarray(weight, dtype=np.float64)
    c = np.concatenate([[0.0], np.cumsum(0.5 * (w[1:] + w[:-1]) * np.diff(wave_nm))])
    return np.ascontiguousarray(c / c[-1]), wave_nm


def flat_iterations(counts_per_pixel: float, max_counts_per_iter: float):
    """``(niter, counts_per_iter)`` of imsim/flat.py:155-157."""
    niter = int(np.ceil(counts_per_pixel / max_counts_per_iter))
    return niter, counts_per_pixel / niter


def flat_sections(nrow: int, ncol: int, nx: int, ny: int, buffer_size: int, x0: int = 1, y0: int = 1):
    """The section grid of imsim/flat.py:182-213 for an image whose first pixel is ``(x0, y0)``: yields
    ``(i, j, (xmin, xmax, ymin, ymax), (bxmin, bxmax, bymin, bymax))`` in the reference's order (x outer, y inner);
    ``dx = ncol // nx``, the last section of a row / column takes the remainder, the bordered bounds may stick
    out of the image."""
    dx, dy = ncol // nx, nrow // ny
    for i in range(nx):
        xmin = i * dx + x0
        xmax = ncol + x0 - 1 if i == nx - 1 else (i + 1) * dx + x0 - 1
        for j in range(ny):
            ymin = j * dy + y0
            ymax = nrow + y0 - 1 if j == ny - 1 else (j + 1) * dy + y0 - 1
            yield i, j, (xmin, xmax, ymin, ymax), (xmin - buffer_size, xmax + buffer_size, ymin - buffer_size,
                                                   ymax + buffer_size)


MAX_UNFUSED_CHUNK = 1 << 27  # photons per accumulate call of the unfused photon-shot branch (4 GiB of SoA)


def build_flat(image: Image, counts_per_pixel: float, sensor: Optional[SiliconSensor], rng=None,
               max_counts_per_iter: float = 1000.0, nx: int = 8, ny: int = 2, buffer_size: int = 5,
               sed_cdf=None, base_level: Optional[Callable] = None, logger=None, fused: bool = True):
    """Add a flat field of ``counts_per_pixel`` electrons to ``image`` in place
    (``LSST_FlatBuilder.addNoise``).  ``sed_cdf = wavelength_cdf(...)`` selects the
    photon-shot branch; otherwise the pixel-area branch.  ``rng``: numpy Generator / seed.
    ``fused`` (photon branch): generate the photons tile by tile inside the deposit kernel
    (``b2_flat_step``; boundary updates fall on iteration ends) instead of materialising photon
    arrays and calling ``accumulate`` (``fused=False``: the reference's literal sequence).
    Returns the number of photons shot (0 in the area branch)."""
    gen = rng if hasattr(rng, "poisson") else np.random.default_rng(rng)  # a Generator (or a stand-in) / a seed
    niter, counts_per_iter = flat_iterations(counts_per_pixel, max_counts_per_iter)
    nrow, ncol = image.array.shape
    x0, y0 = image.xmin, image.ymin
    tot_nphot = 0
    if sed_cdf is not None:
        import torch

        from .photon_pooling import DevicePhotons

        cdf, cdf_wave = (torch.as_tensor(a, device="cuda:%d" % sensor.ctx.device) for a in sed_cdf)
    for i, j, (xmin, xmax, ymin, ymax), (bx0, bx1, by0, by1) in flat_sections(nrow, ncol, nx, ny, buffer_size, x0, y0):
        # section with border (flat.py:209-213); it may stick out of the image like the reference's
        sec = Image(np.zeros((by1 - by0 + 1, bx1 - bx0 + 1), dtype=image.array.dtype), bx0, by0)
        if sed_cdf is None and sensor is not None and fused:
            # pixel-area branch entirely on the device: the section stays in HBM, each iteration computes
            # the areas from the charge collected so far (b2_sensor_pixel_areas) and adds
            # Poisson(counts * base * area / mean(area)) with the exact counter-based sampler (b2_add_sky)
            import torch

            from .sky import add_sky, pixel_areas_device

            dev = "cuda:%d" % sensor.ctx.device
            tdt = torch.float32 if sec.array.dtype == np.float32 else torch.float64
            sec_dev = torch.zeros(sec.array.shape, dtype=tdt, device=dev)
            mod = None
            if base_level is not None:
                mod = torch.as_tensor(np.ascontiguousarray(base_level(sec), dtype=np.float32), device=dev)
            for it in range(niter):
                _lib.check(_lib.load().b2_sensor_bind_image(sensor._h, sec.xmin, sec.ymin, sec.array.shape[1],
                                                            sec.array.shape[0], sec.array.dtype.itemsize,
                                                            C.c_void_p(sec_dev.data_ptr()), 1))
                sensor._bound_shape = (sec.array.shape[0], sec.array.shape[1], sec.array.dtype)
                areas = pixel_areas_device(sensor, use_flux=True)
                add_sky(sensor.ctx, sec_dev, counts_per_iter / float(areas.mean()), seed=int(gen.integers(1 << 62)),
                        areas=areas, modulation=mod)
            sec.array[:, :] = sec_dev.cpu().numpy()
            niter_host = 0
        else:
            niter_host = niter
        for it in range(niter_host):
            if sed_cdf is None:
                area = sensor.calculate_pixel_areas(sec) if sensor is not None else 1.0
                temp = np.full(sec.array.shape, counts_per_iter, dtype=np.float64)
                if base_level is not None:
                    temp *= base_level(sec)
                if not isinstance(area, float):
                    temp *= area.array / np.mean(area.array)
                sec.array[:, :] += gen.poisson(temp).astype(sec.array.dtype)
            elif fused:
                # tile-ordered generation fused with the deposit (b2_flat_step): per-tile Poisson counts
                if it == 0:
                    tile = 32
                    tnx, tny = -(-sec.array.shape[1] // tile), -(-sec.array.shape[0] // tile)
                    wx = np.minimum(tile, sec.array.shape[1] - tile * np.arange(tnx))
                    wy = np.minimum(tile, sec.array.shape[0] - tile * np.arange(tny))
                    tile_area = np.outer(wy, wx).ravel().astype(np.float64)
                    sensor._bind(sec)
                    sensor._last_image = sec
                    accum = 0.0
                cnt = gen.poisson(counts_per_iter * tile_area)
                cum = np.ascontiguousarray(np.concatenate([[0], np.cumsum(cnt)]), dtype=np.int64)
                nphot = int(cnt.sum())
                accum += nphot
                update_after = int(sensor.nrecalc > 0 and accum >= sensor.nrecalc / sensor.strength)
                if update_after:
                    accum = 0.0
                _lib.check(_lib.load().b2_flat_step(
                    sensor.ctx.handle, sensor._h, C.c_void_p(cum.ctypes.data), nphot, tile,
                    _lib.ptr(cdf), _lib.ptr(cdf_wave), int(cdf.shape[0]), int(gen.integers(1 << 62)),
                    sensor._seed & 0xFFFFFFFFFFFFFFFF, tot_nphot, int(it > 0), update_after, None))
                if it == niter - 1:
                    sensor.read_image(sec)
                tot_nphot += nphot
            else:
                # the reference's literal sequence, in device-memory-bounded chunks of photons
                nphot = int(gen.poisson(counts_per_iter * sec.array.size))
                done = 0
                while done < nphot or (nphot == 0 and done == 0):
                    m = min(nphot - done, MAX_UNFUSED_CHUNK)
                    dp = DevicePhotons(m, device="cuda:%d" % sensor.ctx.device,
                                       fields=("x", "y", "flux", "wavelength"))
                    sensor.ctx.flat_photons(dp.x, dp.y, dp.flux, dp.wavelength,
                                            (bx0 - 0.5, bx1 + 0.5, by0 - 0.5, by1 + 0.5), cdf, cdf_wave,
                                            seed=int(gen.integers(1 << 62)), photon_offset=tot_nphot + done)
                    last = done + m >= nphot
                    sensor.accumulate(dp, sec, resume=(it > 0 or done > 0),
                                      sync_image=(it == niter - 1 and last), want_stats=False)
                    done += m
                    if nphot == 0:
                        break
                tot_nphot += nphot
        # copy just the part that is officially part of this section (flat.py:266-267)
        image.array[ymin - y0:ymax - y0 + 1, xmin - x0:xmax - x0 + 1] += \
            sec.array[buffer_size:buffer_size + (ymax - ymin + 1), buffer_size:buffer_size + (xmax - xmin + 1)]
        if logger is not None:
            logger.info("Done section %d,%d: mean level => %s", i, j,
                        image.array[ymin - y0:ymax - y0 + 1, xmin - x0:xmax - x0 + 1].mean())
    return tot_nphot
