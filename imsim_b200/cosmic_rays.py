"""Cosmic rays painted on the device-resident e-image (mirror of ``imsim.cosmic_rays.CosmicRays``,
imsim/cosmic_rays.py:17-146; called from the ``LSST_CCD`` output, imsim/ccd.py:122-135).

A catalogue is a list of cosmic rays, each a list of spans ``(x0, y0, pixel_values)`` harvested from dark frames.
``paint`` draws the number of hits from a Poisson law of ``exptime * ccd_rate * ccd_frac`` and, per hit, a random
catalogue entry and a random starting pixel -- three uniforms in the reference's order (index, x, y); the span
pixels of all hits are then added in ONE kernel launch (``b2_scatter_add``) with numpy's indexing rules, which
the reference's ``try / except IndexError`` implies: negative indices wrap, only overruns are dropped.
"""
from __future__ import annotations

import ctypes as C
from collections import defaultdict, namedtuple
from typing import Optional

import numpy as np

from . import _lib

CR_Span = namedtuple("CR_Span", "x0 y0 pixel_values".split())


class CosmicRays(list):
    def __init__(self, ccd_rate=None, catalog_file=None):
        super().__init__()
        self.num_pix, self.exptime, self.ccd_rate = 4000 * 4000, 1.0, ccd_rate
        if catalog_file is not None:
            self._read_catalog(catalog_file, ccd_rate)

    @classmethod
    def from_spans(cls, fp_id, x0, y0, pixel_values, span_len=None, exptime=1.0, num_pix=4000 * 4000, ccd_rate=None):
        """Build from the columns of the catalogue table (``write_cosmic_ray_catalog``, cosmic_rays.py:149-185);
        ``pixel_values`` is a list of arrays or, with ``span_len``, their concatenation."""
        self = cls()
        if span_len is not None:
            edges = np.concatenate([[0], np.cumsum(span_len)])
            pixel_values = [np.asarray(pixel_values[edges[k]:edges[k + 1]]) for k in range(len(span_len))]
        crs = defaultdict(list)
        for k, fid in enumerate(fp_id):
            crs[int(fid)].append(CR_Span(int(x0[k]), int(y0[k]), np.asarray(pixel_values[k])))
        self.extend(crs.values())
        self.num_pix, self.exptime = num_pix, exptime
        self.ccd_rate = float(len(self)) / exptime if ccd_rate is None else ccd_rate
        return self

    def _read_catalog(self, catalog_file, ccd_rate, extname="COSMIC_RAYS"):
        import astropy.io.fits as fits  # the catalogue is a FITS binary table (cosmic_rays.py:113-126)

        with fits.open(catalog_file) as catalog:
            cr_cat = catalog[extname]
            self.num_pix = cr_cat.header["NUM_PIX"]
            self.exptime = cr_cat.header["EXPTIME"]
            crs = defaultdict(list)
            for span in cr_cat.data:
                crs[span[0]].append(CR_Span(*tuple(span)[1:]))
        self.extend(crs.values())
        self.ccd_rate = float(len(self)) / self.exptime if ccd_rate is None else ccd_rate

    @classmethod
    def read_catalog(cls, catalog_file, ccd_rate=None, extname="COSMIC_RAYS"):
        ret = cls()
        ret._read_catalog(catalog_file, ccd_rate, extname=extname)
        return ret

    # ------------------------------------------------------------------
    def _hits(self, shape, ud, num_crs):
        """(iy, ix, value) of every span pixel of ``num_crs`` hits; uniforms consumed as in paint_cr."""
        ys, xs, vs = [], [], []
        for _ in range(num_crs):
            index = int(ud() * len(self))
            cr = self[index]
            px, py = int(ud() * shape[1]), int(ud() * shape[0])
            for span in cr:
                n = len(span.pixel_values)
                ys.append(np.full(n, py + span.y0 - cr[0].y0, dtype=np.int32))
                xs.append(px + span.x0 - cr[0].x0 + np.arange(n, dtype=np.int32))
                vs.append(np.asarray(span.pixel_values, dtype=np.float32))
        if not ys:
            z = np.zeros(0, np.int32)
            return z, z, np.zeros(0, np.float32)
        return np.concatenate(ys), np.concatenate(xs), np.concatenate(vs)

    def paint(self, ctx, image, rng, exptime=30.0, num_crs: Optional[int] = None):
        """Add cosmic rays to ``image`` (CUDA tensor [ny][nx], float32 / float64) in place.  ``rng``: numpy Generator
        or a callable returning uniforms in [0, 1)."""
        import torch

        ud = rng.random if isinstance(rng, np.random.Generator) else rng
        ny, nx = image.shape
        if num_crs is None:
            ccd_frac = float(ny * nx) / self.num_pix
            gen = rng if isinstance(rng, np.random.Generator) else np.random.default_rng(int(ud() * 2**31))
            num_crs = int(gen.poisson(exptime * self.ccd_rate * ccd_frac))
        iy, ix, val = self._hits((ny, nx), ud, num_crs)
        if iy.size:
            dev = image.device
            d = [torch.as_tensor(a, device=dev) for a in (iy, ix, val)]
            _lib.check(_lib.load().b2_scatter_add(ctx.handle, C.c_void_p(image.data_ptr()), image.element_size(), nx, ny,
                                                  iy.size, *(C.c_void_p(t.data_ptr()) for t in d)))
            self._keep = d
        return image
