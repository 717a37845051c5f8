"""Per-detector device context: telescope + WCS pair + detector geometry +
diffraction set-up uploaded once, then any number of photon-op calls.

Wraps ``b2_ctx`` of include/imsim_b200.h.  PyTorch only supplies the stream and
device buffers; all arithmetic is in csrc/*.cu.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _abi, _lib
from .detector import DetectorGeometry
from .telescope import Telescope
from .wcs import TanSipWCS


def _stream_handle(stream) -> Optional[int]:
    if stream is None:
        return None
    if hasattr(stream, "cuda_stream"):  # torch.cuda.Stream
        return int(stream.cuda_stream)
    return int(stream)


class OpticsContext:
    def __init__(self, device: int = 0, stream=None):
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self.device = int(device)
        self.stream = stream  # None: the legacy default stream
        _lib.check(self._lib.b2_ctx_create(self.device, _stream_handle(stream), C.byref(self._h)))
        self.telescope = None
        self._keep = []

    # -- uploads ----------------------------------------------------------
    def set_stream(self, stream):
        self.stream = stream
        _lib.check(self._lib.b2_ctx_set_stream(self._h, _stream_handle(stream)))

    def set_telescope(self, tel):
        """``tel``: a :class:`Telescope` or an already flattened ``(B2Telescope, extras)``."""
        pod, extras = tel.flatten() if isinstance(tel, Telescope) else tel
        _lib.check(self._lib.b2_telescope_upload(self._h, C.byref(pod)))
        for i, e in enumerate(extras or []):
            if e is None:
                continue
            kind, arr = e
            arr = np.ascontiguousarray(arr, dtype=np.float64)
            _lib.check(self._lib.b2_telescope_set_extra(self._h, i, kind, arr.ctypes.data, arr.size))
        self.telescope = tel

    @property
    def program(self) -> int:
        """Surface program of the uploaded telescope (0 interpreter, 1 Rubin layout; -1 none)."""
        return int(self._lib.b2_telescope_program(self._h))

    #: pixels around the detector over which XyToV is compiled (stamps of objects just off the chip, the
    #: reach of all but the longest diffraction kicks); positions further out take the exact chain
    XYTOV_MARGIN = 512.0
    XYTOV_TOL_PX = 2e-9

    def set_wcs(self, img_wcs, icrf_to_field):
        a = img_wcs.to_pod() if isinstance(img_wcs, TanSipWCS) else img_wcs
        b = icrf_to_field.to_pod() if isinstance(icrf_to_field, TanSipWCS) else icrf_to_field
        _lib.check(self._lib.b2_wcs_upload(self._h, C.byref(a), C.byref(b)))
        self._have_wcs = True
        self.xytov_residual_px = None
        self._compile_xytov()

    def set_detector(self, det):
        d = det.to_pod() if isinstance(det, DetectorGeometry) else det
        _lib.check(self._lib.b2_detector_upload(self._h, C.byref(d)))
        self._det_box = None
        if isinstance(det, DetectorGeometry):
            m = self.XYTOV_MARGIN
            self._det_box = (det.xmin - m, det.xmin + det.nx + m, det.ymin - m, det.ymin + det.ny + m)
        self._compile_xytov()

    def _compile_xytov(self):
        """XyToV as one polynomial over this detector (``b2_xytov_compile``), once both the WCS pair and the
        detector's pixel box are known; ``B2_XYTOV_EXACT=1`` keeps the exact chain.  ``xytov_residual_px``: the
        largest deviation from the exact chain on the check grid (None: not compiled)."""
        import os

        box = getattr(self, "_det_box", None)
        if not getattr(self, "_have_wcs", False) or box is None or os.environ.get("B2_XYTOV_EXACT") == "1":
            return
        r = C.c_double(0.0)
        _lib.check(self._lib.b2_xytov_compile(self._h, box[0], box[1], box[2], box[3], self.XYTOV_TOL_PX, C.byref(r)))
        self.xytov_residual_px = r.value

    def set_diffraction(self, cfg: Optional[_abi.B2Diffraction]):
        if cfg is None:
            cfg = _abi.B2Diffraction()
        _lib.check(self._lib.b2_diffraction_config(self._h, C.byref(cfg)))

    def synchronize(self):
        _lib.check(self._lib.b2_ctx_synchronize(self._h))

    def record_kernel_events(self, on=True):
        """Bracket the main kernel of pool_step / rubin_optics with CUDA events (roofline timing)."""
        _lib.check(self._lib.b2_ctx_record_kernel_events(self._h, int(bool(on))))

    def kernel_ms(self):
        """(total_ms, launches) of the bracketed kernels since the last call."""
        ms, cnt = C.c_double(0.0), C.c_int64(0)
        _lib.check(self._lib.b2_ctx_kernel_ms(self._h, C.byref(ms), C.byref(cnt)))
        return ms.value, cnt.value

    def fma_peak(self, fp64=True) -> float:
        """Measured non-tensor FMA ceiling in TFLOP/s (roofline denominator of the trace kernel)."""
        out = C.c_double(0.0)
        _lib.check(self._lib.b2_fma_peak(self._h, int(bool(fp64)), C.byref(out)))
        return out.value

    def atomic_peak(self, fp64=True, pattern=0, nx=4096, ny=4004, n=1 << 26) -> float:
        """Measured atomicAdd throughput [atomics/s] on an image-sized buffer (0 uniform, 1 hot pixel, 2 stars)."""
        out = C.c_double(0.0)
        _lib.check(self._lib.b2_atomic_peak(self._h, int(bool(fp64)), int(pattern), nx, ny, int(n), C.byref(out)))
        return out.value

    # -- photon ops ---------------------------------------------------------
    def xy_to_v(self, x, y, out=None):
        where = _lib.where_of(x)
        n = x.shape[0]
        if out is None:
            out = tuple(_empty_like(x) for _ in range(3))
        _lib.check(self._lib.b2_xy_to_v(self._h, n, _lib.ptr(x), _lib.ptr(y), _lib.ptr(out[0]), _lib.ptr(out[1]),
                                        _lib.ptr(out[2]), where))
        return out

    def v_to_xy(self, vx, vy, vz, out=None):
        where = _lib.where_of(vx)
        n = vx.shape[0]
        if out is None:
            out = tuple(_empty_like(vx) for _ in range(2))
        _lib.check(self._lib.b2_v_to_xy(self._h, n, _lib.ptr(vx), _lib.ptr(vy), _lib.ptr(vz), _lib.ptr(out[0]),
                                        _lib.ptr(out[1]), where))
        return out

    def trace_rays(self, x, y, z, vx, vy, vz, t, wavelength_m, vignetted, failed):
        """In-place ``batoid.Optic.trace`` of rays given in the stop surface's frame."""
        where = _lib.where_of(x)
        _lib.check(self._lib.b2_trace_rays(self._h, x.shape[0], _lib.ptr(x), _lib.ptr(y), _lib.ptr(z), _lib.ptr(vx),
                                           _lib.ptr(vy), _lib.ptr(vz), _lib.ptr(t), _lib.ptr(wavelength_m),
                                           _lib.ptr(vignetted, np.uint8), _lib.ptr(failed, np.uint8), where))

    def rubin_optics(self, x, y, dxdz, dydz, flux, wavelength, pupil_u, pupil_v, time, gauss=None, time_out=None,
                     options: Optional[_abi.B2OpticsOptions] = None, want_stats=True):
        where = _lib.where_of(x)
        opt = options if options is not None else _abi.B2OpticsOptions()
        stats = _abi.B2OpticsStats() if want_stats else None
        _lib.check(self._lib.b2_rubin_optics(
            self._h, x.shape[0], _lib.ptr(x), _lib.ptr(y), _lib.ptr(dxdz), _lib.ptr(dydz), _lib.ptr(flux),
            _lib.ptr(wavelength), _lib.ptr(pupil_u), _lib.ptr(pupil_v), _lib.ptr(time), _lib.ptr(gauss),
            _lib.ptr(time_out), C.byref(opt), where, C.byref(stats) if want_stats else None))
        return stats

    def rubin_diffraction(self, x, y, wavelength, pupil_u, pupil_v, time, gauss=None,
                          options: Optional[_abi.B2OpticsOptions] = None):
        where = _lib.where_of(x)
        opt = options if options is not None else _abi.B2OpticsOptions()
        _lib.check(self._lib.b2_rubin_diffraction(
            self._h, x.shape[0], _lib.ptr(x), _lib.ptr(y), _lib.ptr(wavelength), _lib.ptr(pupil_u),
            _lib.ptr(pupil_v), _lib.ptr(time), _lib.ptr(gauss), C.byref(opt), where))

    def sample_time_pupil(self, time, pupil_u, pupil_v, t0, exptime, r_inner, r_outer, seed, photon_offset=0):
        ref = time if time is not None else pupil_u
        where = _lib.where_of(ref)
        _lib.check(self._lib.b2_sample_time_pupil(self._h, ref.shape[0], _lib.ptr(time), _lib.ptr(pupil_u),
                                                  _lib.ptr(pupil_v), t0, exptime, r_inner, r_outer, seed,
                                                  photon_offset, where))

    def flat_photons(self, x, y, flux, wavelength, bounds, cdf=None, cdf_wave=None, seed=0, photon_offset=0):
        """Uniform unit-flux photons over ``bounds = (xlo, xhi, ylo, yhi)`` (+ wavelengths from an
        inverse-CDF table) written straight into device arrays (imsim/flat.py:239-259)."""
        ncdf = 0 if cdf is None else int(cdf.shape[0])
        _lib.check(self._lib.b2_flat_photons(self._h, x.shape[0], _lib.ptr(x), _lib.ptr(y), _lib.ptr(flux),
                                             _lib.ptr(wavelength), float(bounds[0]), float(bounds[1]),
                                             float(bounds[2]), float(bounds[3]), _lib.ptr(cdf), _lib.ptr(cdf_wave),
                                             ncdf, int(seed), int(photon_offset)))

    def object_photons(self, x, y, flux, wavelength, obj_x, obj_y, obj_sigma, obj_cum, cdf=None, cdf_wave=None,
                       seed=0, photon_offset=0):
        """Pooled photons of point-like objects behind a Gaussian PSF, generated in HBM
        (``obj_cum``: int64 CUDA tensor of nobj+1 cumulative photon counts)."""
        ncdf = 0 if cdf is None else int(cdf.shape[0])
        _lib.check(self._lib.b2_object_photons(
            self._h, x.shape[0], _lib.ptr(x), _lib.ptr(y), _lib.ptr(flux), _lib.ptr(wavelength), _lib.ptr(obj_x),
            _lib.ptr(obj_y), _lib.ptr(obj_sigma), C.c_void_p(obj_cum.data_ptr()), int(obj_x.shape[0]), _lib.ptr(cdf),
            _lib.ptr(cdf_wave), ncdf, int(seed), int(photon_offset)))

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            self._lib.b2_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _empty_like(a):
    if _lib._is_torch(a):
        import torch

        return torch.empty_like(a)
    return np.empty_like(a)
