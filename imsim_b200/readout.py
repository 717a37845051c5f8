"""Post-path electronics on the device (mirror of ``imsim/readout.py`` ``CcdReadout.build_amp_images`` and
``imsim/bleed_trails.py``): the e-image goes from the sensor's device buffer to int32 amplifier segments
without a host round trip.

Amplifier geometry comes from the camera (``imsim/camera.py:19-160`` wraps lsst.obs.lsst); here it is a list of
``Amp`` records -- ``lsstcam_like_amps`` gives the 16-segment layout of an LSSTCam science CCD (2 rows of 8,
512 x 2002 imaging pixels per e2v segment in a 576 x 2048 raw segment, 509 x 2000 for ITL) with the flips of
``prepare_hdus`` (readout.py:486-520); a live ``imsim.camera.Camera`` can be flattened with ``amps_from_camera``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _abi, _lib


@dataclass
class Amp:
    name: str
    x0: int
    y0: int
    nx: int
    ny: int
    raw_nx: int
    raw_ny: int
    data_x0: int
    data_y0: int
    flip_x: bool
    flip_y: bool
    gain: float = 1.5
    bias_level: float = 1000.0
    read_noise: float = 5.0

    def to_pod(self) -> _abi.B2Amp:
        return _abi.B2Amp(self.x0, self.y0, self.nx, self.ny, self.raw_nx, self.raw_ny, self.data_x0, self.data_y0,
                          int(self.flip_x), int(self.flip_y), self.gain, self.bias_level, self.read_noise)


def lsstcam_like_amps(vendor: str = "e2v", gain=1.5, bias_level=1000.0, read_noise=5.0) -> List[Amp]:
    """16 segments in the HDU order of ``prepare_hdus`` (C10..C17, C07..C00)."""
    if vendor.lower().startswith("e2v"):
        nx, ny, pre, raw_nx, raw_ny = 512, 2002, 10, 576, 2048
    else:
        nx, ny, pre, raw_nx, raw_ny = 509, 2000, 3, 576, 2048
    amps = []
    for k in range(8):  # top row C10..C17, left to right
        amps.append(Amp("C1%d" % k, k * nx, ny, nx, ny, raw_nx, raw_ny, pre, 0, flip_x=vendor.lower().startswith("e2v"),
                        flip_y=True, gain=gain, bias_level=bias_level, read_noise=read_noise))
    for k in range(7, -1, -1):  # bottom row C07..C00
        amps.append(Amp("C0%d" % k, k * nx, 0, nx, ny, raw_nx, raw_ny, pre, 0, flip_x=True, flip_y=False, gain=gain,
                        bias_level=bias_level, read_noise=read_noise))
    return amps


def amps_from_camera(ccd) -> List[Amp]:
    """Flatten an ``imsim.camera.CCD`` (dict of ``Amp`` with galsim.BoundsI fields)."""
    out = []
    for name, a in ccd.items():
        b, rb, db = a.bounds, a.raw_bounds, a.raw_data_bounds
        out.append(Amp(name, b.xmin - 1 if b.xmin >= 1 else b.xmin, b.ymin - 1 if b.ymin >= 1 else b.ymin,
                       b.xmax - b.xmin + 1, b.ymax - b.ymin + 1, rb.xmax - rb.xmin + 1, rb.ymax - rb.ymin + 1,
                       db.xmin - rb.xmin, db.ymin - rb.ymin, bool(a.raw_flip_x), bool(a.raw_flip_y), float(a.gain),
                       float(a.bias_level), float(a.read_noise)))
    return out


def cte_band(npix: int, cti: float, ntransfers: int = 20) -> np.ndarray:
    """Band form of ``cte_matrix`` (imsim/readout.py:153-203): ``band[i, k] = matrix[i, i - k]`` -- the same
    expressions, evaluated with ``scipy.special.binom`` like the reference, without the npix x npix zeros."""
    import scipy.special

    band = np.zeros((npix, ntransfers + 1))
    for i in range(1, npix + 1):
        band[i - 1, 0] = (1.0 - cti) ** i
        jmin = max(1, i - ntransfers)
        j = np.arange(jmin, i)
        band[i - 1, i - j] = scipy.special.binom(i - 1, i - j) * (1.0 - cti) ** j * cti ** (i - j)
    return np.ascontiguousarray(band)


def bleed_eimage(ctx, eimage: np.ndarray, full_well: float, midline_stop: bool = True) -> np.ndarray:
    """``imsim.bleed_trails.bleed_eimage`` for a host float32 array (modified in place and returned)."""
    if eimage.dtype != np.float32 or not eimage.flags.c_contiguous:
        raise _lib.B2Error("bleed_eimage needs a C-contiguous float32 e-image")
    ny, nx = eimage.shape
    _lib.check(_lib.load().b2_bleed_trails(ctx.handle, eimage.ctypes.data, nx, ny, float(full_well), int(midline_stop),
                                           _abi.B2_HOST))
    return eimage


class CcdReadout:
    """``CcdReadout(eimage, ...).build_amp_images(rng)`` (imsim/readout.py:325-480) with the same keyword
    defaults; ``eimage`` is a CUDA float32 tensor [ny][nx] (e.g. ``SiliconSensor.snapshot_image``) or a host
    array (uploaded once).  ``build_amp_images`` returns an int32 CUDA tensor [namp][raw_ny][raw_nx]."""

    def __init__(self, ctx, amps: Sequence[Amp], readout_time=2.0, dark_current=0.02, bias_level=1000.0, scti=1.0e-6,
                 pcti=1.0e-6, full_well: Optional[float] = 1.0e5, read_noise: Optional[float] = None, xtalk=None,
                 exptime=30.0, midline_stop=True, ntransfers=20):
        self.ctx = ctx
        self.amps = list(amps)
        self.exptime, self.readout_time, self.dark_current = exptime, readout_time, dark_current
        self.full_well, self.midline_stop, self.ntransfers = full_well, midline_stop, ntransfers
        pods = []
        for a in self.amps:
            p = a.to_pod()
            if bias_level is not None:
                p.bias_level = float(bias_level)
            if read_noise is not None:
                p.read_noise = float(read_noise)
            pods.append(p)
        self._pods = (_abi.B2Amp * len(pods))(*pods)
        raw_nx, raw_ny = self.amps[0].raw_nx, self.amps[0].raw_ny
        # the tables b2_readout copies to the device on every call live in pinned host memory: a cudaMemcpyAsync
        # from pageable memory of this size waits for the stream, which would stop the caller from queueing the
        # next detector behind this one (visit.DetectorRunner.run_many)
        self.pband = None if pcti == 0 else self._pin(cte_band(raw_ny, pcti, ntransfers))
        self.sband = None if scti == 0 else self._pin(cte_band(raw_nx, scti, ntransfers))
        self.xtalk = None if xtalk is None else self._pin(np.ascontiguousarray(xtalk, dtype=np.float64))
        self.shape = (len(pods), raw_ny, raw_nx)

    def _pin(self, arr: np.ndarray) -> np.ndarray:
        import torch

        t = torch.from_numpy(np.ascontiguousarray(arr)).pin_memory()
        self._pinned = getattr(self, "_pinned", []) + [t]  # owns the memory of the returned view
        return t.numpy()

    def build_amp_images(self, eimage, seed: int = 0, want_segments: bool = False):
        import torch

        dev = "cuda:%d" % self.ctx.device
        e = eimage if hasattr(eimage, "data_ptr") else torch.as_tensor(np.ascontiguousarray(eimage, np.float32), device=dev)
        assert e.dtype == torch.float32 and e.is_contiguous()
        ny, nx = e.shape
        raw = torch.empty(self.shape, dtype=torch.int32, device=dev)
        seg = torch.empty(self.shape, dtype=torch.float32, device=dev) if want_segments else None
        dark_mean = self.dark_current * (self.exptime + self.readout_time)
        _lib.check(_lib.load().b2_readout(
            self.ctx.handle, C.c_void_p(e.data_ptr()), nx, ny, self._pods, len(self.amps),
            self.xtalk.ctypes.data if self.xtalk is not None else None,
            self.pband.ctypes.data if self.pband is not None else None,
            self.sband.ctypes.data if self.sband is not None else None, self.ntransfers,
            float(self.full_well or 0.0), int(self.midline_stop), float(dark_mean), int(seed) & 0xFFFFFFFFFFFFFFFF,
            C.c_void_p(seg.data_ptr()) if seg is not None else None, C.c_void_p(raw.data_ptr())))
        self.eimage = e
        return (raw, seg) if want_segments else raw
