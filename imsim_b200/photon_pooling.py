"""Photon pooling on the B200 (mirror of imsim/photon_pooling.py).

Two layers:

* the batching algebra of ``LSST_PhotonPoolingImageBuilder`` --
  ``partition_objects``, ``make_batches``, ``make_photon_batches``,
  ``make_photon_subbatches``, ``merge_photon_arrays``, ``accumulate_photons``
  (imsim/photon_pooling.py:177-386) -- reproduced exactly as static methods of
  the same class name, so the reference's tests/test_photon_pooling.py reads the
  same against this module;

* :class:`PhotonPool`, the device-resident replacement for the hot loop of
  ``buildImage`` (imsim/photon_pooling.py:141-160): the merged pool is copied to
  HBM once, TimeSampler + PupilAnnulusSampler, RubinDiffractionOptics, FocusDepth,
  Refraction run as two kernels on it and ``SiliconSensor.accumulate`` consumes the
  device arrays directly; the image stays on the device until a checkpoint or the
  end of ``buildImage`` asks for it.
"""
from __future__ import annotations

import dataclasses
import itertools
import warnings
from dataclasses import dataclass
from enum import Enum, auto
from typing import List, Optional

import numpy as np

from . import _abi, _lib
from .photon_array import PhotonArray


class ProcessingMode(Enum):
    """imsim/stamp.py:17-20"""
    FFT = auto()
    PHOT = auto()
    FAINT = auto()


@dataclass
class ObjectInfo:
    """Per-object cache of the rendering decision (imsim/stamp.py:23-34)."""
    index: int
    phot_flux: float
    mode: ProcessingMode


def _uniform_deviate(config, base, logger):
    """``galsim.UniformDeviate(GetRNG(config, base, logger, "LSST_Silicon"))`` when GalSim is
    present; otherwise a numpy generator seeded from ``base['rng']`` / ``base['random_seed']``."""
    try:
        import galsim  # noqa: PLC0415

        rng = galsim.config.GetRNG(config, base, logger, "LSST_Silicon")
        return galsim.UniformDeviate(rng)
    except ImportError:
        rng = (base or {}).get('rng', None)
        if callable(rng):
            return rng
        gen = np.random.default_rng((base or {}).get('random_seed', None) if rng is None else rng)
        return gen.random


class LSST_PhotonPoolingImageBuilder:
    """Pools photons from all objects in ``nbatch`` batches; photons from faint
    objects only appear in one of the batches randomly.  Static batching helpers
    of the reference builder (the GalSim-facing ``setup`` / ``buildImage`` live in
    galsim_plugin.py because they need the GalSim config machinery)."""

    @staticmethod
    def merge_photon_arrays(stamps):
        """Merge the photon arrays of a list of stamps into one PhotonArray
        (imsim/photon_pooling.py:177-192)."""
        n_tot = sum(len(stamp.photons) for stamp in stamps)
        cls = type(stamps[0].photons) if stamps else PhotonArray
        merged = cls(n_tot)
        start = 0
        for stamp in stamps:
            merged.copyFrom(stamp.photons, slice(start, start + stamp.photons.size()))
            start += len(stamp.photons)
        return merged

    @staticmethod
    def accumulate_photons(photons, image, sensor, resume=False, recalc=True):
        """Accumulate a photon array onto a sensor (imsim/photon_pooling.py:195-225)."""
        from .sensor import Image, SiliconSensor  # noqa: PLC0415

        if image.dtype in (np.float32, np.float64):
            if isinstance(sensor, SiliconSensor):
                sensor.accumulate(photons, image, resume=resume, recalc=recalc)
            else:
                sensor.accumulate(photons, image, resume=resume)
        else:
            # integer image: work in a temporary float64 image; resume / recalc do nothing
            if resume:
                warnings.warn("Sensor is not a float. Using temporary ImageD and ignoring resume = True for "
                              "photon accumulation.")
            if not recalc:
                warnings.warn("Sensor is not a float. Using temporary ImageD and ignoring recalc = False for "
                              "photon accumulation.")
            b = image.bounds
            im1 = Image(np.zeros(image.array.shape, dtype=np.float64), int(b.xmin), int(b.ymin))
            sensor.accumulate(photons, im1)
            image.array[:, :] += im1.array.astype(image.array.dtype)

    @staticmethod
    def make_batches(objects, nbatch: int):
        """Yield ``nbatch`` batches of objects; early batches get the remainder
        (imsim/photon_pooling.py:227-247)."""
        base_per_batch = len(objects) // nbatch
        per_batch_remainder = len(objects) % nbatch
        o_iter = iter(objects)
        for i in range(nbatch):
            nobj_per_batch = base_per_batch + 1 if i < per_batch_remainder else base_per_batch
            yield [obj for _, obj in zip(range(nobj_per_batch), o_iter)]

    @staticmethod
    def make_photon_batches(config, base, logger, phot_objects: List[ObjectInfo],
                            faint_objects: List[ObjectInfo], nbatch: int):
        """``nbatch`` copies of the bright objects at ``(f(i+1))//nbatch - (f i)//nbatch`` of
        their flux, faint objects whole into a random batch (imsim/photon_pooling.py:278-313)."""
        if not phot_objects and not faint_objects:
            return []
        batches = [
            [dataclasses.replace(obj, phot_flux=(obj.phot_flux * (i + 1)) // nbatch - (obj.phot_flux * i) // nbatch)
             for obj in phot_objects]
            for i in range(nbatch)]
        ud = _uniform_deviate(config, base, logger)
        for obj in faint_objects:
            batch_index = int(ud() * nbatch)
            batches[batch_index].append(obj)
        return batches

    @staticmethod
    def make_photon_subbatches(batch, nsubbatch):
        """Split a batch into ``nsubbatch`` nearly equal consecutive sub-batches
        (imsim/photon_pooling.py:315-331)."""
        nobj = len(batch)
        nobj_per_subbatch, nobj_extra = divmod(nobj, nsubbatch)
        section_sizes = nobj_extra * [nobj_per_subbatch + 1] + (nsubbatch - nobj_extra) * [nobj_per_subbatch]
        section_indices = [0] + list(itertools.accumulate(section_sizes))
        return [batch[section_indices[i]:section_indices[i + 1]] for i in range(nsubbatch)]

    @staticmethod
    def stamp_bounds(stamp, full_image_bounds):
        """Overlap of a stamp with the full image or None (imsim/photon_pooling.py:333-353)."""
        if stamp is None:
            return None
        bounds = stamp.bounds & full_image_bounds
        if not bounds.isDefined():
            return None
        return bounds

    @staticmethod
    def partition_objects(objects, nbatch):
        """Split objects into (FFT, PHOT, FAINT); PHOT objects with fewer photons than
        batches are drawn like FAINT ones (imsim/photon_pooling.py:355-386)."""
        objects_by_mode = {ProcessingMode.FFT: [], ProcessingMode.PHOT: [], ProcessingMode.FAINT: []}
        for obj in objects:
            if obj.phot_flux < nbatch and obj.mode == ProcessingMode.PHOT:
                mode = ProcessingMode.FAINT
            else:
                mode = obj.mode
            objects_by_mode[mode].append(obj)
        return (objects_by_mode[ProcessingMode.FFT], objects_by_mode[ProcessingMode.PHOT],
                objects_by_mode[ProcessingMode.FAINT])


def photon_batch_counts(phot_flux, modes_faint, nbatch: int, ud, clamp: bool = True):
    """Vectorised ``partition_objects`` + ``make_photon_batches`` for the device pipeline: integer
    photon counts per (batch, object), identical to the list-based reference algebra
    (imsim/photon_pooling.py:74,117,300-311,374-377).  ``phot_flux``: int64 array; ``modes_faint``: bool array
    (True = FAINT); ``ud``: callable returning uniforms in [0, 1), called once per faint object in order.

    As in the reference, the faint partition uses the configured ``nbatch`` (``phot_flux < nbatch``), and the
    number of batches is then clamped to ``max(min(nbatch, n_bright), 1)`` before the flux split (Q9): the
    result has that many rows (``clamp=False``: ``make_photon_batches`` alone, ``nbatch`` rows)."""
    f = np.asarray(phot_flux, dtype=np.int64)
    faint = np.asarray(modes_faint, dtype=bool) | (f < nbatch)
    bright = ~faint
    nb = max(min(int(nbatch), int(bright.sum())), 1) if clamp else int(nbatch)
    counts = np.zeros((nb, f.size), dtype=np.int64)
    i = np.arange(nb, dtype=np.int64)[:, None]
    counts[:, bright] = (f[None, bright] * (i + 1)) // nb - (f[None, bright] * i) // nb
    for k in np.nonzero(faint)[0]:
        counts[int(ud() * nb), k] = f[k]
    return counts


# ---------------------------------------------------------------------------
# device-resident pool
# ---------------------------------------------------------------------------
class DevicePhotons:
    """SoA photon pool in HBM (float64 CUDA tensors); same field names and
    predicates as ``PhotonArray`` so the sensor and the ops take either."""

    FIELDS = ("x", "y", "flux", "dxdz", "dydz", "wavelength", "pupil_u", "pupil_v", "time")

    def __init__(self, n: int, device="cuda:0", fields=FIELDS):
        import torch  # device buffers only

        self.n = int(n)
        self.device = torch.device(device)
        # one allocation, 9 rows: every field is a contiguous 8n-byte row
        self.buf = torch.empty((len(self.FIELDS), self.n), dtype=torch.float64, device=self.device)
        self._has = {f: (f in fields) for f in self.FIELDS}
        for k, f in enumerate(self.FIELDS):
            setattr(self, f, self.buf[k])

    def size(self):
        return self.n

    def __len__(self):
        return self.n

    def hasAllocatedAngles(self):
        return self._has["dxdz"] and self._has["dydz"]

    def hasAllocatedWavelengths(self):
        return self._has["wavelength"]

    def hasAllocatedPupil(self):
        return self._has["pupil_u"] and self._has["pupil_v"]

    def hasAllocatedTimes(self):
        return self._has["time"]

    def upload(self, host: "PinnedPhotons", fields=("x", "y", "flux", "wavelength"), stream=None):
        """Host (pinned) -> device copy of the given fields; returns bytes copied."""
        import torch

        nbytes = 0
        with torch.cuda.stream(stream) if stream is not None else _nullctx():
            for f in fields:
                k = self.FIELDS.index(f)
                self.buf[k].copy_(host.buf[k, : self.n], non_blocking=True)
                nbytes += self.n * 8
        return nbytes


class PinnedPhotons:
    """Pinned host staging buffer with the DevicePhotons layout (numpy views)."""

    def __init__(self, n: int):
        import torch

        self.n = int(n)
        self.buf = torch.empty((len(DevicePhotons.FIELDS), self.n), dtype=torch.float64).pin_memory()
        self.np = self.buf.numpy()
        for k, f in enumerate(DevicePhotons.FIELDS):
            setattr(self, f, self.np[k])


class _nullctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class PhotonPool:
    """Device-resident replacement of the pooled hot loop
    (imsim/photon_pooling.py:141-160) for one detector.

    ``process(photons)`` applies, on HBM-resident arrays,
      TimeSampler + PupilAnnulusSampler      (config/imsim-config.yaml:281-289)
      RubinDiffractionOptics + FocusDepth + Refraction   (:297-320, one kernel)
      SiliconSensor.accumulate(resume, recalc)           (photon_pooling.py:159)
    PhotonDCR is the optional prologue of the optics kernel: arm it on ``self.opt`` with
    ``photon_ops.set_dcr_options`` (one Jacobian and zenith direction for the whole CCD, which is what the
    reference's pooled pipeline does).
    """

    def __init__(self, ctx, sensor, exptime=30.0, t0=0.0, r_inner=2.558, r_outer=4.18, focus_depth=0.0,
                 index_ratio=3.9, seed=1):
        self.ctx = ctx
        self.sensor = sensor
        self.exptime, self.t0 = float(exptime), float(t0)
        self.r_inner, self.r_outer = float(r_inner), float(r_outer)
        self.seed = int(seed)
        self.offset = 0
        self.opt = _abi.B2OpticsOptions()
        self.opt.do_focus_depth = int(focus_depth != 0.0)
        self.opt.focus_depth = float(focus_depth)
        self.opt.do_refraction = 1
        self.opt.index_ratio = float(index_ratio)
        self.opt.seed = self.seed

    def trace(self, dp: DevicePhotons, n: Optional[int] = None, want_stats=False):
        """Samplers + optics only, on the first ``n`` photons of the pool (all by default): the photons are
        left as the op list leaves them for the sensor (x, y, dxdz, dydz, flux)."""
        n = dp.n if n is None else int(n)
        f = (lambda a: a[:n]) if n != dp.n else (lambda a: a)
        self.ctx.sample_time_pupil(f(dp.time), f(dp.pupil_u), f(dp.pupil_v), self.t0, self.exptime, self.r_inner,
                                   self.r_outer, self.seed, self.offset)
        dp._has.update(pupil_u=True, pupil_v=True, time=True)
        self.opt.photon_offset = self.offset
        stats = self.ctx.rubin_optics(f(dp.x), f(dp.y), f(dp.dxdz), f(dp.dydz), f(dp.flux), f(dp.wavelength),
                                      f(dp.pupil_u), f(dp.pupil_v), f(dp.time), options=self.opt,
                                      want_stats=want_stats)
        dp._has.update(dxdz=True, dydz=True)
        self.offset += n
        return stats

    def process(self, dp: DevicePhotons, image, resume: bool, recalc: bool, sample=True, want_stats=False,
                fused=None, write_back=False, prebound=False):
        """Run the chain on a device pool and accumulate onto the sensor's bound image.

        ``fused`` (default: whenever the sensor runs the pooled cadence ``nrecalc == 0``): one
        kernel per batch, photon -> charge deposit without intermediate arrays (``b2_pool_step``);
        otherwise the three separate kernels.  Both give identical images."""
        if fused is None:
            fused = bool(sample) and self.sensor.pod.nrecalc == 0.0
        if fused:
            return self._process_fused(dp, image, resume, recalc, want_stats, write_back, prebound)
        if sample:
            self.ctx.sample_time_pupil(dp.time, dp.pupil_u, dp.pupil_v, self.t0, self.exptime, self.r_inner,
                                       self.r_outer, self.seed, self.offset)
            dp._has.update(pupil_u=True, pupil_v=True, time=True)
        self.opt.photon_offset = self.offset
        stats = self.ctx.rubin_optics(dp.x, dp.y, dp.dxdz, dp.dydz, dp.flux, dp.wavelength, dp.pupil_u, dp.pupil_v,
                                      dp.time, options=self.opt, want_stats=want_stats)
        dp._has.update(dxdz=True, dydz=True)
        self.offset += dp.n
        added = self.sensor.accumulate(dp, image, resume=resume, recalc=recalc, sync_image=False,
                                       want_stats=want_stats, prebound=prebound)
        return added, stats

    def _process_fused(self, dp, image, resume, recalc, want_stats, write_back, prebound=False):
        import ctypes as C

        sensor = self.sensor
        if resume and image is not sensor._last_image:
            raise _lib.B2Error("image must be the same as used for the last accumulate call if resume is True")
        if not resume and not prebound:  # prebound: SiliconSensor.bind_stamp put a zero image on the device
            sensor._bind(image)
        sensor._last_image = image
        self.opt.photon_offset = self.offset
        ostats = _abi.B2OpticsStats() if want_stats else None
        astats = _abi.B2AccumStats() if want_stats else None
        _lib.check(_lib.load().b2_pool_step(
            self.ctx.handle, sensor._h, dp.n, _lib.ptr(dp.x), _lib.ptr(dp.y), _lib.ptr(dp.dxdz), _lib.ptr(dp.dydz),
            _lib.ptr(dp.flux), _lib.ptr(dp.wavelength), C.byref(self.opt), self.t0, self.exptime, self.r_inner,
            self.r_outer, self.seed, sensor._seed & 0xFFFFFFFFFFFFFFFF, self.offset, sensor._photon_offset,
            int(bool(resume)),
            int(bool(recalc)), int(bool(write_back)), C.byref(ostats) if want_stats else None,
            C.byref(astats) if want_stats else None))
        self.offset += dp.n
        sensor._photon_offset += dp.n
        sensor.last_stats = astats
        if write_back:
            dp._has.update(dxdz=True, dydz=True)
        return (astats.added_flux if want_stats else None), ostats

    def run_host_batches(self, batches, image, first_resume=False, fields=("x", "y", "flux", "wavelength"),
                         read_image_every_batch=True):
        """The pooled loop of ``buildImage`` for host-resident batches (what GalSim's shooters produce):
        ``batches`` is a sequence of ``PinnedPhotons``.  Uploads are double-buffered on a copy stream so
        the H2D transfer of batch k+1 overlaps the kernels of batch k; after every batch the image is
        snapshotted on the device and copied to ``image.array`` on a third stream (the reference
        checkpoints ``full_image`` after each batch, imsim/photon_pooling.py:167-168).
        Returns (h2d_bytes, d2h_bytes) moved per batch."""
        import torch

        dev = torch.device("cuda", self.ctx.device)
        compute = torch.cuda.current_stream(dev)
        copy_s, d2h_s = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        nmax = max(b.n for b in batches)
        bufs = [DevicePhotons(nmax, device=dev), DevicePhotons(nmax, device=dev)]
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        free = [torch.cuda.Event(), torch.cuda.Event()]
        for e in free:
            e.record(compute)
        arr = image.array
        tdtype = torch.float32 if arr.dtype == np.float32 else torch.float64
        snap = torch.empty(arr.shape, dtype=tdtype, device=dev)
        # pinned landing buffer for the per-batch image copies (where a checkpoint writer reads it);
        # image.array receives the final state
        host_img = torch.empty(arr.shape, dtype=tdtype).pin_memory()
        snapped, copied = torch.cuda.Event(), torch.cuda.Event()
        copied.record(d2h_s)
        h2d = d2h = 0
        for k, hb in enumerate(batches):
            b = k % 2
            dp = bufs[b]
            dp.n = hb.n
            for j, f in enumerate(DevicePhotons.FIELDS):
                setattr(dp, f, dp.buf[j, : hb.n])
            with torch.cuda.stream(copy_s):
                copy_s.wait_event(free[b])
                h2d = 0
                for f in fields:
                    j = DevicePhotons.FIELDS.index(f)
                    dp.buf[j, : hb.n].copy_(hb.buf[j, : hb.n], non_blocking=True)
                    h2d += hb.n * 8
                ready[b].record(copy_s)
            compute.wait_event(ready[b])
            self.process(dp, image, resume=(first_resume or k > 0), recalc=(first_resume or k > 0))
            free[b].record(compute)
            if read_image_every_batch or k == len(batches) - 1:
                compute.wait_event(copied)  # the previous snapshot has left the device buffer
                self.sensor.snapshot_image(snap)
                snapped.record(compute)
                with torch.cuda.stream(d2h_s):
                    d2h_s.wait_event(snapped)
                    host_img.copy_(snap, non_blocking=True)
                    copied.record(d2h_s)
                d2h = arr.nbytes
        compute.wait_event(copied)
        copied.synchronize()
        arr[...] = host_img.numpy()
        return h2d, d2h
