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

import ctypes as C
import dataclasses
import itertools
import os
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


# ---------------------------------------------------------------------------
# the pooled loop of buildImage on the device, fed with GalSim's host photon arrays
# ---------------------------------------------------------------------------
_POOLED_CHAIN = ("TimeSampler", "PupilAnnulusSampler", "PhotonDCR", "RubinOptics", "RubinDiffractionOptics",
                 "FocusDepth", "Refraction")


class PooledDevicePath:
    """What ``LSST_PhotonPoolingImageBuilder.buildImage`` does per sub-batch -- merge the stamps' photon arrays,
    apply the pooled photon ops, ``accumulate_photons`` onto ``full_image`` (imsim/photon_pooling.py:149-160) --
    with the image, the sensor state and the pool resident in HBM:

    * ``recognise(photon_ops, sensor, local_wcs)`` accepts the op list of config/imsim-config.yaml:281-320
      (TimeSampler, PupilAnnulusSampler, [PhotonDCR], RubinOptics | RubinDiffractionOptics, [FocusDepth], [Refraction],
      in that order) together with a device ``SiliconSensor`` running the pooled cadence (``nrecalc == 0``) and
      returns None for anything else -- the caller then keeps the reference's host loop;
    * ``begin(full_image)`` uploads the image as it stands (FFT objects already drawn, a restored checkpoint);
    * ``add(photon_arrays)`` gathers the stamps' pageable arrays into one device pool (``b2_photons_upload``: the
      merge happens in the pinned ring, no host-side copy of the pool); ``add_begin`` / ``add_wait`` are its two
      halves, the upload running on a helper thread in between;
    * ``step(resume, recalc)`` is one ``b2_pool_step``; ``read_back(full_image)`` copies the image to the host
      (checkpoints, the return of ``buildImage``).
    """

    def __init__(self, pool: PhotonPool, sensor):
        self.pool, self.sensor = pool, sensor
        self.ctx = pool.ctx
        self.dp: Optional[DevicePhotons] = None
        self.image = None
        self.photons = 0
        self.h2d_bytes = 0
        self._worker = None
        self._upload = None
        #: stamps whose photons all carry the same flux send one number instead of the array (B2_FLUX_ARRAYS=1: off)
        self.constant_flux_shortcut = not os.environ.get("B2_FLUX_ARRAYS")

    @classmethod
    def recognise(cls, photon_ops, sensor, local_wcs=None, seed=None):
        from .photon_ops import RubinDiffractionOptics, RubinOptics, set_dcr_options  # noqa: PLC0415
        from .sensor import SiliconSensor  # noqa: PLC0415

        if not isinstance(sensor, SiliconSensor) or sensor.pod.nrecalc != 0.0:
            return None
        names = [type(op).__name__ for op in photon_ops]
        if any(n not in _POOLED_CHAIN for n in names):
            return None
        order = [_POOLED_CHAIN.index(n) for n in names]
        if order != sorted(order) or len(set(names)) != len(names):
            return None
        byname = dict(zip(names, photon_ops))
        optics = byname.get("RubinDiffractionOptics") or byname.get("RubinOptics")
        if optics is None or not isinstance(optics, (RubinOptics, RubinDiffractionOptics)) \
                or "TimeSampler" not in byname or "PupilAnnulusSampler" not in byname:
            return None
        if optics.stamp_center is not None:  # the pooled pipeline builds its ops with stamp_center None (Q2)
            return None
        ts, ps = byname["TimeSampler"], byname["PupilAnnulusSampler"]
        ctx = optics._context()
        sensor.move_to(ctx)
        fd = byname.get("FocusDepth")
        rf = byname.get("Refraction")
        pool = PhotonPool(ctx, sensor, exptime=float(ts.exptime), t0=float(ts.t0), r_inner=float(ps.R_inner),
                          r_outer=float(ps.R_outer), focus_depth=float(fd.depth) if fd is not None else 0.0,
                          index_ratio=float(rf.index_ratio) if rf is not None else 1.0,
                          seed=int(seed) if seed is not None else sensor._seed + 0x5DEECE66D)
        pool.opt.do_refraction = int(rf is not None)
        dcr = byname.get("PhotonDCR")
        if dcr is not None:
            if local_wcs is None:
                return None
            jac = local_wcs.getMatrix() if hasattr(local_wcs, "getMatrix") else np.asarray(local_wcs, float)
            zen = getattr(dcr, "zenith_angle", None)
            par = getattr(dcr, "parallactic_angle", None)
            if zen is None or par is None:
                return None
            rad = (lambda a: float(a.rad) if hasattr(a, "rad") else float(a))
            unit = getattr(dcr, "scale_unit", None)
            set_dcr_options(pool.opt, float(dcr.base_wavelength), rad(zen), rad(par), jac,
                            alpha=float(getattr(dcr, "alpha", 0.0)),
                            scale_unit_rad=rad(unit) if unit is not None else np.pi / (180.0 * 3600.0),
                            pressure=float(getattr(dcr, "pressure", 69.328)),
                            temperature=float(getattr(dcr, "temperature", 293.15)),
                            H2O_pressure=float(getattr(dcr, "H2O_pressure", 1.067)))
        return cls(pool, sensor)

    def begin(self, full_image):
        """Bind ``full_image`` (a galsim.Image or ``sensor.Image``) on the device, pixels as they stand."""
        self.sensor._bind(full_image)
        self.sensor._last_image = None
        self.image = full_image
        self._first = True

    def add(self, photon_arrays):
        """The photons of a sub-batch: a list of (GalSim) PhotonArrays with x, y, flux and wavelengths."""
        n = self.add_begin(photon_arrays)
        self.add_wait()
        return n

    def add_begin(self, photon_arrays):
        """``add_prepare`` + ``add_start``."""
        n = self.add_prepare(photon_arrays)
        self.add_start()
        return n

    def add_prepare(self, photon_arrays):
        """First half of an upload, host work only: the device pool is allocated and the table of the stamps' array
        pointers is built.  May be called while the previous sub-batch is still uploading / waiting for its launch
        (its pool is a different allocation).  The arrays must stay untouched until ``add_wait`` has returned (they
        are kept referenced here)."""
        arrays = [pa for pa in photon_arrays if pa is not None and len(pa) > 0]
        n = int(sum(len(pa) for pa in arrays))
        self._prepared = None
        if n == 0:
            return 0
        fields = ("x", "y", "flux", "wavelength")
        for pa in arrays:
            if not pa.hasAllocatedWavelengths():
                raise _lib.B2Error("pooled photons need wavelengths (the stamps' WavelengthSampler assigns them)")
        dp = DevicePhotons(n, device="cuda:%d" % self.ctx.device)
        nseg = len(arrays)
        # pointer table [field][stamp] without a ctypes object per array
        ptrs = np.empty((len(fields), nseg), dtype=np.uint64)
        keep = [arrays]
        for f, name in enumerate(fields):
            row = ptrs[f]
            for g, pa in enumerate(arrays):
                a = getattr(pa, name)
                if a.dtype != np.float64 or not a.flags.c_contiguous:
                    a = np.ascontiguousarray(a, dtype=np.float64)
                    keep.append(a)
                row[g] = a.__array_interface__["data"][0]
        lens = np.fromiter((len(pa) for pa in arrays), dtype=np.int64, count=nseg)
        lib, handle = _lib.load(), self.ctx.handle
        # GalSim's shooters give every photon of an object the same flux: if that holds for all stamps (one read of
        # the flux arrays on the copy threads, bit for bit), the field is written on the device from one number per
        # stamp instead of crossing PCIe
        fill = None
        if self.constant_flux_shortcut:
            fi = fields.index("flux")
            values = np.empty(nseg, dtype=np.float64)
            flag = C.c_int32(0)
            _lib.check(lib.b2_segments_constant(nseg, ptrs[fi].ctypes.data, lens.ctypes.data, values.ctypes.data,
                                                C.addressof(flag)))
            if flag.value:
                fill = (values, dp.flux.data_ptr())
                ptrs = np.ascontiguousarray(np.delete(ptrs, fi, axis=0))
                fields = tuple(f for f in fields if f != "flux")
        dst = np.array([getattr(dp, name).data_ptr() for name in fields], dtype=np.uint64)
        keep += [ptrs, lens, dst]
        nf = len(fields)

        def upload():
            # ctypes drops the GIL for the duration of the call: the host copy threads and the DMA run while the
            # interpreter builds the next stamps
            rc = lib.b2_photons_upload(handle, nf, nseg, ptrs.ctypes.data, lens.ctypes.data, dst.ctypes.data)
            if rc == 0 and fill is not None:
                rc = lib.b2_fill_segments(handle, nseg, lens.ctypes.data, fill[0].ctypes.data, fill[1])
            return rc

        self._prepared = (upload, dp, keep)
        self.photons += n
        self.h2d_bytes += n * 8 * nf + (8 * nseg if fill is not None else 0)
        return n

    def add_start(self):
        """Second half: start gathering the prepared sub-batch into its device pool on a helper thread and return at
        once; the caller goes on with its own host work and calls ``add_wait`` before ``step``.  Call it after the
        previous sub-batch's ``step``: the upload is ordered on the stream behind what has been queued so far."""
        self.dp = None
        self._upload = None
        prepared, self._prepared = getattr(self, "_prepared", None), None
        if prepared is None:
            return
        upload, dp, keep = prepared
        if self._worker is None:
            from concurrent.futures import ThreadPoolExecutor  # noqa: PLC0415

            self._worker = ThreadPoolExecutor(max_workers=1, thread_name_prefix="b2-upload")
        self._upload = (self._worker.submit(upload), dp, keep)

    def add_wait(self):
        """Wait for the upload started by ``add_begin``; the pool is then ready for ``step``."""
        up, self._upload = getattr(self, "_upload", None), None
        if up is None:
            return
        fut, dp, keep = up
        _lib.check(fut.result())
        keep.clear()
        dp._has.update(wavelength=True)
        self.dp = dp

    def step(self, resume: bool, recalc: bool):
        """SiliconSensor.accumulate(resume, recalc) of the uploaded pool after the pooled ops, one fused launch."""
        if self.dp is None:
            if not resume:  # an empty first sub-batch still initialises the image (as GalSim does)
                from .photon_array import PhotonArray  # noqa: PLC0415

                self.sensor.accumulate(PhotonArray(0), self.image, resume=False, sync_image=False, prebound=True)
            return
        self.pool.process(self.dp, self.image, resume=resume, recalc=recalc, fused=True, prebound=True)
        self.dp = None

    def read_back(self, full_image=None):
        self.sensor.read_image(self.image if full_image is None else full_image)

    # -- the image of a finished batch, travelling to the host while the next batch uploads --------------------
    def snapshot_begin(self):
        """Copy the image as it stands now into a device buffer (in stream order, after the batch just queued) and
        start its transfer to pinned host memory on a side stream.  Returns at once."""
        import torch  # noqa: PLC0415

        from .sensor import _image_parts  # noqa: PLC0415

        arr, _, _ = _image_parts(self.image)
        dev = torch.device("cuda", self.ctx.device)
        if getattr(self, "_snap", None) is None or tuple(self._snap.shape) != arr.shape:
            tdt = torch.float32 if arr.dtype == np.float32 else torch.float64
            self._snap = torch.empty(arr.shape, dtype=tdt, device=dev)
            self._pin = torch.empty(arr.shape, dtype=tdt).pin_memory()
            self._side = torch.cuda.Stream(device=dev)
        main = self.ctx.stream if self.ctx.stream is not None else torch.cuda.default_stream(dev)
        self.sensor.snapshot_image(self._snap)
        ev = torch.cuda.Event()
        ev.record(main)
        self._side.wait_event(ev)
        with torch.cuda.stream(self._side):
            self._pin.copy_(self._snap, non_blocking=True)
            self._done = torch.cuda.Event()
            self._done.record(self._side)
        self._snap_pending = True

    def snapshot_finish(self, full_image=None):
        """Wait for the transfer started by ``snapshot_begin`` and put the pixels into ``full_image.array``."""
        from .sensor import _image_parts  # noqa: PLC0415

        if not getattr(self, "_snap_pending", False):
            return False
        arr, _, _ = _image_parts(self.image if full_image is None else full_image)
        self._done.synchronize()
        src = self._pin.numpy()
        _lib.check(_lib.load().b2_host_memcpy(arr.ctypes.data, src.ctypes.data, arr.nbytes))
        self._snap_pending = False
        return True
