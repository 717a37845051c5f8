"""Synthetic visit driver: the per-detector pooled pipeline of one LSSTCam exposure, sharded by
detector over the GPUs of a box (imsim/ccd.py:72-89 iterates the same 189 CCDs
as output files, config/imsim-config.yaml:326 `output.nproc` is the reference's scaling knob).

For every detector of this rank:
  1. set-up: telescope with the detector's height offset, WCS pair fitted to chief rays, detector affine,
     sensor model by vendor + that detector's tree rings                (imsim/lsst_image.py:93-103)
  2. object table -> per-batch integer photon counts                     (imsim/photon_pooling.py:279-313)
  3. for each of nbatch batches: photons generated in HBM (k_object_photons), then ONE fused kernel
     sampler -> optics -> sensor (b2_pool_step) with recalc=True         (photon_pooling.py:141-160)
  4. image back to the host (where the reference writes the e-image / checkpoint)
No collective on the photon path; ``sharding.gather_visit_metadata`` collects per-CCD records.
"""
from __future__ import annotations

import time
from dataclasses import dataclass
from typing import Dict, Optional

import numpy as np

from . import _lib
from .context import OpticsContext
from .detector import lsstcam_like
from .diffraction import RUBIN_LATITUDE, diffraction_config
from .photon_pooling import DevicePhotons, PhotonPool, photon_batch_counts
from .sensor import Image, SiliconSensor
from .synthetic import gpu_tracer, make_detector_setup

from .detector import ITL_RAFTS  # noqa: E402  (ITL rafts of LSSTCam; the rest of the science rafts carry e2v CCDs)


def vendor_of(det_name: str) -> str:
    return "itl" if det_name[:3] in ITL_RAFTS else "e2v"


def synthetic_objects(n_obj: int, nx: int, ny: int, seed: int, total_photons: float):
    """Object table of a dense field: positions uniform on the CCD, fluxes dN/dm ~ 10^(0.3 m) over 9
    magnitudes, scaled to ``total_photons``; PSF sigma 0.7'' FWHM at 0.2''/px."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(0, nx, n_obj)
    y = rng.uniform(0, ny, n_obj)
    u = rng.uniform(0, 1, n_obj)
    m = np.log10(1 + u * (10 ** (0.3 * 9) - 1)) / 0.3  # 0 (bright) .. 9 (faint), more faint ones
    w = 10 ** (-0.4 * m)
    flux = np.maximum(np.round(w / w.sum() * total_photons), 1).astype(np.int64)
    sigma = np.full(n_obj, 0.7 / 2.355 / 0.2)
    return x, y, flux, sigma


def synthetic_catalog(n_obj: int, nx: int, ny: int, seed: int, total_photons: float, galaxy_fraction: float = 0.6,
                      arcsec_to_pix=None, n_sed: int = 8):
    """A dense field as catalogue rows for stage 1 (``ObjectTable``): stars (DeltaFunction) and galaxies as the
    reference's catalogues describe them -- a de Vaucouleurs bulge, an exponential disc and star-forming knots
    (skyCatalogs ``Add([bulge, disk, knots])``, imsim/skycat.py:163-190; sersic2d / knots rows of instance
    catalogues, imsim/instcat.py:496-545) -- with the flux function of ``synthetic_objects``."""
    from .stage1 import ObjectTable

    rng = np.random.default_rng(seed)
    x, y, flux, _ = synthetic_objects(n_obj, nx, ny, seed, total_photons)
    tab = ObjectTable(arcsec_to_pix=np.eye(2) / 0.2 if arcsec_to_pix is None else np.asarray(arcsec_to_pix, float))
    is_gal = rng.random(n_obj) < galaxy_fraction
    sed = rng.integers(0, n_sed, n_obj)
    st = np.nonzero(~is_gal)[0]
    tab.add_points(x[st], y[st], flux[st], sed=sed[st])
    g = np.nonzero(is_gal)[0]
    ng = g.size
    hlr = np.exp(rng.normal(np.log(0.5), 0.5, ng))
    q, beta = rng.uniform(0.2, 1.0, ng), rng.uniform(0, np.pi, ng)
    g1, g2 = rng.normal(0, 0.01, ng), rng.normal(0, 0.01, ng)
    bt = rng.uniform(0.1, 0.6, ng)  # bulge fraction
    fb = np.maximum(np.round(flux[g] * bt), 1).astype(np.int64)
    fk = np.maximum(np.round(flux[g] * 0.1), 1).astype(np.int64)
    fd = np.maximum(flux[g] - fb - fk, 1)
    tab.add_sersic(x[g], y[g], fb, 0.6 * hlr, 4.0, q=np.minimum(1.0, q + 0.2), beta=beta, g1=g1, g2=g2, sed=sed[g])
    tab.add_sersic(x[g], y[g], fd, hlr, 1.0, q=q, beta=beta, g1=g1, g2=g2, sed=sed[g])
    tab.add_knots(x[g], y[g], fk, hlr, rng.integers(5, 40, ng), q=q, beta=beta, g1=g1, g2=g2,
                  sed=rng.integers(0, n_sed, ng), seed=rng.integers(1 << 62, size=ng))
    return tab


@dataclass
class PreparedDetector:
    """Everything of one detector that the host computes before the first kernel can be queued."""
    det_name: str
    det_index: int
    det: object
    setup: object
    catalogue: bool
    objects: object
    rows: Optional[np.ndarray]
    oflux: np.ndarray
    radial: Optional[np.ndarray]
    wavelength_cdf: Optional[tuple]
    counts: list
    batches: list  # per batch: (n_photons, object indices, their counts)
    nbatch: int
    readout: bool
    sky_level: float
    setup_s: float


@dataclass
class LaunchedDetector:
    """A detector whose work is queued on the stream; ``DetectorRunner.finish`` waits for it."""
    prep: PreparedDetector
    image: Image
    raw: Optional[np.ndarray]
    electrons: object
    events: tuple
    n_total: int
    keep: list


class DetectorRunner:
    """Reusable per-GPU state: one context, one sensor object per vendor model (re-bound per detector).

    A detector goes through three phases so that a visit can be software-pipelined (``run_many``):
    ``prepare`` -- host work only (telescope + WCS fit to chief rays, catalogue rows, the integer photon
    split of imsim/photon_pooling.py:279-313, per-batch index tables); its few chief-ray traces run on a
    second context with its own high-priority stream, so they do not queue behind another detector's kernels;
    ``launch`` -- every upload, kernel and read-back of the detector queued on the compute stream without
    waiting for the GPU, possibly behind a detector that is still running (index tables go through pinned
    staging, the image is bound as zeros on the device, per-detector tables are copied in stream order);
    ``finish`` -- wait, collect the record.  ``run`` is the three in sequence."""

    def __init__(self, device: int, sensor_models: Dict[str, tuple], absorption_table, tree_rings=None,
                 band: str = "r", rot_tel_pos: float = np.radians(60.0), altitude=np.radians(67.0),
                 azimuth=np.radians(213.0), exptime: float = 30.0, seed: int = 1, psf=None):
        import torch

        self.torch = torch
        self.device = device
        dev = torch.device("cuda", device)
        self.ctx = OpticsContext(device=device, stream=torch.cuda.current_stream(dev))
        # set-up queries (chief rays, field angles of the catalogue rows) on their own stream
        self.setup_stream = torch.cuda.Stream(dev, priority=-1)
        self.setup_ctx = OpticsContext(device=device, stream=self.setup_stream)
        self.tracer = gpu_tracer(self.setup_ctx)
        self.sensor_models = sensor_models
        self.absorption = absorption_table
        self.tree_rings = tree_rings or {}
        self.band, self.rot_tel_pos, self.exptime, self.seed = band, rot_tel_pos, exptime, seed
        self.dif = diffraction_config(latitude=RUBIN_LATITUDE, altitude=altitude, azimuth=azimuth)
        self._sensors: Dict[tuple, SiliconSensor] = {}
        self._readouts: Dict[str, object] = {}
        self._pin: Dict[tuple, object] = {}
        self._slot = 0
        self.last_incident_flux = None
        self.last_raw = None
        #: stage-1 PSF (``atmosphere.AtmosphericPSF`` / ``GaussianPSF``): one realisation per visit, shared by the
        #: detectors of this GPU like the reference's ``atm_psf`` input object (imsim/atmPSF.py:339-347)
        self.psf = psf

    def _pinned(self, key, shape, dtype):
        t = self._pin.get((key, shape))
        if t is None:
            t = self._pin[(key, shape)] = self.torch.empty(shape, dtype=dtype).pin_memory()
        return t

    def sensor_for(self, det_name: str) -> SiliconSensor:
        """One sensor object per vendor model, re-used across detectors: only the tree-ring table
        changes (``set_treerings``); the ~2.5 GB of boundary arrays stay allocated."""
        vendor = vendor_of(det_name)
        tr = self.tree_rings.get(det_name)
        func, center = (tr[1], tr[0]) if tr else (None, (0.0, 0.0))
        s = self._sensors.get(vendor)
        if s is None:
            cfg, dat = self.sensor_models[vendor]
            s = SiliconSensor(config=cfg, vertex_data=dat, nrecalc=0, strength=1.0, rng=self.seed, treering_func=func,
                              treering_center=center, absorption_table=self.absorption, context=self.ctx)
            self._sensors[vendor] = s
        else:
            s.set_treerings(func, center)
        return s

    # ------------------------------------------------------------------ phase 1: host
    def prepare(self, det_name: str, objects, nbatch: int = 10, wavelength_cdf=None, det_index: int = 0,
                readout: bool = False, sky_level: float = 0.0) -> PreparedDetector:
        """``objects``: ``(x, y, flux, sigma)`` arrays, a ``stage1.ObjectTable``, or a callable returning one of
        them (evaluated here, so that building the catalogue also overlaps the previous detector's kernels)."""
        t0 = time.perf_counter()
        if callable(objects):
            objects = objects()
        det = lsstcam_like(det_name)
        su = make_detector_setup(self.tracer, det_name, band=self.band, rot_tel_pos=self.rot_tel_pos, detector=det)
        catalogue = not (isinstance(objects, tuple) and len(objects) == 4)
        rows = radial = None
        if catalogue:
            rows, oflux = objects.build()
            rows = rows.copy()
            sc = self.setup_ctx  # the tracer left this detector's telescope there
            sc.set_wcs(su.img_wcs, su.icrf_to_field)
            sc.set_detector(su.detector)
            vx, vy, vz = sc.xy_to_v(np.ascontiguousarray(rows["x"]), np.ascontiguousarray(rows["y"]))
            rows["tanx"], rows["tany"] = vx / -vz, vy / -vz  # tangents of the field angle
            radial = objects.radial_tables()
        else:
            oflux = objects[2]
        gen = np.random.default_rng(self.seed + det_index)
        counts = photon_batch_counts(oflux, np.zeros(len(oflux), bool), nbatch, gen.random)
        batches = []
        nbatch = counts.shape[0]  # clamped to the number of bright objects, like the reference
        for k in range(nbatch):
            cnt = np.asarray(counts[k])
            idx = np.nonzero(cnt)[0]
            cnt = cnt[idx].astype(np.int64)
            batches.append((int(cnt.sum()), idx.astype(np.int64), cnt))
        return PreparedDetector(det_name, det_index, det, su, catalogue, objects, rows, oflux, radial, wavelength_cdf,
                                counts, batches, nbatch, readout, sky_level, time.perf_counter() - t0)

    # ------------------------------------------------------------------ phase 2: queue the device work
    def launch(self, p: PreparedDetector) -> LaunchedDetector:
        torch = self.torch
        dev = torch.device("cuda", self.device)
        det, su, det_index = p.det, p.setup, p.det_index
        self.ctx.set_telescope(su.telescope)
        self.ctx.set_wcs(su.img_wcs, su.icrf_to_field)
        self.ctx.set_detector(su.detector)
        self.ctx.set_diffraction(self.dif)
        sensor = self.sensor_for(p.det_name)
        # an independent sensor stream per detector image (the reference: sensor.updateRNG(rng) per image,
        # imsim/photon_pooling.py:71, with the per-file seed of imsim/ccd.py:30)
        sensor.updateRNG((self.seed * 0x9E3779B97F4A7C15 + 1000003 * (det_index + 1)) & 0x7FFFFFFFFFFFFFFF)
        # the returned e-image (and raw segments) land in pinned buffers of the runner; three slots rotate, so the
        # result handed to the caller stays valid while the next two detectors are in flight
        slot = self._slot
        self._slot = (self._slot + 1) % 3
        host_img = self._pinned(("img", slot), (det.ny, det.nx), torch.float32)
        image = Image(host_img.numpy(), 0, 0)
        sensor.bind_stamp(0, 0, det.nx, det.ny, dtype=np.float32)  # zeros, on the device: no 66 MB upload
        pool = PhotonPool(self.ctx, sensor, exptime=self.exptime, seed=self.seed + 1000 * det_index)
        keep = []
        cdf = cdfw = None
        if p.catalogue:
            # stage 1 from catalogue rows (stage1.ObjectTable): profiles, per-object SEDs, PSF kicks
            from .stage1 import Stage1

            cdf_np, cdfw_np = p.wavelength_cdf if p.wavelength_cdf is not None else (None, None)
            stage1 = Stage1(self.ctx, p.rows, cdf_np, cdfw_np, p.radial)
            if self.psf is not None:
                self.psf.upload(self.ctx, p.objects.arcsec_to_pix)
            keep.append(stage1)
        else:
            ox, oy, _, osig = p.objects
            d_ox, d_oy, d_os = (_lib.h2d_async(np.asarray(a, np.float64), dev) for a in (ox, oy, osig))
            if p.wavelength_cdf is not None:
                cdf, cdfw = (_lib.h2d_async(np.asarray(a, np.float64), dev) for a in p.wavelength_cdf)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n_total = 0
        first = True
        for n, idx, cnt in p.batches:
            if n == 0:
                continue
            dp = DevicePhotons(n, device=dev, fields=("x", "y", "flux", "wavelength"))
            if p.catalogue:
                stage1.shoot(dp, cnt, seed=self.seed + 7 * det_index, photon_offset=n_total, select=idx)
            else:
                cum_h = np.zeros(cnt.size + 1, dtype=np.int64)
                np.cumsum(cnt, out=cum_h[1:])
                cum = _lib.h2d_async(cum_h, dev)
                sel = _lib.h2d_async(idx, dev)
                self.ctx.object_photons(dp.x, dp.y, dp.flux, dp.wavelength, d_ox[sel].contiguous(),
                                        d_oy[sel].contiguous(), d_os[sel].contiguous(), cum, cdf, cdfw,
                                        seed=self.seed + 7 * det_index, photon_offset=n_total)
            pool.process(dp, image, resume=not first, recalc=not first, fused=True, prebound=True)
            first = False
            n_total += n
        raw = None
        e = torch.empty((det.ny, det.nx), dtype=torch.float32, device=dev)
        if first:
            e.zero_()  # no photons at all: nothing was accumulated on the bound image
        else:
            sensor.snapshot_image(e)
        # collected source charge, before sky / dark current; read back through a pinned scalar (an .item() would
        # wait for everything queued on the stream, including the next detector)
        electrons = self._pinned(("electrons", slot), (1,), torch.float64)
        electrons.copy_(e.sum(dtype=torch.float64).reshape(1), non_blocking=True)
        if p.sky_level > 0.0:
            # sky background through the sensor model (imsim/lsst_image.py:128-199): level x pixel areas (tree
            # rings + the accumulated charge), exact Poisson noise, all on the device
            from .sky import add_sky, pixel_areas_device

            areas = pixel_areas_device(sensor, use_flux=True)
            add_sky(self.ctx, e, p.sky_level, seed=self.seed + 17 * det_index, areas=areas)
        if p.readout:
            # post-path on the device (imsim/readout.py:414-480): the e-image goes from the sensor's buffer to
            # int32 amplifier segments without visiting the host; both are then copied back
            from .readout import CcdReadout, lsstcam_like_amps

            vendor = vendor_of(p.det_name)
            ro = self._readouts.get(vendor)
            if ro is None:
                ro = self._readouts[vendor] = CcdReadout(self.ctx, lsstcam_like_amps(vendor), exptime=self.exptime,
                                                         midline_stop=(vendor == "e2v"))
            amp = lsstcam_like_amps(vendor)[0]
            ex = e[: 2 * amp.ny, : 8 * amp.nx].contiguous() if (2 * amp.ny, 8 * amp.nx) != tuple(e.shape) else e
            draw = ro.build_amp_images(ex, seed=self.seed + 13 * det_index)
            praw = self._pinned(("raw", slot), tuple(draw.shape), torch.int32)
            praw.copy_(draw, non_blocking=True)
            raw = praw.numpy()
        # with the readout, e is the e-image after bleed trails and dark current, like CcdReadout.eimage
        host_img.copy_(e, non_blocking=True)
        e1.record()
        return LaunchedDetector(p, image, raw, electrons, (e0, e1), n_total, keep)

    # ------------------------------------------------------------------ phase 3: wait and collect
    def finish(self, h: LaunchedDetector):
        e0, e1 = h.events
        e1.synchronize()
        p = h.prep
        self.last_raw = h.raw
        # per-object sum of the photon fluxes shot over all batches: the ``incident_flux`` column of the
        # photon_pooling_truth output (imsim/photon_pooling.py:472-511, stamp.py:743)
        self.last_incident_flux = np.asarray(p.counts, dtype=np.int64).sum(axis=0).astype(np.float64)
        rec = {"det_name": p.det_name, "device": self.device, "photons": h.n_total, "nbatch": p.nbatch,
               "electrons": float(h.electrons[0]), "gpu_ms": float(e0.elapsed_time(e1)),
               "setup_ms": 1e3 * p.setup_s}
        h.keep.clear()
        return rec, h.image

    def run(self, det_name: str, objects, nbatch: int = 10, wavelength_cdf=None, det_index: int = 0,
            readout: bool = False, sky_level: float = 0.0):
        """One detector start to finish; returns ``(record, image)``."""
        return self.finish(self.launch(self.prepare(det_name, objects, nbatch=nbatch, wavelength_cdf=wavelength_cdf,
                                                    det_index=det_index, readout=readout, sky_level=sky_level)))

    def run_many(self, jobs):
        """Software-pipelined visit loop over ``jobs`` (an iterable of keyword dicts for ``prepare``).  Two
        detectors are kept queued on the stream: while the GPU works through detector k (with k+1 already behind
        it), the host prepares detector k+2, so neither the host work nor the hand-over between detectors leaves
        the GPU idle.  Yields ``(record, image)`` in order; an image (and ``last_raw``) stays valid until the
        caller asks for the next result."""
        it = iter(jobs)
        queue = []
        for job in it:
            prep = self.prepare(**job)          # host work, overlapping the kernels already queued
            if len(queue) == 2:
                result = self.finish(queue.pop(0))
                queue.append(self.launch(prep))  # the stream holds two detectors again before the caller works
                yield result
            else:
                queue.append(self.launch(prep))
        for h in queue:
            yield self.finish(h)


class VisitLanes:
    """Several ``DetectorRunner`` s of one GPU, each on its own stream and driven by its own host thread, so that one
    detector's FP64-bound ray trace shares the SMs with another's boundary update and its memory-bound kernels
    (measured: two lanes finish a catalogue-field visit 6 % sooner than one, three do no better).  Detectors are
    independent (seeds follow ``det_index``), so the images do not depend on the number of lanes."""

    def __init__(self, device: int, lanes: int, *runner_args, **runner_kw):
        import torch

        self.torch, self.device = torch, device
        self.streams = [torch.cuda.Stream(device=device) for _ in range(max(1, lanes))]
        self.runners = []
        for s in self.streams:
            with torch.cuda.stream(s):
                self.runners.append(DetectorRunner(device, *runner_args, **runner_kw))
        psf = runner_kw.get("psf")
        if psf is not None and hasattr(psf, "screens"):  # moved to the device once, before the lanes start
            psf.upload(self.runners[0].ctx, None)
        torch.cuda.synchronize(device)

    def run(self, jobs, cost=None, on_result=None):
        """``jobs``: a list of keyword dicts for ``DetectorRunner.prepare``; ``cost(job)`` balances the lanes (LPT;
        default: equal costs).  ``on_result(record, image, raw)`` is called on the lane's thread while the image
        is valid.  Returns the records in the order of ``jobs``."""
        import threading

        torch = self.torch
        jobs = list(jobs)
        lanes = len(self.runners)
        load = [0.0] * lanes
        lane_of = [0] * len(jobs)
        for i in sorted(range(len(jobs)), key=lambda i_: -(cost(jobs[i_]) if cost else 1.0)):
            k = int(np.argmin(load))
            lane_of[i] = k
            load[k] += cost(jobs[i]) if cost else 1.0
        recs, errs = [None] * len(jobs), []

        def work(k):
            try:
                torch.cuda.set_device(self.device)  # the current device is per thread
                mine = [i for i in range(len(jobs)) if lane_of[i] == k]
                runner = self.runners[k]
                with torch.cuda.stream(self.streams[k]):
                    for i, (rec, image) in zip(mine, runner.run_many(jobs[i] for i in mine)):
                        recs[i] = rec
                        if on_result is not None:
                            on_result(rec, image, runner.last_raw)
            except BaseException as e:  # noqa: BLE001 -- re-raised on the calling thread
                errs.append(e)

        threads = [threading.Thread(target=work, args=(k,), name="b2-visit-lane-%d" % k) for k in range(lanes)]
        for t in threads:
            t.start()
        for t in threads:
            t.join()
        if errs:
            raise errs[0]
        return recs
