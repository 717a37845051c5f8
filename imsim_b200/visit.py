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
from typing import Dict

import numpy as np

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


class DetectorRunner:
    """Reusable per-GPU state: one context, one sensor object per vendor model (re-bound per detector)."""

    def __init__(self, device: int, sensor_models: Dict[str, tuple], absorption_table, tree_rings=None,
                 band: str = "r", rot_tel_pos: float = np.radians(60.0), altitude=np.radians(67.0),
                 azimuth=np.radians(213.0), exptime: float = 30.0, seed: int = 1, psf=None):
        import torch

        self.torch = torch
        self.device = device
        self.ctx = OpticsContext(device=device, stream=torch.cuda.current_stream(torch.device("cuda", device)))
        self.tracer = gpu_tracer(self.ctx)
        self.sensor_models = sensor_models
        self.absorption = absorption_table
        self.tree_rings = tree_rings or {}
        self.band, self.rot_tel_pos, self.exptime, self.seed = band, rot_tel_pos, exptime, seed
        self.dif = diffraction_config(latitude=RUBIN_LATITUDE, altitude=altitude, azimuth=azimuth)
        self._sensors: Dict[tuple, SiliconSensor] = {}
        self._readouts: Dict[str, object] = {}
        self._pin: Dict[tuple, object] = {}
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

    def run(self, det_name: str, objects, nbatch: int = 10, wavelength_cdf=None, det_index: int = 0,
            readout: bool = False, sky_level: float = 0.0) -> dict:
        torch = self.torch
        dev = torch.device("cuda", self.device)
        t0 = time.perf_counter()
        det = lsstcam_like(det_name)
        su = make_detector_setup(self.tracer, det_name, band=self.band, rot_tel_pos=self.rot_tel_pos, detector=det)
        self.ctx.set_telescope(su.telescope)
        self.ctx.set_wcs(su.img_wcs, su.icrf_to_field)
        self.ctx.set_detector(su.detector)
        self.ctx.set_diffraction(self.dif)
        sensor = self.sensor_for(det_name)
        # the returned e-image (and raw segments) live in pinned buffers of the runner, reused by the next run()
        image = Image(self._pinned("img", (det.ny, det.nx), torch.float32).numpy(), 0, 0)
        image.array[:, :] = 0.0
        pool = PhotonPool(self.ctx, sensor, exptime=self.exptime, seed=self.seed + 1000 * det_index)
        gen = np.random.default_rng(self.seed + det_index)
        catalogue = not (isinstance(objects, tuple) and len(objects) == 4)
        cdf = cdfw = None
        if catalogue:
            # stage 1 from catalogue rows (stage1.ObjectTable): profiles, per-object SEDs, PSF kicks
            from .stage1 import Stage1

            rows, oflux = objects.build()
            rows = rows.copy()
            vx, vy, vz = self.ctx.xy_to_v(np.ascontiguousarray(rows["x"]), np.ascontiguousarray(rows["y"]))
            rows["tanx"], rows["tany"] = vx / -vz, vy / -vz  # tangents of the field angle
            cdf_np, cdfw_np = wavelength_cdf if wavelength_cdf is not None else (None, None)
            stage1 = Stage1(self.ctx, rows, cdf_np, cdfw_np, objects.radial_tables())
            if self.psf is not None:
                self.psf.upload(self.ctx, objects.arcsec_to_pix)
        else:
            ox, oy, oflux, osig = objects
            d_ox = torch.as_tensor(ox, device=dev)
            d_oy = torch.as_tensor(oy, device=dev)
            d_os = torch.as_tensor(osig, device=dev)
            if wavelength_cdf is not None:
                cdf, cdfw = (torch.as_tensor(a, device=dev) for a in wavelength_cdf)
        counts = photon_batch_counts(oflux, np.zeros(len(oflux), bool), nbatch, gen.random)
        t_setup = time.perf_counter() - t0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n_total = 0
        for k in range(nbatch):
            cnt = counts[k]
            idx = np.nonzero(cnt)[0]
            cnt = cnt[idx]
            n = int(cnt.sum())
            if n == 0:
                continue
            dp = DevicePhotons(n, device=dev, fields=("x", "y", "flux", "wavelength"))
            if catalogue:
                stage1.shoot(dp, cnt, seed=self.seed + 7 * det_index, photon_offset=n_total, select=idx)
            else:
                cum = torch.as_tensor(np.concatenate([[0], np.cumsum(cnt)]), device=dev)
                sel = torch.as_tensor(idx, device=dev)
                self.ctx.object_photons(dp.x, dp.y, dp.flux, dp.wavelength, d_ox[sel].contiguous(),
                                        d_oy[sel].contiguous(), d_os[sel].contiguous(), cum, cdf, cdfw,
                                        seed=self.seed + 7 * det_index, photon_offset=n_total)
            pool.process(dp, image, resume=(k > 0), recalc=(k > 0), fused=True)
            n_total += n
        raw = None
        e = torch.empty((det.ny, det.nx), dtype=torch.float32, device=dev)
        sensor.snapshot_image(e)
        electrons = e.sum(dtype=torch.float64)  # collected source charge, before sky / dark current
        if sky_level > 0.0:
            # sky background through the sensor model (imsim/lsst_image.py:128-199): level x pixel areas (tree
            # rings + the accumulated charge), exact Poisson noise, all on the device
            from .sky import add_sky, pixel_areas_device

            areas = pixel_areas_device(sensor, use_flux=True)
            add_sky(self.ctx, e, sky_level, seed=self.seed + 17 * det_index, areas=areas)
        if readout:
            # post-path on the device (imsim/readout.py:414-480): the e-image goes from the sensor's buffer to
            # int32 amplifier segments without visiting the host; both are then copied back
            from .readout import CcdReadout, lsstcam_like_amps

            vendor = vendor_of(det_name)
            ro = self._readouts.get(vendor)
            if ro is None:
                ro = self._readouts[vendor] = CcdReadout(self.ctx, lsstcam_like_amps(vendor), exptime=self.exptime,
                                                         midline_stop=(vendor == "e2v"))
            amp = lsstcam_like_amps(vendor)[0]
            ex = e[: 2 * amp.ny, : 8 * amp.nx].contiguous() if (2 * amp.ny, 8 * amp.nx) != tuple(e.shape) else e
            draw = ro.build_amp_images(ex, seed=self.seed + 13 * det_index)
            praw = self._pinned("raw", tuple(draw.shape), torch.int32)
            praw.copy_(draw, non_blocking=True)
            raw = praw.numpy()
        # with the readout, e is the e-image after bleed trails and dark current, like CcdReadout.eimage
        self._pinned("img", (det.ny, det.nx), torch.float32).copy_(e, non_blocking=True)
        e1.record()
        torch.cuda.synchronize(dev)
        self.last_raw = raw
        # per-object sum of the photon fluxes shot over all batches: the ``incident_flux`` column of the
        # photon_pooling_truth output (imsim/photon_pooling.py:472-511, stamp.py:743)
        self.last_incident_flux = np.asarray(counts, dtype=np.int64).sum(axis=0).astype(np.float64)
        rec = {"det_name": det_name, "device": self.device, "photons": n_total, "nbatch": nbatch,
               "electrons": float(electrons.item()), "gpu_ms": float(e0.elapsed_time(e1)),
               "setup_ms": 1e3 * t_setup}
        return rec, image
