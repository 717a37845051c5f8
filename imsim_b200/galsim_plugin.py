"""GalSim config registration of the B200 hot path.

List this module after ``imsim`` in a config's ``modules:`` and the same YAML
runs with the photon-shooting path swapped:

    modules: [imsim, imsim_b200.galsim_plugin]

It re-registers, under the reference's own names,

* photon ops ``RubinOptics``, ``RubinDiffractionOptics``, ``RubinDiffraction``
  (imsim/photon_ops.py:400-451),
* the sensor type ``Silicon`` (galsim.config sensor builder used by
  config/imsim-config.yaml:230-235),
* input ``tree_rings`` and value types ``TreeRingCenter`` / ``TreeRingFunc``
  (imsim/treerings.py:220-243).

* image type ``LSST_PhotonPoolingImage`` (imsim/photon_pooling.py:469): imSim's builder with ``buildImage``
  re-stated so that the pooled hot loop (photon_pooling.py:141-160) keeps ``full_image``, the sensor state and
  the photon pool in HBM -- the stamps' photon arrays go straight from GalSim's shooters into the device pool,
  one fused launch per sub-batch, the image comes back at checkpoints and at the end.  Object loading,
  partitioning, FFT objects, checkpoint files and the faint-object rule stay imSim's own code (inherited),
* image type ``LSST_Flat`` (imsim/flat.py:300): imSim's builder with the section loop of ``addNoise`` on the
  device (``B200FlatBuilder``: pixel-area branch and ``sed`` branch; other sensors / noise types / checkpointed
  runs keep imSim's loop),
* stamp type ``LSST_Photons`` (imsim/stamp.py:747): imSim's builder, re-registered so that a config naming it gets
  the device sensor behind it.

``B200SiliconSensor`` subclasses ``galsim.SiliconSensor`` so the ``isinstance`` test at
photon_pooling.py:209 keeps passing ``recalc`` on the host route as well.

GalSim, batoid and the LSST stack are not available in the build container: the module is exercised against the
stand-in config engine of tests/stubs (tests/test_gpu_plugin_pooling.py, tests/test_gpu_plugin_flat.py), not against
a real GalSim.
"""
from __future__ import annotations

import os

import galsim
import numpy as np
from galsim.config import (GetAllParams, GetInputObj, InputLoader, PhotonOpBuilder, RegisterInputType,
                           RegisterPhotonOpType, RegisterValueType)
from galsim.config.util import get_cls_params

from . import photon_ops as _ops
from .diffraction import RUBIN_LATITUDE
from .photon_pooling import PooledDevicePath
from .sensor import SiliconSensor as _B200Sensor
from .treerings import TreeRings

_TYPES = {"CelestialCoord": galsim.CelestialCoord, "Angle": galsim.Angle, "PositionD": galsim.PositionD}


def _resolve(params):
    """type names -> the types of the GalSim that is imported now (also on a re-import of this module)"""
    return {k: _TYPES.get(v if isinstance(v, str) else getattr(v, "__name__", None), v) for k, v in params.items()}


for _cls in (_ops.RubinOptics, _ops.RubinDiffractionOptics, _ops.RubinDiffraction):
    _cls._req_params = _resolve(_cls._req_params)
    _cls._opt_params = _resolve(_cls._opt_params)


def photon_op_type(identifier: str, input_type=None):
    """Same decorator as imsim/photon_ops.py:361-386."""

    def decorator(deserializer):
        class Factory(PhotonOpBuilder):
            def buildPhotonOp(self, config, base, logger):
                return deserializer(config, base, logger)

        RegisterPhotonOpType(identifier, Factory(), input_type=input_type)
        return deserializer

    return decorator


def config_kwargs(config, base, cls, base_args=()):
    req, opt, single, _takes_rng = get_cls_params(cls)
    kwargs, _safe = GetAllParams(config, base, req, opt, single)
    kwargs.update({key: base[key] for key in base_args})
    return kwargs


def _camera(name):
    from imsim.camera import get_camera  # the reference's cached camera lookup

    return get_camera(name)


_N_GPUS = None


def _device(base):
    """One GPU per worker process.  A launcher that pins one process per GPU says which (``LOCAL_RANK`` of torchrun /
    srun wrappers, or ``b2_device`` in the config's base dict); otherwise det_num % n_gpus: GalSim forks its
    workers per output file, i.e. per detector."""
    global _N_GPUS
    if _N_GPUS is None:
        import torch

        _N_GPUS = max(torch.cuda.device_count(), 1)  # asked once: the NVML query costs ~20 ms a call
    if "b2_device" in base:
        return int(base["b2_device"]) % _N_GPUS
    local_rank = os.environ.get("LOCAL_RANK")
    if local_rank is not None and local_rank.lstrip("-").isdigit():
        return int(local_rank) % _N_GPUS
    return int(base.get("det_num", base.get("file_num", 0))) % _N_GPUS


_rubin_optics_base_args = ("stamp_center",)


@photon_op_type("RubinOptics", input_type="telescope")
def deserialize_rubin_optics(config, base, _logger):
    kwargs = config_kwargs(config, base, _ops.RubinOptics, base_args=_rubin_optics_base_args)
    return _ops.RubinOptics(telescope=base["det_telescope"], icrf_to_field=base["_icrf_to_field"],
                            img_wcs=base["current_image"].wcs, camera=_camera(kwargs.pop("camera")),
                            device=_device(base), **kwargs)


@photon_op_type("RubinDiffractionOptics", input_type="telescope")
def deserialize_rubin_diffraction_optics(config, base, _logger):
    kwargs = config_kwargs(config, base, _ops.RubinDiffractionOptics, _rubin_optics_base_args)
    telescope = base["det_telescope"]
    rubin_diffraction = _ops.RubinDiffraction(
        telescope=telescope, latitude=kwargs.pop("latitude", RUBIN_LATITUDE), altitude=kwargs.pop("altitude"),
        azimuth=kwargs.pop("azimuth"), img_wcs=base["current_image"].wcs, icrf_to_field=base["_icrf_to_field"],
        disable_field_rotation=kwargs.pop("disable_field_rotation", False), device=_device(base))
    return _ops.RubinDiffractionOptics(telescope=telescope, camera=_camera(kwargs.pop("camera")),
                                       rubin_diffraction=rubin_diffraction, device=_device(base), **kwargs)


@photon_op_type("RubinDiffraction", input_type="telescope")
def deserialize_rubin_diffraction(config, base, _logger):
    kwargs = config_kwargs(config, base, _ops.RubinDiffraction)
    return _ops.RubinDiffraction(telescope=base["det_telescope"], icrf_to_field=base["_icrf_to_field"],
                                 img_wcs=base["current_image"].wcs, device=_device(base), **kwargs)


class B200SiliconSensor(_B200Sensor, galsim.SiliconSensor):
    """The device sensor, recognisable as a ``galsim.SiliconSensor`` (photon_pooling.py:209)."""

    def __init__(self, *args, **kwargs):
        _B200Sensor.__init__(self, *args, **kwargs)  # galsim's __init__ (C++ Silicon) is not run


class _SiliconSensorBuilder(galsim.config.sensor.SensorBuilder):
    def buildSensor(self, config, base, logger):
        opt = {"name": str, "strength": float, "diffusion_factor": float, "qdist": int, "nrecalc": float,
               "treering_func": None, "treering_center": galsim.PositionD, "transpose": bool}
        kwargs, _safe = GetAllParams(config, base, opt=opt)
        kwargs["rng"] = galsim.config.GetRNG(config, base, logger, "SiliconSensor")
        kwargs["device"] = _device(base)
        return B200SiliconSensor(**kwargs)


galsim.config.sensor.RegisterSensorType("Silicon", _SiliconSensorBuilder())


def TreeRingCenter(config, base, value_type):
    tree_rings = GetInputObj("tree_rings", config, base, "TreeRingCenter")
    kwargs, safe = GetAllParams(config, base, req={"det_name": str})
    return tree_rings.get_center(kwargs["det_name"]), safe


def TreeRingFunc(config, base, value_type):
    tree_rings = GetInputObj("tree_rings", config, base, "TreeRingCenter")
    kwargs, safe = GetAllParams(config, base, req={"det_name": str})
    return tree_rings.get_func(kwargs["det_name"]), safe


def _tree_rings_with_data_dir(file_name, only_dets=None, logger=None, defer_load=True):
    from imsim.meta_data import data_dir

    return TreeRings(file_name, only_dets=only_dets, logger=logger, defer_load=defer_load, data_dir=data_dir)


_tree_rings_with_data_dir._req_params = TreeRings._req_params
_tree_rings_with_data_dir._opt_params = TreeRings._opt_params
RegisterInputType("tree_rings", InputLoader(_tree_rings_with_data_dir, takes_logger=True))
RegisterValueType("TreeRingCenter", TreeRingCenter, [galsim.PositionD], input_type="tree_rings")
RegisterValueType("TreeRingFunc", TreeRingFunc, [object], input_type="tree_rings")


# ---------------------------------------------------------------------------
# LSST_PhotonPoolingImage with the pooled loop on the device
# ---------------------------------------------------------------------------
from imsim.photon_pooling import LSST_PhotonPoolingImageBuilder as _ImsimPoolingBuilder  # noqa: E402


class B200PhotonPoolingImageBuilder(_ImsimPoolingBuilder):
    """``LSST_PhotonPoolingImage``.  ``setup``, ``addNoise``, object loading / partitioning / batching, stamp
    building and checkpoint I/O are imSim's (inherited); ``buildImage`` follows imsim/photon_pooling.py:29-174 step
    by step but hands the per-sub-batch work -- merge, photon ops, ``accumulate_photons`` -- to
    ``PooledDevicePath`` whenever the configured op list and sensor are the ones it knows, and to the inherited
    host code otherwise."""

    #: set by buildImage: "device" or "host" (which route the last image took), photons pooled, bytes uploaded
    last_route = None

    def _draw_fft_batches(self, base, logger, fft_objects, full_image, book, chk_name, image_num):
        """FFT-rendered objects, batch by batch on the host (photon_pooling.py:76-114)."""
        nbatch = max(min(self.nbatch_fft, len(fft_objects)), 1)
        for batch_num, batch in enumerate(self.make_batches(fft_objects, nbatch), start=1):
            if nbatch > 1:
                logger.warning("Start FFT batch %d/%d with %d objects", batch_num, nbatch, len(batch))
            stamps, current_vars = self.build_stamps(base, logger, batch)
            base['index_key'] = 'image_num'
            for obj, stamp in zip(batch, stamps):
                overlap = self.stamp_bounds(stamp, full_image.bounds)
                if overlap is None:
                    continue
                full_image[overlap] += stamp[overlap]
                book["obj_nums"].append(obj.index)
            noisy = [k for k, v in enumerate(current_vars) if v != 0]
            book["stamps"].extend(stamps[k] for k in noisy)
            book["vars"].extend(current_vars[k] for k in noisy)
            if self.checkpoint is not None:
                self.save_checkpoint(self.checkpoint, chk_name, base, full_image, book["stamps"], book["vars"],
                                     book["obj_nums"], book["photon_batch"])
                logger.warning('File %d: Completed batch %d, and wrote checkpoint data to %s',
                               base.get('file_num', 0), batch_num, self.checkpoint.file_name)

    def buildImage(self, config, base, image_num, _obj_num, logger):
        self._set_config_image_pos(config, base)
        book = {"stamps": [], "vars": [], "obj_nums": [], "photon_batch": 0}
        full_image = None
        chk_name = "buildImage_photonpooling_" + self.det_name
        if self.checkpoint is not None:
            (full_image, book["vars"], book["stamps"], book["obj_nums"],
             book["photon_batch"]) = self.load_checkpoint(self.checkpoint, chk_name, base, logger)
        todo = sorted(frozenset(range(self.nobjects)) - frozenset(book["obj_nums"]))
        if full_image is None:
            full_image = self._create_full_image(config, base)

        sensor = base.get('sensor', None)
        rng = galsim.config.GetRNG(config, base, logger, "LSST_Silicon")
        if sensor is not None:
            sensor.updateRNG(rng)

        fft_objects, phot_objects, faint_objects = self.partition_objects(
            self.load_objects(todo, config, base, logger), self.nbatch)
        logger.info("Found %d FFT objects, %d photon shooting objects and %d faint objects", len(fft_objects),
                    len(phot_objects), len(faint_objects))
        if self.checkpoint is not None:
            if not fft_objects:
                logger.warning('All FFT objects already rendered for this image.')
            else:
                logger.warning("%d objects already rendered", len(book["obj_nums"]))
        self._draw_fft_batches(base, logger, fft_objects, full_image, book, chk_name, image_num)

        # photon batches: nbatch clipped to the bright objects, nsubbatch to the shortest batch (Q9)
        nbatch = max(min(self.nbatch, len(phot_objects)), 1)
        batches = self.make_photon_batches(config, base, logger, phot_objects, faint_objects, nbatch)
        done = book["photon_batch"]
        if done > 0:
            logger.warning("Photon batches [0, %d) / %d already rendered - skipping", done, nbatch)
            batches = batches[done:]
        nsub = max(min(self.nsubbatch, min((len(b) for b in batches), default=0)), 1)
        logger.warning("Splitting photon batches into %d subbatches.", nsub)

        base["image_pos"] = None
        base["stamp_center"] = None
        ops_cfg = {"photon_ops": base.get("stamp", {}).get("photon_ops", [])}
        photon_ops = galsim.config.BuildPhotonOps(ops_cfg, 'photon_ops', base, logger)
        local_wcs = base['wcs'].local(full_image.true_center)
        if sensor is None:
            sensor = galsim.Sensor()

        path = None
        if full_image.dtype in (np.float32, np.float64) and any(batches):
            path = PooledDevicePath.recognise(photon_ops, sensor, local_wcs)
        self.last_route = "host" if path is None else "device"
        if path is not None:
            path.begin(full_image)
        # The device route is software-pipelined against the interpreter.  The gather + upload of sub-batch s runs on
        # a helper thread while this loop builds the stamps of sub-batch s + 1; the fused launch of s is queued when
        # that upload has finished (``drain``), right before the upload of s + 1 (whose pointer tables are ready by then) starts.  A checkpoint's image is
        # snapshotted on the device right after the last launch of its batch -- what is saved is exactly the state
        # after that batch -- travels to the host on a side stream and is written to the file while the next
        # upload runs (``flush_pending``).  The order of the device work is the reference's.
        pending = None  # batch number whose checkpoint awaits its image
        inflight = None  # [resume, recalc, batch number to checkpoint after this launch or None]
        image_current = False  # full_image.array holds the final pixels already
        import os as _os
        import time as _time

        prof = {} if _os.environ.get("B2_PLUGIN_PROFILE") else None  # host seconds per phase of the loop

        def lap(name, t0):
            if prof is not None:
                prof[name] = prof.get(name, 0.0) + _time.perf_counter() - t0
            return _time.perf_counter()

        def flush_pending():
            nonlocal pending
            if pending is not None:
                path.snapshot_finish(full_image)
                self.save_checkpoint(self.checkpoint, chk_name, base, full_image, book["stamps"], book["vars"],
                                     book["obj_nums"], pending)
                pending = None

        def drain():
            nonlocal inflight, pending, image_current
            if inflight is None:
                return
            resume_, recalc_, closes = inflight
            inflight = None
            t = _time.perf_counter()
            path.add_wait()
            t = lap("wait for the upload", t)
            path.step(resume=resume_, recalc=recalc_)
            t = lap("launch", t)
            image_current = False
            if closes is not None:
                flush_pending()  # (only if a batch had no sub-batch after it to do this)
                t = _time.perf_counter()
                path.snapshot_begin()
                pending = closes
                lap("snapshot_begin", t)

        for batch_num, batch in enumerate(batches, start=done):
            if not batch:
                continue
            if nbatch > 1:
                logger.warning("Starting photon batch %d/%d.", batch_num + 1, nbatch)
            base['index_key'] = 'image_num'
            for sub_num, sub in enumerate(self.make_photon_subbatches(batch, nsub)):
                t = _time.perf_counter()
                stamps, current_vars = self.build_stamps(base, logger, sub)
                t = lap("build_stamps", t)
                resume = batch_num > done or sub_num > 0
                recalc = sub_num == 0
                if path is not None:
                    t = _time.perf_counter()
                    path.add_prepare([stamp.photons for stamp in stamps])  # while the previous upload runs
                    del stamps
                    lap("add_prepare", t)
                    drain()
                    t = _time.perf_counter()
                    path.add_start()
                    t = lap("add_start", t)
                    flush_pending()
                    lap("checkpoint (image to host + save)", t)
                    inflight = [resume, recalc, None]
                else:
                    photons = self.merge_photon_arrays(stamps)
                    del stamps
                    for op in photon_ops:
                        op.applyTo(photons, local_wcs, rng)
                    self.accumulate_photons(photons, full_image, sensor, resume=resume, recalc=recalc)
                    del photons
                book["vars"].extend(v for v in current_vars if v != 0)
            if self.checkpoint is not None:
                if path is not None:
                    if inflight is not None:
                        inflight[2] = batch_num + 1
                else:
                    self.save_checkpoint(self.checkpoint, chk_name, base, full_image, book["stamps"], book["vars"],
                                         book["obj_nums"], batch_num + 1)
        if path is not None:
            drain()
            if pending is not None:
                flush_pending()
                image_current = True  # the last checkpoint's image is the final one
            if not image_current:
                path.read_back(full_image)
            self.last_pooled_photons, self.last_h2d_bytes = path.photons, path.h2d_bytes
            if prof is not None:
                logger.warning("pooled device route, host seconds: %s", {k: round(v, 4) for k, v in prof.items()})
                self.last_profile = prof

        current_var = galsim.config.FlattenNoiseVariance(base, full_image, book["stamps"], tuple(book["vars"]), logger)
        return full_image, current_var


galsim.config.RegisterImageType('LSST_PhotonPoolingImage', B200PhotonPoolingImageBuilder())

# LSST_Photons: imSim's own builder behind the device sensor (re-registered so that listing this module alone in
# ``modules:`` still provides every type name of the path).  LSST_Flat: imSim's set-up, the section loop of addNoise
# on the device.
try:
    from imsim.flat import LSST_FlatBuilder as _ImsimFlatBuilder  # noqa: E402
    from imsim.stamp import LSST_PhotonsBuilder as _ImsimPhotonsBuilder  # noqa: E402
except ImportError:  # a stripped-down imsim: the two types stay whatever ``imsim`` registered
    _ImsimFlatBuilder = _ImsimPhotonsBuilder = None


def _sed_bandpass_cdf(sed, bandpass, n=2048):
    """CDF table of the photon wavelengths ``galsim.WavelengthSampler(sed, bandpass)`` draws: density
    sed(w) * bandpass(w) on the bandpass's range (flat.py:172-175)."""
    from .flat import wavelength_cdf

    lo, hi = float(bandpass.blue_limit), float(bandpass.red_limit)
    wave = np.union1d(np.linspace(lo, hi, n), np.asarray(getattr(bandpass, "wave_list", ()), dtype=float))
    wave = wave[(wave >= lo) & (wave <= hi)]
    return wavelength_cdf(wave, np.asarray(sed(wave), dtype=float) * np.asarray(bandpass(wave), dtype=float))


if _ImsimFlatBuilder is not None:
    class B200FlatBuilder(_ImsimFlatBuilder):
        """``LSST_Flat`` (imsim/flat.py:15-300).  ``setup`` / ``buildImage`` are imSim's; ``addNoise`` keeps its
        sequence -- sections with a border, ``niter`` iterations of ``max_counts_per_iter``, pixel areas from the charge
        collected so far times the WCS sky image in the area branch, photons with SED wavelengths through
        ``SiliconSensor.accumulate`` in the ``sed`` branch (flat.py:131-279) -- but runs it through
        ``imsim_b200.flat.build_flat``: sections stay in HBM, the Poisson realisation and the photon generation
        happen on the device.  Anything the device loop does not cover (another sensor class, a noise type other
        than Poisson, checkpointed sections) goes to imSim's own loop, which then drives the device sensor per
        section."""

        last_route = None

        def addNoise(self, image, config, base, image_num, obj_num, current_var, logger):
            from .flat import build_flat

            sensor = base.get('sensor', None)
            noise_type = (base.get('image', {}).get('noise', {}) or {}).get('type', 'Poisson')
            if not isinstance(sensor, _B200Sensor) or getattr(self, 'checkpoint', None) is not None \
                    or noise_type != 'Poisson':
                self.last_route = "host"
                return super().addNoise(image, config, base, image_num, obj_num, current_var, logger)
            self.last_route = "device"
            base['current_noise_image'] = base['current_image']
            rng = galsim.config.GetRNG(config, base, logger=logger, tag='LSST_Flat')
            sensor.updateRNG(rng)
            seed = (int(rng.raw()) << 32) | int(rng.raw())
            b = image.bounds
            sed_cdf = base_level = None
            if self.sed is None:
                # relative pixel areas from the WCS, mean 1 over the bordered image (flat.py:160-168)
                sky = galsim.ImageF(b.withBorder(self.buffer_size), wcs=image.wcs)
                sky.wcs.makeSkyImage(sky, sky_level=1.)
                rel = np.asarray(sky.array, dtype=np.float64) / float(sky.array.mean())
                sx0, sy0 = sky.bounds.xmin, sky.bounds.ymin

                def base_level(sec):
                    ny_, nx_ = sec.array.shape
                    return rel[sec.ymin - sy0:sec.ymin - sy0 + ny_, sec.xmin - sx0:sec.xmin - sx0 + nx_]
            else:
                if 'bandpass' not in base:
                    raise RuntimeError('Using sed with flat builder requires a valid bandpass')
                sed_cdf = _sed_bandpass_cdf(self.sed, base['bandpass'])
            from .sensor import Image as _Image

            target = _Image(image.array, b.xmin, b.ymin)
            nphot = build_flat(target, self.counts_per_pixel, sensor, rng=np.random.default_rng(seed),
                               max_counts_per_iter=self.max_counts_per_iter, nx=self.nx, ny=self.ny,
                               buffer_size=self.buffer_size, sed_cdf=sed_cdf, base_level=base_level, logger=logger)
            if self.sed is not None:
                logger.info('Accumulated %s photons in total.', nphot)

    galsim.config.RegisterImageType('LSST_Flat', B200FlatBuilder())
if _ImsimPhotonsBuilder is not None:
    galsim.config.RegisterStampType('LSST_Photons', _ImsimPhotonsBuilder())
