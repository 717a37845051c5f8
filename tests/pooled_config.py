"""The image section of config/imsim-config-photon-pooling.yaml against the stand-in config engine of tests/stubs
-- TEST / BENCH INFRASTRUCTURE shared by tests/test_gpu_plugin_pooling.py and bench.py's e2e leg.

``make_run`` assembles what ``galsim config.yaml`` has in ``base`` when ``LSST_PhotonPoolingImage.buildImage`` is
called for one CCD: det_telescope, _icrf_to_field, wcs, the sensor built by the registered ``Silicon`` type from
``image.sensor`` (nrecalc: 0), the photon-op list of config/imsim-config.yaml:281-320 under ``stamp.photon_ops``,
and a catalogue whose stamps hand over ordinary (pageable) numpy PhotonArrays -- GalSim's shooters in a real run.
"""
import importlib
import os
import sys
import tempfile

import numpy as np

STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")
_SENSOR_DIR = {}


def load_plugin():
    """Import imsim_b200.galsim_plugin against the stand-in ``galsim`` / ``imsim`` (a real GalSim is not installable
    here); returns (plugin module, galsim stand-in)."""
    if STUBS not in sys.path:
        sys.path.insert(0, STUBS)
    import galsim

    mod = importlib.import_module("imsim_b200.galsim_plugin")
    return mod, galsim


def sensor_model_files(name="lsst_e2v_50_4"):
    """Write ``<name>.cfg`` / ``<name>.dat`` (the reference's data/sensor_models files) to a temporary directory and
    return the path prefix the ``Silicon`` sensor type is given as ``name``."""
    if name not in _SENSOR_DIR:
        from imsim_b200 import workload_data as wd

        cfg, dat = wd.sensor_model(name)
        d = tempfile.mkdtemp(prefix="b2_sensor_")
        with open(os.path.join(d, name + ".cfg"), "w") as f:
            for k, v in cfg.items():
                f.write("%s = %r\n" % (k, v))
        with open(os.path.join(d, name + ".dat"), "w") as f:
            f.write("X0 Y0 Theta X Y\n")
            np.savetxt(f, dat, fmt="%.10g")
        _SENSOR_DIR[name] = os.path.join(d, name)
    return _SENSOR_DIR[name]


def make_run(setup, catalog, photon_source, nbatch=10, nsubbatch=1, diffraction=True, sensor_nrecalc=0.0,
             xsize=None, ysize=None, seed=7, exptime=30.0):
    """(builder, config, base) ready for ``builder.setup(...)`` + ``builder.buildImage(...)``."""
    plugin, galsim = load_plugin()
    import imsim.camera as cam

    from imsim_b200 import workload_data as wd

    cam._CAMERAS["LsstCamSim"] = {setup.det_name: setup.detector}
    tr = wd.tree_ring_table(setup.det_name if setup.det_name in ("R22_S11",) else "R22_S11")
    optics = {"type": "RubinDiffractionOptics" if diffraction else "RubinOptics", "boresight": galsim.CelestialCoord(),
              "camera": "LsstCamSim", "det_name": setup.det_name}
    if diffraction:
        optics.update(altitude=np.radians(67.0), azimuth=np.radians(213.0))
    image_cfg = {"type": "LSST_PhotonPoolingImage", "nbatch": nbatch, "nsubbatch": nsubbatch, "det_name": setup.det_name,
                 "xsize": xsize or setup.detector.nx, "ysize": ysize or setup.detector.ny,
                 "sensor": {"type": "Silicon", "name": sensor_model_files("lsst_e2v_50_4"), "strength": 1.0,
                            "nrecalc": sensor_nrecalc, "treering_func": tr[1],
                            "treering_center": galsim.PositionD(*tr[0])}}
    base = {"det_name": setup.det_name, "det_num": 94, "det_telescope": setup.telescope,
            "_icrf_to_field": setup.icrf_to_field, "wcs": setup.img_wcs, "rng": galsim.BaseDeviate(seed),
            "stamp_center": None, "image": image_cfg, "_catalog": catalog, "_photon_source": photon_source,
            "stamp": {"type": "LSST_Photons", "photon_ops": [
                {"type": "TimeSampler", "t0": 0.0, "exptime": exptime},
                {"type": "PupilAnnulusSampler", "R_outer": 4.18, "R_inner": 2.55},
                optics,
                {"type": "Refraction", "index_ratio": 3.9}]}}

    class _Img:  # base['current_image'].wcs is what the optics deserialisers read (imsim/photon_ops.py:408)
        wcs = setup.img_wcs

    base["current_image"] = _Img()
    base["sensor"] = galsim.config.REGISTRY["sensor"]["Silicon"].buildSensor(image_cfg["sensor"], base, None)
    builder = galsim.config.REGISTRY["image"]["LSST_PhotonPoolingImage"]
    return builder, image_cfg, base


class Quiet:
    def info(self, *a):
        pass

    warning = debug = error = info
