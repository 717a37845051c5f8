"""imsim_b200.galsim_plugin against a stand-in for GalSim's config registry (tests/stubs): the plugin registers
the reference's type names and its deserialisers build working device ops from a config dict + ``base``.
GalSim itself is not available in the build container; this checks the wiring, not GalSim."""
import importlib
import os
import sys

import numpy as np
import pytest

import helpers

STUBS = os.path.join(os.path.dirname(__file__), "stubs")


@pytest.fixture()
def plugin(monkeypatch):
    for name in [m for m in sys.modules if m == "galsim" or m.startswith("galsim.") or m == "imsim" or
                 m.startswith("imsim.")]:
        monkeypatch.delitem(sys.modules, name)
    monkeypatch.syspath_prepend(STUBS)
    sys.modules.pop("imsim_b200.galsim_plugin", None)
    mod = importlib.import_module("imsim_b200.galsim_plugin")
    yield mod
    sys.modules.pop("imsim_b200.galsim_plugin", None)
    for name in [m for m in sys.modules if m == "galsim" or m.startswith("galsim.") or m == "imsim" or
                 m.startswith("imsim.")]:
        sys.modules.pop(name, None)


def test_plugin_registers_the_reference_type_names(plugin):
    import galsim

    reg = galsim.config.REGISTRY
    # imsim/photon_ops.py:400-451, imsim/treerings.py:241-243, galsim sensor type used by imsim-config.yaml:230-235
    assert set(reg["photon_op"]) == {"RubinOptics", "RubinDiffractionOptics", "RubinDiffraction"}
    assert all(inp == "telescope" for _, inp in reg["photon_op"].values())
    assert set(reg["value"]) == {"TreeRingCenter", "TreeRingFunc"} and "tree_rings" in reg["input"]
    assert "Silicon" in reg["sensor"]
    # parameter tables mirror the reference (including the duplicated altitude / azimuth, photon_ops.py:173-186)
    from imsim_b200 import photon_ops as ops

    assert set(ops.RubinOptics._req_params) == {"boresight", "camera", "det_name"}
    assert ops.RubinOptics._req_params["boresight"] is galsim.CelestialCoord
    assert {"altitude", "azimuth"} <= set(ops.RubinDiffractionOptics._req_params)
    assert {"altitude", "azimuth", "latitude", "disable_field_rotation", "shift_photons"} <= \
        set(ops.RubinDiffractionOptics._opt_params)
    assert issubclass(plugin.B200SiliconSensor, galsim.SiliconSensor)  # photon_pooling.py:209
    # an unknown parameter is rejected by the config layer, like GalSim does
    builder, _ = reg["photon_op"]["RubinOptics"]
    with pytest.raises(KeyError):
        builder.buildPhotonOp({"type": "RubinOptics", "camera": "LsstCamSim", "det_name": "R22_S11", "bogus": 1},
                              {"stamp_center": None}, None)
    # tree-ring value types look their detector up in the input object
    from imsim_b200.treerings import TreeRings

    class FakeTR:
        def get_center(self, det):
            return ("center", det)

        def get_func(self, det):
            return ("func", det)

    base = {"_input_objs": {"tree_rings": FakeTR()}}
    assert plugin.TreeRingCenter({"det_name": "R22_S11"}, base, None)[0] == ("center", "R22_S11")
    assert plugin.TreeRingFunc({"det_name": "R22_S11"}, base, None)[0] == ("func", "R22_S11")
    assert TreeRings._req_params == plugin._tree_rings_with_data_dir._req_params


@pytest.mark.gpu
def test_deserialised_ops_trace_photons(plugin):
    """Build RubinDiffractionOptics and the Silicon sensor from config dicts and run them on a PhotonArray."""
    import galsim
    import imsim.camera as cam

    from imsim_b200 import OpticsContext, PhotonArray
    from imsim_b200.sensor import Image
    from imsim_b200.synthetic import gpu_tracer, make_detector_setup

    su = make_detector_setup(gpu_tracer(OpticsContext(device=0)), "R22_S11", rot_tel_pos=0.3)
    cam._CAMERAS["LsstCamSim"] = {"R22_S11": su.detector}

    class Img:
        wcs = su.img_wcs

    base = {"det_telescope": su.telescope, "_icrf_to_field": su.icrf_to_field, "current_image": Img(),
            "stamp_center": None, "det_num": 94, "rng": galsim.BaseDeviate(7)}
    builder, _ = galsim.config.REGISTRY["photon_op"]["RubinDiffractionOptics"]
    op = builder.buildPhotonOp({"type": "RubinDiffractionOptics", "boresight": galsim.CelestialCoord(), "camera": "LsstCamSim",
                                "det_name": "R22_S11", "altitude": 1.1, "azimuth": 0.4}, base, None)
    p = helpers.test_photon_arrays(n=20000, center=(2000.0, 2000.0))
    pa = PhotonArray(20000, x=p["x"].copy(), y=p["y"].copy(), flux=p["flux"].copy(), wavelength=p["wavelength"].copy(),
                     pupil_u=p["pupil_u"].copy(), pupil_v=p["pupil_v"].copy(), time=np.random.default_rng(1).uniform(0, 30, 20000))
    op.applyTo(pa, rng=galsim.BaseDeviate(3))
    ok = pa.flux > 0
    # (a few photons grazing a spider vane are kicked thousands of pixels: the diffraction spikes)
    assert ok.mean() > 0.8 and np.median(np.abs(pa.x[ok] - p["x"][ok])) < 3
    assert np.all(np.abs(pa.dxdz[ok]) < 0.5)
    cfg_dir = helpers.sensor_model_files("lsst_e2v_50_4") if hasattr(helpers, "sensor_model_files") else None
    sb = galsim.config.REGISTRY["sensor"]["Silicon"]
    if cfg_dir is not None:
        sensor = sb.buildSensor({"type": "Silicon", "name": cfg_dir, "nrecalc": 10000.0}, base, None)
        assert isinstance(sensor, galsim.SiliconSensor)
        img = Image(np.zeros((4004, 4096), np.float32), 0, 0)
        added = sensor.accumulate(pa, img)
        assert 0.9 * ok.sum() < added <= ok.sum() and abs(img.array.sum() - added) < 1
