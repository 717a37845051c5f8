"""``LSST_Flat`` of the plugin (galsim_plugin.B200FlatBuilder): imSim's builder with the section loop of ``addNoise``
(imsim/flat.py:131-279) on the device.  Stand-in config engine (tests/stubs); the checks are the reference's own
(tests/test_flats.py: mean level, sub-Poisson variance from brighter-fatter) plus equality with ``build_flat``, which
tests/test_flats.py / test_gpu_flat.py pin to the reference's control flow."""
import numpy as np
import pytest

import helpers
import pooled_config as pc

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, scope="module")
def _stand_in_engine_only_here():
    import sys

    yield
    if pc.STUBS in sys.path:
        sys.path.remove(pc.STUBS)
    for name in [m for m in sys.modules if m in ("galsim", "imsim", "imsim_b200.galsim_plugin")
                 or m.startswith("galsim.") or m.startswith("imsim.")]:
        sys.modules.pop(name, None)


def _builder(plugin, counts, sed=None):
    b = plugin.B200FlatBuilder()
    b.counts_per_pixel, b.max_counts_per_iter = counts, 1000.0
    b.nx, b.ny, b.buffer_size, b.sed, b.checkpoint = 2, 2, 5, sed, None
    return b


def _sensor(plugin, galsim, nrecalc):
    tr = helpers.tree_ring_table("R22_S11")
    return plugin.B200SiliconSensor(name=pc.sensor_model_files("lsst_itl_50_8"), strength=1.0, nrecalc=nrecalc,
                                    treering_func=tr[1], treering_center=galsim.PositionD(*tr[0]),
                                    rng=galsim.BaseDeviate(3))


def test_area_branch_runs_on_the_device_and_equals_build_flat():
    from imsim_b200.flat import build_flat, flat_nrecalc
    from imsim_b200.sensor import Image

    plugin, galsim = pc.load_plugin()
    counts = 40000.0
    nx, ny = 160, 128
    sensor = _sensor(plugin, galsim, flat_nrecalc(nx, ny, 2, 2))
    image = galsim.ImageF(nx, ny, wcs=galsim.PixelScale(0.2))
    base = {"sensor": sensor, "current_image": image, "rng": galsim.BaseDeviate(11),
            "image": {"noise": {"type": "Poisson"}}}
    b = _builder(plugin, counts)
    b.addNoise(image, {}, base, 0, 0, 0, pc.Quiet())
    assert b.last_route == "device"
    a = image.array[20:-20, 20:-20].astype(np.float64)
    np.testing.assert_allclose(a.mean(), counts, rtol=2e-3)  # tree rings move the mean of a window by ~1e-3
    assert a.var() < counts  # brighter-fatter: sub-Poisson (tests/test_flats.py:70-77)
    # the same call made by hand: same seeds, same image
    rng = galsim.BaseDeviate(11)
    sensor.updateRNG(rng)
    seed = (int(rng.raw()) << 32) | int(rng.raw())
    want = Image(np.zeros((ny, nx), np.float32), 1, 1)
    sky = np.full((ny + 10, nx + 10), 0.2 ** 2, np.float32)  # makeSkyImage of the bordered image, sky_level = 1
    rel = float((sky.astype(np.float64) / float(sky.mean()))[0, 0])
    build_flat(want, counts, sensor, rng=np.random.default_rng(seed), max_counts_per_iter=1000.0, nx=2, ny=2,
               buffer_size=5, base_level=lambda sec: np.full(sec.array.shape, rel))
    np.testing.assert_array_equal(image.array, want.array)


def test_sed_branch_shoots_photons_with_the_bandpass_weighted_sed():
    plugin, galsim = pc.load_plugin()

    class Band:  # the part of galsim.Bandpass the builder reads
        blue_limit, red_limit = 900.0, 1000.0
        wave_list = np.array([900.0, 930.0, 960.0, 1000.0])

        def __call__(self, w):
            return np.interp(w, [900.0, 1000.0], [0.5, 1.0])

    sed = lambda w: np.interp(w, [930.0, 940.0, 950.0, 960.0], [0.0, 1.0, 1.0, 0.0], left=0.0, right=0.0)  # noqa: E731
    cdf, wave = plugin._sed_bandpass_cdf(sed, Band())
    assert cdf[0] == 0.0 and abs(cdf[-1] - 1.0) < 1e-12 and np.all(np.diff(cdf) >= 0)
    # no photons outside the SED's support, half of them below its centre (the bandpass tilts it slightly redwards)
    assert np.interp(930.0, wave, cdf) < 1e-3 and np.interp(960.0, wave, cdf) > 1.0 - 1e-3
    assert 0.45 < np.interp(945.0, wave, cdf) < 0.5
    counts = 3000.0
    nx, ny = 96, 80
    sensor = _sensor(plugin, galsim, 0.0)
    image = galsim.ImageF(nx, ny, wcs=galsim.PixelScale(0.2))
    base = {"sensor": sensor, "current_image": image, "rng": galsim.BaseDeviate(5), "bandpass": Band(),
            "image": {"noise": {"type": "Poisson"}}}
    b = _builder(plugin, counts, sed=sed)
    b.addNoise(image, {}, base, 0, 0, 0, pc.Quiet())
    assert b.last_route == "device"
    a = image.array[10:-10, 10:-10].astype(np.float64)
    # 930-960 nm: a fifth of the photons leave through the back of the 100 um sensor (tests/test_flats.py:167-216)
    assert 0.7 * counts < a.mean() < 0.9 * counts
    with pytest.raises(RuntimeError):
        base.pop("bandpass")
        b.addNoise(image, {}, base, 0, 0, 0, pc.Quiet())


def test_other_sensors_and_noise_types_take_imsims_own_loop():
    plugin, galsim = pc.load_plugin()
    image = galsim.ImageF(32, 32, wcs=galsim.PixelScale(0.2))
    b = _builder(plugin, 100.0)
    for base in ({"sensor": galsim.Sensor(), "current_image": image, "image": {"noise": {"type": "Poisson"}}},
                 {"sensor": _sensor(plugin, galsim, 0.0), "current_image": image, "image": {"noise": {"type": "CCD"}}}):
        with pytest.raises(AttributeError):  # the stand-in imsim builder has no addNoise: the call went to super()
            b.addNoise(image, {}, base, 0, 0, 0, pc.Quiet())
        assert b.last_route == "host"
