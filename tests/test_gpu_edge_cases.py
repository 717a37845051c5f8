"""Edge cases and error behaviour of the host-facing API (the cases the reference handles in
imsim/photon_ops.py:139-140, imsim/photon_pooling.py:195-225, galsim/sensor.py)."""
import warnings

import numpy as np
import pytest

import helpers
from imsim_b200 import B2Error, PhotonArray
from imsim_b200.photon_pooling import LSST_PhotonPoolingImageBuilder as Builder
from imsim_b200.sensor import Image, Sensor, SiliconSensor

pytestmark = pytest.mark.gpu


def _sensor(**kw):
    cfg, dat = helpers.sensor_model("lsst_itl_50_4")
    return SiliconSensor(config=cfg, vertex_data=dat, rng=1, absorption_table=helpers.absorption(), **kw)


def _ops(stamp_center=None, shift_photons=False):
    from imsim_b200 import RubinDiffraction, RubinDiffractionOptics, RubinOptics

    su = helpers.oracle_setup()
    cam = {"R22_S11": su.detector}
    plain = RubinOptics(su.telescope, None, su.img_wcs, stamp_center, su.icrf_to_field, "R22_S11", cam,
                        shift_photons=shift_photons)
    rd = RubinDiffraction(su.telescope, helpers.RUBIN_LAT, np.radians(67.0), np.radians(213.0), su.img_wcs,
                          su.icrf_to_field, stamp_center=stamp_center, shift_photons=shift_photons)
    both = RubinDiffractionOptics(su.telescope, None, stamp_center, "R22_S11", cam, rd, shift_photons=shift_photons)
    return su, plain, rd, both


def test_empty_photon_arrays():
    s = _sensor()
    img = Image(np.zeros((8, 8), np.float32))
    assert s.accumulate(PhotonArray(0), img) == 0.0
    su, plain, rd, both = _ops()
    pa = PhotonArray(0)
    pa.pupil_u, pa.pupil_v, pa.time, pa.wavelength  # allocate
    plain.applyTo(pa)
    rd.applyTo(pa)
    assert pa.size() == 0


def test_ops_assert_on_missing_pupil_or_time():
    su, plain, rd, both = _ops()
    pa = PhotonArray(10, x=np.zeros(10), y=np.zeros(10), flux=np.ones(10), wavelength=np.full(10, 600.0))
    for op in (plain, rd, both):
        with pytest.raises(AssertionError):
            op.applyTo(pa)
    pa.pupil_u = 3.0
    pa.pupil_v = 0.5
    with pytest.raises(AssertionError):  # times still missing (photon_ops.py:140)
        plain.applyTo(pa)


def test_photon_ops_host_api_pupil_untouched_and_shift_symmetry():
    """tests/test_photon_ops.py:173-196 (pupil coordinates unchanged; photons stay near their
    stamp-relative position) and the stamp_center quirk Q3 of the reference."""
    p = helpers.test_photon_arrays(n=5000)
    center = (809.5, 3432.5)
    su, plain, rd, both = _ops(stamp_center=center, shift_photons=True)

    def fresh():  # PhotonArray keeps the arrays it is given: copy so the objects do not alias
        return PhotonArray(5000, **{k: v.copy() for k, v in p.items()})

    pa = fresh()
    u0, v0 = pa.pupil_u.copy(), pa.pupil_v.copy()
    plain.applyTo(pa)
    np.testing.assert_array_equal(pa.pupil_u, u0)
    np.testing.assert_array_equal(pa.pupil_v, v0)
    ok = pa.flux > 0
    assert ok.mean() > 0.9
    # stamp-relative coordinates in and out: photons stay within a few pixels (aberrations + defocus)
    assert np.abs(pa.x[ok] - p["x"][ok]).max() < 20 and np.abs(pa.y[ok] - p["y"][ok]).max() < 20
    assert plain.last_stats.n_vignetted == (~ok).sum()
    # same photons given in full-image coordinates with no stamp_center land at the same place
    su2, plain2, _, _ = _ops(stamp_center=None)
    pb = fresh()
    pb.x += center[0]
    pb.y += center[1]
    plain2.applyTo(pb)
    np.testing.assert_allclose(pb.x[ok] - center[0], pa.x[ok], atol=1e-8)
    # combined op == diffraction then optics to 6 decimals with the same seed (tests/test_photon_ops.py:281-318)
    pc, pd = fresh(), fresh()
    both.applyTo(pc, rng=42)
    rd.applyTo(pd, rng=42)
    plain.applyTo(pd, rng=42)
    okc = (pc.flux > 0) & (pd.flux > 0)
    np.testing.assert_array_almost_equal(pc.x[okc], pd.x[okc], decimal=6)
    np.testing.assert_array_almost_equal(pc.dxdz[okc], pd.dxdz[okc], decimal=6)


def test_resume_requires_same_image():
    s = _sensor()
    rng = np.random.default_rng(0)
    pa = PhotonArray(100, x=rng.uniform(1, 8, 100), y=rng.uniform(1, 8, 100), flux=np.ones(100))
    a, b = Image(np.zeros((8, 8), np.float32)), Image(np.zeros((8, 8), np.float32))
    with pytest.raises(B2Error):
        s.accumulate(pa, a, resume=True)  # no previous call
    s.accumulate(pa, a)
    with pytest.raises(B2Error):
        s.accumulate(pa, b, resume=True)  # another image
    before = a.array.sum()
    added = s.accumulate(pa, a, resume=True)
    assert a.array.sum() == before + added and 90 <= added <= 100  # a few diffuse off the 8x8 image


def test_photons_off_image_and_plain_sensor():
    s = _sensor(nrecalc=0)
    n = 1000
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-50, -10, n // 2), rng.uniform(2, 7, n // 2)])
    y = rng.uniform(2, 7, n)
    pa = PhotonArray(n, x=x, y=y, flux=np.ones(n))
    img = Image(np.zeros((8, 8), np.float64))
    added = s.accumulate(pa, img)
    assert added == img.array.sum() == n // 2
    # galsim.Sensor: PhotonArray.addTo semantics, half-integer rounding, fractional and negative flux
    plain = Sensor()
    img2 = Image(np.zeros((4, 4), np.float64), 1, 1)
    pb = PhotonArray(5, x=np.array([1.0, 1.49, 1.5, 4.49, 4.5]), y=np.array([1.0, 1.0, 1.0, 4.0, 4.0]),
                     flux=np.array([1.0, 0.25, -2.0, 3.0, 7.0]))
    got = plain.accumulate(pb, img2)
    assert got == pytest.approx(2.25)  # the last photon rounds to x=5: off the image
    assert img2.array[0, 0] == 1.25 and img2.array[0, 1] == -2.0 and img2.array[3, 3] == 3.0


def test_integer_image_goes_through_a_temporary(recwarn):
    """imsim/photon_pooling.py:213-225: integer images use a temporary float image, warn about
    resume / recalc being ignored."""
    s = _sensor(nrecalc=0)
    rng = np.random.default_rng(2)
    pa = PhotonArray(500, x=rng.uniform(2, 7, 500), y=rng.uniform(2, 7, 500), flux=np.ones(500))
    img = Image(np.zeros((8, 8), np.int32))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        Builder.accumulate_photons(pa, img, s, resume=True, recalc=False)
        assert len(w) == 2
    assert img.array.sum() == 500
    fimg = Image(np.zeros((8, 8), np.float32))
    Builder.accumulate_photons(pa, fimg, s, resume=False, recalc=True)
    assert fimg.array.sum() == 500 and s.last_stats.n_updates == 0


def test_no_angles_no_wavelengths_and_strength_scaling():
    """PhotonArray without dxdz/dydz/wavelength: conversion at 1 micron (achromatic shooting,
    tests/test_sensor_models.py); strength rescales num_elec and nrecalc like GalSim."""
    s = _sensor(strength=2.0, nrecalc=10000)
    assert s.pod.num_elec == 50000.0 and s.pod.nrecalc == 5000.0
    rng = np.random.default_rng(3)
    n = 20000
    pa = PhotonArray(n, x=rng.normal(0, 1, n), y=rng.normal(0, 1, n), flux=np.ones(n))
    img = Image(np.zeros((17, 17), np.float32), -8, -8)
    added = s.accumulate(pa, img)
    assert added == img.array.sum() == n
    assert s.last_stats.n_updates == 4 and s.last_stats.n_dropped_bottom == 0


def test_restart_from_a_checkpointed_image_rebuilds_the_boundaries():
    """checkpoints hold the image but not the sensor's pixel boundaries; after a restart the first
    accumulate runs with resume=False on the restored image and rebuilds them from it in one step
    (imsim/photon_pooling.py:159,447-466).  The distortions are linear in the charge, so the restarted run
    lands (to rounding of the float32 boundary points) the same electrons as the uninterrupted one."""
    import torch

    from imsim_b200 import OpticsContext
    from imsim_b200.photon_pooling import DevicePhotons, PhotonPool
    from imsim_b200.synthetic import synthetic_photons

    su = helpers.oracle_setup()
    cfg, dat = helpers.sensor_model("lsst_e2v_50_4")
    tr = helpers.tree_ring_table()
    n = 1_500_000
    x, y, wl, flux = synthetic_photons(2 * n, kind="stars", n_stars=12, seed=21)

    def make():
        ctx = OpticsContext(device=0, stream=torch.cuda.current_stream())
        ctx.set_telescope(su.telescope)
        ctx.set_wcs(su.img_wcs, su.icrf_to_field)
        ctx.set_detector(su.detector)
        ctx.set_diffraction(helpers.default_diffraction())
        sensor = SiliconSensor(config=cfg, vertex_data=dat, nrecalc=0, strength=1.0, rng=77, treering_func=tr[1],
                               treering_center=tr[0], absorption_table=helpers.absorption(), context=ctx)
        return ctx, sensor, PhotonPool(ctx, sensor, exptime=30.0, seed=5)

    def batch(k):
        dp = DevicePhotons(n)
        for f, a in (("x", x), ("y", y), ("wavelength", wl), ("flux", flux)):
            getattr(dp, f).copy_(torch.as_tensor(a[k * n:(k + 1) * n]))
        return dp

    # uninterrupted: two batches, boundaries updated from the first batch's charge at the start of the second
    ctx, sensor, pool = make()
    img = Image(np.zeros((su.detector.ny, su.detector.nx), np.float32), 0, 0)
    pool.process(batch(0), img, resume=False, recalc=False)
    sensor.read_image(img)
    checkpoint = img.array.copy()
    pool.process(batch(1), img, resume=True, recalc=True)
    sensor.read_image(img)
    full = img.array.copy()
    sensor.close()
    # restart: a fresh sensor, the checkpointed image, resume=False, the pool's photon counter restored
    ctx2, sensor2, pool2 = make()
    pool2.offset = n
    sensor2._photon_offset = n
    img2 = Image(checkpoint.copy(), 0, 0)
    pool2.process(batch(1), img2, resume=False, recalc=False)
    sensor2.read_image(img2)
    assert checkpoint.sum() > 0.9 * n and abs(img2.array.sum() - full.sum()) <= 2
    diff = np.abs(img2.array.astype(np.float64) - full)
    # identical up to the few photons within a float32 ulp of a moved boundary
    assert diff.sum() <= 20 and diff.max() <= 2, (diff.sum(), diff.max())
    assert not np.array_equal(full, checkpoint)


def test_empty_first_batch_then_resume_on_a_reused_sensor():
    """A first pooled sub-batch may be empty (all stamps None).  The empty non-resume call must still bind and
    initialise the new image, so that the following resume=True call neither fails nor lands on the buffers of
    the previous CCD (and read-back must not overwrite what the new image already held)."""
    s = _sensor()
    rng = np.random.default_rng(5)
    n = 20000
    first = Image(np.zeros((32, 32), np.float32))
    s.accumulate(PhotonArray(n, x=rng.uniform(5, 27, n), y=rng.uniform(5, 27, n), flux=np.ones(n)), first)
    assert first.array.sum() > 0
    # same-sized new image that already holds something (an FFT object drawn earlier)
    second = Image(np.zeros((32, 32), np.float32))
    second.array[3, 4] = 123.0
    assert s.accumulate(PhotonArray(0), second, resume=False) == 0.0
    m = 5000
    pa = PhotonArray(m, x=rng.uniform(10, 20, m), y=rng.uniform(10, 20, m), flux=np.ones(m))
    added = s.accumulate(pa, second, resume=True)
    assert added == m
    assert second.array[3, 4] == 123.0
    assert second.array.sum() == pytest.approx(123.0 + m)
    assert second.array[:8].sum() == 123.0  # nothing of the first image's 5..27 field leaked in
    with pytest.raises(B2Error):
        s.accumulate(pa, Image(np.zeros((32, 32), np.float32)), resume=True)
