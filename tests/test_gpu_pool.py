"""The fused pool step (one kernel: sampler -> optics -> sensor fast path) against the separate
kernels and, through them, against the oracle."""
import numpy as np
import pytest

import helpers
from imsim_b200.sensor import Image, SiliconSensor

pytestmark = pytest.mark.gpu


def _setup(nrecalc=0):
    import torch

    from imsim_b200 import OpticsContext
    from imsim_b200.photon_pooling import DevicePhotons, PhotonPool, PinnedPhotons
    from imsim_b200.synthetic import synthetic_photons

    su = helpers.oracle_setup()
    ctx = OpticsContext(device=0, stream=torch.cuda.current_stream())
    ctx.set_telescope(su.telescope)
    ctx.set_wcs(su.img_wcs, su.icrf_to_field)
    ctx.set_detector(su.detector)
    ctx.set_diffraction(helpers.default_diffraction())
    cfg, dat = helpers.sensor_model("lsst_e2v_50_4")
    tr = helpers.tree_ring_table()
    sensor = SiliconSensor(config=cfg, vertex_data=dat, nrecalc=nrecalc, rng=77, treering_func=tr[1],
                           treering_center=tr[0], absorption_table=helpers.absorption(), context=ctx)
    n = 400000
    x, y, wl, flux = synthetic_photons(n, kind="stars", n_stars=40, seed=3)
    pin = PinnedPhotons(n)
    pin.x[:], pin.y[:], pin.wavelength[:], pin.flux[:] = x, y, wl, flux
    pool = PhotonPool(ctx, sensor, exptime=30.0, focus_depth=-0.6, index_ratio=3.9, seed=5)
    return su, ctx, sensor, pool, pin, n, DevicePhotons


def test_fused_pool_step_equals_separate_kernels():
    import torch

    images = {}
    traced = {}
    stats = {}
    for fused in (False, True):
        su, ctx, sensor, pool, pin, n, DevicePhotons = _setup()
        img = Image(np.zeros((su.detector.ny, su.detector.nx), np.float32), 0, 0)
        for batch in range(3):
            dp = DevicePhotons(n)
            dp.upload(pin)
            added, ost = pool.process(dp, img, resume=batch > 0, recalc=batch > 0, want_stats=True, fused=fused,
                                      write_back=True)
            st = sensor.last_stats.as_dict()
        sensor.read_image(img)
        torch.cuda.synchronize()
        images[fused] = img.array.copy()
        traced[fused] = [getattr(dp, f).cpu().numpy() for f in ("x", "y", "dxdz", "dydz", "flux")]
        stats[fused] = (added, ost.n_vignetted, st)
        sensor.close()
    assert np.array_equal(images[True], images[False])
    # The traced photons agree to rounding, not always to the last bit: the two kernels inline the same source, but
    # the compiler is free to contract a * b + c * d one way or the other in each of them, and a one-ulp change of a
    # direction cosine is amplified by the spider kick of the few photons that graze a vane.  Bound: 1e-13 of the
    # pixel coordinate (the parity bar against the reference arithmetic is 1e-10), on well under 1 % of the photons.
    for name, a, b in zip(("x", "y", "dxdz", "dydz", "flux"), traced[True], traced[False]):
        ok = np.isfinite(a) & np.isfinite(b)
        assert np.array_equal(np.isfinite(a), np.isfinite(b))
        tol = 4e-10 if name in ("x", "y") else (1e-15 if name != "flux" else 0.0)
        assert np.abs(a[ok] - b[ok]).max() <= tol, name
        assert np.count_nonzero(a[ok] != b[ok]) < 0.01 * a.size, name
    assert stats[True] == stats[False]
    assert images[True].sum() > 0.8 * 3 * 400000


def test_fused_pool_step_rejects_nrecalc_cadence():
    from imsim_b200 import B2Error

    su, ctx, sensor, pool, pin, n, DevicePhotons = _setup(nrecalc=10000)
    dp = DevicePhotons(n)
    dp.upload(pin)
    img = Image(np.zeros((su.detector.ny, su.detector.nx), np.float32), 0, 0)
    with pytest.raises(B2Error):
        pool.process(dp, img, resume=False, recalc=False, fused=True)
    # the default picks the separate kernels for nrecalc > 0
    added, _ = pool.process(dp, img, resume=False, recalc=False, want_stats=True)
    assert added > 0 and sensor.last_stats.n_updates >= 30


def test_full_size_pool_properties():
    """BASELINE's full size (2^25 photons on a 4096 x 4004 CCD), checked through size-independent properties:
    * conservation: the flux the kernel reports as added equals the image sum, and every photon is accounted for
      (landed + vignetted + off the chip / through the substrate);
    * linearity with brighter-fatter off: the image of the pool equals the sum of the images of its two halves
      (photon offsets keep every photon's random draws the same);
    * the fused kernel and the three separate kernels give identical images at this size too."""
    import torch

    from imsim_b200 import OpticsContext
    from imsim_b200.photon_pooling import DevicePhotons, PhotonPool
    from imsim_b200.synthetic import synthetic_photons

    n = 1 << 25
    su = helpers.oracle_setup()
    ctx = OpticsContext(device=0, stream=torch.cuda.current_stream())
    ctx.set_telescope(su.telescope)
    ctx.set_wcs(su.img_wcs, su.icrf_to_field)
    ctx.set_detector(su.detector)
    ctx.set_diffraction(helpers.default_diffraction())
    cfg, dat = helpers.sensor_model("lsst_e2v_50_4")
    tr = helpers.tree_ring_table()
    x, y, wl, flux = synthetic_photons(n, kind="stars", seed=11)
    src = DevicePhotons(n, fields=("x", "y", "flux", "wavelength"))
    for f, a in (("x", x), ("y", y), ("wavelength", wl), ("flux", flux)):
        getattr(src, f).copy_(torch.as_tensor(a))

    def run(strength, fused, parts, recalc=False):
        sensor = SiliconSensor(config=cfg, vertex_data=dat, nrecalc=0, strength=strength, rng=77, treering_func=tr[1],
                               treering_center=tr[0], absorption_table=helpers.absorption(), context=ctx)
        pool = PhotonPool(ctx, sensor, exptime=30.0, seed=5)
        img = Image(np.zeros((su.detector.ny, su.detector.nx), np.float64), 0, 0)
        added = vig = 0
        lo = 0
        for k, m in enumerate(parts):
            dp = DevicePhotons(m)
            for f in ("x", "y", "wavelength", "flux"):
                getattr(dp, f).copy_(getattr(src, f)[lo:lo + m])
            a, ost = pool.process(dp, img, resume=k > 0, recalc=(recalc and k > 0), want_stats=True, fused=fused)
            added += a
            vig += ost.n_vignetted
            lo += m
        sensor.read_image(img)
        st = sensor.last_stats.as_dict()
        sensor.close()
        return img.array.copy(), added, vig, st

    whole, added, vig, st = run(1e-12, True, [n])
    assert abs(whole.sum() - added) < 1e-6 * added and whole.min() >= 0
    assert 0.9 * n < added <= n - vig  # r band: nearly every unvignetted photon converts on the chip
    halves, added2, vig2, _ = run(1e-12, True, [n // 2, n - n // 2])
    assert vig2 == vig and added2 == added
    assert np.array_equal(whole, halves)
    unfused, added3, vig3, _ = run(1e-12, False, [n])
    assert np.array_equal(whole, unfused) and added3 == added and vig3 == vig
    # brighter-fatter acts at batch starts in the pooled cadence (nrecalc = 0, recalc=True): one batch from an empty
    # image is unaffected; in two batches the second sees the first's charge -- boundaries move, photons are not lost
    bf1, added4, vig4, _ = run(1.0, True, [n])
    assert np.array_equal(bf1, whole)
    bf2, added5, vig5, st5 = run(1.0, True, [n // 2, n - n // 2], recalc=True)
    assert vig5 == vig and abs(added5 - added) < 2e-5 * added and not np.array_equal(bf2, whole)
    assert st5["n_updates"] == 1
