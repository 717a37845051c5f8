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
    for a, b in zip(traced[True], traced[False]):
        assert np.array_equal(a, b)
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
