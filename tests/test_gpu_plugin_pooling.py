"""``LSST_PhotonPoolingImage`` of the plugin, end to end on the image section of imsim-config-photon-pooling.yaml
(stand-in config engine, tests/pooled_config.py): the builder must take the device route for the configured op
list, keep the reference's batching (nbatch photon batches, faint objects whole in one batch), and produce the same
image -- statistically, the device route draws its samplers from Philox -- as the reference's host loop
(merge_photon_arrays -> op.applyTo ... -> accumulate_photons) run through the very same builder."""
import numpy as np
import pytest

import helpers
import pooled_config as pc

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, scope="module")
def _stand_in_engine_only_here():
    """The stand-in ``galsim`` / ``imsim`` of tests/stubs must not leak into other test modules (the package
    looks for a real GalSim in a few places)."""
    import sys

    yield
    if pc.STUBS in sys.path:
        sys.path.remove(pc.STUBS)
    for name in [m for m in sys.modules if m in ("galsim", "imsim", "imsim_b200.galsim_plugin")
                 or m.startswith("galsim.") or m.startswith("imsim.")]:
        sys.modules.pop(name, None)


def _catalog(rng, n, nx, ny):
    flux = np.concatenate([rng.integers(20000, 90000, n - 6), rng.integers(1, 9, 6)])  # six faint ones (< nbatch)
    return [dict(x=float(rng.uniform(150, nx - 150)), y=float(rng.uniform(150, ny - 150)), flux=int(f)) for f in flux]


def _source(catalog, seed):
    """stamps' photons: Gaussian star images, r-band wavelengths, unit flux -- ordinary numpy arrays"""
    from imsim_b200.photon_array import PhotonArray

    def shoot(obj):
        c = catalog[obj.index]
        g = np.random.default_rng(seed + 7919 * obj.index + int(obj.phot_flux))
        n = int(obj.phot_flux)
        return PhotonArray(n, x=c["x"] + g.normal(0, 1.6, n), y=c["y"] + g.normal(0, 1.6, n), flux=np.ones(n),
                           wavelength=g.uniform(550.0, 690.0, n))

    return shoot


def _run(route, catalog, su, nbatch=4, nsubbatch=3):
    builder, cfg, base = pc.make_run(su, catalog, _source(catalog, 5), nbatch=nbatch, nsubbatch=nsubbatch,
                                     sensor_nrecalc=0.0 if route == "device" else 0.0)
    if route == "host":  # make the op list unrecognisable: an extra (identity) op in front
        class Identity:
            def applyTo(self, photon_array, local_wcs=None, rng=None):
                pass

        import galsim

        galsim.Identity = Identity
        base["stamp"]["photon_ops"].insert(0, {"type": "Identity"})
    builder.setup(cfg, base, 0, 0, [], pc.Quiet())
    image, var = builder.buildImage(cfg, base, 0, 0, pc.Quiet())
    assert builder.last_route == route and var == 0.0
    return image, builder


def test_pooled_builder_takes_the_device_route_and_matches_the_host_loop():
    su = helpers.oracle_setup("R22_S11")
    rng = np.random.default_rng(1)
    cat = _catalog(rng, 40, su.detector.nx, su.detector.ny)
    total = sum(c["flux"] for c in cat)
    img_d, b = _run("device", cat, su)
    # x, y, wavelength as arrays; the flux, constant per stamp, as one number per stamp and batch
    assert b.last_pooled_photons == total and 24 * total < b.last_h2d_bytes < 24.1 * total
    img_h, _ = _run("host", cat, su)
    a, h = img_d.array.astype(np.float64), img_h.array.astype(np.float64)
    assert a.shape == (su.detector.ny, su.detector.nx)
    assert 0.9 * total < a.sum() <= total and abs(a.sum() - h.sum()) < 6 * np.sqrt(total - min(a.sum(), h.sum()) + 1)
    # object by object: same electrons (up to the photons lost to vignetting) and the same centroid
    for c in cat[:34]:
        x0, y0 = int(c["x"]), int(c["y"])
        sa, sh = a[y0 - 12:y0 + 13, x0 - 12:x0 + 13], h[y0 - 12:y0 + 13, x0 - 12:x0 + 13]
        assert abs(sa.sum() - sh.sum()) < 6 * np.sqrt(c["flux"] * 0.2 + 1)
        gy, gx = np.mgrid[y0 - 12:y0 + 13, x0 - 12:x0 + 13]
        for g in (gx, gy):
            assert abs((sa * g).sum() / sa.sum() - (sh * g).sum() / sh.sum()) < 0.06
    assert np.all(a == np.round(a))  # unit-flux photons: integer electrons


def test_pooled_builder_batches_like_the_reference():
    """nbatch photon batches with recalc at each batch start; nbatch clipped to the bright objects (Q9)."""
    su = helpers.oracle_setup("R22_S11")
    rng = np.random.default_rng(2)
    cat = _catalog(rng, 9, su.detector.nx, su.detector.ny)  # 3 bright objects + 6 faint
    image, b = _run("device", cat, su, nbatch=10, nsubbatch=50)
    sensor_stats = b.last_pooled_photons
    assert sensor_stats == sum(c["flux"] for c in cat)
    assert image.array.sum() > 0.9 * sensor_stats


def test_checkpoints_hold_the_image_after_their_batch():
    """With a checkpointer configured the builder writes one checkpoint per photon batch (photon_pooling.py:167-168).
    On the device route the image of batch k travels to the host while batch k+1 uploads; what is saved must still be
    the state after batch k exactly, and the last checkpoint the final image."""
    su = helpers.oracle_setup("R22_S11")
    rng = np.random.default_rng(3)
    cat = _catalog(rng, 30, su.detector.nx, su.detector.ny)
    builder, cfg, base = pc.make_run(su, cat, _source(cat, 5), nbatch=5, nsubbatch=2)
    builder.setup(cfg, base, 0, 0, [], pc.Quiet())
    saved = []

    class Ck:
        file_name = "(memory)"

    builder.checkpoint = Ck()
    builder.load_checkpoint = lambda *a, **k: (None, [], [], [], 0)
    builder.save_checkpoint = lambda ck, name, b, img, stamps, vars_, objs, nb: saved.append((nb, img.array.copy()))
    try:
        image, _ = builder.buildImage(cfg, base, 0, 0, pc.Quiet())
    finally:
        builder.checkpoint = None
        del builder.load_checkpoint, builder.save_checkpoint
    assert builder.last_route == "device"
    assert [nb for nb, _ in saved] == [0, 1, 2, 3, 4, 5]  # after the (empty) FFT batch, then after each photon batch
    sums = [float(a.sum(dtype=np.float64)) for _, a in saved]
    assert sums[0] == 0.0 and all(b > a for a, b in zip(sums, sums[1:]))
    total = sum(c["flux"] for c in cat)
    for k in range(1, 6):
        assert abs(sums[k] / sums[5] - k / 5.0) < 0.01  # every batch carries a fifth of every bright object
    assert 0.9 * total < sums[5] <= total
    np.testing.assert_array_equal(saved[-1][1], image.array)
    # the same run without checkpoints gives the same final image (same seeds, same batches)
    builder2, cfg2, base2 = pc.make_run(su, cat, _source(cat, 5), nbatch=5, nsubbatch=2)
    builder2.setup(cfg2, base2, 0, 0, [], pc.Quiet())
    image2, _ = builder2.buildImage(cfg2, base2, 0, 0, pc.Quiet())
    np.testing.assert_array_equal(image2.array, image.array)


def test_constant_flux_stamps_send_one_number_each(monkeypatch):
    """Stamps whose photons all carry the same flux are written on the device from one value per stamp; a stamp with
    varying fluxes switches the whole field back to the upload.  Same pool either way, bit for bit."""
    import torch

    from imsim_b200 import OpticsContext
    from imsim_b200.photon_array import PhotonArray
    from imsim_b200.photon_pooling import PhotonPool, PooledDevicePath
    from imsim_b200.sensor import SiliconSensor

    ctx = OpticsContext(device=0, stream=torch.cuda.current_stream())
    cfg, dat = helpers.sensor_model("lsst_itl_50_4")
    sensor = SiliconSensor(config=cfg, vertex_data=dat, nrecalc=0, rng=1, absorption_table=helpers.absorption(), context=ctx)
    rng = np.random.default_rng(0)

    def stamps(vary):
        out = []
        for k, n in enumerate((700_000, 3, 250_000, 1)):
            f = np.full(n, 0.25 * (k + 1))
            if vary and k == 2:
                f[n // 2] = 7.0
            out.append(PhotonArray(n, x=rng.uniform(0, 100, n), y=rng.uniform(0, 100, n), flux=f,
                                   wavelength=rng.uniform(500, 700, n)))
        return out

    for vary in (False, True):
        arrays = stamps(vary)
        want = np.concatenate([a.flux for a in arrays])
        got = {}
        for shortcut in (True, False):
            path = PooledDevicePath(PhotonPool(ctx, sensor), sensor)
            path.constant_flux_shortcut = shortcut
            n = path.add(arrays)
            torch.cuda.synchronize()
            assert n == want.size
            got[shortcut] = (path.dp.flux.cpu().numpy(), path.dp.x.cpu().numpy(), path.h2d_bytes)
        np.testing.assert_array_equal(got[True][0], want)
        np.testing.assert_array_equal(got[False][0], want)
        np.testing.assert_array_equal(got[True][1], got[False][1])
        assert (got[True][2] < got[False][2]) == (not vary)  # fewer bytes only when every stamp is constant
