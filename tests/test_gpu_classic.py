"""The classic per-object pipeline (LSST_Image + LSST_Silicon stamps) on the device."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _runner(nrecalc=10000, psf_fwhm=0.7):
    import torch

    from imsim_b200 import OpticsContext
    from imsim_b200.atmosphere import GaussianPSF
    from imsim_b200.detector import lsstcam_like
    from imsim_b200.diffraction import RUBIN_LATITUDE, diffraction_config
    from imsim_b200.sensor import SiliconSensor
    from imsim_b200.synthetic import gpu_tracer, make_detector_setup

    ctx = OpticsContext(device=0, stream=torch.cuda.current_stream())
    det = lsstcam_like("R22_S11")
    su = make_detector_setup(gpu_tracer(ctx), "R22_S11", band="r", rot_tel_pos=np.radians(30.0), detector=det)
    ctx.set_telescope(su.telescope)
    ctx.set_wcs(su.img_wcs, su.icrf_to_field)
    ctx.set_detector(su.detector)
    ctx.set_diffraction(diffraction_config(latitude=RUBIN_LATITUDE, altitude=np.radians(67.0), azimuth=np.radians(213.0)))
    cfg, dat = helpers.sensor_model("lsst_e2v_50_4")
    tr = helpers.tree_ring_table("R22_S11")
    sensor = SiliconSensor(config=cfg, vertex_data=dat, nrecalc=nrecalc, strength=1.0, rng=5, treering_func=tr[1],
                           treering_center=tr[0], absorption_table=helpers.absorption(), context=ctx)
    return ctx, det, sensor, GaussianPSF(psf_fwhm)


def test_classic_pipeline_objects_land_in_their_stamps():
    from imsim_b200.flat import wavelength_cdf
    from imsim_b200.lsst_image import ClassicImageBuilder
    from imsim_b200.sensor import Image
    from imsim_b200.stage1 import ObjectTable

    ctx, det, sensor, psf = _runner()
    tab = ObjectTable()
    #            bright star, star split in chunks, faint star (no optics / silicon), star hanging over the edge
    tab.add_points([1000.0, 2000.0, 3000.0, 3.0], [1000.0, 2500.0, 500.0, 2000.0], [1, 1, 1, 1])
    tab.add_sersic(1500.0, 3200.0, 1, 1.2, 1.0, q=0.5, beta=0.8)
    tab.add_knots(2600.0, 1500.0, 1, 0.8, 15, seed=4)
    rows, _ = tab.build()
    flux = np.array([300000, 250000, 60, 100000, 150000, 50000], dtype=np.float64)
    wave = np.linspace(550, 690, 15)
    cdf, cw = wavelength_cdf(wave, np.ones_like(wave))
    b = ClassicImageBuilder(ctx, sensor, rows, tab.radial_tables(), tab.sersic_n, cdf[None], cw[None], psf=psf,
                            maxN=100000, seed=3)
    image = Image(np.zeros((det.ny, det.nx), np.float32), 0, 0)
    st = b.build(image, flux, phot_flux=flux.astype(np.int64))
    assert st == {**st, "phot": 5, "faint": 1, "skipped": 0, "photons": int(flux.sum())}
    a = image.array.astype(np.float64)
    # everything an object deposits lies inside its stamp; r band: ~all photons convert
    tot = 0.0
    for j in range(rows.size):
        x0, y0, s = b.stamp_bounds(j, flux[j])
        sub = a[max(y0, 0):y0 + s, max(x0, 0):x0 + s]
        tot += sub.sum()
        lo = 0.45 if j == 3 else 0.93  # the star at x = 3 loses about half its light off the chip
        assert lo * flux[j] < sub.sum() <= flux[j] * 1.0001, (j, sub.sum(), flux[j])
    assert abs(a.sum() - tot) < 1e-6 * tot  # nothing outside the stamps
    # the faint star skipped optics and silicon: its photons are binned at the catalogue position
    yy, xx = np.mgrid[480:521, 2980:3021]
    fs = a[480:521, 2980:3021]
    assert fs.sum() == 60 and abs((fs * xx).sum() / 60 - 3000.0) < 1.0 and abs((fs * yy).sum() / 60 - 500.0) < 1.0
    assert np.all(fs == np.round(fs))
    # the galaxy is elongated along beta
    gy, gx = np.mgrid[3200 - 60:3200 + 61, 1500 - 60:1500 + 61]
    g = a[3200 - 60:3200 + 61, 1500 - 60:1500 + 61]
    mx, my = (g * gx).sum() / g.sum(), (g * gy).sum() / g.sum()
    ixx, iyy, ixy = (g * (gx - mx) ** 2).sum(), (g * (gy - my) ** 2).sum(), (g * (gx - mx) * (gy - my)).sum()
    ang = 0.5 * np.arctan2(2 * ixy, ixx - iyy)
    assert abs(ang - 0.8) < 0.15 and g.sum() > 0.9 * flux[4]


def test_classic_brighter_fatter_acts_within_the_stamp():
    """nrecalc = 10000 electrons: a 4e5 e- star is wider with brighter-fatter than without; chunking by maxN
    (resume) does not change the statistics."""
    from imsim_b200.flat import wavelength_cdf
    from imsim_b200.lsst_image import ClassicImageBuilder
    from imsim_b200.sensor import Image
    from imsim_b200.stage1 import ObjectTable

    wave = np.linspace(550, 690, 15)
    cdf, cw = wavelength_cdf(wave, np.ones_like(wave))
    tab = ObjectTable()
    tab.add_points([2000.0], [2000.0], [1])
    rows, _ = tab.build()
    flux = np.array([1200000.0])
    widths = {}
    for name, strength, maxN in (("bf", 1.0, 1000000), ("bf_chunks", 1.0, 100000), ("off", 1e-12, 1000000)):
        ctx, det, sensor, psf = _runner()
        sensor.close()
        from imsim_b200.sensor import SiliconSensor

        cfg, dat = helpers.sensor_model("lsst_e2v_50_4")
        sensor = SiliconSensor(config=cfg, vertex_data=dat, nrecalc=10000, strength=strength, rng=5,
                               absorption_table=helpers.absorption(), context=ctx)
        b = ClassicImageBuilder(ctx, sensor, rows, None, None, cdf[None], cw[None], psf=psf, maxN=maxN, seed=8)
        image = Image(np.zeros((det.ny, det.nx), np.float32), 0, 0)
        st = b.build(image, flux, phot_flux=flux.astype(np.int64))
        g = image.array[1960:2041, 1960:2041].astype(np.float64)
        gy, gx = np.mgrid[1960:2041, 1960:2041]
        mx, my = (g * gx).sum() / g.sum(), (g * gy).sum() / g.sum()
        core = (np.hypot(gx - mx, gy - my) < 6.0)  # the wings (diffraction spikes) are not charge dependent
        widths[name] = ((g * core * ((gx - mx) ** 2 + (gy - my) ** 2)).sum() / (g * core).sum(), g.sum())
    assert widths["bf"][1] > 0.95 * 1.2e6
    # brighter-fatter broadens the core by a few per cent at ~1e5 e- peak (doc/validation/brighter-fatter.rst)
    print({k: v[0] for k, v in widths.items()})
    assert widths["bf"][0] > widths["off"][0] * 1.01
    assert abs(widths["bf_chunks"][0] / widths["bf"][0] - 1.0) < 0.005


def test_batched_build_matches_the_per_object_loop():
    """``build`` (all objects in three launches, chunk loop on the device) against ``build_per_object`` (host-driven
    bind / optics / accumulate per object).  The two seed their photons differently, so the comparison is
    statistical here (per-object electron counts and centroids); that a stamp is bit-identical for identical
    photons is tests/test_gpu_stamps.py."""
    from imsim_b200.flat import wavelength_cdf
    from imsim_b200.lsst_image import ClassicImageBuilder
    from imsim_b200.sensor import Image
    from imsim_b200.stage1 import ObjectTable

    ctx, det, sensor, psf = _runner()
    tab = ObjectTable()
    xs = [800.0, 1500.0, 2300.0, 3100.0, 3900.0, 4090.0]
    ys = [700.0, 1600.0, 2400.0, 3300.0, 500.0, 3990.0]
    tab.add_points(xs, ys, [1] * 6)
    rows, _ = tab.build()
    flux = np.array([200000, 90, 150000, 40000, 30, 120000], dtype=np.float64)
    wave = np.linspace(550, 690, 15)
    cdf, cw = wavelength_cdf(wave, np.ones_like(wave))
    out = {}
    for name in ("build", "build_per_object"):
        b = ClassicImageBuilder(ctx, sensor, rows, None, None, cdf[None], cw[None], psf=psf, maxN=100000, seed=3)
        image = Image(np.zeros((det.ny, det.nx), np.float32), 0, 0)
        st = getattr(b, name)(image, flux, phot_flux=flux.astype(np.int64))
        assert (st["phot"], st["faint"], st["photons"]) == (4, 2, int(flux.sum()))
        a = image.array.astype(np.float64)
        per = []
        for j in range(rows.size):
            x0, y0, s = b.stamp_bounds(j, flux[j])
            sub = a[max(y0, 0):y0 + s, max(x0, 0):x0 + s]
            gy, gx = np.mgrid[max(y0, 0):max(y0, 0) + sub.shape[0], max(x0, 0):max(x0, 0) + sub.shape[1]]
            per.append((sub.sum(), (sub * gx).sum() / sub.sum(), (sub * gy).sum() / sub.sum()))
        out[name] = np.array(per)
    for j in range(rows.size):
        e0, e1 = out["build"][j, 0], out["build_per_object"][j, 0]
        assert abs(e0 - e1) <= 5.0 * np.sqrt(max(flux[j] - min(e0, e1), 1.0)) + 1.0, (j, e0, e1)  # lost-photon noise
        assert np.all(np.abs(out["build"][j, 1:] - out["build_per_object"][j, 1:]) < 0.05 + 12.0 / np.sqrt(flux[j]))
