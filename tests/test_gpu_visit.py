"""Device-side photon generation and the per-detector visit runner."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def test_object_photons_statistics():
    import torch

    from imsim_b200 import OpticsContext
    from imsim_b200.flat import wavelength_cdf
    from imsim_b200.photon_pooling import DevicePhotons

    ctx = OpticsContext(device=0, stream=torch.cuda.current_stream())
    ox = np.array([100.0, 2000.5, 3999.0])
    oy = np.array([50.0, 1000.25, 3900.0])
    cnt = np.array([200000, 1, 300000])
    sig = np.array([1.5, 2.0, 0.5])
    n = int(cnt.sum())
    cum = torch.as_tensor(np.concatenate([[0], np.cumsum(cnt)]), device="cuda")
    wave = np.linspace(500, 700, 21)
    cdf = wavelength_cdf(wave, np.linspace(1, 3, 21))
    dp = DevicePhotons(n)
    ctx.object_photons(dp.x, dp.y, dp.flux, dp.wavelength, *(torch.as_tensor(a, device="cuda") for a in (ox, oy, sig)),
                       cum, *(torch.as_tensor(a, device="cuda") for a in cdf), seed=3)
    x, y, wl, fl = (t.cpu().numpy() for t in (dp.x, dp.y, dp.wavelength, dp.flux))
    assert np.all(fl == 1.0)
    a, b = slice(0, 200000), slice(200001, n)
    assert abs(x[a].mean() - 100.0) < 0.02 and abs(y[a].std() - 1.5) < 0.02
    assert abs(x[b].mean() - 3999.0) < 0.01 and abs(x[b].std() - 0.5) < 0.01
    assert abs(x[200000] - 2000.5) < 12 and abs(y[200000] - 1000.25) < 12
    assert wl.min() >= 500 and wl.max() <= 700
    # pdf rises linearly 1 -> 3: mean wavelength = 500 + 200 * (1/2 + 1/12 * 2/2) = 616.67
    assert abs(wl.mean() - (500 + 200 * (0.5 + (3 - 1) / (6.0 * (3 + 1)) * 1.0))) < 0.5


def test_detector_runner_conserves_photons():
    from imsim_b200.flat import wavelength_cdf
    from imsim_b200.visit import DetectorRunner, synthetic_objects, vendor_of

    assert vendor_of("R22_S11") == "e2v" and vendor_of("R01_S00") == "itl"
    models = {"e2v": helpers.sensor_model("lsst_e2v_50_4"), "itl": helpers.sensor_model("lsst_itl_50_4")}
    runner = DetectorRunner(0, models, helpers.absorption(), tree_rings={"R22_S11": helpers.tree_ring_table()})
    objs = synthetic_objects(2000, 4096, 4004, seed=1, total_photons=3e6)
    wave = np.linspace(550, 690, 15)
    rec, image = runner.run("R22_S11", objs, nbatch=10, wavelength_cdf=wavelength_cdf(wave, np.ones_like(wave)))
    assert rec["photons"] == int(objs[2].sum()) and rec["nbatch"] == 10
    assert np.array_equal(runner.last_incident_flux, objs[2].astype(np.float64))  # truth column incident_flux
    # same detector again with the electronics readout on the device: 16 int32 segments carrying the e-image
    rec2, image2 = runner.run("R22_S11", objs, nbatch=10, wavelength_cdf=wavelength_cdf(wave, np.ones_like(wave)),
                              readout=True)
    raw = runner.last_raw
    assert raw.shape == (16, 2048, 576) and raw.dtype == np.int32
    npix = 16 * 2002 * 512
    adu = (raw[:, :2002, 10:522].astype(np.float64) - 1000.0).sum()  # bias 1000 ADU
    # the returned e-image is the one the readout digitised (after bleed trails and the dark current, 0.02 e-/s x
    # 32 s per pixel); gain 1.5 e-/ADU, truncation to int loses 0.5 ADU per pixel on average; read noise (5 ADU
    # rms) and the charge deferred into the overscan by the CTI stay within 1e5 ADU
    assert abs(image2.array.sum(dtype=np.float64) - (rec2["electrons"] + 0.64 * npix)) < 0.002 * 0.64 * npix
    want = image2.array[:4004].sum(dtype=np.float64) / 1.5 - 0.5 * npix
    assert abs(adu - want) < 1.0e5, (adu, want)
    # r-band photons: nearly all convert; a few are vignetted or fall off the chip near the edges
    assert 0.9 * rec["photons"] < rec["electrons"] <= rec["photons"]
    # the brightest object shows up where it was put (optics keep photons within a few pixels)
    k = int(np.argmax(objs[2]))
    cx, cy = int(round(objs[0][k])), int(round(objs[1][k]))
    if 10 < cx < 4086 and 10 < cy < 3994:
        stamp = image.array[cy - 8:cy + 9, cx - 8:cx + 9]
        assert stamp.sum() > 0.7 * objs[2][k]


def test_detector_runner_from_catalogue_rows_behind_the_atmosphere():
    """Stage 1 (catalogue rows + atmospheric PSF) -> fused pooled step, all on the device."""
    from imsim_b200.atmosphere import AtmosphericPSF
    from imsim_b200.flat import wavelength_cdf
    from imsim_b200.stage1 import ObjectTable
    from imsim_b200.visit import DetectorRunner, synthetic_catalog

    models = {"e2v": helpers.sensor_model("lsst_e2v_50_4"), "itl": helpers.sensor_model("lsst_itl_50_4")}
    psf = AtmosphericPSF(1.2, 0.7, "r", rng=3, screen_size=102.4, screen_scale=0.1, device="cuda:0")
    runner = DetectorRunner(0, models, helpers.absorption(), psf=psf)
    wave = np.linspace(550, 690, 15)
    seds = [wavelength_cdf(wave, 1.0 + 0.1 * k * (wave - 550) / 140) for k in range(8)]
    cdf = (np.array([c for c, _ in seds]), np.array([w for _, w in seds]))
    cat = synthetic_catalog(300, 4096, 4004, seed=2, total_photons=4e6)
    rows, flux = cat.build()
    assert set(np.unique(rows["kind"])) == {0, 2, 3} and len(cat) > 300
    rec, image = runner.run("R22_S11", cat, nbatch=5, wavelength_cdf=cdf)
    assert rec["photons"] == int(flux.sum())
    assert 0.85 * rec["photons"] < rec["electrons"] <= rec["photons"]
    # a bright star and a bright extended galaxy on an empty field: the galaxy image is wider
    tab = ObjectTable()
    tab.add_points([1000.0], [1000.0], [400000])
    tab.add_sersic(3000.0, 2500.0, 400000, 1.5, 1.0, q=0.4, beta=0.3)
    rec, image = runner.run("R22_S11", tab, nbatch=2, wavelength_cdf=cdf)

    def width(cx, cy, h=40):
        st = image.array[cy - h:cy + h + 1, cx - h:cx + h + 1].astype(np.float64)
        yy, xx = np.mgrid[-h:h + 1, -h:h + 1]
        mx, my = (st * xx).sum() / st.sum(), (st * yy).sum() / st.sum()
        return np.sqrt((st * ((xx - mx) ** 2 + (yy - my) ** 2)).sum() / st.sum()), st.sum()

    ws, fs = width(1000, 1000)
    wg, fg = width(3000, 2500)
    assert fs > 0.9 * 400000 * 0.95 and fg > 0.8 * 400000 * 0.95
    # star: atmosphere 0.75'' FWHM + optics ~ 2-3 px rms; galaxy: exponential hlr 1.5'' = 7.5 px adds ~ 10 px rms
    assert 1.0 < ws < 6.0 and wg > ws + 4.0


def test_itl_detector_full_chain():
    """An ITL CCD (4072 x 4000, 509 x 2000 segments, no midline bleed stop) through the whole chain: catalogue rows,
    Gaussian PSF, optics, silicon, sky, readout."""
    from imsim_b200.atmosphere import GaussianPSF
    from imsim_b200.flat import wavelength_cdf
    from imsim_b200.visit import DetectorRunner, synthetic_catalog, vendor_of

    assert vendor_of("R01_S00") == "itl"
    models = {"e2v": helpers.sensor_model("lsst_e2v_50_4"), "itl": helpers.sensor_model("lsst_itl_50_4")}
    runner = DetectorRunner(0, models, helpers.absorption(), psf=GaussianPSF(0.7))
    wave = np.linspace(550, 690, 15)
    seds = [wavelength_cdf(wave, 1.0 + 0.1 * k * (wave - 550) / 140) for k in range(8)]
    cdf = (np.array([c for c, _ in seds]), np.array([w for _, w in seds]))
    cat = synthetic_catalog(400, 4072, 4000, seed=5, total_photons=2e6)
    rec, image = runner.run("R01_S00", cat, nbatch=4, wavelength_cdf=cdf, readout=True, sky_level=300.0)
    assert image.array.shape == (4000, 4072) and runner.last_raw.shape == (16, 2048, 576)
    _, flux = cat.build()
    assert rec["photons"] == int(flux.sum()), (rec["photons"], flux.sum())
    # R01 sits 1.6 deg off axis, where the Rubin-like prescription vignettes about half the pupil
    assert 0.3 * rec["photons"] < rec["electrons"] <= rec["photons"], (rec["photons"], rec["electrons"])
    sky = image.array.astype(np.float64).sum() - rec["electrons"]
    npix = 4000 * 4072
    assert abs(sky / npix - (300.0 + 0.64)) < 0.05  # sky through the pixel areas + dark current
    raw = runner.last_raw
    # imaging area of an ITL segment: 3 prescan columns, 509 data columns, 2000 rows
    data = raw[:, :2000, 3:512].astype(np.float64)
    assert abs(data.mean() - (1000.0 + (300.64 + rec["electrons"] / npix) / 1.5 - 0.5)) < 0.5
    assert abs(raw[:, 2010:, 520:].astype(np.float64).mean() - 999.5) < 0.1  # overscan corner: bias only


def test_pipelined_visit_equals_one_detector_at_a_time():
    """``run_many`` (host preparation of detector k+1 overlapping the kernels of detector k, results landing in
    alternating pinned slots) gives exactly the images, raw segments and records of ``run`` called in turn --
    across both vendors, with catalogue rows and with plain point-source tables."""
    from imsim_b200.atmosphere import GaussianPSF
    from imsim_b200.flat import wavelength_cdf
    from imsim_b200.visit import DetectorRunner, synthetic_catalog, synthetic_objects

    models = {"e2v": helpers.sensor_model("lsst_e2v_50_4"), "itl": helpers.sensor_model("lsst_itl_50_4")}
    wave = np.linspace(550, 690, 15)
    seds = [wavelength_cdf(wave, 1.0 + 0.1 * k * (wave - 550) / 140) for k in range(8)]
    cdf = (np.array([c for c, _ in seds]), np.array([w for _, w in seds]))
    dets = ["R22_S11", "R01_S00", "R22_S12", "R22_S11"]
    tr = {d: helpers.tree_ring_table() for d in dets if d.startswith("R22")}

    def jobs(catalogue):
        for i, d in enumerate(dets):
            nx, ny = (4072, 4000) if d.startswith("R01") else (4096, 4004)
            if catalogue:
                objects = (lambda i=i, nx=nx, ny=ny: synthetic_catalog(300, nx, ny, seed=i, total_photons=1.5e6))
                yield dict(det_name=d, objects=objects, nbatch=3, wavelength_cdf=cdf, det_index=i, readout=True,
                           sky_level=100.0)
            else:
                objects = synthetic_objects(500, nx, ny, seed=i, total_photons=1e6)
                yield dict(det_name=d, objects=objects, nbatch=4, wavelength_cdf=(cdf[0][0], cdf[1][0]), det_index=i)

    for catalogue in (False, True):
        runner = DetectorRunner(0, models, helpers.absorption(), tree_rings=tr, psf=GaussianPSF(0.7))
        serial = []
        for job in jobs(catalogue):
            rec, image = runner.run(**job)
            serial.append((rec, image.array.copy(), None if runner.last_raw is None else runner.last_raw.copy()))
        runner2 = DetectorRunner(0, models, helpers.absorption(), tree_rings=tr, psf=GaussianPSF(0.7))
        n = 0
        for (rec, image), (srec, simg, sraw) in zip(runner2.run_many(jobs(catalogue)), serial):
            assert rec["det_name"] == srec["det_name"] and rec["photons"] == srec["photons"] > 0
            assert rec["electrons"] == srec["electrons"]
            assert np.array_equal(image.array, simg)
            if catalogue:
                assert np.array_equal(runner2.last_raw, sraw)
            n += 1
        assert n == len(dets)
    assert list(runner2.run_many([])) == []


def test_visit_lanes_give_the_images_of_one_runner():
    """Two detectors at a time on separate streams and host threads (visit.VisitLanes): same records, e-images and
    raw segments as one runner working through the list."""
    import zlib

    from imsim_b200.atmosphere import GaussianPSF
    from imsim_b200.flat import wavelength_cdf
    from imsim_b200.visit import DetectorRunner, VisitLanes, synthetic_catalog

    models = {"e2v": helpers.sensor_model("lsst_e2v_50_4"), "itl": helpers.sensor_model("lsst_itl_50_4")}
    wave = np.linspace(550, 690, 15)
    seds = [wavelength_cdf(wave, 1.0 + 0.1 * k * (wave - 550) / 140) for k in range(8)]
    cdf = (np.array([c for c, _ in seds]), np.array([w for _, w in seds]))
    dets = ["R22_S11", "R01_S00", "R22_S12", "R22_S11", "R01_S00"]
    tr = {d: helpers.tree_ring_table() for d in dets if d.startswith("R22")}
    jobs = []
    for i, d in enumerate(dets):
        nx, ny = (4072, 4000) if d.startswith("R01") else (4096, 4004)
        jobs.append(dict(det_name=d, objects=(lambda i=i, nx=nx, ny=ny: synthetic_catalog(300, nx, ny, seed=i,
                                                                                         total_photons=1.5e6)),
                         nbatch=3, wavelength_cdf=cdf, det_index=i, readout=True, sky_level=100.0))
    crc = lambda a: zlib.crc32(np.ascontiguousarray(a).tobytes())  # noqa: E731
    runner = DetectorRunner(0, models, helpers.absorption(), tree_rings=tr, psf=GaussianPSF(0.7))
    want = {}
    for rec, image in runner.run_many(jobs):
        want[(rec["det_name"], rec["photons"])] = (rec["electrons"], crc(image.array), crc(runner.last_raw))
    got = {}

    def collect(rec, image, raw):
        got[(rec["det_name"], rec["photons"])] = (rec["electrons"], crc(image.array), crc(raw))

    lanes = VisitLanes(0, 2, models, helpers.absorption(), tree_rings=tr, psf=GaussianPSF(0.7))
    recs = lanes.run(jobs, cost=lambda j: 2.0 if j["det_name"].startswith("R22") else 1.0, on_result=collect)
    assert [r["det_name"] for r in recs] == dets
    assert got == want
