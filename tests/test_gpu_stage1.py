"""Stage 1 on the device (b2_stage1_photons) against the numpy restatement on injected uniforms, the Philox
streams against their numpy mirror, and the statistics the profiles / PSF must reproduce."""
import numpy as np
import pytest

from imsim_b200 import _abi

pytestmark = pytest.mark.gpu


def _setup(psf=None, n_sed=3):
    import torch

    from imsim_b200 import OpticsContext
    from imsim_b200.flat import wavelength_cdf
    from imsim_b200.stage1 import ObjectTable, Stage1

    ctx = OpticsContext(device=0, stream=torch.cuda.current_stream())
    a2p = np.array([[4.9, 0.3], [-0.25, 5.05]])
    tab = ObjectTable(arcsec_to_pix=a2p)
    rng = np.random.default_rng(2)
    tab.add_points(rng.uniform(0, 4000, 5), rng.uniform(0, 4000, 5), [3000, 10, 1, 2500, 700], sed=[0, 1, 2, 0, 1],
                   tanx=rng.uniform(-0.02, 0.02, 5), tany=rng.uniform(-0.02, 0.02, 5))
    tab.add_gaussians([100.0, 900.0], [50.0, 10.0], [4000, 5000], [0.3, 1.1], sed=1)
    tab.add_sersic(500.0, 600.0, 20000, 0.8, 4.0, q=0.6, beta=0.4, g1=0.02, g2=-0.01, mu=1.1, sed=2, tanx=0.01, tany=-0.02)
    tab.add_sersic(1500.0, 1600.0, 15000, 1.3, 1.0, q=0.9, beta=2.0, sed=0)
    tab.add_knots(2500.0, 700.0, 9000, 0.9, 17, q=0.5, beta=1.0, sed=1, seed=99)
    tab.add_streak(3000.0, 3000.0, 6000, 40.0, 0.4, position_angle=0.7, sed=2)
    objects, flux = tab.build()
    wave = np.linspace(540.0, 700.0, 33)
    cdfs, waves = [], []
    for k in range(n_sed):
        c, w = wavelength_cdf(wave, 1.0 + 0.5 * np.sin(wave / (20.0 + 10 * k)))
        cdfs.append(c)
        waves.append(w)
    st = Stage1(ctx, objects, np.array(cdfs), np.array(waves), tab.radial_tables(), psf=None)
    if psf is not None:
        psf.upload(ctx, a2p)
    return ctx, tab, objects, flux.astype(np.int64), np.array(cdfs), np.array(waves), st, a2p


def _shoot(st, counts, rand=None, seed=11, offset=0):
    import torch

    from imsim_b200.photon_pooling import DevicePhotons

    n = int(counts.sum())
    dp = DevicePhotons(n, device="cuda:0", fields=("x", "y", "flux", "wavelength"))
    r = None if rand is None else torch.as_tensor(np.ascontiguousarray(rand), device="cuda:0")
    st.shoot(dp, counts, seed=seed, photon_offset=offset, rand=r)
    torch.cuda.synchronize()
    return dp.x.cpu().numpy(), dp.y.cpu().numpy(), dp.flux.cpu().numpy(), dp.wavelength.cpu().numpy()


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
def test_stage1_matches_oracle_on_injected_uniforms(dtype):
    from imsim_b200.atmosphere import AtmosphericPSF
    from oracle import stage1 as orc

    psf = AtmosphericPSF(1.1, 0.8, "r", rng=7, screen_size=51.2, screen_scale=0.1, dtype=dtype)
    ctx, tab, objects, counts, cdf, cdfw, st, a2p = _setup(psf)
    pod = psf.to_pod(a2p)
    n = int(counts.sum())
    r = np.random.default_rng(0).random((_abi.B2_STAGE1_NRAND, n))
    x, y, f, w = _shoot(st, counts, rand=r)
    ox, oy, of, ow = orc.stage1_photons(objects, counts, r, cdf, cdfw, pod, psf.screens, psf.second_kick[0],
                                        tab.radial_tables())
    assert np.array_equal(f, of)
    np.testing.assert_allclose(w, ow, rtol=1e-13)
    # log / sincos / pow differ from numpy's by an ulp or two: 1e-9 px on offsets of up to ~1e3 px
    np.testing.assert_allclose(x, ox, rtol=0, atol=1e-9)
    np.testing.assert_allclose(y, oy, rtol=0, atol=1e-9)
    assert np.abs(x - np.repeat(objects["x"], counts)).max() > 5.0  # the kicks and profiles did something


def test_philox_streams_match_numpy_mirror():
    from oracle import stage1 as orc

    ctx, tab, objects, counts, cdf, cdfw, st, a2p = _setup()
    n = int(counts.sum())
    x, y, f, w = _shoot(st, counts, seed=0x1234567890ABCDEF, offset=10**12)
    r = orc.stage1_uniforms(0x1234567890ABCDEF, 10**12, n)
    ox, oy, of, ow = orc.stage1_photons(objects, counts, r, cdf, cdfw, None, None, None, tab.radial_tables())
    np.testing.assert_allclose(x, ox, rtol=0, atol=1e-9)
    np.testing.assert_allclose(y, oy, rtol=0, atol=1e-9)
    np.testing.assert_allclose(w, ow, rtol=1e-13)
    # a sub-range of the objects and an offset reproduce the same photons (batches: photon_pooling.py:300-304)
    sel = np.array([7, 8])
    c2 = counts[sel]
    import torch

    from imsim_b200.photon_pooling import DevicePhotons

    dp = DevicePhotons(int(c2.sum()), device="cuda:0", fields=("x", "y", "flux", "wavelength"))
    st.shoot(dp, c2, seed=5, select=sel)
    torch.cuda.synchronize()
    xs = dp.x.cpu().numpy()
    assert abs(np.median(xs[: c2[0]]) - objects["x"][7]) < 1.0 and abs(np.median(xs[c2[0]:]) - objects["x"][8]) < 1.0


def test_profile_statistics():
    """What GalSim's shooters guarantee statistically: half the photons inside the half-light ellipse, the
    sheared second moments, knots on exactly n_knots points, a uniform box, SED-weighted wavelengths."""
    import torch

    from imsim_b200 import OpticsContext
    from imsim_b200.flat import wavelength_cdf
    from imsim_b200.stage1 import ObjectTable, Stage1, shear_matrix

    ctx = OpticsContext(device=0, stream=torch.cuda.current_stream())
    tab = ObjectTable(arcsec_to_pix=np.eye(2) / 0.2)
    q, beta = 0.5, 0.6
    tab.add_sersic(0.0, 0.0, 1, 1.0, 1.0, q=q, beta=beta)
    tab.add_sersic(0.0, 0.0, 1, 0.7, 4.0)
    tab.add_knots(0.0, 0.0, 1, 1.0, 25, seed=3)
    tab.add_streak(0.0, 0.0, 1, 10.0, 2.0)
    tab.add_gaussians([0.0], [0.0], [1], [0.5])
    objects, _ = tab.build()
    wave = np.linspace(500.0, 700.0, 201)
    cdf, cw = wavelength_cdf(wave, (wave - 500.0))  # linear ramp: mean = 500 + 200 * 2/3
    st = Stage1(ctx, objects, cdf[None], cw[None], tab.radial_tables())
    m = 400000
    counts = np.full(5, m)
    x, y, f, w = _shoot(st, counts)
    x, y = x.reshape(5, m) * 0.2, y.reshape(5, m) * 0.2  # arcsec
    # exponential disc, sheared: undo the shear, median radius = hlr, and E[x x^T] = <r^2>/2 S S^T
    S = shear_matrix(q=q, beta=beta)
    u = np.linalg.solve(S, np.vstack([x[0], y[0]]))
    assert abs(np.median(np.hypot(*u)) - 1.0) < 0.01
    r0 = 1.0 / 1.6783469900166605
    cov = np.cov(np.vstack([x[0], y[0]]))
    np.testing.assert_allclose(cov, 3.0 * r0 * r0 * (S @ S.T), rtol=0.03)  # <r^2> = 6 r0^2 for an exponential
    # de Vaucouleurs: median radius only (the second moment is dominated by the truncated tail)
    assert abs(np.median(np.hypot(x[1], y[1])) - 0.7) < 0.01
    # knots: exactly 25 distinct positions, about equally populated, Gaussian with hlr = 1
    pts = np.unique(np.round(np.vstack([x[2], y[2]]).T, 9), axis=0, return_counts=True)
    assert pts[0].shape[0] == 25 and pts[1].min() > 0.8 * m / 25
    assert 0.3 < np.median(np.hypot(pts[0][:, 0], pts[0][:, 1])) < 2.0
    # box 10 x 2 arcsec
    assert abs(x[3].min() + 5) < 1e-3 and abs(x[3].max() - 5) < 1e-3 and abs(y[3].min() + 1) < 1e-3
    assert abs(x[3].var() - 100 / 12) < 0.1 and abs(y[3].var() - 4 / 12) < 0.01
    # Gaussian sigma 0.5 arcsec
    assert abs(x[4].std() - 0.5) < 0.005 and abs(y[4].std() - 0.5) < 0.005
    assert abs(w.mean() - (500.0 + 200.0 * 2 / 3)) < 0.3


def test_atmospheric_psf_has_the_target_fwhm():
    """The first kick (screens below kcrit / r0) plus the second kick reproduce the long-exposure von Karman
    PSF the reference asks for: FWHM = rawSeeing * airmass^0.6 * (lam_eff / 500)^-0.3 (atmPSF.py:128)."""
    import torch

    from imsim_b200 import OpticsContext
    from imsim_b200.atmosphere import AtmosphericPSF
    from imsim_b200.stage1 import ObjectTable, Stage1

    ctx = OpticsContext(device=0, stream=torch.cuda.current_stream())
    psf = AtmosphericPSF(1.0, 0.9, "i", rng=12, screen_size=409.6, screen_scale=0.1, device="cuda:0", gauss_fwhm=0.0)
    tab = ObjectTable()
    tab.add_points([0.0], [0.0], [1])
    objects, _ = tab.build()
    st = Stage1(ctx, objects)
    psf.upload(ctx, np.eye(2))  # positions in arcsec
    m = 4_000_000
    # several independent 30 s realisations (time offsets) to average the speckle pattern
    rr = []
    for k in range(4):
        x, y, f, w = _shoot(st, np.array([m]), seed=100 + k, offset=k * m)
        rr.append(np.hypot(x - x.mean(), y - y.mean()))
    r = np.concatenate(rr)
    # radial profile -> FWHM: density in annuli relative to the core
    edges = np.linspace(0.0, 1.5, 61)
    h, _ = np.histogram(r, bins=edges)
    dens = h / (np.pi * (edges[1:] ** 2 - edges[:-1] ** 2))
    core = dens[:2].mean()
    mid = 0.5 * (edges[1:] + edges[:-1])
    half = mid[np.argmax(dens < 0.5 * core)]
    fwhm = 2.0 * half
    print("target FWHM %.3f measured %.3f (r0_500 %.3f, L0 %.1f)" % (psf.targetFWHM, fwhm, psf.kw["r0_500"], psf.kw["L0"][0]))
    assert abs(fwhm / psf.targetFWHM - 1.0) < 0.2
    # the second kick alone is sub-arcsecond and the screens alone are narrower than the total
    assert np.median(r) < psf.targetFWHM
