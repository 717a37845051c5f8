"""Flats on the device (imsim/flat.py:133-281): the statistical checks of the reference's
tests/test_flats.py, on both branches (pixel areas / photon shooting)."""
import numpy as np
import pytest

import helpers
from imsim_b200.flat import build_flat, flat_nrecalc, wavelength_cdf
from imsim_b200.sensor import Image, SiliconSensor
from imsim_b200.treerings import RadialTable

pytestmark = pytest.mark.gpu


def _cov(a):
    f = a - a.mean()
    return (np.mean(f[1:, :] * f[:-1, :]), np.mean(f[:, 1:] * f[:, :-1]), np.mean(f[1:, 1:] * f[:-1, :-1]))


@pytest.mark.parametrize("fused", [True, False])
def test_silicon_flat_area_branch(fused):
    """tests/test_flats.py:69-111: BF correlates neighbours, more along y, variance drops below the mean.
    fused: areas and Poisson realisation on the device; otherwise numpy Poisson on the host."""
    cfg, dat = helpers.sensor_model("lsst_itl_50_8")
    sensor = SiliconSensor(config=cfg, vertex_data=dat, rng=1234, absorption_table=helpers.absorption())
    tot = 80_000
    size = 1024 if fused else 256  # the covariance estimates carry a noise of var / size: go large where it is cheap
    img = Image(np.zeros((size, size), np.float32), 1, 1)
    build_flat(img, tot, sensor, rng=1234, max_counts_per_iter=4000, fused=fused)
    a = img.array.astype(float)
    np.testing.assert_allclose(a.mean(), tot, rtol=1e-2)
    np.testing.assert_allclose(a.var(), tot, rtol=1e-1)
    assert a.var() < tot
    cov10, cov01, cov11 = _cov(a)
    # expectation k * N^2 with the kernel's neighbour area coefficients: ~0.015, ~0.005, ~0.003 (x N)
    if fused:  # at 1024^2 the estimates are good to ~1e-3 tot: the reference's own thresholds (test_flats.py:106-110)
        assert cov10 > 1e-2 * tot and cov01 > 3e-3 * tot and cov11 > 2e-3 * tot
    else:
        assert cov10 > 5e-3 * tot and cov01 > 1e-3 * tot and cov11 > 0
    assert cov10 > cov01 > cov11


@pytest.mark.parametrize("fused", [True, False])
def test_treerings_flat_area_branch(fused):
    """tests/test_flats.py:113-165: cosine tree rings of amplitude A, period P (+ BF) give
    var ~ 1/2 (N A 2pi / P)^2 + N to 3 %, and large covariances in every direction."""
    cfg, dat = helpers.sensor_model("lsst_itl_50_8")
    amp, period = 0.26, 87
    sensor = SiliconSensor(config=cfg, vertex_data=dat, rng=5, treering_func=SiliconSensor.simple_treerings(amp, period),
                           treering_center=(-100.0, -100.0), absorption_table=helpers.absorption())
    tot = 100_000
    img = Image(np.zeros((256, 256), np.float32), 1, 1)
    build_flat(img, tot, sensor, rng=1234, max_counts_per_iter=10_000, fused=fused)
    a = img.array.astype(float)
    pred_var = 0.5 * (tot * amp * 2 * np.pi / period) ** 2 + tot
    np.testing.assert_allclose(a.mean(), tot, rtol=1e-2)
    np.testing.assert_allclose(a.var(), pred_var, rtol=3e-2)
    cov10, cov01, cov11 = _cov(a)
    assert cov10 > 0.5 * tot and cov01 > 0.5 * tot and cov11 > 0.5 * tot


@pytest.mark.parametrize("fused", [False, True])
def test_photon_shot_flat_with_sed(fused):
    """tests/test_flats.py:167-216: with wavelengths, red light is lost out of the back of the sensor;
    the photon branch reproduces the BF sign."""
    cfg, dat = helpers.sensor_model("lsst_itl_50_8")
    n = 64
    results = {}
    for band, (lo, hi) in {"r": (550.0, 690.0), "y": (930.0, 1050.0)}.items():
        sensor = SiliconSensor(config=cfg, vertex_data=dat, rng=11, nrecalc=flat_nrecalc(n, n, 1, 1),
                               absorption_table=helpers.absorption())
        wave = np.linspace(lo, hi, 50)
        cdf = wavelength_cdf(wave, np.ones_like(wave))
        img = Image(np.zeros((n, n), np.float32), 1, 1)
        nphot = build_flat(img, 4000, sensor, rng=3, max_counts_per_iter=1000, nx=1, ny=1, sed_cdf=cdf, fused=fused)
        results[band] = (img.array.astype(float), nphot)
        assert nphot == pytest.approx(4000 * (n + 10) ** 2, rel=0.01)
    r, y = results["r"][0], results["y"][0]
    np.testing.assert_allclose(r.mean(), 4000, rtol=0.02)
    assert y.mean() < 0.9 * r.mean()  # photons lost out the back in y band
    assert r.var() == pytest.approx(r.mean(), rel=0.15)


def test_fused_flat_brighter_fatter_statistics():
    """Photon-shot flat through the fused tile kernel, r band, 30 ke-/px on 128^2: Poisson mean,
    sub-Poisson variance and positive neighbour covariances stronger along y (tests/test_flats.py:69-111)."""
    cfg, dat = helpers.sensor_model("lsst_itl_50_8")
    n = 128
    sensor = SiliconSensor(config=cfg, vertex_data=dat, rng=21, nrecalc=flat_nrecalc(n + 10, n + 10, 1, 1),
                           absorption_table=helpers.absorption())
    wave = np.linspace(550.0, 690.0, 30)
    img = Image(np.zeros((n, n), np.float32), 1, 1)
    tot = 30_000
    build_flat(img, tot, sensor, rng=5, max_counts_per_iter=2000, nx=1, ny=1, sed_cdf=wavelength_cdf(wave, np.ones(30)))
    a = img.array.astype(float)
    np.testing.assert_allclose(a.mean(), tot, rtol=0.01)
    assert 0.8 * tot < a.var() < tot
    cov10, cov01, cov11 = _cov(a)
    assert cov10 > 2e-3 * tot and cov10 > cov01 > -2e-3 * tot
