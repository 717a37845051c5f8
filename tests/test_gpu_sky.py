"""Sky background with exact Poisson noise on the device."""
import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu


def _ctx():
    import torch

    from imsim_b200 import OpticsContext

    return OpticsContext(device=0, stream=torch.cuda.current_stream())


@pytest.mark.parametrize("mean", [0.3, 7.5, 29.9, 30.1, 450.0, 12345.6])
def test_poisson_deviates_are_poisson(mean):
    import torch
    from scipy import stats

    from imsim_b200.sky import add_sky

    ctx = _ctx()
    n = 4_000_000
    img = torch.zeros(n, dtype=torch.float64, device="cuda")
    add_sky(ctx, img, mean, seed=17)
    k = img.cpu().numpy()
    assert np.all(k == np.round(k)) and k.min() >= 0
    se = np.sqrt(mean / n)
    assert abs(k.mean() - mean) < 5 * se
    assert abs(k.var() / mean - 1.0) < 5 * np.sqrt(2.0 / n + 1.0 / (mean * n))
    skew = ((k - mean) ** 3).mean() / mean**1.5
    assert abs(skew - mean**-0.5) < 6 * np.sqrt(6.0 / n) + 0.002
    # chi-square of the histogram against the Poisson pmf over the central 99.9 %
    lo, hi = int(stats.poisson.ppf(5e-4, mean)), int(stats.poisson.ppf(1 - 5e-4, mean))
    edges = np.unique(np.round(np.linspace(lo, hi + 1, min(hi - lo + 2, 60))).astype(int))
    obs, _ = np.histogram(k, bins=edges - 0.5)
    exp = n * np.diff(stats.poisson.cdf(edges - 1, mean))
    chi2 = ((obs - exp) ** 2 / exp).sum()
    assert chi2 < stats.chi2.ppf(1 - 1e-4, len(exp) - 1), (chi2, len(exp))
    # a different seed gives a different, a repeated seed the same realisation
    img2 = torch.zeros(n, dtype=torch.float64, device="cuda")
    add_sky(ctx, img2, mean, seed=17)
    assert torch.equal(img, img2)
    add_sky(ctx, img2.zero_(), mean, seed=18)
    assert not torch.equal(img, img2)


def test_sky_follows_pixel_areas_and_modulation():
    """Tree rings modulate the pixel areas; the sky level follows them (config/imsim-config.yaml:222-228)."""
    import torch

    from imsim_b200.sensor import Image, SiliconSensor
    from imsim_b200.sky import add_sky, pixel_areas_device

    ctx = _ctx()
    cfg, dat = helpers.sensor_model("lsst_e2v_50_4")
    tr = helpers.tree_ring_table("R22_S11")
    sensor = SiliconSensor(config=cfg, vertex_data=dat, nrecalc=0, rng=1, treering_func=tr[1], treering_center=tr[0],
                           absorption_table=helpers.absorption(), context=ctx)
    img = Image(np.zeros((600, 800), np.float32), 3000, 3000)
    host = sensor.calculate_pixel_areas(img, use_flux=False)
    sensor._bind(img)
    areas = pixel_areas_device(sensor, use_flux=False)
    a = areas.cpu().numpy()
    assert np.array_equal(a, np.asarray(getattr(host, "array", host), dtype=np.float64))
    assert 1e-6 < a.std() < 0.05 and abs(a.mean() - 1.0) < 1e-3
    e = torch.zeros((600, 800), dtype=torch.float64, device="cuda")
    mod = torch.ones((600, 800), device="cuda")
    mod[:, 400:] = 0.5  # e.g. a vignetting map
    level = 1.0e7  # high enough for the 2e-5 tree-ring modulation to stand out of the shot noise
    add_sky(ctx, e, level, seed=5, areas=areas, modulation=mod)
    k = e.cpu().numpy().astype(np.float64)
    assert abs(k[:, :400].mean() / (level * a[:, :400].mean()) - 1.0) < 5e-6
    assert abs(k[:, 400:].mean() / (0.5 * level * a[:, 400:].mean()) - 1.0) < 7e-6
    # the tree-ring pattern is in the sky: correlation of (counts / level - 1) with (area - 1)
    r = k[:, :400] / level - 1.0
    c = np.corrcoef(r.ravel(), (a[:, :400] - 1.0).ravel())[0, 1]
    want = a[:, :400].std() / np.hypot(a[:, :400].std(), 1.0 / np.sqrt(level))
    assert want > 0.04 and abs(c - want) < 0.01, (c, want)
