"""Host logic of stage 1 (object tables, instance-catalogue rows, atmosphere parameters) -- CPU only."""
import numpy as np
import pytest

from imsim_b200 import _abi
from imsim_b200.atmosphere import (AtmosphericPSF, kolmogorov_fwhm, r0_500_for_seeing, second_kick_table, vk_seeing,
                                   von_karman_screen)
from imsim_b200.stage1 import (ObjectTable, add_instcat_object, lens_matrix, lens_params, read_instcat_objects,
                                sersic_radial_table, shear_matrix)


def test_shear_and_lens_matrices():
    s = shear_matrix(q=0.5, beta=np.radians(30.0))
    assert abs(np.linalg.det(s) - 1.0) < 1e-14  # galsim.Shear preserves area
    w = np.linalg.eigvalsh(s)
    assert abs(w[0] / w[1] - 0.5) < 1e-12  # axis ratio q
    v = np.linalg.eigh(s)[1][:, 1]
    assert abs(np.arctan(v[1] / v[0]) - np.radians(30.0)) < 1e-12  # major axis at beta
    g1, g2, mu = lens_params(0.01, -0.02, 0.05)  # imsim/instcat.py:438-444
    assert abs(g1 - 0.01 / 0.95) < 1e-15 and abs(mu - 1.0 / (0.95**2 - 5e-4)) < 1e-12
    assert abs(np.linalg.det(lens_matrix(g1, g2, mu)) - mu) < 1e-12  # magnification = area ratio


@pytest.mark.parametrize("n", [0.5, 1.0, 2.5, 4.0])
def test_sersic_radial_table(n):
    from scipy.special import gammainc, gammaincinv

    tab = sersic_radial_table(n)
    t = np.linspace(0.0, 14.0, tab.size)
    assert tab[0] == 0.0 and np.all(np.diff(tab) > 0)
    assert abs(np.interp(np.log(2.0), t, tab) - 1.0) < 1e-5  # half the flux inside the half-light radius
    b = gammaincinv(2 * n, 0.5)
    for u in (0.1, 0.9, 0.999):
        r = np.interp(-np.log1p(-u), t, tab)
        assert abs(gammainc(2 * n, b * r ** (1 / n)) - u) < 2e-5
    if n == 1.0:  # exponential disc: hlr = 1.67835 scale lengths
        assert abs(1.0 / np.interp(-np.log1p(-(1 - 2 / np.e)), t, tab) - 1.6783469900166605) < 1e-3


INSTCAT = """rightascension 60.0
object 1 60.01 -30.02 17.5 starSED/kurucz.txt.gz 0 0 0 0 0 0 point none CCM 0.03 3.1
object 2 60.02 -30.01 22.1 galaxySED/a.spec.gz 0.7 0.01 0.02 0.03 0 0 sersic2d 1.2 0.6 35.0 4.0 CCM 0.1 3.1 CCM 0.03 3.1
object 3 60.03 -30.03 23.0 galaxySED/b.spec.gz 0.7 0 0 0 0 0 knots 1.0 0.5 10.0 12 none CCM 0.03 3.1
object 4 60.03 -30.03 23.0 galaxySED/b.spec.gz 0.7 0 0 0 0 0 sersic2d 0.5 1.0 10.0 1 none CCM 0.03 3.1
object 5 60.04 -30.00 55.0 starSED/x.txt.gz 0 0 0 0 0 0 point none none
object 6 60.05 -30.00 20.0 starSED/x.txt.gz 0 0 0 0 0 0 streak 30.0 0.5 45.0 none none
"""


def test_instcat_rows(tmp_path):
    f = tmp_path / "cat.txt"
    f.write_text(INSTCAT)
    objs = read_instcat_objects(str(f))
    assert [o.objid for o in objs] == ["1", "2", "3", "6"]  # a < b and magnorm >= 50 are skipped (instcat.py:273-283)
    assert objs[1].lens == (0.01, -0.02, 0.03)  # flip_g2 (instcat.py:229,262)
    assert objs[1].objinfo == ["sersic2d", "1.2", "0.6", "35.0", "4.0"] and objs[1].dust[0] == "CCM"
    assert objs[0].dust == ["none", "CCM", "0.03", "3.1"]
    tab = ObjectTable(arcsec_to_pix=np.eye(2) / 0.2)
    for k, o in enumerate(objs):
        assert add_instcat_object(tab, o, 100.0 * k, 50.0, 1000.0 + k, sed=k)
    rows, flux = tab.build()
    assert rows.dtype == _abi.OBJECT_DTYPE and rows.size == 4
    assert list(rows["kind"]) == [_abi.PROF_DELTA, _abi.PROF_RADIAL, _abi.PROF_KNOTS, _abi.PROF_BOX]
    assert list(flux) == [1000.0, 1001.0, 1002.0, 1003.0] and list(rows["sed"]) == [0, 1, 2, 3]
    # sersic: hlr = sqrt(a b), q = b/a, beta = 90 - pa (flip_g2), then the lens matrix, in pixels
    g1, g2, mu = lens_params(0.01, -0.02, 0.03)
    want = np.eye(2) / 0.2 @ lens_matrix(g1, g2, mu) @ shear_matrix(q=0.5, beta=np.radians(55.0)) * np.sqrt(0.72)
    np.testing.assert_allclose(rows["m"][1].reshape(2, 2), want, rtol=1e-14)
    assert rows["n_knots"][2] == 12 and tab.sersic_n == [4.0]
    assert rows["p0"][3] == 30.0 and rows["p1"][3] == 0.5
    with pytest.raises(RuntimeError):
        o = objs[0]
        o.objinfo = ["image.fits", "0.2", "0"]
        add_instcat_object(tab, o, 0, 0, 1)


def test_seeing_relations_follow_the_reference():
    # atmPSF.py:211-237: r0_500 from bisection reproduces the target von Karman FWHM
    for L0, target in ((25.0, 0.7), (12.0, 1.1), (80.0, 0.5)):
        r0 = r0_500_for_seeing(622.2, L0, target)
        assert abs(vk_seeing(r0, 622.2, L0) - target) < 1e-9
    assert abs(kolmogorov_fwhm(0.2, 500.0) - 0.9758634299 * 500e-9 / 0.2 * 206264.80624709636) < 1e-12
    psf = AtmosphericPSF(1.2, 0.7, "r", rng=5, screen_size=25.6, screen_scale=0.1)
    assert abs(psf.targetFWHM - 0.7 * 1.2**0.6 * (622.2 / 500) ** -0.3) < 1e-15  # atmPSF.py:128
    kw = psf.kw
    assert kw["altitude"] == [0.2, 2.58, 5.16, 7.73, 12.89, 15.46] and abs(sum(kw["r0_weights"]) - 1) < 1e-12
    assert 10.0 <= kw["L0"][0] <= 100.0 and all(0 <= s <= 20 for s in kw["speed"])
    assert abs(psf.r0_500_effective - kw["r0_500"]) < 1e-12  # weights sum to one
    pod = psf.to_pod()
    assert pod.n_screens == 6 and pod.npix == 256 and abs(pod.altitude[1] - 2580.0) < 1e-9
    assert abs(pod.r_inner / pod.r_outer - 0.61) < 1e-12 and pod.exponent == -0.3 and pod.base_wavelength == 622.2
    # the screens carry only k <= kcrit / r0 (atmPSF.py:173-189)
    f = np.fft.fftfreq(256, 0.1)
    k2 = (2 * np.pi) ** 2 * (f[:, None] ** 2 + f[None, :] ** 2)
    power = np.abs(np.fft.fft2(psf.screens[0].astype(np.float64))) ** 2
    assert power[k2 > psf.kmax**2 * 1.0001].max() < 1e-6 * power.max()


def test_von_karman_screen_structure_function():
    r0, L0 = 0.2, 25.0
    s = von_karman_screen(1024, 0.1, r0, L0, np.random.default_rng(3), dtype=np.float64) * (2 * np.pi / 500.0)
    for d in (2, 4, 8):
        rho = d * 0.1
        D = 0.5 * (np.mean((s[:, d:] - s[:, :-d]) ** 2) + np.mean((s[d:, :] - s[:-d, :]) ** 2))
        # von Karman: Kolmogorov 6.88 (rho/r0)^(5/3) reduced by ~ 1 - 1.485 (rho/L0)^(1/3) (Tokovinin 2002)
        want = 6.8839 * (rho / r0) ** (5 / 3) * (1 - 1.485 * (rho / L0) ** (1 / 3))
        assert abs(D / want - 1) < 0.12, (rho, D, want)


def test_second_kick_table():
    tab, delta, tmax = second_kick_table(622.2, 0.15, 8.36, 0.61, 0.2)
    assert tab[0] == 0.0 and np.all(np.diff(tab) >= 0) and delta == 0.0
    t = np.linspace(0, tmax, tab.size)
    med = np.interp(np.log(2.0), t, tab)
    assert 0.15 < med < 0.6  # sub-arcsecond halo of the modes above kcrit / r0
    # with no turbulence above kcrit the table is the annular Airy pattern: 50 % inside ~ 0.5 lam / D
    airy, _, _ = second_kick_table(622.2, 1e6, 8.36, 0.61, 0.2)
    assert np.interp(np.log(2.0), t, airy) < 0.05


def test_atmosphere_parameter_draws_match_the_reference_source():
    """tests/golden/atmosphere.npz: imsim/atmPSF.py:211-296 executed on recorded deviates
    (tests/golden/make_golden_atmosphere.py).  Same draws in, same atmosphere parameters out."""
    import os

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "atmosphere.npz"))
    for (r0, lam, L0), want in zip(g["vk_grid"], g["vk_seeing"]):
        assert abs(vk_seeing(r0, lam, L0) - want) <= 1e-15 * max(1.0, want)
    for (lam, L0, target), want in zip(g["r0_grid"], g["r0_500"]):
        assert abs(r0_500_for_seeing(lam, L0, target) - want) <= 1e-12
    bands = {365.49: "u", 480.03: "g", 622.2: "r", 754.06: "i", 868.21: "z", 991.66: "y"}

    class Replay:
        """numpy-Generator look-alike replaying the recorded deviates (Gaussian and uniform streams apart,
        like the reference's two deviates on one rng are in the stand-in)."""

        def __init__(self, gauss, unif):
            self.gauss, self.unif, self.ng, self.nu = gauss, unif, 0, 0

        def standard_normal(self):
            self.ng += 1
            return self.gauss[self.ng - 1]

        def random(self):
            self.nu += 1
            return self.unif[self.nu - 1]

    for k in range(int(g["n_cases"])):
        wl, airmass, seeing = g["case%d_in" % k]
        psf = AtmosphericPSF.__new__(AtmosphericPSF)
        psf.rng = Replay(g["case%d_gauss" % k], g["case%d_unif" % k])
        psf.wlen_eff, psf.screen_size, psf.screen_scale = wl, 819.2, 0.1
        psf.targetFWHM = seeing * airmass ** 0.6 * (wl / 500.0) ** (-0.3)
        assert bands[float(wl)] in "ugrizy"
        kw = psf._getAtmKwargs()
        assert [psf.rng.ng, psf.rng.nu] == list(g["case%d_used" % k])  # same number of draws, rejections included
        assert abs(kw["r0_500"] - float(g["case%d_r0_500" % k])) <= 1e-12
        np.testing.assert_array_equal(np.asarray(kw["L0"]), g["case%d_L0" % k])
        np.testing.assert_array_equal(np.asarray(kw["speed"]), g["case%d_speed" % k])
        np.testing.assert_allclose(np.degrees(kw["direction"]), g["case%d_direction_deg" % k], rtol=1e-15, atol=0)
        np.testing.assert_array_equal(np.asarray(kw["altitude"]), g["case%d_altitude" % k])
        np.testing.assert_allclose(np.asarray(kw["r0_weights"]), g["case%d_weights" % k], rtol=1e-15, atol=0)
