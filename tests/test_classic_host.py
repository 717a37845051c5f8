"""Host logic of the classic per-object pipeline: stamp sizes (imsim/stamp_utils.py control flow) -- CPU only."""
import numpy as np

from imsim_b200 import _abi
from imsim_b200.stage1 import ObjectTable
from imsim_b200.stamp_utils import (FT_DEFAULT, get_gal_stamp_size, get_stamp_size, get_star_stamp_size,
                                    kolmogorov_radius)


def test_star_stamp_sizes():
    # below the default folding threshold nothing changes (stamp_utils.py:127-133); brighter stars get
    # e-folding-quantised thresholds and monotonically larger, even stamps, capped at Nmax = 4096
    base = get_star_stamp_size(1e3, 800.0)
    assert base == get_star_stamp_size(1e5, 800.0) == get_star_stamp_size(50.0, 800.0)  # 800/1e5 = 8e-3 >= 5e-3
    assert base % 2 == 0 and 30 <= base <= 80
    sizes = [get_star_stamp_size(f, 800.0) for f in (2e5, 1e6, 1e7, 1e9, 1e12)]
    assert all(b >= a for a, b in zip(sizes, sizes[1:])) and sizes[0] > base and sizes[-1] == 4096
    assert get_star_stamp_size(1e6, 0.0) == base  # sky level 0: folding_threshold = 0 -> default (stamp_utils.py:128-132)
    # same e-folding bin -> same size (np.exp(np.floor(np.log(ft))))
    assert get_star_stamp_size(1.0e6, 800.0) == get_star_stamp_size(1.2e6, 800.0)
    # worse seeing / higher airmass -> larger stamps
    assert get_star_stamp_size(1e6, 800.0, rawSeeing=1.2) > get_star_stamp_size(1e6, 800.0, rawSeeing=0.5)
    # Kolmogorov wings: 1 - E ~ theta^(-5/3)
    r1, r2 = kolmogorov_radius(0.7, 1 - 1e-4), kolmogorov_radius(0.7, 1 - 1e-5)
    assert abs(r2 / r1 - 10 ** 0.6) < 0.05 * 10 ** 0.6
    assert abs(kolmogorov_radius(0.7, 0.5) - 0.5 * 0.7 * 1.1) < 0.1  # half-light radius ~ 0.55 FWHM


def test_galaxy_stamp_sizes():
    tab = ObjectTable()
    tab.add_sersic(0, 0, 1, 1.0, 1.0, q=0.5, beta=0.3)
    tab.add_sersic(0, 0, 1, 1.0, 4.0)
    tab.add_sersic(0, 0, 1, 2.0, 4.0)
    tab.add_knots(0, 0, 1, 1.0, 20)
    tab.add_streak(0, 0, 1, 30.0, 0.5)
    tab.add_points([0.0], [0.0], [1])
    rows, _ = tab.build()
    kw = dict(radial_tables=tab.radial_tables(), sersic_n=tab.sersic_n)
    s = [get_stamp_size(r, 1e4, 800.0, **kw) for r in rows]
    assert all(v % 2 == 0 for v in s[:5])
    assert s[1] > s[0] and s[2] > 1.8 * s[1] * 0.9  # de Vaucouleurs wings; size scales with the half-light radius
    assert s[4] >= 30.0 / 0.2  # the streak fits
    assert s[5] == get_star_stamp_size(1e4, 800.0)  # DeltaFunction -> star branch (stamp_utils.py:53-66)
    # tiny fluxes: 32 x 32 (stamp.py:208-210)
    assert all(get_stamp_size(r, 5.0, 800.0, **kw) == 32 for r in rows)
    # bright extended objects grow until the edge surface brightness is below sqrt(noise_var) / 8
    bright = get_gal_stamp_size(rows[1], 1e9, 800.0, **kw)
    assert bright > s[1] and bright <= 4096
    assert FT_DEFAULT == 5e-3 and _abi.PROF_RADIAL == 2


def test_catalogue_wide_stamp_sizes_equal_the_per_object_ones():
    from imsim_b200.stamp_utils import get_stamp_sizes
    from imsim_b200.visit import synthetic_catalog

    for n_obj, total in ((300, 5e7), (120, 5e10)):
        cat = synthetic_catalog(n_obj, 4096, 4004, seed=7, total_photons=total)
        rows, flux = cat.build()
        flux = flux.astype(float)
        flux[::17] = 3.0  # tiny fluxes take the fixed 32-pixel stamp
        kw = dict(radial_tables=cat.radial_tables(), sersic_n=cat.sersic_n)
        want = [get_stamp_size(rows[j], float(flux[j]), 800.0, **kw) for j in range(rows.size)]
        np.testing.assert_array_equal(get_stamp_sizes(rows, flux, 800.0, **kw), want)
