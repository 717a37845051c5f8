"""Oracle against the LIVE GalSim / batoid (tests/golden/live_reference.npz, written by
tools/make_reference_fixtures.py on a machine that has the libraries).  The build container has neither, so the
fixture is absent here and these tests skip; they are the place where the "parity unpinned" rows of DESIGN.md
section 6 get closed."""
import os
import pickle

import numpy as np
import pytest

from imsim_b200 import _abi
from imsim_b200.wcs import TanSipWCS

FIX = os.path.join(os.path.dirname(__file__), "golden", "live_reference.npz")
pytestmark = pytest.mark.skipif(not os.path.exists(FIX), reason="no live-reference fixture (needs GalSim + batoid to generate)")


@pytest.fixture(scope="module")
def live():
    return np.load(FIX, allow_pickle=False)


@pytest.mark.parametrize("tag", ["nominal", "rotated", "shifted", "zernike"])
def test_trace_against_batoid(live, tag):
    from oracle import oracle as orc

    tel = pickle.loads(live["trace_%s_telescope_pickle" % tag].tobytes())
    i = {k: live["trace_%s_in_%s" % (tag, k)] for k in ("x", "y", "z", "vx", "vy", "vz", "t", "wavelength")}
    ref = orc.trace_rays(*tel.flatten(), i["x"], i["y"], i["z"], i["vx"], i["vy"], i["vz"], i["t"], i["wavelength"])
    vig = live["trace_%s_out_vignetted" % tag].astype(np.uint8)
    assert np.array_equal(ref[7], vig)
    ok = live["trace_%s_out_failed" % tag] == 0
    for k, name in enumerate(("x", "y", "z", "vx", "vy", "vz", "t")):
        want = live["trace_%s_out_%s" % (tag, name)]
        scale = max(1.0, float(np.abs(want[ok]).max())) if name != "z" else 1.0
        assert np.abs(ref[k][ok] - want[ok]).max() / scale < 1e-10, name


def test_tansip_against_galsim(live):
    from oracle import oracle as orc

    ab = np.zeros((2, 4, 4))
    a = live["wcs_ab"]
    ab[:, : a.shape[1], : a.shape[2]] = a
    w = TanSipWCS(crpix=live["wcs_crpix"], cd=live["wcs_cd"], center=tuple(live["wcs_center"]), ab=ab, order=3).to_pod()
    for k in range(0, live["wcs_x"].size, 97):
        ra, dec = orc.tansip_fwd(w, float(live["wcs_x"][k]), float(live["wcs_y"][k]))
        assert abs(ra - live["wcs_ra"][k]) < 1e-13 and abs(dec - live["wcs_dec"][k]) < 1e-13
        x, y = orc.tansip_inv(w, float(live["wcs_ra"][k]), float(live["wcs_dec"][k]))
        assert abs(x - live["wcs_back_x"][k]) < 4e-7 and abs(y - live["wcs_back_y"][k]) < 4e-7


def test_pixel_areas_against_galsim(live):
    import helpers
    from oracle import oracle as orc

    for name in ("lsst_itl_50_8", "lsst_e2v_50_8"):
        cfg, dat = helpers.sensor_model(name)
        key = "areas_%s_notr" % name
        if key not in live.files:
            continue
        s = orc.Sensor(cfg, dat)
        img = np.ascontiguousarray(live[key + "_image"], dtype=np.float32)
        s.bind_image(img, 1, 1)
        np.testing.assert_allclose(s.pixel_areas(), live[key], rtol=1e-6)


def test_fixture_carries_the_sampler_tables(live):
    # second-kick profile, stamp sizes and the screen spectrum are compared in tests/test_stage1_host.py style checks once
    # the fixture exists; here only their presence and sanity
    assert live["kick2_xvalue"][0] > live["kick2_xvalue"][-1] > 0
    assert live["size_star"].shape[1] == 4 and np.all(np.diff(live["size_star"][:, 1]) >= 0)
    assert _abi.B2_ABI_VERSION == 2
