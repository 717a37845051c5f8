"""Host-side pieces that need neither GPU nor oracle: tree-ring reader, cfg parser,
telescope flattening, WCS fitting, detector geometry, extractors on duck-typed fakes."""
import os
import types

import numpy as np
import pytest

import helpers
from imsim_b200 import _abi
from imsim_b200.detector import lsstcam_like, lsstcam_science_detectors
from imsim_b200.telescope import CoordSys, rot_z, rubin_like
from imsim_b200.treerings import TreeRings
from imsim_b200.wcs import field_wcs, fit_tan_sip, tan_deproject, tan_project


def _write_tree_ring_file(tmp_path):
    lines = []
    for det in ("R22_S11", "R34_S22"):
        lines += helpers.tree_ring_block(det, "tree_ring_parameters_19mar18.txt")
    fn = tmp_path / "tr.txt"
    fn.write_text("".join(lines))
    return str(fn)


def test_tree_rings_reader(tmp_path):
    fn = _write_tree_ring_file(tmp_path)
    g = helpers.golden("tree_rings.npz")
    tr = TreeRings(fn, only_dets=["R22_S11", "R34_S22"], defer_load=False)
    for i, det in enumerate(("R22_S11", "R34_S22")):
        c = tr.get_center(det)
        assert (c.x - 2048.5, c.y - 2048.5) == pytest.approx(tuple(g["known_centers"][i]), abs=0.05)
        assert float(tr.get_func(det)(5280.0)) == pytest.approx(float(g["known_values"][i]), abs=5e-7)
        assert len(tr.get_func(det)) == 2667 and tr.get_func(det).x_max == 8000.0
    # deferred load, unknown detector, missing file (tests/test_tree_rings.py:54-88)
    tr2 = TreeRings(fn, defer_load=True)
    assert tr2.info == {}
    assert float(tr2.get_func("R22_S11")(5280.0)) == pytest.approx(.0030205, abs=5e-7)
    with pytest.warns(UserWarning):
        assert tr2.get_func("R99_S99") is None
    with pytest.raises(OSError):
        TreeRings("invalid.txt")
    # update_info_block drops the cached table
    tr2.update_info_block("R22_S11", A=0.0, B=0.0)
    assert float(tr2.get_func("R22_S11")(5280.0)) == 0.0
    # write: untouched files come back character by character; edits land in the parameter line
    tr3 = TreeRings(fn)
    out = tmp_path / "copy.txt"
    tr3.write(str(out))
    assert out.read_text() == open(fn).read()
    with pytest.raises(FileExistsError):
        tr3.write(str(out))
    tr3.update_info_block("R34_S22", Cx=12.25, A=1.5e-3)
    tr3.write(str(out), overwrite=True)
    tr4 = TreeRings(str(out))
    assert tr4.get_center("R34_S22").x == pytest.approx(2048.5 + 12.2, abs=0.06)
    assert tr4.info_blocks["R34_S22"][1].split()[6] == "1.50e-03"
    assert tr4.info_blocks["R22_S11"] == tr3.info_blocks["R22_S11"] and len(tr4.info_blocks) == 2


def test_detector_geometry_golden():
    det = lsstcam_like("R22_S11")
    x, y = det.focal_to_pixel(np.array([-1000.0, 3000.0]), np.array([1000.0, 2000.0]))
    np.testing.assert_allclose(x, [-97952.5, 302047.5])
    np.testing.assert_allclose(y, [102001.5, 202001.5])
    np.testing.assert_allclose(det.jhat(), [[0, 1], [1, 0]])
    fx, fy = det.pixel_to_focal(x, y)
    np.testing.assert_allclose(fx, [-1000.0, 3000.0], atol=1e-9)
    assert len(lsstcam_science_detectors()) == 189


def test_telescope_flatten_and_rotator():
    tel = rubin_like("r")
    pod, extras = tel.flatten()
    assert pod.n_surfaces == 12 and pod.n_media == 2 and all(e is None for e in extras)
    assert all(pod.surf[i].rot_identity for i in range(12))
    # the stop -> M1 transform is a pure shift along z
    assert pod.surf[0].dr[2] == pytest.approx(-0.4393899)
    rot = rubin_like("r", rot_tel_pos=np.radians(30.0))
    p2, _ = rot.flatten()
    # camera items share the rotated frame: only the first camera surface carries the rotation
    ident = [p2.surf[i].rot_identity for i in range(12)]
    assert ident[:4] == [1, 1, 1, 0]
    np.testing.assert_allclose(np.array(p2.surf[3].drot[:]).reshape(3, 3), rot_z(np.radians(30.0)), atol=1e-15)
    for i in range(4, 12):  # within the camera the relative rotation is the identity up to rounding
        np.testing.assert_allclose(np.array(p2.surf[i].drot[:]).reshape(3, 3), np.eye(3), atol=1e-15)
    # detector shift (telescope_loader.py:399-405)
    sh = tel.with_locally_shifted_item("Detector", [0, 0, -1e-5])
    assert sh.items[-1].coord_sys.origin[2] == pytest.approx(tel.items[-1].coord_sys.origin[2] - 1e-5)
    with pytest.raises(KeyError):
        tel.with_locally_shifted_item("nope", [0, 0, 0])


def test_fit_tan_sip_recovers_a_known_wcs():
    rng = np.random.default_rng(0)
    truth = field_wcs((1.1, -0.4), 0.3, distortion=1e-3, seed=3)
    x = rng.uniform(-0.03, 0.03, 400)
    y = rng.uniform(-0.03, 0.03, 400)
    from imsim_b200.synthetic import _tansip_forward_host

    ra, dec = _tansip_forward_host(truth, x, y)
    fit = fit_tan_sip(x, y, ra, dec, order=3, center=truth.center)
    ra2, dec2 = _tansip_forward_host(fit, x, y)
    np.testing.assert_allclose(ra2, ra, atol=1e-12)
    np.testing.assert_allclose(dec2, dec, atol=1e-12)
    xi, eta = tan_project(ra, dec, *truth.center)
    r3, d3 = tan_deproject(xi, eta, *truth.center)
    np.testing.assert_allclose(r3, ra, atol=1e-14)
    assert fit.to_pod().order == 3


def test_extract_on_duck_typed_batoid():
    from imsim_b200 import extract

    def mk(cls_name, **kw):
        return type(cls_name, (), kw)()

    air = mk("Air", pressure=69.328, temperature=293.15, h2o_pressure=1.067)
    silica = mk("SellmeierMedium", coefs=[0.6961663, 0.4079426, 0.8974794, 0.00467914826, 0.0135120631, 97.9340025])
    cs0 = mk("CoordSys", origin=np.zeros(3), rot=np.eye(3))
    cs1 = mk("CoordSys", origin=np.array([0, 0, 1.0]), rot=np.eye(3))
    clear = mk("ClearAnnulus", original=mk("ObscAnnulus", inner=2.558, outer=4.18, x=0.0, y=0.0))
    m1 = mk("Mirror", name="M1", surface=mk("Asphere", R=19.835, conic=-1.215, coefs=[0.0, -1.38e-9]),
            coordSys=cs0, inMedium=air, outMedium=air, obscuration=clear, skip=False)
    lens = mk("Lens", name="L", skip=False, items=[
        mk("RefractiveInterface", name="L_in", surface=mk("Sphere", R=2.8), coordSys=cs1, inMedium=air,
           outMedium=silica, obscuration=mk("ClearCircle", original=mk("ObscCircle", radius=0.7, x=0.0, y=0.0)),
           skip=False)])
    det = mk("Detector", name="D", surface=mk("Plane"), coordSys=cs1, inMedium=air, outMedium=air, obscuration=None,
             skip=False)
    optic = mk("CompoundOptic", name="T", items=[m1, lens, det], inMedium=air,
               stopSurface=mk("Interface", surface=mk("Plane"), coordSys=cs0))
    tel = extract.telescope_from_batoid(optic)
    assert [it.name for it in tel.items] == ["M1", "L_in", "D"]
    assert tel.items[0].obscurations[0].negate and tel.items[0].obscurations[0].kind == "annulus"
    pod, _ = tel.flatten()
    assert pod.n_surfaces == 3 and pod.n_media == 2 and pod.surf[1].interact == _abi.INT_REFRACT
    with pytest.raises(extract.ExtractError):
        extract.surface_from_batoid(mk("Tilted"))
    # an OPDScreen inserted in front of M1 (tests/test_telescope_loader.py:641-653) becomes a 'pass' interface
    xy = np.zeros((3, 3))
    xy[1, 0] = 1e-6
    screen = mk("OPDScreen", name="Screen", surface=mk("Plane"), screen=mk("Zernike", _xycoef=xy), coordSys=cs0,
                inMedium=air, outMedium=air, obscuration=clear, skip=False)
    optic2 = mk("CompoundOptic", name="T", items=[screen, m1, lens, det], inMedium=air,
                stopSurface=mk("Interface", surface=mk("Plane"), coordSys=cs0))
    tel2 = extract.telescope_from_batoid(optic2)
    pod2, extras2 = tel2.flatten()
    assert tel2.items[0].interact == "pass" and pod2.surf[0].interact == _abi.INT_PASS
    assert pod2.surf[0].surf_kind == _abi.SURF_PLANE and pod2.surf[0].extra_kind == _abi.EXTRA_POLY2D
    assert extras2[0][1].reshape(3, 3)[1, 0] == 1e-6 and pod2.surf[0].medium_in == pod2.surf[0].medium_out
    with pytest.raises(extract.ExtractError):
        extract.telescope_from_batoid(mk("CompoundOptic", name="T", inMedium=air, items=[
            mk("OPDScreen", name="S", surface=mk("Sphere", R=3.0), screen=mk("Plane"), coordSys=cs0, inMedium=air,
               outMedium=air, obscuration=None, skip=False), det],
            stopSurface=mk("Interface", surface=mk("Plane"), coordSys=cs0)))
    # galsim-like WCS
    w = mk("GSFitsWCS", wcs_type="TAN-SIP", pv=None, ab=np.zeros((2, 4, 4)), crpix=np.array([1.0, 2.0]),
           cd=np.eye(2) * 5e-5, center=mk("CelestialCoord", ra=mk("Angle", rad=0.1), dec=mk("Angle", rad=-0.2)))
    t = extract.tansip_from_galsim(w)
    assert t.order == 3 and t.center == (0.1, -0.2)


def test_sensor_cfg_parser(tmp_path):
    from imsim_b200.sensor import calculate_diff_step, read_config_file

    fn = tmp_path / "s.cfg"
    fn.write_text("# comment\nNumVertices = 4  # trailing\nPixelBoundaryLowerLeft = 10.0 10.0\nVbb = -50.0\n"
                  "outputfiledir = data/x\n")
    cfg = read_config_file(str(fn))
    assert cfg["NumVertices"] == 4 and cfg["PixelBoundaryLowerLeft"] == [10.0, 10.0] and cfg["Vbb"] == -50.0
    assert cfg["outputfiledir"] == "data/x"
    full, _ = helpers.sensor_model("lsst_itl_50_4")
    assert calculate_diff_step(full) == pytest.approx(4.4286, abs=1e-3)


def test_flat_control_flow_matches_the_reference_source():
    """tests/golden/flat_control_flow.npz: ``LSST_FlatBuilder.addNoise`` (imsim/flat.py:133-281), its source
    executed against recording stand-ins (tests/golden/make_golden_flat.py).  Same section grid, borders,
    iteration count and level per iteration, photon counts and position ranges, ``resume`` pattern here."""
    import os

    from imsim_b200.flat import build_flat, flat_iterations, flat_sections
    from imsim_b200.sensor import Image

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "flat_control_flow.npz"))
    for k in range(int(g["n_cases"])):
        sed, nrow, ncol, nx, ny, buf, counts, mx = g["c%d_in" % k]
        nrow, ncol, nx, ny, buf = int(nrow), int(ncol), int(nx), int(ny), int(buf)
        niter, per_iter = flat_iterations(counts, mx)
        secs = list(flat_sections(nrow, ncol, nx, ny, buf))
        if sed == 0:
            # pixel-area branch: one calculate_pixel_areas + one noise call per (section, iteration), on the
            # bordered bounds, at counts_per_iter; the charge seen by the k-th call is what k iterations left
            calls = []

            class Sensor:
                def calculate_pixel_areas(self, sec):
                    h, w = sec.array.shape
                    calls.append((sec.xmin, sec.xmin + w - 1, sec.ymin, sec.ymin + h - 1, float(sec.array.sum())))
                    return 1.0

            class Identity:  # the golden run's noise builder adds nothing: a Poisson deviate equal to its mean
                def poisson(self, lam):
                    return lam

            image = Image(np.zeros((nrow, ncol)), 1, 1)
            build_flat(image, counts, Sensor(), rng=Identity(), max_counts_per_iter=mx, nx=nx, ny=ny, buffer_size=buf,
                       fused=False)
            want = g["c%d_areas" % k]
            assert len(calls) == len(want) == len(secs) * niter
            np.testing.assert_allclose(np.array(calls), want, rtol=1e-12)
            np.testing.assert_allclose(g["c%d_noise" % k][:, 4], per_iter, rtol=1e-15)  # level per iteration
            np.testing.assert_allclose(image.array, g["c%d_image" % k], rtol=1e-12)
        else:
            # photon-shot branch: per (section, iteration) one accumulate on the bordered section with
            # counts_per_iter x bordered area photons, uniform over [bxmin - 0.5, bxmax + 0.5), resume = it > 0
            want = g["c%d_accumulate" % k]
            assert len(want) == len(secs) * niter
            row = 0
            for _, _, _, (bx0, bx1, by0, by1) in secs:
                for it in range(niter):
                    w = want[row]
                    assert tuple(w[:4]) == (bx0, bx1, by0, by1)
                    assert w[4] == int(per_iter * (bx1 - bx0 + 1) * (by1 - by0 + 1)) and w[5] == (it > 0)
                    # the range build_flat hands to k_flat_photons (flat.py:250-251 of the reference)
                    assert abs(w[6] - (bx0 - 0.5)) < 1e-9 and abs(w[7] - (bx1 + 0.5)) < 1e-9
                    assert abs(w[8] - (by0 - 0.5)) < 1e-9 and abs(w[9] - (by1 + 0.5)) < 1e-9
                    row += 1
