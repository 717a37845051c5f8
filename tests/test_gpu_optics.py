"""Parity of the CUDA optics path (through the C ABI) against the CPU oracle on
identical inputs with injected random draws.

Tolerances (BASELINE.json north_star): ray-traced positions, directions and
times agree to 1e-10 relative in FP64 mode; vignetting flags identical.
"""
import numpy as np
import pytest

import helpers
from imsim_b200 import _abi
from imsim_b200.telescope import paraboloid_test_telescope, rubin_like

pytestmark = pytest.mark.gpu

RTOL = 1e-10


def _ctx(setup=None, dif=None):
    from imsim_b200 import OpticsContext

    ctx = OpticsContext(device=0)
    if setup is not None:
        ctx.set_telescope(setup.telescope)
        ctx.set_wcs(setup.img_wcs, setup.icrf_to_field)
        ctx.set_detector(setup.detector)
    ctx.set_diffraction(dif)
    return ctx


def _close(a, b, scale=None, rtol=RTOL):
    a, b = np.asarray(a), np.asarray(b)
    s = np.maximum(np.abs(b), 1e-300) if scale is None else scale
    err = np.max(np.abs(a - b) / s)
    assert err < rtol, "max relative error %.3e" % err
    return err


def test_xy_to_v_and_inverse_match_oracle():
    from oracle import oracle as orc

    su = helpers.oracle_setup()
    ctx = _ctx(su)
    rng = np.random.default_rng(3)
    x = rng.uniform(0, 4096, 20000)
    y = rng.uniform(0, 4004, 20000)
    vx, vy, vz = ctx.xy_to_v(x, y)
    ox, oy, oz = orc.xy_to_v(su.img_wcs.to_pod(), su.icrf_to_field.to_pod(), x, y)
    # direction cosines: absolute 1e-12 on O(0.03) components is far below 1e-10 relative of the pixel position
    np.testing.assert_allclose(vx, ox, rtol=0, atol=2e-14)
    np.testing.assert_allclose(vy, oy, rtol=0, atol=2e-14)
    np.testing.assert_allclose(vz, oz, rtol=1e-14)
    # XyToV.inverse o XyToV = id (tests/test_photon_ops.py:429-446)
    x2, y2 = ctx.v_to_xy(vx, vy, vz)
    np.testing.assert_allclose(x2, x, rtol=0, atol=4e-7)  # 1e-10 relative of 4000 px
    np.testing.assert_allclose(y2, y, rtol=0, atol=4e-7)
    ox2, oy2 = orc.v_to_xy(su.img_wcs.to_pod(), su.icrf_to_field.to_pod(), vx, vy, vz)
    np.testing.assert_allclose(x2, ox2, rtol=0, atol=4e-7)
    np.testing.assert_allclose(y2, oy2, rtol=0, atol=4e-7)


@pytest.mark.parametrize("rot", [0.0, np.radians(60.0)])
def test_trace_rays_matches_oracle(rot):
    from oracle import oracle as orc

    tel = rubin_like("r", rot_tel_pos=rot)
    ctx = _ctx()
    ctx.set_telescope(tel)
    rng = np.random.default_rng(5)
    n = 50000
    r = np.sqrt(rng.uniform(2.3**2, 4.3**2, n))  # includes vignetted rays
    ph = rng.uniform(0, 2 * np.pi, n)
    x, y, z = r * np.cos(ph), r * np.sin(ph), np.zeros(n)
    thx, thy = rng.uniform(-0.031, 0.031, n), rng.uniform(-0.031, 0.031, n)
    wl = rng.uniform(320e-9, 1050e-9, n)
    g = 1 / np.sqrt(1 + thx**2 + thy**2)
    nair = tel.in_medium.n(wl)
    vx, vy, vz = thx * g / nair, thy * g / nair, -g / nair
    t = np.zeros(n)
    bt, ex = tel.flatten()
    ref = orc.trace_rays(bt, ex, x, y, z, vx, vy, vz, t, wl)
    arrs = [np.ascontiguousarray(a.copy()) for a in (x, y, z, vx, vy, vz, t, wl)]
    vig = np.zeros(n, np.uint8)
    fail = np.zeros(n, np.uint8)
    ctx.trace_rays(*arrs, vig, fail)
    assert np.array_equal(vig, ref[7]), "vignetting flags differ for %d rays" % np.sum(vig != ref[7])
    assert np.array_equal(fail, ref[8])
    ok = fail == 0
    # positions on the focal plane are O(0.3 m); compare relative to the focal-plane scale
    _close(arrs[0][ok], ref[0][ok], scale=0.3)
    _close(arrs[1][ok], ref[1][ok], scale=0.3)
    assert np.abs(arrs[2][ok]).max() < 1e-15  # on the detector plane (photon_ops.py:494)
    for k in (3, 4, 5):
        _close(arrs[k][ok], ref[k][ok], scale=1.0)
    _close(arrs[6][ok], ref[6][ok])  # time of flight


@pytest.mark.parametrize("rot", [0.0, np.radians(-25.0)])
def test_surface_program_equals_interpreter(rot, monkeypatch):
    """The Rubin layout compiled as straight-line code (program 1) and the generic interpreter over the
    surface list (program 0) are the same arithmetic: identical flags, results equal to rounding."""
    tel = rubin_like("i", rot_tel_pos=rot, detector_z_offset=1.5e-5)
    rng = np.random.default_rng(17)
    n = 200000
    r = np.sqrt(rng.uniform(2.3**2, 4.3**2, n))
    ph = rng.uniform(0, 2 * np.pi, n)
    thx, thy = rng.uniform(-0.031, 0.031, n), rng.uniform(-0.031, 0.031, n)
    wl = rng.uniform(320e-9, 1050e-9, n)
    g = 1 / np.sqrt(1 + thx**2 + thy**2)
    nair = tel.in_medium.n(wl)
    base = [r * np.cos(ph), r * np.sin(ph), np.zeros(n), thx * g / nair, thy * g / nair, -g / nair, np.zeros(n), wl]
    out = {}
    for prog in (1, 0):
        if prog == 0:
            monkeypatch.setenv("B2_PROGRAM", "0")
        ctx = _ctx()
        ctx.set_telescope(tel)
        assert ctx.program == prog
        arrs = [np.ascontiguousarray(a.copy()) for a in base]
        vig, fail = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        ctx.trace_rays(*arrs, vig, fail)
        out[prog] = arrs, vig, fail
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])
    assert 0.05 < out[1][1].mean() < 0.6
    ok = out[1][2] == 0
    for k, scale in ((0, 0.3), (1, 0.3), (3, 1.0), (4, 1.0), (5, 1.0), (6, 30.0)):
        _close(out[1][0][k][ok], out[0][0][k][ok], scale=scale, rtol=1e-14)
    # a perturbed mirror leaves the layout: interpreter
    monkeypatch.delenv("B2_PROGRAM", raising=False)
    ctx = _ctx()
    poly = np.zeros((4, 4))
    poly[2, 0] = poly[0, 2] = 1e-8
    ctx.set_telescope(tel.with_surface_perturbation("M2", poly=poly, poly_scale=1 / 1.71))
    assert ctx.program == 0


def test_paraboloid_focus_exact():
    tel = paraboloid_test_telescope(10.0)
    ctx = _ctx()
    ctx.set_telescope(tel)
    rng = np.random.default_rng(0)
    n = 4096
    x, y = rng.uniform(-3, 3, n), rng.uniform(-3, 3, n)
    z, vx, vy, vz, t = np.zeros(n), np.zeros(n), np.zeros(n), -np.ones(n), np.zeros(n)
    vig, fail = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    ctx.trace_rays(x, y, z, vx, vy, vz, t, np.full(n, 500e-9), vig, fail)
    assert np.abs(x).max() < 1e-14 and np.abs(y).max() < 1e-14
    assert np.ptp(t) < 1e-13 and abs(t[0] - 30.0) < 1e-13  # equal optical path (Fermat)


@pytest.mark.parametrize("mode", ["optics", "diffraction_optics", "diffraction_optics_norot"])
def test_rubin_optics_matches_oracle(mode):
    from oracle import oracle as orc

    su = helpers.oracle_setup()
    dif = None if mode == "optics" else helpers.default_diffraction(field_rotation=(mode == "diffraction_optics"))
    ctx = _ctx(su, dif)
    p = helpers.test_photon_arrays(n=100000, t=12.0, center=(809.5, 3432.5))
    rng = np.random.default_rng(11)
    p["time"] = rng.uniform(0, 30, p["x"].size)
    p["wavelength"] = rng.uniform(540, 700, p["x"].size)
    gauss = rng.standard_normal(p["x"].size)
    opt = _abi.B2OpticsOptions()
    ref = orc.rubin_optics(*su.telescope.flatten(), su.img_wcs.to_pod(), su.icrf_to_field.to_pod(),
                           su.detector.to_pod(), dif, opt, p["x"], p["y"], p["flux"], p["wavelength"], p["pupil_u"],
                           p["pupil_v"], p["time"], gauss, want_time=True)
    x, y, flux = p["x"].copy(), p["y"].copy(), p["flux"].copy()
    dxdz, dydz, tout = np.empty_like(x), np.empty_like(x), np.empty_like(x)
    stats = ctx.rubin_optics(x, y, dxdz, dydz, flux, p["wavelength"], p["pupil_u"], p["pupil_v"], p["time"],
                             gauss=gauss, time_out=tout, options=opt)
    assert np.array_equal(flux, ref["flux"])  # vignetting identical
    assert stats.n_vignetted == ref["stats"].n_vignetted and stats.n_failed == 0 and stats.n_offdetector_z == 0
    ok = flux > 0
    assert ok.mean() > 0.9
    # The spider kick is ill-conditioned for photons grazing a vane edge: phi* ~ 1/delta with
    # delta = ||n.p - d| - w| a cancelling difference, so a 1e-15 m rounding difference in delta
    # moves a photon at delta = 1e-6 m by ~1e-5 px.  The 1e-10 bar applies to photons whose kick is
    # below 100 px (delta > ~4e-4 m); the grazing ones are held to 1e-9 of their kick and counted.
    if dif is not None:
        nd = orc.rubin_optics(*su.telescope.flatten(), su.img_wcs.to_pod(), su.icrf_to_field.to_pod(),
                              su.detector.to_pod(), None, opt, p["x"], p["y"], p["flux"], p["wavelength"],
                              p["pupil_u"], p["pupil_v"], p["time"])
        kick = np.hypot(ref["x"] - nd["x"], ref["y"] - nd["y"])
    else:
        kick = np.zeros_like(x)
    small = ok & (kick < 100.0)
    graze = ok & ~small
    assert small.sum() > 0.99 * ok.sum()
    _close(x[small], ref["x"][small], scale=4000.0)
    _close(y[small], ref["y"][small], scale=4000.0)
    _close(dxdz[small], ref["dxdz"][small], scale=1.0)
    _close(dydz[small], ref["dydz"][small], scale=1.0)
    _close(tout[small], ref["time_out"][small])
    if graze.any():
        _close(x[graze], ref["x"][graze], scale=np.maximum(kick[graze], 4000.0), rtol=1e-9)
        _close(y[graze], ref["y"][graze], scale=np.maximum(kick[graze], 4000.0), rtol=1e-9)
    print("grazing photons (kick > 100 px): %d of %d" % (graze.sum(), ok.sum()))
    # photons land where the WCS says (within the PSF + diffraction spikes): tests/test_photon_ops.py:173-196
    if mode == "optics":
        assert np.abs(x[ok] - p["x"][ok]).max() < 20 and np.abs(y[ok] - p["y"][ok]).max() < 20


def test_fused_focus_depth_and_refraction():
    from oracle import oracle as orc

    su = helpers.oracle_setup()
    ctx = _ctx(su, None)
    p = helpers.test_photon_arrays(n=20000, center=(2000.0, 2000.0))
    opt = _abi.B2OpticsOptions()
    opt.do_focus_depth, opt.focus_depth = 1, -0.6
    opt.do_refraction, opt.index_ratio = 1, 3.9
    opt.shift_in, opt.shift_out = 1, 1
    opt.stamp_center[0], opt.stamp_center[1] = 12.0, -7.0
    ref = orc.rubin_optics(*su.telescope.flatten(), su.img_wcs.to_pod(), su.icrf_to_field.to_pod(),
                           su.detector.to_pod(), None, opt, p["x"], p["y"], p["flux"], p["wavelength"], p["pupil_u"],
                           p["pupil_v"], p["time"])
    x, y, flux = p["x"].copy(), p["y"].copy(), p["flux"].copy()
    dxdz, dydz = np.empty_like(x), np.empty_like(x)
    ctx.rubin_optics(x, y, dxdz, dydz, flux, p["wavelength"], p["pupil_u"], p["pupil_v"], p["time"], options=opt)
    ok = flux > 0
    _close(x[ok], ref["x"][ok], scale=4000.0)
    _close(dxdz[ok], ref["dxdz"][ok], scale=1.0)
    # Refraction shrinks the slopes by ~1/3.9
    assert np.abs(dxdz[ok]).max() < 0.15


def test_rubin_diffraction_matches_oracle_and_modular_equals_combined():
    """RubinDiffraction.applyTo parity, and the reference's own invariant
    (tests/test_photon_ops.py:281-318): combined op == diffraction then optics, 6 decimals."""
    from oracle import oracle as orc

    su = helpers.oracle_setup()
    dif = helpers.default_diffraction()
    ctx = _ctx(su, dif)
    p = helpers.test_photon_arrays(n=30000, t=3.0, center=(1500.0, 2500.0))
    gauss = np.random.default_rng(42).standard_normal(p["x"].size)
    opt = _abi.B2OpticsOptions()
    rx, ry = orc.rubin_diffraction(su.telescope.flatten()[0], su.img_wcs.to_pod(), su.icrf_to_field.to_pod(), dif,
                                   opt, p["x"], p["y"], p["wavelength"], p["pupil_u"], p["pupil_v"], p["time"], gauss)
    x, y = p["x"].copy(), p["y"].copy()
    ctx.rubin_diffraction(x, y, p["wavelength"], p["pupil_u"], p["pupil_v"], p["time"], gauss=gauss, options=opt)
    # grazing photons: tolerance conditioned on the kick size (see test_rubin_optics_matches_oracle)
    kick = np.hypot(rx - p["x"], ry - p["y"])
    small = kick < 100.0
    assert small.mean() > 0.99
    _close(x[small], rx[small], scale=4000.0)
    _close(y[small], ry[small], scale=4000.0)
    if (~small).any():
        _close(x[~small], rx[~small], scale=np.maximum(kick[~small], 4000.0), rtol=1e-9)
        _close(y[~small], ry[~small], scale=np.maximum(kick[~small], 4000.0), rtol=1e-9)
    # modular: diffraction (above) then plain optics
    ctx2 = _ctx(su, None)
    flux = p["flux"].copy()
    dxdz, dydz = np.empty_like(x), np.empty_like(x)
    ctx2.rubin_optics(x, y, dxdz, dydz, flux, p["wavelength"], p["pupil_u"], p["pupil_v"], p["time"], options=opt)
    # combined
    xc, yc, fc = p["x"].copy(), p["y"].copy(), p["flux"].copy()
    ac, bc = np.empty_like(x), np.empty_like(x)
    ctx.rubin_optics(xc, yc, ac, bc, fc, p["wavelength"], p["pupil_u"], p["pupil_v"], p["time"], gauss=gauss,
                     options=opt)
    ok = (fc > 0) & (flux > 0)
    np.testing.assert_array_almost_equal(xc[ok], x[ok], decimal=6)
    np.testing.assert_array_almost_equal(yc[ok], y[ok], decimal=6)
    np.testing.assert_array_almost_equal(ac[ok], dxdz[ok], decimal=6)


def test_zernike_and_bicubic_perturbations():
    from oracle import oracle as orc

    rng = np.random.default_rng(9)
    poly = np.zeros((5, 5))
    poly[2, 0], poly[0, 2], poly[1, 1], poly[3, 1], poly[0, 4] = 3e-7, -2e-7, 1e-7, 4e-8, -3e-8
    tel = rubin_like("r").with_surface_perturbation("M1", poly=poly, poly_scale=1 / 4.18)
    xs = np.linspace(-1.8, 1.8, 41)
    X, Y = np.meshgrid(xs, xs)
    bic = dict(xs=xs, ys=xs, zs=1e-7 * np.cos(2 * X) * np.sin(Y), dzdxs=-2e-7 * np.sin(2 * X) * np.sin(Y),
               dzdys=1e-7 * np.cos(2 * X) * np.cos(Y), d2zdxdys=-2e-7 * np.sin(2 * X) * np.cos(Y))
    tel = tel.with_surface_perturbation("M2", bicubic=bic)
    ctx = _ctx()
    ctx.set_telescope(tel)
    n = 20000
    r = np.sqrt(rng.uniform(2.6**2, 4.1**2, n))
    ph = rng.uniform(0, 2 * np.pi, n)
    x, y, z = r * np.cos(ph), r * np.sin(ph), np.zeros(n)
    thx, thy = rng.uniform(-0.02, 0.02, n), rng.uniform(-0.02, 0.02, n)
    wl = np.full(n, 622e-9)
    g = 1 / np.sqrt(1 + thx**2 + thy**2)
    nair = tel.in_medium.n(wl)
    vx, vy, vz, t = thx * g / nair, thy * g / nair, -g / nair, np.zeros(n)
    bt, ex = tel.flatten()
    ref = orc.trace_rays(bt, ex, x, y, z, vx, vy, vz, t, wl)
    base = orc.trace_rays(*rubin_like("r").flatten(), x, y, z, vx, vy, vz, t, wl)
    arrs = [np.ascontiguousarray(a.copy()) for a in (x, y, z, vx, vy, vz, t, wl)]
    vig, fail = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    ctx.trace_rays(*arrs, vig, fail)
    ok = (fail == 0) & (ref[8] == 0)
    assert ok.mean() > 0.99
    _close(arrs[0][ok], ref[0][ok], scale=0.3)
    _close(arrs[1][ok], ref[1][ok], scale=0.3)
    # and the perturbation does something (microns on the focal plane)
    assert np.abs(ref[0][ok] - base[0][ok]).max() > 1e-7


def test_fused_photon_dcr():
    """PhotonDCR prologue: GPU == oracle, and both equal an independent numpy evaluation of
    galsim.dcr's formulas on the pre-trace shift (checked through the plain-optics difference)."""
    from imsim_b200.photon_ops import get_refraction, set_dcr_options
    from oracle import oracle as orc

    su = helpers.oracle_setup()
    ctx = _ctx(su, None)
    p = helpers.test_photon_arrays(n=20000, center=(1200.0, 900.0))
    p["wavelength"] = np.random.default_rng(8).uniform(380, 1000, p["x"].size)
    opt = _abi.B2OpticsOptions()
    zen, q = np.radians(40.0), np.radians(25.0)
    J = np.array([[0.2 * np.cos(0.3), -0.2 * np.sin(0.3)], [0.2 * np.sin(0.3), 0.2 * np.cos(0.3)]])
    set_dcr_options(opt, 622.0, zen, q, J, center=(1200.0, 900.0), alpha=-0.05)
    ref = orc.rubin_optics(*su.telescope.flatten(), su.img_wcs.to_pod(), su.icrf_to_field.to_pod(),
                           su.detector.to_pod(), None, opt, p["x"], p["y"], p["flux"], p["wavelength"], p["pupil_u"],
                           p["pupil_v"], p["time"])
    x, y, flux = p["x"].copy(), p["y"].copy(), p["flux"].copy()
    dxdz, dydz = np.empty_like(x), np.empty_like(x)
    ctx.rubin_optics(x, y, dxdz, dydz, flux, p["wavelength"], p["pupil_u"], p["pupil_v"], p["time"], options=opt)
    ok = flux > 0
    _close(x[ok], ref["x"][ok], scale=4000.0)
    _close(y[ok], ref["y"][ok], scale=4000.0)
    # independent numpy evaluation of the DCR shift: feed pre-shifted photons to the op without DCR
    s = (get_refraction(p["wavelength"], zen) - get_refraction(622.0, zen)) * 206264.80624709636
    du, dv = -s * np.sin(q), s * np.cos(q)
    Ji = np.linalg.inv(J)
    sc = (p["wavelength"] / 622.0) ** -0.05
    xs = sc * (p["x"] - 1200.0) + 1200.0 + Ji[0, 0] * du + Ji[0, 1] * dv
    ys = sc * (p["y"] - 900.0) + 900.0 + Ji[1, 0] * du + Ji[1, 1] * dv
    assert np.abs(xs - p["x"]).max() > 0.5  # DCR moves blue photons by pixels at 40 deg zenith angle
    x2, y2, f2 = xs.copy(), ys.copy(), p["flux"].copy()
    ctx.rubin_optics(x2, y2, dxdz, dydz, f2, p["wavelength"], p["pupil_u"], p["pupil_v"], p["time"],
                     options=_abi.B2OpticsOptions())
    ok = (flux > 0) & (f2 > 0)
    np.testing.assert_allclose(x[ok], x2[ok], rtol=0, atol=4e-7)
    np.testing.assert_allclose(y[ok], y2[ok], rtol=0, atol=4e-7)


def test_opd_screen():
    """batoid.OPDScreen on a plane (tests/test_telescope_loader.py:641-653): parity with the oracle, the
    physics of a thin phase plate (a tilt W = a x deflects by a, a defocus W = a r^2 focuses at 1 / (2 a)),
    and a null screen changing nothing."""
    from oracle import oracle as orc

    from imsim_b200.telescope import (AIR, CoordSys, Interface, Obscuration, Surface, Telescope)

    # 1. plate + detector plane 10 m below, in vacuum-like air: analytic checks
    def plate(poly, scale=1.0):
        stop = CoordSys(np.zeros(3), np.eye(3))
        items = [Interface("Screen", Surface("plane", poly=np.asarray(poly, float), poly_scale=scale), "pass",
                           CoordSys(np.array([0.0, 0.0, -1.0]), np.eye(3)), AIR, AIR,
                           [Obscuration("circle", (5.0, 0.0, 0.0), negate=True)]),
                 Interface("Detector", Surface("plane"), "detector", CoordSys(np.array([0.0, 0.0, -11.0]), np.eye(3)),
                           AIR, AIR, [])]
        return Telescope(stop, items)

    rng = np.random.default_rng(2)
    n = 5000
    x0, y0 = rng.uniform(-3, 3, n), rng.uniform(-3, 3, n)
    wl = np.full(n, 600e-9)
    nair = AIR.n(wl)

    def shoot(tel):
        ctx = _ctx()
        ctx.set_telescope(tel)
        assert ctx.program == 0
        arrs = [x0.copy(), y0.copy(), np.zeros(n), np.zeros(n), np.zeros(n), -1.0 / nair, np.zeros(n), wl.copy()]
        vig, fail = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        ctx.trace_rays(*arrs, vig, fail)
        ref = orc.trace_rays(*tel.flatten(), x0, y0, np.zeros(n), np.zeros(n), np.zeros(n), -1.0 / nair, np.zeros(n), wl)
        for k in range(7):
            _close(arrs[k], ref[k], scale=max(1.0, float(np.abs(ref[k]).max())), rtol=1e-13)
        assert np.array_equal(vig, ref[7]) and np.array_equal(fail, ref[8])
        return arrs

    a = 1e-4
    tilt = np.zeros((2, 2))
    tilt[1, 0] = a  # W = a x
    out = shoot(plate(tilt))
    np.testing.assert_allclose(out[0] - x0, 10.0 * a / np.sqrt(1 - a * a), rtol=1e-12)  # deflected by asin(a) over 10 m
    np.testing.assert_allclose(out[1], y0, atol=1e-15)
    np.testing.assert_allclose(out[6] - (11.0 * nair), a * x0 + 10.0 * nair * (1 / np.sqrt(1 - a * a) - 1), rtol=0, atol=1e-12)
    foc = np.zeros((3, 3))
    foc[2, 0] = foc[0, 2] = -1.0 / 20.0  # W = -r^2 / (2 f), f = 10 m: a converging plate
    out = shoot(plate(foc))
    r0 = np.hypot(x0, y0)
    r_out = r0 - 10.0 * np.tan(np.arcsin(r0 / 10.0))  # sin(theta) = r / f exactly for this plate
    np.testing.assert_allclose(out[0], x0 * r_out / r0, rtol=0, atol=1e-12)
    np.testing.assert_allclose(out[1], y0 * r_out / r0, rtol=0, atol=1e-12)
    assert np.abs(out[0][r0 < 0.3]).max() < 2e-4  # paraxial focus 10 m behind the plate
    null = shoot(plate(np.zeros((2, 2))))
    np.testing.assert_array_equal(null[0], x0)
    # 2. a Zernike-like screen in front of M1 of the Rubin-like telescope (telescope_loader's use): oracle parity,
    #    and the image shifts by focal length x wavefront tilt
    tel = rubin_like("r")
    scr = np.zeros((3, 3))
    scr[1, 0] = 2e-7  # tilt in units of R_outer
    scr[1, 1] = 1e-7
    tels = tel.with_inserted_screen("M1", poly=scr, poly_scale=1 / 4.18,
                                    obscurations=[Obscuration("annulus", (2.558, 4.18, 0.0, 0.0), negate=True)])
    assert [it.name for it in tels.items][:2] == ["Screen", "M1"]
    rr = np.sqrt(rng.uniform(2.6**2, 4.1**2, n))
    ph = rng.uniform(0, 2 * np.pi, n)
    nair = tel.in_medium.n(wl)  # the telescope's own entrance medium (|v| = 1 / n)
    base = [rr * np.cos(ph), rr * np.sin(ph), np.zeros(n), np.full(n, 1e-3) / nair, np.zeros(n), -np.sqrt(1 - 1e-6) / nair,
            np.zeros(n), wl]
    res = {}
    for name, t in (("plain", tel), ("screen", tels)):
        ctx = _ctx()
        ctx.set_telescope(t)
        arrs = [np.ascontiguousarray(b.copy()) for b in base]
        vig, fail = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
        ctx.trace_rays(*arrs, vig, fail)
        ref = orc.trace_rays(*t.flatten(), *[b.copy() for b in base])
        ok = fail == 0
        _close(arrs[0][ok], ref[0][ok], scale=0.3)
        _close(arrs[1][ok], ref[1][ok], scale=0.3)
        assert np.array_equal(vig, ref[7])
        res[name] = arrs, vig
    good = (res["plain"][1] == 0) & (res["screen"][1] == 0)
    shift = np.median(res["screen"][0][0][good] - res["plain"][0][0][good])
    # wavefront tilt 2e-7 / 4.18 rad x effective focal length 10.31 m ~ 4.9e-7 m, sign set by the three mirrors
    assert 3e-7 < abs(shift) < 7e-7


def test_time_and_pupil_samplers():
    """galsim.TimeSampler / PupilAnnulusSampler (config/imsim-config.yaml:281-289): times uniform over the exposure,
    pupil positions uniform over the annulus, reproducible from (seed, photon offset)."""
    ctx = _ctx()
    n = 2_000_000
    t, u, v = np.empty(n), np.empty(n), np.empty(n)
    ctx.sample_time_pupil(t, u, v, 5.0, 30.0, 2.558, 4.18, 99, 0)
    assert t.min() >= 5.0 and t.max() <= 35.0 and abs(t.mean() - 20.0) < 0.03 and abs(t.var() - 900.0 / 12) < 0.3
    r2 = u * u + v * v
    assert r2.min() >= 2.558**2 * (1 - 1e-12) and r2.max() <= 4.18**2 * (1 + 1e-12)
    # uniform in r^2 and in azimuth
    assert abs(r2.mean() - 0.5 * (2.558**2 + 4.18**2)) < 0.01
    ph = np.arctan2(v, u)
    assert abs(np.cos(ph).mean()) < 3e-3 and abs(np.sin(2 * ph).mean()) < 3e-3
    assert abs(np.corrcoef(t, r2)[0, 1]) < 3e-3
    # the second half regenerated with an offset equals the first call's second half
    t2, u2, v2 = np.empty(n // 2), np.empty(n // 2), np.empty(n // 2)
    ctx.sample_time_pupil(t2, u2, v2, 5.0, 30.0, 2.558, 4.18, 99, n // 2)
    assert np.array_equal(t2, t[n // 2:]) and np.array_equal(u2, u[n // 2:]) and np.array_equal(v2, v[n // 2:])


@pytest.mark.parametrize("program", ["1", "0"])
def test_opd_zemax_on_device(program, monkeypatch):
    """Ray trace PINNED on the device: the CUDA trace (surface program and interpreter) of the LSST v3.3
    prescription against the Zemax wavefront of the reference's tests/test_opd.py:16-95, same tolerances."""
    from test_oracle_golden import check_opd_against_zemax

    monkeypatch.setenv("B2_PROGRAM", program)
    ctx = _ctx()

    def trace(tel, x, y, z, vx, vy, vz, t, wl):
        ctx.set_telescope(tel)
        assert ctx.program == (1 if program == "1" else 0)
        a = [np.ascontiguousarray(v, dtype=np.float64).copy() for v in (x, y, z, vx, vy, vz, t)]
        w = np.full(a[0].size, wl)
        vig = np.zeros(a[0].size, np.uint8)
        fail = np.zeros(a[0].size, np.uint8)
        ctx.trace_rays(*a, w, vig, fail)
        return (*a, vig, fail)

    rms = check_opd_against_zemax(trace)
    assert rms < 0.005  # nm


def test_compiled_xy_to_v_reproduces_the_exact_chain(monkeypatch):
    """XyToV compiled per detector (b2_xytov_compile): the polynomial is adopted only below 2e-9 px on its check
    grid; here it is compared with the exact chain (B2_XYTOV_EXACT=1, and the oracle) on random positions inside
    the box, and positions outside the box must take the exact chain."""
    from imsim_b200 import OpticsContext
    from oracle import oracle as orc

    for det in ("R22_S11", "R41_S20"):
        su = helpers.oracle_setup(det)
        ctx = _ctx(su)
        assert ctx.xytov_residual_px is not None and ctx.xytov_residual_px < 2e-9, ctx.xytov_residual_px
        monkeypatch.setenv("B2_XYTOV_EXACT", "1")
        exact = _ctx(su)
        monkeypatch.delenv("B2_XYTOV_EXACT")
        assert exact.xytov_residual_px is None
        rng = np.random.default_rng(11)
        m = OpticsContext.XYTOV_MARGIN
        x = rng.uniform(-m, su.detector.nx + m, 50000)
        y = rng.uniform(-m, su.detector.ny + m, 50000)
        x[:100] = rng.uniform(-5000, -m - 1, 100)  # outside the box: exact chain
        v = ctx.xy_to_v(x, y)
        ve = exact.xy_to_v(x, y)
        vo = orc.xy_to_v(su.img_wcs.to_pod(), su.icrf_to_field.to_pod(), x, y)
        rad_per_px = 0.2 / 206264.8
        for k in range(2):
            assert np.abs(v[k] - ve[k]).max() < 2e-9 * rad_per_px
            assert np.abs(v[k] - vo[k]).max() < 2e-9 * rad_per_px
            np.testing.assert_array_equal(v[k][:100], ve[k][:100])
