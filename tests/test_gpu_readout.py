"""Post-path electronics on the device against the reference's own outputs (tests/golden/readout.npz) and
the oracle composition of build_amp_images."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "readout.npz")


def _ctx():
    import torch

    from imsim_b200 import OpticsContext

    return OpticsContext(device=0, stream=torch.cuda.current_stream())


def test_bleed_trails_bit_exact_against_reference():
    from imsim_b200.readout import bleed_eimage

    g = np.load(GOLD)
    ctx = _ctx()
    for tag in "ab":
        fw = float(g["bleed_fw_" + tag])
        for mode, ms in (("mid", True), ("nomid", False)):
            out = bleed_eimage(ctx, g["bleed_in_" + tag].copy(), fw, ms)
            assert np.array_equal(out, g["bleed_%s_%s" % (mode, tag)]), (tag, mode)
    # an unsaturated image is untouched; an empty-range edge case: saturation in row 0 and the last row only
    img = np.full((64, 33), 500.0, np.float32)
    assert np.array_equal(bleed_eimage(ctx, img.copy(), 1e5), img)
    from oracle import readout as R

    img[0, 3] = 5e5
    img[63, 4] = 7e5
    img[31, 5] = 3e5
    img[32, 5] = 3e5
    for ms in (True, False):
        assert np.array_equal(bleed_eimage(ctx, img.copy(), 1e5, ms), R.bleed_eimage(img, 1e5, ms))


def _small_amps(nry=72, nrx=56, namp=4):
    """4 amps tiling a (2 * 60) x (2 * 50) e-image, raw segments 72 x 56 with prescan 4 / overscan, all four
    flip combinations, different gains."""
    from imsim_b200.readout import Amp

    ny, nx = 60, 50
    amps = []
    k = 0
    for row in range(2):
        for col in range(2):
            amps.append(Amp("A%d" % k, col * nx, row * ny, nx, ny, nrx, nry, 4, 2 * row, flip_x=bool(k & 1),
                            flip_y=bool(k & 2), gain=1.3 + 0.1 * k, bias_level=1000.0 + 10 * k, read_noise=0.0))
            k += 1
    return amps, ny, nx


def _oracle_build(e, amps, xtalk, pcti, scti, bias=True):
    """CcdReadout.build_amp_images (readout.py:414-480) without dark current and read noise."""
    from oracle import readout as R

    arrs = []
    for a in amps:
        d = e[a.y0:a.y0 + a.ny, a.x0:a.x0 + a.nx] / a.gain  # float32 / python float -> float32
        assert d.dtype == np.float32
        if a.flip_x:
            d = d[:, ::-1]
        if a.flip_y:
            d = d[::-1, :]
        arrs.append(d)
    arrs = R.apply_crosstalk(arrs, xtalk)
    segs = []
    for d, a in zip(arrs, amps):
        s = np.zeros((a.raw_ny, a.raw_nx), np.float32)
        s[a.data_y0:a.data_y0 + a.ny, a.data_x0:a.data_x0 + a.nx] += d  # float64 (after crosstalk) into ImageF
        segs.append(s)
    segs = R.apply_cte(segs, pcti, scti)
    raw = [np.array(s + np.float32(a.bias_level), dtype=np.int32) for s, a in zip(segs, amps)]
    return np.array(segs), np.array(raw)


@pytest.mark.parametrize("with_xtalk", [False, True])
def test_readout_chain_matches_oracle(with_xtalk):
    from imsim_b200.readout import CcdReadout

    ctx = _ctx()
    amps, ny, nx = _small_amps()
    rng = np.random.default_rng(4)
    e = rng.poisson(900.0, (2 * ny, 2 * nx)).astype(np.float32)
    e[rng.integers(0, 2 * ny, 20), rng.integers(0, 2 * nx, 20)] += rng.uniform(1e4, 9e4, 20).astype(np.float32)
    xt = None
    if with_xtalk:
        xt = rng.normal(0, 3e-4, (4, 4))
        np.fill_diagonal(xt, 0.0)
    ro = CcdReadout(ctx, amps, dark_current=0.0, bias_level=None, scti=3e-5, pcti=1e-4, full_well=None, read_noise=0.0,
                    xtalk=xt)
    raw, seg = ro.build_amp_images(e, want_segments=True)
    oseg, oraw = _oracle_build(e, amps, xt, 1e-4, 3e-5)
    assert np.array_equal(seg.cpu().numpy(), oseg)
    assert np.array_equal(raw.cpu().numpy(), oraw)
    assert raw.cpu().numpy()[:, :, :4].max() <= 1031 and oraw.max() > 5000  # prescan holds bias + deferred charge only


def test_cte_matches_reference_apply_cte():
    """The banded kernel against the reference's dense ``cte_matrix @ column`` (golden vectors)."""
    from imsim_b200.readout import Amp, CcdReadout

    g = np.load(GOLD)
    ctx = _ctx()
    amps_in = g["amps_in"]
    namp, nry, nrx = amps_in.shape
    # amps whose imaging area is the whole raw segment, gain 1, no flips: segments == input
    e = np.concatenate(list(amps_in), axis=1)
    amps = [Amp("A%d" % k, k * nrx, 0, nrx, nry, nrx, nry, 0, 0, False, False, gain=1.0, bias_level=0.0, read_noise=0.0)
            for k in range(namp)]
    for tag in ("both", "p_only", "s_only"):
        p, s = g["cte_%s_cti" % tag]
        ro = CcdReadout(ctx, amps, dark_current=0.0, bias_level=None, scti=s, pcti=p, full_well=None, read_noise=0.0)
        raw, seg = ro.build_amp_images(np.ascontiguousarray(e), want_segments=True)
        assert np.array_equal(seg.cpu().numpy(), g["cte_" + tag]), tag
    xt = g["xtalk"]
    ro = CcdReadout(ctx, amps, dark_current=0.0, bias_level=None, scti=0.0, pcti=0.0, full_well=None, read_noise=0.0,
                    xtalk=xt)
    raw, seg = ro.build_amp_images(np.ascontiguousarray(e), want_segments=True)
    assert np.array_equal(seg.cpu().numpy(), g["xtalk_out"].astype(np.float32))


def test_full_size_readout_statistics():
    """A 4096 x 4004 e2v e-image: dark current and read noise have the right statistics, saturated stars bleed,
    charge is conserved up to the part that escapes into the electronics."""
    import torch

    from imsim_b200.readout import CcdReadout, lsstcam_like_amps

    ctx = _ctx()
    amps = lsstcam_like_amps("e2v", gain=1.0, bias_level=1000.0, read_noise=5.0)
    ny, nx = 4004, 4096
    e = torch.full((ny, nx), 200.0, dtype=torch.float32, device="cuda")
    e[1000:1003, 2000:2003] = 4.0e5  # 9 px x 3e5 excess: a trail of ~27 px per column
    ro = CcdReadout(ctx, amps, scti=0.0, pcti=0.0, full_well=1.0e5, read_noise=5.0)
    raw = ro.build_amp_images(e, seed=3)
    torch.cuda.synchronize()
    eb = ro.eimage.cpu().numpy()
    assert eb.max() <= 1.0e5 + 10 and (eb[:, 2001] >= 1.0e5).sum() >= 10  # trail along the column (+ dark current)
    dark = eb[:, :1000] - 200.0
    assert abs(dark.mean() - 0.64) < 0.01 and abs(dark.var() - 0.64) < 0.02  # Poisson(0.02 * 32)
    r = raw.cpu().numpy()
    assert r.shape == (16, 2048, 576)
    over = r[:, :, 530:].astype(np.float64)  # serial overscan: bias + read noise, truncated to int
    assert abs(over.mean() - 999.5) < 0.05 and abs(over.std() - 5.0) < 0.1
    data = r[3, 100:1900, 20:500].astype(np.float64)
    assert abs(data.mean() - (1200.64 - 0.5)) < 0.2


def test_cosmic_rays_match_reference_paint():
    """The one-launch scatter against the reference's own paint_cr loop (tests/golden/cosmic_rays.npz): same draws
    (injected), same wrap-around at the image edges, same skipped pixels."""
    import torch

    from imsim_b200.cosmic_rays import CosmicRays

    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "cosmic_rays.npz"))
    crs = CosmicRays.from_spans(g["fp_id"], g["x0"], g["y0"], g["pixel_values"], span_len=g["span_len"], exptime=100.0)
    assert len(crs) == 40 and abs(crs.ccd_rate - 0.4) < 1e-12
    it = iter(g["uniforms"])
    img = torch.as_tensor(g["image_in"].copy(), device="cuda")
    crs.paint(_ctx(), img, lambda: next(it), num_crs=int(g["num_crs"]))
    torch.cuda.synchronize()
    assert np.array_equal(img.cpu().numpy(), g["image_out"])
    # the number of hits follows exptime * rate * area fraction (cosmic_rays.py:66-69)
    big = torch.zeros((4000, 4000), dtype=torch.float32, device="cuda")
    crs.paint(_ctx(), big, np.random.default_rng(3), exptime=3000.0)
    hit = int((big > 0).sum())
    assert 0.5 * 1200 * 20 < hit < 2.0 * 1200 * 60  # ~1200 hits of 20..60 pixels each
