"""The device-side nrecalc cadence of per-object stamps (csrc/stamps.cu, b2_sensor_accumulate_stamps) against the
host-driven path it replaces: every stamp must come out bit-identical to bind(zeros) + accumulate on the same
photons with the same draws -- boundary updates at the same photons, same polygon tests, same image -- and the
stamps must land in the full image like ``full_image[bounds] += stamp[bounds]`` (imsim/lsst_image.py:359-368).
The host-driven path itself is pinned against the oracle in tests/test_gpu_sensor.py."""
import numpy as np
import pytest

import helpers
from imsim_b200.sensor import Image, SiliconSensor

pytestmark = pytest.mark.gpu


def _sensor(model="lsst_itl_50_4", nrecalc=10000, treerings=True, **kw):
    cfg, dat = helpers.sensor_model(model)
    tr = helpers.tree_ring_table() if treerings else None
    return SiliconSensor(config=cfg, vertex_data=dat, nrecalc=nrecalc, rng=11, absorption_table=helpers.absorption(),
                         treering_func=tr[1] if tr else None, treering_center=tr[0] if tr else (0.0, 0.0), **kw)


def _objects(rng, nobj, full_nx, full_ny, big=False):
    """Stars of assorted brightness, some hanging over the edge of the full image, one of them empty."""
    jobs, xs, ys, ws, fl, ax, ay = [], [], [], [], [], [], []
    p0 = 0
    for k in range(nobj):
        n = int(rng.choice([0, 300, 4000, 25000, 120000 if big else 60000]))
        size = int(rng.choice([9, 16, 33, 48]))
        cx, cy = rng.uniform(-5, full_nx + 5), rng.uniform(-5, full_ny + 5)
        icx, icy = int(np.floor(cx + 0.5)), int(np.floor(cy + 0.5))
        jobs.append((p0, n, icx - size // 2, icy - size // 2, size, size, int(k % 7 == 3)))
        xs.append(cx + rng.normal(0, 1.3, n))
        ys.append(cy + rng.normal(0, 1.3, n))
        ws.append(rng.uniform(400.0, 1000.0, n))
        fl.append(np.ones(n))
        ax.append(rng.normal(0, 0.15, n))
        ay.append(rng.normal(0, 0.15, n))
        p0 += n
    cat = lambda v: np.concatenate(v) if v else np.zeros(0)
    return jobs, cat(xs), cat(ys), cat(ws), cat(fl), cat(ax), cat(ay)


def _device_photons(x, y, wl, flux, dxdz, dydz):
    import torch

    from imsim_b200.photon_pooling import DevicePhotons

    dp = DevicePhotons(x.size, device="cuda:0")
    for name, v in (("x", x), ("y", y), ("wavelength", wl), ("flux", flux), ("dxdz", dxdz), ("dydz", dydz)):
        getattr(dp, name).copy_(torch.as_tensor(v))
    dp._has.update(dxdz=True, dydz=True, wavelength=True)
    return dp


#: one thread block per stamp; every silicon stamp on a cluster of 8 / of 4 blocks (B2_STAMP_HEAVY = 1: all are "heavy")
TEAMS = [("1", None), ("8", "1"), ("4", "1")]


def _team_env(monkeypatch, cluster, heavy):
    monkeypatch.setenv("B2_STAMP_CLUSTER", cluster)
    if heavy is not None:
        monkeypatch.setenv("B2_STAMP_HEAVY", heavy)


@pytest.mark.parametrize("cluster,heavy", TEAMS)
@pytest.mark.parametrize("model,nrecalc,dtype", [("lsst_itl_50_4", 10000, np.float32), ("lsst_e2v_50_8", 3000, np.float64),
                                                 ("lsst_itl_50_4", 0, np.float32)])
def test_stamps_equal_host_driven_accumulate(model, nrecalc, dtype, cluster, heavy, monkeypatch):
    import torch

    _team_env(monkeypatch, cluster, heavy)

    from imsim_b200 import PhotonArray

    rng = np.random.default_rng(4)
    full_nx, full_ny = 150, 120
    jobs, x, y, wl, flux, dxdz, dydz = _objects(rng, 24, full_nx, full_ny)
    n = x.size
    rand4 = np.vstack([rng.standard_normal(n), rng.standard_normal(n), rng.uniform(size=n), rng.uniform(size=n)])
    sensor = _sensor(model, nrecalc)
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    full = torch.zeros((full_ny, full_nx), dtype=tdt, device="cuda:0")
    full[5, 7] = 3.0  # what is already there stays
    dp = _device_photons(x, y, wl, flux, dxdz, dydz)
    stats, added = sensor.accumulate_stamps(jobs, dp, full, 1, 1, rand4=torch.as_tensor(rand4, device="cuda:0"),
                                            want_added=True)
    got = full.cpu().numpy()
    # reference: one host-driven accumulate per object on its own stamp image
    ref_sensor = _sensor(model, nrecalc)
    want = np.zeros((full_ny, full_nx), dtype=dtype)
    want[5, 7] = 3.0
    n_upd = 0
    for k, (p0, m, xmin, ymin, nx, ny, plain) in enumerate(jobs):
        sl = slice(p0, p0 + m)
        stamp = Image(np.zeros((ny, nx), dtype=dtype), xmin, ymin)
        if plain:
            ix, iy = np.floor(x[sl] + 0.5).astype(int) - xmin, np.floor(y[sl] + 0.5).astype(int) - ymin
            ok = (ix >= 0) & (ix < nx) & (iy >= 0) & (iy < ny)
            np.add.at(stamp.array, (iy[ok], ix[ok]), 1.0)
            a = float(ok.sum())
        else:
            pa = PhotonArray(m, x=x[sl].copy(), y=y[sl].copy(), flux=flux[sl].copy(), dxdz=dxdz[sl].copy(),
                             dydz=dydz[sl].copy(), wavelength=wl[sl].copy())
            a = ref_sensor.accumulate(pa, stamp, rand4=np.ascontiguousarray(rand4[:, sl]))
            n_upd += ref_sensor.last_stats.n_updates if m else 0  # an empty call leaves last_stats alone
        assert added[k] == a, (k, added[k], a)
        x0, x1 = max(xmin, 1), min(xmin + nx, 1 + full_nx)
        y0, y1 = max(ymin, 1), min(ymin + ny, 1 + full_ny)
        if x0 < x1 and y0 < y1:
            want[y0 - 1:y1 - 1, x0 - 1:x1 - 1] += stamp.array[y0 - ymin:y1 - ymin, x0 - xmin:x1 - xmin]
    np.testing.assert_array_equal(got, want)
    assert stats.n_updates == n_upd
    if nrecalc:
        assert n_upd > 10
    assert stats.added_flux == added.sum()


def test_stamps_in_several_arena_loads(monkeypatch):
    """A job list larger than the arena runs as several launches with the same result."""
    import torch

    rng = np.random.default_rng(9)
    jobs, x, y, wl, flux, dxdz, dydz = _objects(rng, 60, 200, 200)
    sensor = _sensor()
    dp = _device_photons(x, y, wl, flux, dxdz, dydz)
    full_a = torch.zeros((200, 200), dtype=torch.float32, device="cuda:0")
    sa = sensor.accumulate_stamps(jobs, dp, full_a)
    sensor.updateRNG(11)
    monkeypatch.setenv("B2_STAMP_ARENA_MB", "64")  # the floor; 60 stamps of up to 48 x 48 pixels still need > 1 load?
    full_b = torch.zeros((200, 200), dtype=torch.float32, device="cuda:0")
    sb = sensor.accumulate_stamps(jobs, dp, full_b)
    assert torch.equal(full_a, full_b) and sa.added_flux == sb.added_flux


def test_stamps_leave_the_bound_image_alone():
    import torch

    from imsim_b200 import PhotonArray

    rng = np.random.default_rng(2)
    sensor = _sensor(nrecalc=0)
    img = Image(np.zeros((40, 40), np.float32))
    n = 5000
    pa = PhotonArray(n, x=rng.uniform(10, 30, n), y=rng.uniform(10, 30, n), flux=np.ones(n))
    sensor.accumulate(pa, img)
    before = img.array.copy()
    jobs, x, y, wl, flux, dxdz, dydz = _objects(rng, 5, 64, 64)
    full = torch.zeros((64, 64), dtype=torch.float32, device="cuda:0")
    sensor.accumulate_stamps(jobs, _device_photons(x, y, wl, flux, dxdz, dydz), full)
    sensor.accumulate(pa, img, resume=True)
    assert img.array.sum() == before.sum() + n


@pytest.mark.parametrize("cluster,heavy", [("1", None), ("8", "1")])
def test_one_wide_stamp_with_more_charged_pixels_than_the_list_holds(cluster, heavy, monkeypatch):
    """A flat-ish 400 x 400 stamp updated once after 8e5 electrons: far more charged pixels than the per-block list
    takes, so the update scans the box of the pending charge -- same image as the host-driven accumulate; a second,
    bright star on a small stamp takes the list path beside it."""
    import torch

    from imsim_b200 import PhotonArray

    _team_env(monkeypatch, cluster, heavy)
    rng = np.random.default_rng(12)
    n1, n2 = 1_000_000, 150_000
    x = np.concatenate([rng.uniform(0.6, 400.4, n1), 500.0 + rng.normal(0, 1.2, n2)])
    y = np.concatenate([rng.uniform(0.6, 400.4, n1), 40.0 + rng.normal(0, 1.2, n2)])
    n = n1 + n2
    wl = rng.uniform(400.0, 1000.0, n)
    flux = np.ones(n)
    dxdz, dydz = rng.normal(0, 0.15, n), rng.normal(0, 0.15, n)
    jobs = [(0, n1, 1, 1, 400, 400, 0), (n1, n2, 480, 20, 40, 40, 0)]
    nrecalc = 800_000
    sensor = _sensor("lsst_itl_50_4", nrecalc)
    full = torch.zeros((420, 560), dtype=torch.float32, device="cuda:0")
    stats = sensor.accumulate_stamps(jobs, _device_photons(x, y, wl, flux, dxdz, dydz), full, 1, 1)
    ref_sensor = _sensor("lsst_itl_50_4", nrecalc)
    ref_sensor.updateRNG(11)
    want = np.zeros((420, 560), np.float32)
    n_upd, p_off = 0, 0
    for p0, m, xmin, ymin, nx, ny, _ in jobs:
        sl = slice(p0, p0 + m)
        stamp = Image(np.zeros((ny, nx), np.float32), xmin, ymin)
        pa = PhotonArray(m, x=x[sl].copy(), y=y[sl].copy(), flux=flux[sl].copy(), dxdz=dxdz[sl].copy(),
                         dydz=dydz[sl].copy(), wavelength=wl[sl].copy())
        ref_sensor.accumulate(pa, stamp)
        n_upd += ref_sensor.last_stats.n_updates
        want[ymin - 1:ymin - 1 + ny, xmin - 1:xmin - 1 + nx] += stamp.array
    np.testing.assert_array_equal(full.cpu().numpy(), want)
    assert stats.n_updates == n_upd and n_upd >= 1
