"""The pipelined B2_HOST route (csrc/hostpipe.cu): large pageable numpy arrays -- what a GalSim run hands to
``applyTo`` / ``accumulate`` (imsim/photon_ops.py:81, imsim/photon_pooling.py:210) -- go through a ring of
pinned slots in chunks, host copy threads, DMA and kernel overlapped.  The route must not change one bit of the
results of the single staged copy."""
import os

import numpy as np
import pytest

import helpers
from imsim_b200 import _abi

pytestmark = pytest.mark.gpu


class _Env:
    def __init__(self, **kw):
        self.kw, self.old = kw, {}

    def __enter__(self):
        for k, v in self.kw.items():
            self.old[k] = os.environ.get(k)
            os.environ[k] = str(v)

    def __exit__(self, *a):
        for k, v in self.old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


PIPE_OFF = dict(B2_PIPE_MIN=10**15)


def _optics(ctx, p, gauss, opt, time_out=False):
    x, y, flux = p["x"].copy(), p["y"].copy(), p["flux"].copy()
    dxdz, dydz = np.full(x.size, np.nan), np.full(x.size, np.nan)
    tout = np.full(x.size, np.nan) if time_out else None
    st = ctx.rubin_optics(x, y, dxdz, dydz, flux, p["wavelength"], p["pupil_u"], p["pupil_v"], p["time"], gauss=gauss,
                          time_out=tout, options=opt)
    return (x, y, dxdz, dydz, flux) + ((tout,) if time_out else ()), (st.n_vignetted, st.n_failed, st.n_offdetector_z)


@pytest.mark.parametrize("inject", [True, False])
@pytest.mark.parametrize("n,env", [(50001, dict(B2_PIPE_MIN=1, B2_PIPE_CHUNK=4096)),      # 13 chunks, ragged tail
                                   (4097, dict(B2_PIPE_MIN=1, B2_PIPE_CHUNK=4096)),       # tail of one photon
                                   (3000, dict(B2_PIPE_MIN=1, B2_PIPE_CHUNK=4096)),       # a single short chunk
                                   (600000, dict())])                                      # defaults: 2 chunks of 2^19
def test_rubin_optics_pipelined_equals_staged(n, env, inject):
    from imsim_b200 import OpticsContext

    su = helpers.oracle_setup()
    ctx = OpticsContext(device=0)
    ctx.set_telescope(su.telescope)
    ctx.set_wcs(su.img_wcs, su.icrf_to_field)
    ctx.set_detector(su.detector)
    ctx.set_diffraction(helpers.default_diffraction())
    rng = np.random.default_rng(n)
    p = helpers.test_photon_arrays(n=n, center=(2000.0, 1900.0))
    p["x"] += rng.normal(0, 500, n)
    p["y"] += rng.normal(0, 500, n)
    p["time"] = rng.uniform(0, 30, n)
    gauss = rng.standard_normal(n) if inject else None
    opt = _abi.B2OpticsOptions()
    opt.do_refraction, opt.index_ratio, opt.seed, opt.photon_offset = 1, 3.9, 1234, 77
    with _Env(**PIPE_OFF):
        ref, ref_stats = _optics(ctx, p, gauss, opt, time_out=inject)
    with _Env(**env):
        out, stats = _optics(ctx, p, gauss, opt, time_out=inject)
    assert stats == ref_stats and ref_stats[0] > 0
    for a, b in zip(out, ref):
        assert np.array_equal(a, b, equal_nan=True)
    assert not np.isnan(out[2]).all()


@pytest.mark.parametrize("inject", [True, False])
def test_accumulate_pipelined_upload_equals_staged(inject):
    from imsim_b200 import OpticsContext, PhotonArray
    from imsim_b200.sensor import Image, SiliconSensor

    ctx = OpticsContext(device=0)
    cfg, dat = helpers.sensor_model("lsst_e2v_50_4")
    tr = helpers.tree_ring_table()
    aw, al = helpers.absorption()
    n = 70001
    rng = np.random.default_rng(5)
    pa = PhotonArray(n, x=rng.normal(200, 30, n), y=rng.normal(180, 30, n), flux=np.ones(n),
                     dxdz=rng.normal(0, 0.05, n), dydz=rng.normal(0, 0.05, n), wavelength=rng.uniform(550, 690, n))
    rand4 = np.vstack([rng.standard_normal(n), rng.standard_normal(n), rng.uniform(size=n), rng.uniform(size=n)]) \
        if inject else None
    images = []
    for env in (PIPE_OFF, dict(B2_PIPE_MIN=1, B2_PIPE_CHUNK=8192)):
        sensor = SiliconSensor(config=cfg, vertex_data=dat, nrecalc=20000, rng=3, treering_func=tr[1],
                               treering_center=tr[0], absorption_table=(aw, al), context=ctx)
        img = Image(np.zeros((400, 400), np.float32), 1, 1)
        with _Env(**env):
            added = sensor.accumulate(pa, img, rand4=rand4)
        images.append((added, img.array.copy()))
    assert images[0][0] == images[1][0] > 0.9 * n
    assert np.array_equal(images[0][1], images[1][1])


def test_copy_through_ring_round_trip():
    """b2_copy_through_ring: a pageable host array to the device and back (large: through the pinned ring, also with
    non-temporal stores; small: plain copy); the size must be a multiple of 8 bytes."""
    import torch

    from imsim_b200 import OpticsContext, _lib

    ctx = OpticsContext(device=0)
    lib = _lib.load()
    rng = np.random.default_rng(0)
    for n, env in ((1 << 22) + 3, {}), ((1 << 22) + 3, dict(B2_COPY_NT=1)), (1000, {}):
        with _Env(**env):
            a = rng.standard_normal(n).astype(np.float32)[: n - (n % 2)]  # 8-byte multiple
            dev = torch.empty(a.size, dtype=torch.float32, device="cuda:0")
            _lib.check(lib.b2_copy_through_ring(ctx.handle, a.ctypes.data, dev.data_ptr(), a.nbytes, 1))
            torch.cuda.synchronize()
            assert np.array_equal(dev.cpu().numpy(), a)
            dev.mul_(2.0)
            torch.cuda.synchronize()
            back = np.empty_like(a)
            _lib.check(lib.b2_copy_through_ring(ctx.handle, back.ctypes.data, dev.data_ptr(), back.nbytes, 0))
            assert np.array_equal(back, 2.0 * a)
    with pytest.raises(_lib.B2Error):
        _lib.check(lib.b2_copy_through_ring(ctx.handle, back.ctypes.data, dev.data_ptr(), 12, 0))
