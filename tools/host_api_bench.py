"""Throughput of the plain host-array calls a `galsim config.yaml` run makes (pageable numpy arrays in, results
back in the same arrays): RubinDiffractionOptics.applyTo -> b2_rubin_optics, SiliconSensor.accumulate ->
b2_sensor_accumulate; single staged copy (B2_PIPE_MIN huge) against the pipelined pinned ring of
csrc/hostpipe.cu (copy threads: B2_HOST_THREADS, one setting per process).
usage: [B2_HOST_THREADS=k] python tools/host_api_bench.py [log2_n]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from imsim_b200 import OpticsContext, PhotonArray, _abi  # noqa: E402
from imsim_b200.sensor import Image, SiliconSensor  # noqa: E402
from imsim_b200.synthetic import gpu_tracer, make_detector_setup, synthetic_photons  # noqa: E402

n = 1 << (int(sys.argv[1]) if len(sys.argv) > 1 else 24)
ctx = OpticsContext(device=0)
su = make_detector_setup(gpu_tracer(ctx), "R22_S11", rot_tel_pos=np.radians(60.0))
ctx.set_telescope(su.telescope)
ctx.set_wcs(su.img_wcs, su.icrf_to_field)
ctx.set_detector(su.detector)
ctx.set_diffraction(helpers.default_diffraction())
rng = np.random.default_rng(1)
hx, hy, hwl, hflux = synthetic_photons(n, su.detector.nx, su.detector.ny, seed=0, kind="stars")
pu, pv, t = rng.uniform(-4, 4, n), rng.uniform(-4, 4, n), rng.uniform(0, 30, n)
cfg, dat = helpers.sensor_model("lsst_e2v_50_4")
tr = helpers.tree_ring_table()
sensor = SiliconSensor(config=cfg, vertex_data=dat, nrecalc=0, rng=3, treering_func=tr[1], treering_center=tr[0],
                       absorption_table=helpers.absorption(), context=ctx)
opt = _abi.B2OpticsOptions()
opt.do_refraction, opt.index_ratio = 1, 3.9
img = Image(np.zeros((su.detector.ny, su.detector.nx), np.float32), 0, 0)


def one(reps=3):
    best_o = best_a = 1e9
    for _ in range(reps):
        x, y, flux = hx.copy(), hy.copy(), hflux.copy()
        dxdz, dydz = np.empty(n), np.empty(n)
        t0 = time.perf_counter()
        ctx.rubin_optics(x, y, dxdz, dydz, flux, hwl, pu, pv, t, options=opt)
        t1 = time.perf_counter()
        pa = PhotonArray(n, x=x, y=y, flux=flux, dxdz=dxdz, dydz=dydz, wavelength=hwl)
        img.array[:] = 0
        t2 = time.perf_counter()
        sensor.accumulate(pa, img, sync_image=False, want_stats=False)
        ctx.synchronize()
        t3 = time.perf_counter()
        best_o, best_a = min(best_o, t1 - t0), min(best_a, t3 - t2)
    return n / best_o, n / best_a


out = {"photons": n}
os.environ["B2_PIPE_MIN"] = str(10**15)
out["staged"] = one()
del os.environ["B2_PIPE_MIN"]
out["pipelined"] = one()
out["host_threads"] = int(os.environ.get("B2_HOST_THREADS", 0)) or min(8, os.cpu_count() or 1)
print(json.dumps(out))
