"""microbenchmarks (the roofline denominators) and N-sweeps of the hot kernels on one B200.
Writes one JSON document to stdout.  usage: python tools/sweep.py > gpurun_out/sweeps.json"""
import json
import os
import sys

os.environ["B2_TIMING"] = "1"
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from imsim_b200 import OpticsContext  # noqa: E402
from imsim_b200._lib import timing_report  # noqa: E402
from imsim_b200.photon_pooling import DevicePhotons, PhotonPool  # noqa: E402
from imsim_b200.sensor import Image, SiliconSensor  # noqa: E402
from imsim_b200.synthetic import gpu_tracer, make_detector_setup, synthetic_photons  # noqa: E402

ctx = OpticsContext(device=0, stream=torch.cuda.current_stream())
out = {"device": torch.cuda.get_device_name(0)}
# ---- microbenchmarks
a = torch.empty(1 << 28, dtype=torch.float64, device="cuda")
b = torch.empty_like(a)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
best = 1e9
for _ in range(5):
    ev[0].record()
    b.copy_(a)
    ev[1].record()
    torch.cuda.synchronize()
    best = min(best, ev[0].elapsed_time(ev[1]))
out["hbm_copy_GBps"] = 2 * a.numel() * 8 / best / 1e6
del a, b
out["fma_peak_TFLOPs"] = {"fp64": ctx.fma_peak(True), "fp32": ctx.fma_peak(False)}
out["atomic_add_per_s"] = {}
for fp64 in (True, False):
    for pat, name in ((0, "uniform_random"), (1, "one_hot_pixel"), (2, "1000_stars_sigma1.5")):
        out["atomic_add_per_s"]["%s_%s" % ("double" if fp64 else "float", name)] = ctx.atomic_peak(fp64, pat)
# ---- N sweep of the fused pool step and of the separate kernels (C1 / C2 shapes)
su = make_detector_setup(gpu_tracer(ctx), "R22_S11", rot_tel_pos=np.radians(60.0))
ctx.set_telescope(su.telescope)
ctx.set_wcs(su.img_wcs, su.icrf_to_field)
ctx.set_detector(su.detector)
ctx.set_diffraction(helpers.default_diffraction())
cfg, dat = helpers.sensor_model("lsst_e2v_50_4")
tr = helpers.tree_ring_table("R22_S11")
sensor = SiliconSensor(config=cfg, vertex_data=dat, nrecalc=0, strength=1.0, rng=1, treering_func=tr[1],
                       treering_center=tr[0], absorption_table=helpers.absorption(), context=ctx)
pool = PhotonPool(ctx, sensor, exptime=30.0, seed=5)
img = Image(np.zeros((su.detector.ny, su.detector.nx), np.float32), 0, 0)
out["sweep"] = []
for kind in ("stars", "uniform"):
    for n in (100_000, 1_000_000, 10_000_000, 100_000_000):
        x, y, wl, flux = synthetic_photons(min(n, 1 << 24), kind=kind, seed=1)
        reps = -(-n // x.size)
        src = DevicePhotons(n, fields=("x", "y", "flux", "wavelength"))
        for f, arr in (("x", x), ("y", y), ("wavelength", wl), ("flux", flux)):
            getattr(src, f).copy_(torch.as_tensor(np.tile(arr, reps)[:n]))
        rec = {"workload": kind, "photons": n}
        for fused in (True, False):
            dp = DevicePhotons(n)
            times = []
            for it in range(5):
                for f in ("x", "y", "wavelength", "flux"):
                    getattr(dp, f).copy_(getattr(src, f))
                torch.cuda.synchronize()
                timing_report()
                ev[0].record()
                pool.process(dp, img, resume=it > 0, recalc=it > 0, fused=fused)
                ev[1].record()
                torch.cuda.synchronize()
                times.append((ev[0].elapsed_time(ev[1]), timing_report()))
            ms, rep = sorted(times[2:], key=lambda t: t[0])[len(times[2:]) // 2]
            rec["fused" if fused else "separate"] = {"ms": ms, "photons_per_s": n / ms * 1e3,
                                                     "kernels_ms": {k: round(v[1], 4) for k, v in rep.items()}}
        out["sweep"].append(rec)
        del src, dp
print(json.dumps(out, indent=1))
