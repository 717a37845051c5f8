"""Photon-shot flat throughput (config 4 of BASELINE.json: examples/flat_with_sed.yaml shape).
usage: python tools/flat_bench.py [section_px] [counts]"""
import os
import sys
import time

os.environ["B2_TIMING"] = "1"
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from imsim_b200._lib import timing_report  # noqa: E402
from imsim_b200.flat import build_flat, flat_nrecalc, wavelength_cdf  # noqa: E402
from imsim_b200.sensor import Image, SiliconSensor  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
counts = float(sys.argv[2]) if len(sys.argv) > 2 else 5000.0
cfg, dat = helpers.sensor_model("lsst_e2v_50_4")
tr = helpers.tree_ring_table("R22_S11")
wave = np.linspace(930.0, 960.0, 31)  # y-band triangular sed x bandpass of flat_with_sed.yaml
cdf = wavelength_cdf(wave, 1.0 - np.abs(wave - 945.0) / 15.0 + 1e-3)
for fused in (True, False):
    sensor = SiliconSensor(config=cfg, vertex_data=dat, rng=1, nrecalc=flat_nrecalc(n + 10, n + 10, 1, 1),
                           treering_func=tr[1], treering_center=tr[0], absorption_table=helpers.absorption())
    img = Image(np.zeros((n, n), np.float32), 1, 1)
    build_flat(img, 1000.0, sensor, rng=1, max_counts_per_iter=1000, nx=1, ny=1, sed_cdf=cdf, fused=fused)  # warm-up
    torch.cuda.synchronize()
    timing_report()
    img = Image(np.zeros((n, n), np.float32), 1, 1)
    t0 = time.perf_counter()
    nphot = build_flat(img, counts, sensor, rng=2, max_counts_per_iter=1000, nx=1, ny=1, sed_cdf=cdf, fused=fused)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    rep = timing_report()
    gpu_ms = sum(v[1] for v in rep.values())
    print("fused=%s section %dx%d counts %.0f: %d photons, wall %.3f s (%.3e photons/s), kernels %.1f ms (%.3e photons/s)"
          % (fused, n, n, counts, nphot, wall, nphot / wall, gpu_ms, nphot / gpu_ms * 1e3))
    print("   ", {k: round(v[1], 2) for k, v in rep.items()}, " mean level %.1f" % img.array.mean())
    sensor.close()
