"""Per-kernel device time of one visit CCD (B2_TIMING events).  usage: python tools/visit_kernel_breakdown.py [--catalog]"""
import os
import sys

os.environ["B2_TIMING"] = "1"
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from imsim_b200._lib import timing_report  # noqa: E402
from imsim_b200.flat import wavelength_cdf  # noqa: E402
from imsim_b200.visit import DetectorRunner, synthetic_catalog, synthetic_objects  # noqa: E402

catalog = "--catalog" in sys.argv
models = {"e2v": helpers.sensor_model("lsst_e2v_50_4"), "itl": helpers.sensor_model("lsst_itl_50_4")}
tr = helpers.tree_ring_table("R22_S11")
psf = None
wave = np.linspace(550.0, 690.0, 29)
cdf = wavelength_cdf(wave, np.ones_like(wave))
if catalog:
    from imsim_b200.atmosphere import AtmosphericPSF

    psf = AtmosphericPSF(1.2, 0.7, "r", rng=271828, device="cuda:0")
    seds = [wavelength_cdf(wave, 1.0 + 0.8 * np.sin(wave / (15.0 + 5 * k))) for k in range(8)]
    cdf = (np.array([c for c, _ in seds]), np.array([w for _, w in seds]))
runner = DetectorRunner(0, models, helpers.absorption(), tree_rings={"R22_S11": tr, "R22_S12": tr}, psf=psf)
for k, d in enumerate(["R22_S11", "R22_S12", "R22_S11"]):
    objs = (synthetic_catalog if catalog else synthetic_objects)(20000, 4096, 4004, seed=k, total_photons=1e8)
    timing_report()
    rec, _ = runner.run(d, objs, nbatch=10, wavelength_cdf=cdf, det_index=k, readout=catalog, sky_level=800.0 if catalog else 0.0)
    rep = timing_report()
    tot = sum(v[1] for v in rep.values())
    print(d, "gpu_ms %.1f  kernels %.1f ms:" % (rec["gpu_ms"], tot),
          {k2: (v[0], round(v[1], 2)) for k2, v in sorted(rep.items(), key=lambda kv: -kv[1][1])})
