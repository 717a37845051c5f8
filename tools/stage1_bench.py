"""Throughput of k_stage1_photons with full-size phase screens.  usage: python tools/stage1_bench.py [n]"""
import os
import sys
import time

os.environ["B2_TIMING"] = "1"
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from imsim_b200 import OpticsContext  # noqa: E402
from imsim_b200._lib import timing_report  # noqa: E402
from imsim_b200.atmosphere import AtmosphericPSF, GaussianPSF  # noqa: E402
from imsim_b200.flat import wavelength_cdf  # noqa: E402
from imsim_b200.photon_pooling import DevicePhotons  # noqa: E402
from imsim_b200.stage1 import Stage1  # noqa: E402
from imsim_b200.visit import synthetic_catalog  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 1 << 25
ctx = OpticsContext(device=0, stream=torch.cuda.current_stream())
t0 = time.perf_counter()
cat = synthetic_catalog(20000, 4096, 4004, seed=1, total_photons=n)
rows, flux = cat.build()
print("catalogue: %d rows in %.2f s" % (rows.size, time.perf_counter() - t0))
wave = np.linspace(550.0, 690.0, 29)
seds = [wavelength_cdf(wave, 1.0 + 0.8 * np.sin(wave / (15.0 + 5 * k))) for k in range(8)]
st = Stage1(ctx, rows, np.array([c for c, _ in seds]), np.array([w for _, w in seds]), cat.radial_tables())
counts = flux.astype(np.int64)
dp = DevicePhotons(int(counts.sum()), device="cuda:0", fields=("x", "y", "flux", "wavelength"))
for name, mk in (("no psf", None), ("gaussian", lambda: GaussianPSF(0.7)),
                 ("atmosphere f32 8192^2 x 6", lambda: AtmosphericPSF(1.2, 0.7, "r", rng=1, device="cuda:0")),
                 ("atmosphere f64 8192^2 x 6", lambda: AtmosphericPSF(1.2, 0.7, "r", rng=1, device="cuda:0", dtype=np.float64))):
    t0 = time.perf_counter()
    psf = mk() if mk else None
    if psf is not None:
        psf.upload(ctx)
    torch.cuda.synchronize()
    tb = time.perf_counter() - t0
    for k in range(3):
        st.shoot(dp, counts, seed=k)
    torch.cuda.synchronize()
    timing_report()
    for k in range(5):
        st.shoot(dp, counts, seed=10 + k)
    torch.cuda.synchronize()
    rep = timing_report()["k_stage1_photons"]
    ms = rep[1] / rep[0]
    print("%-28s build %.2f s   %.3f ms / %d photons = %.3e photons/s" % (name, tb, ms, dp.n, dp.n / ms * 1e3))
    del psf
