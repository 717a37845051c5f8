"""Time k_accumulate alone on device-resident photons.  usage: python tools/accum_bench.py [n]"""
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
from imsim_b200.photon_pooling import DevicePhotons, PinnedPhotons  # noqa: E402
from imsim_b200.sensor import Image, SiliconSensor  # noqa: E402
from imsim_b200.synthetic import synthetic_photons  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
ctx = OpticsContext(device=0, stream=torch.cuda.current_stream())
cfg, dat = helpers.sensor_model("lsst_e2v_50_4")
tr = helpers.tree_ring_table("R22_S11")
for kind in ("uniform", "stars"):
    for strength in (1e-12, 1.0):
        sensor = SiliconSensor(config=cfg, vertex_data=dat, nrecalc=0, strength=strength, rng=1, treering_func=tr[1],
                               treering_center=tr[0], absorption_table=helpers.absorption(), context=ctx)
        x, y, wl, flux = synthetic_photons(n, kind=kind, seed=0)
        rng = np.random.default_rng(1)
        pin = PinnedPhotons(n)
        pin.x[:], pin.y[:], pin.wavelength[:], pin.flux[:] = x, y, wl, flux
        pin.dxdz[:] = rng.normal(0, 0.05, n)
        pin.dydz[:] = rng.normal(0, 0.05, n)
        dp = DevicePhotons(n)
        dp.upload(pin, fields=("x", "y", "flux", "wavelength", "dxdz", "dydz"))
        dp._has.update(pupil_u=False, pupil_v=False, time=False)
        img = Image(np.zeros((4004, 4096), np.float32), 0, 0)
        sensor.accumulate(dp, img, resume=False, sync_image=False, want_stats=False)
        for i in range(6):
            sensor.accumulate(dp, img, resume=True, recalc=True, sync_image=False, want_stats=False)
        torch.cuda.synchronize()
        timing_report()
        st = None
        for i in range(3):
            sensor.accumulate(dp, img, resume=True, recalc=True, sync_image=False, want_stats=(i == 2))
        st = sensor.last_stats
        torch.cuda.synchronize()
        rep = timing_report()
        ms = rep["k_accumulate"][1] / rep["k_accumulate"][0]
        print("%-8s strength=%-6g n=%d  k_accumulate %.3f ms  %.3e photons/s | poly tests %.3f%% neighbour %.3f%% | "
              "update %.3f ms bounds %.3f ms" % (kind, strength, n, ms, n / ms * 1e3, 100.0 * st.n_polygon_tests / n,
                                                  100.0 * st.n_neighbor_search / n,
                                                  rep["update_distortions(total)"][1] / 3, rep["k_update_bounds"][1] / 3))
        sensor.close()
