"""Classic per-object pipeline (config 1: single CCD, instance-catalogue-like field, BF + tree rings, nrecalc 1e4).
usage: python tools/classic_bench.py [n_objects] [total_photons]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from imsim_b200 import OpticsContext  # noqa: E402
from imsim_b200.atmosphere import AtmosphericPSF  # noqa: E402
from imsim_b200.detector import lsstcam_like  # noqa: E402
from imsim_b200.diffraction import RUBIN_LATITUDE, diffraction_config  # noqa: E402
from imsim_b200.flat import wavelength_cdf  # noqa: E402
from imsim_b200.lsst_image import ClassicImageBuilder  # noqa: E402
from imsim_b200.sensor import Image, SiliconSensor  # noqa: E402
from imsim_b200.synthetic import gpu_tracer, make_detector_setup  # noqa: E402
from imsim_b200.visit import synthetic_catalog  # noqa: E402

n_obj = int(sys.argv[1]) if len(sys.argv) > 1 else 1998   # examples/example_instance_catalog.txt has 1998 entries
total = float(sys.argv[2]) if len(sys.argv) > 2 else 5e7
ctx = OpticsContext(device=0, stream=torch.cuda.current_stream())
det = lsstcam_like("R22_S11")
su = make_detector_setup(gpu_tracer(ctx), "R22_S11", band="r", rot_tel_pos=np.radians(30.0), detector=det)
ctx.set_telescope(su.telescope)
ctx.set_wcs(su.img_wcs, su.icrf_to_field)
ctx.set_detector(su.detector)
ctx.set_diffraction(diffraction_config(latitude=RUBIN_LATITUDE, altitude=np.radians(67.0), azimuth=np.radians(213.0)))
cfg, dat = helpers.sensor_model("lsst_e2v_50_4")
tr = helpers.tree_ring_table("R22_S11")
sensor = SiliconSensor(config=cfg, vertex_data=dat, nrecalc=10000, strength=1.0, rng=5, treering_func=tr[1],
                       treering_center=tr[0], absorption_table=helpers.absorption(), context=ctx)
cat = synthetic_catalog(n_obj, det.nx, det.ny, seed=3, total_photons=total)
rows, flux = cat.build()
wave = np.linspace(550.0, 690.0, 29)
seds = [wavelength_cdf(wave, 1.0 + 0.8 * np.sin(wave / (15.0 + 5 * k))) for k in range(8)]
psf = AtmosphericPSF(1.2, 0.7, "r", rng=1, device="cuda:0")
b = ClassicImageBuilder(ctx, sensor, rows, cat.radial_tables(), cat.sersic_n, np.array([c for c, _ in seds]),
                        np.array([w for _, w in seds]), psf=psf, seed=2)
import json  # noqa: E402

res = {}
for which, reps in (("build", 3), ("build_per_object", 1)):
  for rep in range(reps):
    image = Image(np.zeros((det.ny, det.nx), np.float32), 0, 0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = getattr(b, which)(image, flux, phot_flux=flux.astype(np.int64))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res[which] = {"seconds": dt, "photons_per_s": st["photons"] / dt, "host_setup_s": st.get("host_setup_seconds"),
                  "electrons": float(image.array.sum(dtype=np.float64)), "objects": int(rows.size),
                  "photons": int(st["photons"]), "phases": st.get("phases")}
    print(which, "classic pipeline: %d catalogue rows (%d phot, %d faint), %d photons: %.2f s = %.3e photons/s, %.2f ms / object; "
          "electrons %.4e" % (rows.size, st["phot"], st["faint"], st["photons"], dt, st["photons"] / dt,
                              1e3 * dt / rows.size, image.array.sum(dtype=np.float64)), "host set-up %s s" % st.get("host_setup_seconds"))
print(json.dumps({"classic_bench": res}))
