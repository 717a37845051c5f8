"""examples/flat.yaml (BASELINE config 4 as named): 4096 x 4096 sky flat by the pixel-area branch, counts_per_pixel
e-/px in iterations of 1000, 8 x 2 sections.  usage: python tools/flat_area_bench.py [counts] [size]"""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from imsim_b200.flat import build_flat  # noqa: E402
from imsim_b200.sensor import Image, SiliconSensor  # noqa: E402

counts = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0e5
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
cfg, dat = helpers.sensor_model("lsst_e2v_50_4")
tr = helpers.tree_ring_table("R22_S11")
sensor = SiliconSensor(config=cfg, vertex_data=dat, rng=1, treering_func=tr[1], treering_center=tr[0],
                       absorption_table=helpers.absorption())
for fused, c in ((True, 2000.0), (True, counts), (False, min(counts, 4000.0))):
    img = Image(np.zeros((n, n), np.float32), 1, 1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    build_flat(img, c, sensor, rng=2, nx=8, ny=2, fused=fused)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    a = img.array.astype(np.float64)
    print("fused=%s %dx%d counts %.0f: %.2f s (%.1f ms per section-iteration), mean %.1f var/mean %.4f"
          % (fused, n, n, c, dt, 1e3 * dt / (16 * np.ceil(c / 1000)), a.mean(), a.var() / a.mean()))
