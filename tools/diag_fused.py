"""diagnostic: largest difference between the fused pool step and the separate kernels (traced arrays)"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import test_gpu_pool as T
from imsim_b200.sensor import Image
res = {}
for fused in (False, True):
    su, ctx, sensor, pool, pin, n, DevicePhotons = T._setup()
    img = Image(np.zeros((su.detector.ny, su.detector.nx), np.float32), 0, 0)
    dp = DevicePhotons(n); dp.upload(pin)
    pool.process(dp, img, resume=False, recalc=False, want_stats=True, fused=fused, write_back=True)
    torch.cuda.synchronize()
    res[fused] = [getattr(dp, f).cpu().numpy() for f in ("x", "y", "dxdz", "dydz", "flux")]
    print("xytov resid", ctx.xytov_residual_px)
for name, a, b in zip(("x", "y", "dxdz", "dydz", "flux"), res[True], res[False]):
    d = np.abs(a - b)
    print(name, "max abs diff", np.nanmax(d), "n differing", int((a != b).sum()), "of", a.size)
