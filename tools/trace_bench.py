"""Time k_rubin_optics alone (device-resident pool).  usage: B2_OPTICS_OCC=3 python tools/trace_bench.py [n]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from imsim_b200 import OpticsContext, _abi  # noqa: E402
from imsim_b200.photon_pooling import DevicePhotons, PinnedPhotons  # noqa: E402
from imsim_b200.synthetic import gpu_tracer, make_detector_setup, synthetic_photons  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
ctx = OpticsContext(device=0, stream=torch.cuda.current_stream())
su = make_detector_setup(gpu_tracer(ctx), "R22_S11", rot_tel_pos=np.radians(60.0))
ctx.set_telescope(su.telescope)
ctx.set_wcs(su.img_wcs, su.icrf_to_field)
ctx.set_detector(su.detector)
ctx.set_diffraction(helpers.default_diffraction())
x, y, wl, flux = synthetic_photons(n, kind="stars", seed=0)
pin = PinnedPhotons(n)
pin.x[:], pin.y[:], pin.wavelength[:], pin.flux[:] = x, y, wl, flux
src = DevicePhotons(n)
src.upload(pin)
dp = DevicePhotons(n)
ctx.sample_time_pupil(dp.time, dp.pupil_u, dp.pupil_v, 0.0, 30.0, 2.558, 4.18, 1, 0)
opt = _abi.B2OpticsOptions()
opt.do_refraction, opt.index_ratio, opt.seed = 1, 3.9, 5
times = []
for i in range(8):
    for f in ("x", "y", "wavelength", "flux"):
        getattr(dp, f).copy_(getattr(src, f))
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    ctx.rubin_optics(dp.x, dp.y, dp.dxdz, dp.dydz, dp.flux, dp.wavelength, dp.pupil_u, dp.pupil_v, dp.time,
                     options=opt, want_stats=False)
    b.record()
    torch.cuda.synchronize()
    times.append(a.elapsed_time(b))
ms = float(np.median(times[3:]))
print("OCC=%s n=%d  %.3f ms  %.3e photons/s  vignetted frac %.4f" % (os.environ.get("B2_OPTICS_OCC", "default"), n, ms,
                                                                      n / ms * 1e3, float((dp.flux == 0).double().mean())))

# ---- decomposition: no diffraction; trace only (k_trace_rays on stop-plane rays)
def timeit(fn, reps=6):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts[2:]))

ctx.set_diffraction(None)
def run_nodif():
    ctx.rubin_optics(dp.x, dp.y, dp.dxdz, dp.dydz, dp.flux, dp.wavelength, dp.pupil_u, dp.pupil_v, dp.time,
                     options=opt, want_stats=False)
for f in ("x", "y", "wavelength", "flux"):
    getattr(dp, f).copy_(getattr(src, f))
ms = timeit(run_nodif)
print("  no diffraction: %.3f ms  %.3e photons/s" % (ms, n / ms * 1e3))
vx, vy, vz = (torch.empty(n, dtype=torch.float64, device="cuda") for _ in range(3))
for f in ("x", "y"):
    getattr(dp, f).copy_(getattr(src, f))
ms = timeit(lambda: ctx.xy_to_v(dp.x, dp.y, out=(vx, vy, vz)))
print("  xy_to_v only:   %.3f ms  %.3e photons/s" % (ms, n / ms * 1e3))
z = torch.zeros(n, dtype=torch.float64, device="cuda")
t = torch.zeros(n, dtype=torch.float64, device="cuda")
wl_m = dp.wavelength * 1e-9
vig = torch.zeros(n, dtype=torch.uint8, device="cuda")
fail = torch.zeros(n, dtype=torch.uint8, device="cuda")
rx, ry = dp.pupil_u.clone(), dp.pupil_v.clone()
def run_trace():
    rx.copy_(dp.pupil_u); ry.copy_(dp.pupil_v); z.zero_(); t.zero_()
    a, b, c = vx.clone(), vy.clone(), vz.clone()
    ctx.trace_rays(rx, ry, z, a, b, c, t, wl_m, vig, fail)
base = timeit(lambda: (rx.copy_(dp.pupil_u), ry.copy_(dp.pupil_v), z.zero_(), t.zero_(), vx.clone(), vy.clone(), vz.clone()))
ms = timeit(run_trace) - base
print("  trace_rays only: %.3f ms  %.3e photons/s" % (ms, n / ms * 1e3))
