"""A small pass through every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):
    compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
from imsim_b200 import OpticsContext  # noqa: E402
from imsim_b200.atmosphere import AtmosphericPSF  # noqa: E402
from imsim_b200.flat import build_flat, flat_nrecalc, wavelength_cdf  # noqa: E402
from imsim_b200.photon_pooling import DevicePhotons, PhotonPool  # noqa: E402
from imsim_b200.readout import CcdReadout, lsstcam_like_amps  # noqa: E402
from imsim_b200.sensor import Image, SiliconSensor  # noqa: E402
from imsim_b200.sky import add_sky, pixel_areas_device  # noqa: E402
from imsim_b200.stage1 import ObjectTable, Stage1  # noqa: E402
from imsim_b200.synthetic import gpu_tracer, make_detector_setup  # noqa: E402

ctx = OpticsContext(device=0, stream=torch.cuda.current_stream())
su = make_detector_setup(gpu_tracer(ctx), "R22_S11", rot_tel_pos=0.4)
ctx.set_telescope(su.telescope)
ctx.set_wcs(su.img_wcs, su.icrf_to_field)
ctx.set_detector(su.detector)
ctx.set_diffraction(helpers.default_diffraction())
cfg, dat = helpers.sensor_model("lsst_e2v_50_4")
tr = helpers.tree_ring_table("R22_S11")
wave = np.linspace(550.0, 690.0, 15)
cdf, cw = wavelength_cdf(wave, np.ones_like(wave))
# stage 1 with a small atmosphere -> fused pool step (nrecalc 0) and separate kernels (nrecalc 2000) on a sub-image
psf = AtmosphericPSF(1.2, 0.7, "r", rng=1, screen_size=25.6, screen_scale=0.1, device="cuda:0")
tab = ObjectTable()
tab.add_points([700.0, 820.0], [650.0, 900.0], [1, 1])
tab.add_sersic(760.0, 760.0, 1, 1.0, 4.0, q=0.6, beta=0.3)
tab.add_knots(800.0, 700.0, 1, 0.8, 9, seed=2)
rows, _ = tab.build()
st = Stage1(ctx, rows, cdf[None], cw[None], tab.radial_tables(), psf=psf)
for nrecalc in (0, 2000):
    sensor = SiliconSensor(config=cfg, vertex_data=dat, nrecalc=nrecalc, strength=1.0, rng=3, treering_func=tr[1],
                           treering_center=tr[0], absorption_table=helpers.absorption(), context=ctx)
    pool = PhotonPool(ctx, sensor, exptime=30.0, seed=5)
    img = Image(np.zeros((500, 400), np.float32), 560, 520)
    for k in range(2):
        counts = np.array([6000, 5000, 7000, 3000])
        dp = DevicePhotons(int(counts.sum()))
        st.shoot(dp, counts, seed=9, photon_offset=k * 21000)
        pool.process(dp, img, resume=k > 0, recalc=(k > 0 and nrecalc == 0), fused=(nrecalc == 0))
    sensor.read_image(img)
    print("nrecalc", nrecalc, "electrons", img.array.sum())
    if nrecalc == 0:
        e = torch.empty((500, 400), dtype=torch.float32, device="cuda")
        sensor.snapshot_image(e)
        add_sky(ctx, e, 300.0, seed=4, areas=pixel_areas_device(sensor))
        print("with sky", float(e.sum()))
    sensor.close()
# photon-shot flat, fused and not, with a boundary update
sensor = SiliconSensor(config=cfg, vertex_data=dat, rng=1, nrecalc=flat_nrecalc(74, 74, 1, 1) / 20, treering_func=tr[1],
                       treering_center=tr[0], absorption_table=helpers.absorption(), context=ctx)
for fused in (True, False):
    f = Image(np.zeros((64, 64), np.float32), 1, 1)
    build_flat(f, 2000.0, sensor, rng=1, max_counts_per_iter=500, nx=1, ny=1, sed_cdf=(cdf, cw), fused=fused)
    print("flat fused", fused, f.array.mean())
# readout of a small CCD-like image
amps = lsstcam_like_amps("e2v")
e = torch.poisson(torch.full((4004, 4096), 50.0, device="cuda")).float()
e[100:104, 300:900] = 2.0e5
xt = np.full((16, 16), 1e-4)
np.fill_diagonal(xt, 0.0)
raw = CcdReadout(ctx, amps, xtalk=xt).build_amp_images(e, seed=1)
torch.cuda.synchronize()
print("raw", tuple(raw.shape), int(raw.sum()))
# round 2: per-object stamps with the cadence inside the kernel (csrc/stamps.cu), the compiled XyToV (set_detector above
# compiled it), the pooled upload of host segments and the threaded host copy
sensor = SiliconSensor(config=cfg, vertex_data=dat, nrecalc=1500, strength=1.0, rng=3, treering_func=tr[1],
                       treering_center=tr[0], absorption_table=helpers.absorption(), context=ctx)
rng = np.random.default_rng(1)
jobs, p0, xs, ys = [], 0, [], []
for k, (n, size) in enumerate([(9000, 24), (400, 12), (0, 10), (6000, 31), (2500, 16)]):
    cx, cy = rng.uniform(20, 80, 2)
    jobs.append((p0, n, int(cx) - size // 2, int(cy) - size // 2, size, size, int(k == 4)))
    xs.append(cx + rng.normal(0, 1.2, n))
    ys.append(cy + rng.normal(0, 1.2, n))
    p0 += n
dp = DevicePhotons(p0)
dp.x.copy_(torch.as_tensor(np.concatenate(xs)))
dp.y.copy_(torch.as_tensor(np.concatenate(ys)))
dp.flux.fill_(1.0)
dp.wavelength.fill_(620.0)
dp.dxdz.zero_()
dp.dydz.zero_()
dp._has.update(dxdz=True, dydz=True, wavelength=True)
full = torch.zeros((100, 100), dtype=torch.float32, device="cuda")
stt = sensor.accumulate_stamps(jobs, dp, full, 1, 1)
print("stamps", float(full.sum()), stt.n_updates)
# the same stamps on thread-block clusters (every silicon stamp counted as heavy)
os.environ["B2_STAMP_HEAVY"] = "1"
sensor.updateRNG(3)
full2 = torch.zeros((100, 100), dtype=torch.float32, device="cuda")
stt2 = sensor.accumulate_stamps(jobs, dp, full2, 1, 1)
print("stamps on clusters", float(full2.sum()), stt2.n_updates, bool(torch.equal(full, full2)))
del os.environ["B2_STAMP_HEAVY"]
import ctypes as C  # noqa: E402

from imsim_b200 import _lib  # noqa: E402

segs = [rng.uniform(0, 1, m) for m in (1000, 0, 70000, 333)]
n = sum(len(a) for a in segs)
dst = torch.empty(n, dtype=torch.float64, device="cuda")
ptrs = np.array([a.ctypes.data for a in segs], dtype=np.uint64)
lens = np.array([len(a) for a in segs], dtype=np.int64)
dptr = np.array([dst.data_ptr()], dtype=np.uint64)
_lib.check(_lib.load().b2_photons_upload(ctx.handle, 1, len(segs), ptrs.ctypes.data, lens.ctypes.data, dptr.ctypes.data))
torch.cuda.synchronize()
assert np.array_equal(dst.cpu().numpy(), np.concatenate(segs))
a, b = rng.uniform(0, 1, 300000), np.empty(300000)
_lib.check(_lib.load().b2_host_memcpy(b.ctypes.data, a.ctypes.data, a.nbytes))
assert np.array_equal(a, b)
print("upload / host copy ok")
