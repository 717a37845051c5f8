#!/usr/bin/env python
"""bench.py -- photons/s through optics + silicon sensor on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # B200 arm (default N=1)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm (oracle port, all host threads)

Workload (config.workload = "C2-pooled"): one LSSTCam e2v science CCD (R22_S11, 4096 x 4004),
SiliconSensor lsst_e2v_50_4 with brighter-fatter (strength 1) and tree rings on, bright-star
dominated photon pool.  A step is one photon batch of the pooled pipeline
(imsim/photon_pooling.py:141-160): TimeSampler + PupilAnnulusSampler ->
RubinDiffractionOptics + FocusDepth + Refraction -> SiliconSensor.accumulate(resume, recalc=True),
i.e. pixel boundaries are recomputed from the accumulated charge at every step (nrecalc = 0,
config/imsim-config-photon-pooling.yaml:33).

`value`  : photons/s with the pool already resident in HBM (CUDA events, max over ranks).
`e2e`    : same chain through the host-facing API: pinned host arrays -> H2D -> kernels -> D2H of
           the image and the deposited flux, every step.
`roofline`: dominant kernel (k_rubin_optics) against HBM (algorithmic 96 B/photon) and, because
           that kernel is FP64-pipe bound, against the measured FP64 FMA ceiling (`roofline_fp64`).
`visit`  : the other half of BASELINE.json's metric, LSSTCam visits per hour on these N GPUs: the full chain
           of a synthetic 189-CCD visit sharded by detector over the ranks (`visit_for_line`; --no-visit-line
           skips it; `--visit ...` runs other variants of it alone).
With N > 1 every rank simulates its own detector (weak scaling, no data-path collective).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "photons/sec/GPU (optics+sensor)"
UNIT = "photons/s"
POOL = 1 << 25  # photons per step (33.5 M): every SoA array is 268 MB > L2 (126 MB)
ALG_BYTES_TRACE = 96.0  # read x,y,wl,u,v,t,flux (56 B) + write x,y,dxdz,dydz,flux (40 B)
# FP64 operations per photon of the reference arithmetic, INSTRUMENTED: oracle/count_flops.py compiles the
# oracle with a counting scalar (add 1308 + mul 2012 + div 323 + sqrt 121 per photon for RubinDiffractionOptics
# + Refraction; 15 transcendental calls not counted); pinned by tests/test_flop_count.py
ALG_FLOP_TRACE = 3765.0
# fused k_pool_step = trace + sensor fast path: SURVEY 8d adds 56 B (6 reads + 1 atomic RMW) and ~60 FLOP
ALG_BYTES_POOL = 96.0 + 56.0
ALG_FLOP_POOL = ALG_FLOP_TRACE + 60.0
# DRAM traffic of k_pool_step<4, LSST program> from `ncu --set full` (profiles/r01_k_pool_step_program_details.csv):
# (1.173 GB read + 0.032 GB written) / 33554432 photons
NCU_TRAFFIC_BYTES_PER_PHOTON_POOL = 35.9
DETECTORS = ["R22_S11", "R21_S11", "R23_S11", "R12_S11", "R32_S11", "R22_S00", "R22_S22", "R11_S11"]


def dist_info():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


class ClockSampler:
    """nvidia-smi clocks / throttle reasons.  Started before the warm-up (nvidia-smi needs ~0.3 s to
    produce its first line) and sampled every 20 ms; the summary uses the samples that fall inside
    the timed window, or, if the window is shorter than the sampling period, the samples taken while
    the GPU was under load (warm-up + timed region)."""

    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.samples = []  # (wall time, fields)
        self.proc = None
        self.window = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.samples and time.time() - t0 < 3.0:  # wait for the first line
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def mark(self, t_begin, t_end):
        self.window = (t_begin, t_end)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            pass
        rows = []
        for ts, s in self.samples:
            f = [t.strip() for t in s.split(",")]
            if len(f) < 10:
                continue
            try:
                rows.append((ts, float(f[1]), float(f[2]), float(f[3]), f[5:9], float(f[9])))
            except ValueError:
                continue
        sel, how = [], "timed window"
        if self.window:
            sel = [r for r in rows if self.window[0] - 0.02 <= r[0] <= self.window[1] + 0.02]
        if not sel:
            sel, how = [r for r in rows if r[5] > 50.0], "under load (warm-up + timed region)"
        if not sel:
            sel, how = rows, "all samples"
        reasons = set()
        for r in sel:
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median([r[1] for r in sel])) if sel else None,
                "sm_max_mhz": max(r[2] for r in sel) if sel else None,
                "power_w_max": max(r[3] for r in sel) if sel else None, "reasons": sorted(reasons),
                "samples": len(sel), "selection": how}


def sensor_inputs():
    from imsim_b200 import workload_data as wd

    cfg, dat = wd.sensor_model("lsst_e2v_50_4")
    tr = wd.tree_ring_table("R22_S11")
    aw, al = wd.absorption()
    return cfg, dat, tr, (aw, al)


def cpu_baseline(n_sample, threads, det_name="R22_S11", seed=0):
    """Time the CPU oracle (restatement of the reference path, kind='port') on a bounded sample
    of the same workload: optics (+diffraction, FocusDepth, Refraction) then sensor (BF + tree rings)."""
    import helpers  # tests/helpers.py: oracle set-up (this leg is the one place bench.py may run oracle/)
    from imsim_b200 import _abi
    from imsim_b200.synthetic import make_detector_setup, synthetic_photons
    from oracle import oracle as orc

    orc.set_threads(threads)
    su = make_detector_setup(helpers.oracle_tracer(), det_name, rot_tel_pos=np.radians(60.0))
    dif = helpers.default_diffraction()
    cfg, dat, tr, (aw, al) = sensor_inputs()
    rng = np.random.default_rng(seed)
    x, y, wl, flux = synthetic_photons(n_sample, kind="stars", seed=seed)
    u_r = np.sqrt(rng.uniform(2.558**2, 4.18**2, n_sample))
    ph = rng.uniform(0, 2 * np.pi, n_sample)
    pu, pv = u_r * np.cos(ph), u_r * np.sin(ph)
    t = rng.uniform(0, 30, n_sample)
    gauss = rng.standard_normal(n_sample)
    rand4 = np.vstack([rng.standard_normal(n_sample), rng.standard_normal(n_sample), rng.uniform(size=n_sample),
                       rng.uniform(size=n_sample)])
    opt = _abi.B2OpticsOptions()
    opt.do_refraction, opt.index_ratio = 1, 3.9
    pod = helpers.sensor_pod(cfg, nrecalc=0, treering=tr, n_abs=len(aw))
    sens = orc.Sensor(pod, dat, tr[1].x, tr[1].f, True, aw, al)
    img = np.zeros((4004, 4096), np.float32)
    sens.bind_image(img, 0, 0)
    tel = su.telescope.flatten()
    # untimed: boundary initialisation of the full CCD (set-up, done once per image)
    sens.accumulate(x[:10], y[:10], flux[:10], rand4[:, :10].copy())
    t0 = time.perf_counter()
    out = orc.rubin_optics(*tel, su.img_wcs.to_pod(), su.icrf_to_field.to_pod(), su.detector.to_pod(), dif, opt, x, y,
                           flux, wl, pu, pv, t, gauss)
    t1 = time.perf_counter()
    sens.accumulate(out["x"], out["y"], out["flux"], rand4, dxdz=out["dxdz"], dydz=out["dydz"], wavelength=wl,
                    resume=True, recalc=False)
    t2 = time.perf_counter()
    return {"value": n_sample / (t2 - t0), "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d photons of the C2-pooled workload: oracle optics %.2f s + sensor %.2f s (boundary "
                      "recalculation of the full CCD excluded)" % (n_sample, t1 - t0, t2 - t1),
            "optics_photons_per_s": n_sample / (t1 - t0), "sensor_photons_per_s": n_sample / (t2 - t1)}


def run_reference(args):
    """--impl reference: the CPU arm.  The reference itself (imSim over GalSim/batoid C++) cannot be
    installed offline (no wheels for galsim/batoid/lsst.* in /opt/wheelhouse), so
    this times the oracle restatement of its path with every host thread."""
    rank, world, _ = dist_info()
    if rank != 0:
        return 0
    from oracle import oracle as orc

    # every host thread this process may run on: torchrun exports OMP_NUM_THREADS=1 to its workers, which would
    # silently turn the "all host threads" arm into a single-thread one for N > 1
    try:
        threads = len(os.sched_getaffinity(0))
    except AttributeError:
        threads = os.cpu_count() or 1
    threads = max(threads, orc.max_threads())
    n = args.ref_sample
    vals = []
    for _ in range(args.warmup):
        cpu_baseline(min(n, 200000), threads)
    for k in range(args.steps):
        vals.append(cpu_baseline(n, threads, seed=k))
    v = float(np.mean([c["value"] for c in vals]))
    ms = n / v * 1e3
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(n, note="bounded sample per step on host cores"),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": vals[-1]["sample"]},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def simulate_visit(opts, rank, world, local, barrier=None):
    """This rank's share of one synthetic LSSTCam visit (C5): the detectors LPT assigns to it, software-pipelined
    (visit.DetectorRunner.run_many), simulated ``opts.visit_repeat`` times.  No collective in here unless
    ``barrier`` is given (called before each repetition).  Returns (per-repetition wall times, records of the last
    repetition)."""
    import torch

    from imsim_b200 import workload_data as helpers
    from imsim_b200.detector import lsstcam_science_detectors
    from imsim_b200.flat import wavelength_cdf
    from imsim_b200.sharding import lpt_partition
    from imsim_b200.visit import DetectorRunner, synthetic_catalog, synthetic_objects

    dets = lsstcam_science_detectors()[: opts.visit_ccds]
    rng = np.random.default_rng(5)
    costs = {d: float(opts.visit_photons * rng.lognormal(0.0, 0.3)) for d in dets}  # bright-star CCDs cost more
    mine = lpt_partition(costs, world)[rank]
    models = {"e2v": helpers.sensor_model("lsst_e2v_50_4"), "itl": helpers.sensor_model("lsst_itl_50_4")}
    tr = helpers.tree_ring_table("R22_S11")
    psf = None
    wave = np.linspace(550.0, 690.0, 29)
    cdf = wavelength_cdf(wave, np.ones_like(wave))
    if opts.visit_catalog:
        # stage 1 from catalogue rows: one atmosphere realisation per visit (6 screens of 8192^2 at 0.1 m), 8 SEDs
        from imsim_b200.atmosphere import AtmosphericPSF

        psf = AtmosphericPSF(1.2, 0.7, "r", rng=271828, device="cuda:%d" % local)
        seds = [wavelength_cdf(wave, 1.0 + 0.8 * np.sin(wave / (15.0 + 5 * k))) for k in range(8)]
        cdf = (np.array([c for c, _ in seds]), np.array([w for _, w in seds]))
    # lanes: independent DetectorRunners on their own streams, driven by one host thread each, so that one
    # detector's FP64-bound ray trace shares the SMs with another's boundary update and memory-bound kernels
    lanes = 1 if opts.visit_serial else max(1, int(getattr(opts, "visit_lanes", 0) or os.environ.get("B2_VISIT_LANES", "1")))
    lanes = min(lanes, max(1, len(mine)))
    if lanes == 1:
        runner = DetectorRunner(local, models, helpers.absorption(), tree_rings={d: tr for d in mine}, psf=psf)
    else:
        from imsim_b200.visit import VisitLanes

        visit_lanes = VisitLanes(local, lanes, models, helpers.absorption(), tree_rings={d: tr for d in mine}, psf=psf)

    def job(d):
        make = synthetic_catalog if opts.visit_catalog else synthetic_objects
        i = dets.index(d)
        # the catalogue is built inside prepare(), i.e. while the previous detector's kernels run
        return dict(det_name=d, objects=lambda: make(20000, 4096, 4004, seed=i, total_photons=costs[d]), nbatch=10,
                    wavelength_cdf=cdf, det_index=i, readout=opts.visit_readout, sky_level=opts.visit_sky)

    # The visit is simulated visit_repeat times by the same process and the LAST one is reported: the first
    # carries the one-time initialisation of a process (2.5 GB of boundary arrays per vendor model, the packed
    # phase screens, pinned result buffers, the caching allocator's first cudaMallocs), which a production run
    # pays once for all its visits, as it does the generation of the atmosphere.  first_visit_wall_s keeps it.
    walls, recs = [], []
    for rep in range(max(1, opts.visit_repeat)):
        if barrier is not None:
            barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if opts.visit_serial:  # one detector at a time, host and device in turn (the pre-pipelining behaviour)
            recs = [runner.run(**job(d))[0] for d in mine]
        elif lanes == 1:
            recs = [rec for rec, _ in runner.run_many(job(d) for d in mine)]
        else:
            recs = visit_lanes.run([job(d) for d in mine], cost=lambda j: costs[j["det_name"]])
        torch.cuda.synchronize()
        walls.append(time.perf_counter() - t0)
    return walls, recs


def run_visit(args):
    """C5: one synthetic LSSTCam visit, sharded by detector (weak scaling unit = CCD)."""
    import torch

    from imsim_b200.sharding import gather_visit_metadata

    rank, world, local = dist_info()
    barrier = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        barrier = dist.barrier
    torch.cuda.set_device(local)
    walls, recs = simulate_visit(args, rank, world, local, barrier=barrier)
    wall = walls[-1]
    wt = torch.tensor([wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(wt, op=dist.ReduceOp.MAX)
    allrec = gather_visit_metadata(recs)
    if rank == 0:
        photons = sum(r["photons"] for r in allrec)
        gpu_s = max(sum(r["gpu_ms"] for r in allrec if r["device"] == g) for g in range(world)) * 1e-3
        print(json.dumps({"mode": "visit", "stage1": "catalogue+atmosphere" if args.visit_catalog else "gaussian-points",
                          "readout_on_device": bool(args.visit_readout), "sky_level": args.visit_sky,
                          "n_gpus": world, "ccds": len(allrec), "photons": photons,
                          "wall_s_max_rank": float(wt.item()), "gpu_s_max_rank": gpu_s,
                          "visits_per_hour_wall": 3600.0 / float(wt.item()),
                          "visits_per_hour_gpu_time": 3600.0 / gpu_s,
                          "photons_per_s_wall": photons / float(wt.item()),
                          "setup_s_total": sum(r["setup_ms"] for r in allrec) * 1e-3,
                          "pipelined": not args.visit_serial, "visits_simulated": len(walls),
                          "first_visit_wall_s_rank0": walls[0],
                          "note": "synthetic 20k-object field per CCD generated on device; host work per CCD = WCS "
                                  "fit + object batching (Python) + 66 MB image readback"}))
    if world > 1:
        dist.destroy_process_group()
    return 0


def visit_for_line(args, rank, world, local, dev):
    """The second half of BASELINE.json's metric -- LSSTCam visits per hour on these GPUs -- for the bench line:
    the full chain (catalogue objects behind the atmosphere -> optics -> silicon -> sky -> electronics readout) of a
    synthetic 189-CCD visit, sharded by detector over the ranks, second visit of the process.  Ranks do not
    synchronise inside the simulation, and a rank that fails still takes part in the two reductions below."""
    import torch

    opts = argparse.Namespace(visit_catalog=True, visit_readout=True, visit_sky=800.0, visit_ccds=args.visit_ccds,
                              visit_photons=args.visit_photons, visit_repeat=2, visit_serial=False,
                              visit_lanes=getattr(args, "visit_lanes", 0) or int(os.environ.get("B2_VISIT_LANES", "2")))
    mx = [0.0, 0.0, 0.0, 0.0]  # wall of the reported visit, wall of the first, GPU time, failure flag
    sm = [0.0, 0.0]            # photons, CCDs
    err = ""
    try:
        walls, recs = simulate_visit(opts, rank, world, local)
        mx = [walls[-1], walls[0], sum(r["gpu_ms"] for r in recs) * 1e-3, 0.0]
        sm = [float(sum(r["photons"] for r in recs)), float(len(recs))]
    except Exception as e:  # the headline line must survive a failure of the extra measurement
        mx[3], err = 1.0, "%s: %s" % (type(e).__name__, e)
    tmx = torch.tensor(mx, dtype=torch.float64, device=dev)
    tsm = torch.tensor(sm, dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist

        dist.all_reduce(tmx, op=dist.ReduceOp.MAX)
        dist.all_reduce(tsm, op=dist.ReduceOp.SUM)
    mx, sm = tmx.tolist(), tsm.tolist()
    if mx[3] != 0.0 or mx[0] <= 0.0:
        return {"error": err or "a rank failed"}
    if int(sm[1]) != 189:
        err = "partial visit (%d of 189 CCDs): visits_per_hour is per this many CCDs" % int(sm[1])
    return {"visits_per_hour": 3600.0 / mx[0], "note": err, "wall_s_max_rank": mx[0], "gpu_s_max_rank": mx[2],
            "first_visit_wall_s_max_rank": mx[1], "ccds": int(sm[1]), "photons": int(sm[0]), "n_gpus": world,
            "visits_simulated": 2, "reported": "second visit of the process (the first pays the one-time allocations)",
            "chain": "catalogue objects (stars, bulge / disc / knots galaxies, 8 SEDs) -> six-screen atmosphere + second "
                     "kick -> RubinDiffractionOptics -> SiliconSensor (brighter-fatter + tree rings, 10 batches) -> "
                     "sky through the pixel areas -> bleed / dark / CTI / noise -> int32 segments, e-image and raw "
                     "segments to pinned host buffers; 189 CCDs by LPT over the ranks, no collective on the path",
            "lanes_per_gpu": int(opts.visit_lanes),
            "gpu_s_note": "gpu_s_max_rank sums the per-CCD device times; with more than one lane they overlap"}


def plugin_e2e(su, P, K, rank, hx, hy, hwl, hflux, n_obj=1000):
    """K pooled photon batches of P photons through ``B200PhotonPoolingImageBuilder.buildImage`` (the registered
    LSST_PhotonPoolingImage type), driven by the stand-in config engine of tests/stubs (GalSim itself cannot be
    installed here): the stamps hand over ordinary pageable numpy PhotonArrays, the builder gathers them into the
    HBM pool (pinned ring), runs one fused launch per batch and copies the image back to the host after every
    batch.  Wall clock around the call, device drained at the end."""
    import torch

    import pooled_config as pc
    from imsim_b200.photon_array import PhotonArray

    per = P // n_obj
    cat = [dict(x=0.0, y=0.0, flux=per * K) for _ in range(n_obj)]
    # pageable host arrays (np.array copies): what GalSim's shooters leave in a stamp's PhotonArray
    arrs = [PhotonArray(per, x=np.array(hx[k * per:(k + 1) * per]), y=np.array(hy[k * per:(k + 1) * per]),
                        flux=np.array(hflux[k * per:(k + 1) * per]), wavelength=np.array(hwl[k * per:(k + 1) * per]))
            for k in range(n_obj)]
    builder, cfg, base = pc.make_run(su, cat, lambda obj: arrs[obj.index], nbatch=K, nsubbatch=1, seed=11 + rank)

    class EveryBatch:  # a checkpointer that keeps nothing: makes the builder bring the image back after each batch
        file_name = "(memory)"
        saves = 0

    ck = EveryBatch()
    builder.setup(cfg, base, 0, 0, [], pc.Quiet())
    builder.checkpoint = ck
    builder.load_checkpoint = lambda *a, **k: (None, [], [], [], 0)

    def save(*a, **k):
        ck.saves += 1

    builder.save_checkpoint = save
    out = {}
    for rep in range(2):  # the first call pays the one-time allocations (pinned ring, 2.5 GB of sensor state)
        ck.saves = 0
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        image, _ = builder.buildImage(cfg, base, 0, 0, pc.Quiet())
        torch.cuda.synchronize()
        out = {"seconds": time.perf_counter() - t0}
    n = per * n_obj
    assert builder.last_route == "device", builder.last_route
    assert ck.saves >= K, ck.saves  # one per photon batch (+ one after the, here empty, FFT batch)
    if getattr(builder, "last_profile", None):
        out["host_phase_seconds"] = {k: round(v, 4) for k, v in builder.last_profile.items()}
    out.update(value=n * K / out["seconds"], photons_per_step=n, electrons=float(image.array.sum(dtype=np.float64)),
               h2d_bytes_per_step=int(builder.last_h2d_bytes // K + image.array.nbytes // K),
               d2h_bytes_per_step=int(image.array.nbytes), steps=K, route=builder.last_route,
               api="imsim_b200.galsim_plugin.B200PhotonPoolingImageBuilder.buildImage (registered as "
                   "LSST_PhotonPoolingImage), %d stamps per batch with pageable numpy PhotonArrays, stand-in config "
                   "engine (tests/stubs)" % n_obj)
    return out


def config_results(local, scale=1.0):
    """One short driver-timed number per BASELINE.json config (N = 1 only; the headline line is C2-pooled and
    `visit` is C5).  Each entry: what ran, how long, and the rate in the unit natural to it."""
    import torch

    from imsim_b200 import OpticsContext
    from imsim_b200 import workload_data as wd
    from imsim_b200.atmosphere import AtmosphericPSF
    from imsim_b200.detector import lsstcam_like
    from imsim_b200.flat import build_flat, flat_nrecalc, wavelength_cdf
    from imsim_b200.lsst_image import ClassicImageBuilder
    from imsim_b200.sensor import Image, SiliconSensor
    from imsim_b200.stage1 import ObjectTable
    from imsim_b200.synthetic import gpu_tracer, make_detector_setup
    from imsim_b200.visit import DetectorRunner, synthetic_catalog

    out = {}
    dev = "cuda:%d" % local
    cfg, dat = wd.sensor_model("lsst_e2v_50_4")
    tr = wd.tree_ring_table("R22_S11")
    wave = np.linspace(550.0, 690.0, 29)
    seds = [wavelength_cdf(wave, 1.0 + 0.8 * np.sin(wave / (15.0 + 5 * k))) for k in range(8)]
    sed_cdf, sed_wave = np.array([c for c, _ in seds]), np.array([w for _, w in seds])

    def timed(fn, reps=2):
        best = None
        for _ in range(reps):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            best = (dt, r) if best is None or dt < best[0] else best
        return best

    def guard(name, fn):
        try:
            out[name] = fn()
        except Exception as e:  # a config that fails must not take the headline line with it
            out[name] = {"error": "%s: %s" % (type(e).__name__, e)}

    ctx = OpticsContext(device=local, stream=torch.cuda.current_stream())
    det = lsstcam_like("R22_S11")
    su = make_detector_setup(gpu_tracer(ctx), "R22_S11", band="r", rot_tel_pos=np.radians(30.0), detector=det)
    ctx.set_telescope(su.telescope)
    ctx.set_wcs(su.img_wcs, su.icrf_to_field)
    ctx.set_detector(su.detector)
    ctx.set_diffraction(wd.default_diffraction())
    psf = AtmosphericPSF(1.2, 0.7, "r", rng=1, device=dev)

    def classic(rows, flux, radial, sersic_n, label):
        sensor = SiliconSensor(config=cfg, vertex_data=dat, nrecalc=10000, strength=1.0, rng=5, treering_func=tr[1],
                               treering_center=tr[0], absorption_table=wd.absorption(), context=ctx)
        b = ClassicImageBuilder(ctx, sensor, rows, radial, sersic_n, sed_cdf, sed_wave, psf=psf, seed=2)
        image = Image(np.zeros((det.ny, det.nx), np.float32), 0, 0)
        dt, st = timed(lambda: b.build(image, flux, phot_flux=flux.astype(np.int64)))
        sensor.close()
        return {"value": st["photons"] / dt, "unit": UNIT, "seconds": dt, "photons": int(st["photons"]),
                "objects": int(rows.size), "ms_per_object": 1e3 * dt / rows.size, "what": label}

    def c1():
        cat = synthetic_catalog(1998, det.nx, det.ny, seed=3, total_photons=5e7 * scale)
        rows, flux = cat.build()
        return classic(rows, flux, cat.radial_tables(), cat.sersic_n,
                       "classic per-object pipeline (LSST_Image / LSST_Silicon): catalogue the size of "
                       "examples/example_instance_catalog.txt (1998 entries -> 4410 rows: stars, bulge / disc / knots), "
                       "atmosphere + optics + SiliconSensor per stamp with nrecalc = 1e4, tree rings")

    def c2():
        rng = np.random.default_rng(4)
        n = int(1000 * scale) or 1
        tab = ObjectTable()
        tab.add_points(rng.uniform(100, det.nx - 100, n), rng.uniform(100, det.ny - 100, n), np.ones(n, int))
        rows, _ = tab.build()
        return classic(rows, np.full(n, 1.0e5), None, None,
                       "bright-star stamps: %d stars x 1e5 e-, brighter-fatter recomputed every 1e4 e- inside each stamp "
                       "(device-side cadence, csrc/stamps.cu), tree rings" % n)

    def c3():
        runner = DetectorRunner(local, {"e2v": (cfg, dat), "itl": wd.sensor_model("lsst_itl_50_4")}, wd.absorption(),
                                tree_rings={"R22_S11": tr}, psf=psf)
        nobj, total = int(1e5 * scale), 3e8 * scale
        job = dict(det_name="R22_S11", objects=lambda: synthetic_catalog(nobj, det.nx, det.ny, seed=9, total_photons=total),
                   nbatch=10, wavelength_cdf=(sed_cdf, sed_wave), det_index=0)
        dt, rec = timed(lambda: runner.run(**job)[0])
        return {"value": rec["photons"] / dt, "unit": UNIT, "seconds": dt, "photons": int(rec["photons"]),
                "objects": nobj, "gpu_ms": rec["gpu_ms"], "setup_ms": rec["setup_ms"],
                "what": "pooled dense field (LSST_PhotonPoolingImage cadence): %d catalogue objects on one CCD, stage 1 on "
                        "the device, 10 photon batches with the boundary recalculation at each batch start" % nobj}

    def c4_area():
        sensor = SiliconSensor(config=cfg, vertex_data=dat, rng=1, treering_func=tr[1], treering_center=tr[0],
                               absorption_table=wd.absorption())
        n, counts = 4096, 1.0e5 * min(scale, 1.0)
        img = Image(np.zeros((n, n), np.float32), 1, 1)
        dt, _ = timed(lambda: build_flat(img, counts, sensor, rng=2, nx=8, ny=2, fused=True), reps=1)
        sensor.close()
        return {"value": n * n * counts / dt, "unit": "electrons/s", "seconds": dt, "level_e_per_pixel": counts,
                "what": "examples/flat.yaml: 4096 x 4096 flat at 1e5 e-/pixel through calculate_pixel_areas + exact Poisson, "
                        "8 x 2 sections, 1000 e- per iteration"}

    def c4_photon():
        n, counts = 2048, 2000.0 * scale
        w = np.linspace(930.0, 960.0, 31)
        cdf = wavelength_cdf(w, 1.0 - np.abs(w - 945.0) / 15.0 + 1e-3)
        sensor = SiliconSensor(config=cfg, vertex_data=dat, rng=1, nrecalc=flat_nrecalc(n + 10, n + 10, 1, 1),
                               treering_func=tr[1], treering_center=tr[0], absorption_table=wd.absorption())
        img = Image(np.zeros((n, n), np.float32), 1, 1)
        build_flat(img, 1000.0, sensor, rng=1, max_counts_per_iter=1000, nx=1, ny=1, sed_cdf=cdf, fused=True)
        img = Image(np.zeros((n, n), np.float32), 1, 1)
        dt, nphot = timed(lambda: build_flat(img, counts, sensor, rng=2, max_counts_per_iter=1000, nx=1, ny=1,
                                             sed_cdf=cdf, fused=True), reps=1)
        sensor.close()
        return {"value": nphot / dt, "unit": UNIT, "seconds": dt, "photons": int(nphot),
                "what": "examples/flat_with_sed.yaml: photon-shot flat, %d x %d section, %.0f e-/pixel, y-band sed, photons "
                        "generated inside the deposit kernel (k_flat_step)" % (n, n, counts)}

    guard("C1", c1)
    guard("C2-stamps", c2)
    guard("C3", c3)
    guard("C4-area", c4_area)
    guard("C4-photon", c4_photon)
    return out


def e2e_line(pinned_value, h2d, d2h, steps, plug):
    """`e2e`: through the plugin's image builder when that leg ran (host numpy arrays in, host image out, every
    step); the pinned-buffer route of PhotonPool.run_host_batches is kept beside it."""
    pinned = {"value": pinned_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
              "steps": steps, "api": "PhotonPool.run_host_batches (caller fills pinned buffers)"}
    if plug and "value" in plug:
        return {"value": plug["value"], "unit": UNIT, "h2d_bytes_per_step": plug["h2d_bytes_per_step"],
                "d2h_bytes_per_step": plug["d2h_bytes_per_step"], "steps": plug["steps"], "api": plug["api"],
                "wall_s": plug["seconds"], "pinned_route": pinned}
    pinned["plugin_route"] = plug
    return pinned


def workload_config(pool, note=""):
    return {"workload": "C2-pooled: single e2v CCD R22_S11 4096x4004, SiliconSensor lsst_e2v_50_4 brighter-fatter "
                        "(strength 1, boundaries recomputed every step = photon batch) + tree rings, bright-star "
                        "dominated pool (1000 stars, 5 mag range, sigma 1.5 px), RubinDiffractionOptics + "
                        "FocusDepth + Refraction, r band",
            "photons_per_step": pool, "detector": "R22_S11 (+7 neighbours for N>1, one per rank)",
            "l2_policy": "inputs larger than L2 (each SoA array %d MB)" % (pool * 8 // 2**20),
            "note": (note + "; " if note else "") + "the ops work in place: before every timed step the inputs are restored "
                    "by a device-to-device copy (`refill`) that the CUDA events do not bracket"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pool", type=int, default=POOL)
    ap.add_argument("--ref-sample", type=int, default=2_000_000)
    ap.add_argument("--cpu-sample", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--unfused", action="store_true", help="time the three separate kernels instead of b2_pool_step")
    ap.add_argument("--visit", action="store_true",
                    help="extra mode (not the headline line): simulate a synthetic LSSTCam visit, 189 CCDs sharded "
                         "by detector over the ranks, --visit-photons per CCD; prints one JSON line")
    ap.add_argument("--visit-photons", type=float, default=1e8)
    ap.add_argument("--visit-sky", type=float, default=0.0,
                    help="with --visit: sky level [e-/pixel] added through the sensor's pixel areas with exact Poisson "
                         "noise on the device (a 30 s r-band dark sky is ~ 800)")
    ap.add_argument("--visit-readout", action="store_true",
                    help="with --visit: also run the electronics readout (bleed trails, dark current, crosstalk-free "
                         "amp split, CTI, bias, read noise -> int32 segments) on the device for every CCD")
    ap.add_argument("--visit-catalog", action="store_true",
                    help="with --visit: stage 1 from catalogue rows (stars + bulge/disc/knots galaxies, per-object SEDs) "
                         "behind the atmospheric PSF (6 phase screens + second kick) instead of Gaussian point sources")
    ap.add_argument("--visit-ccds", type=int, default=189)
    ap.add_argument("--visit-repeat", type=int, default=2,
                    help="with --visit: simulate the visit this many times in the process and report the last "
                         "(steady state; the first also pays the one-time allocations)")
    ap.add_argument("--visit-lanes", type=int, default=0,
                    help="detectors simulated concurrently per GPU on separate streams (default: B2_VISIT_LANES or 1)")
    ap.add_argument("--visit-serial", action="store_true",
                    help="with --visit: no software pipelining (prepare, launch and finish each detector in turn)")
    ap.add_argument("--no-configs", action="store_true",
                    help="skip the short per-config runs (C1, C2-stamps, C3, C4-area, C4-photon) added to the line at N = 1")
    ap.add_argument("--no-plugin-e2e", action="store_true",
                    help="skip the e2e leg through the plugin's LSST_PhotonPoolingImage builder (e2e is then the "
                         "pinned-buffer route)")
    ap.add_argument("--no-visit-line", action="store_true",
                    help="skip the synthetic full-chain visit that adds `visit` (visits per hour) to the bench line")
    ap.add_argument("--kernel-timing", action="store_true",
                    help="diagnostic: bracket every kernel with CUDA events (B2_TIMING=1) and print the breakdown "
                         "to stderr; adds event overhead, do not quote `value` from such a run")
    args = ap.parse_args()
    if args.kernel_timing:
        os.environ["B2_TIMING"] = "1"
    if args.impl == "reference":
        return run_reference(args)
    if args.visit:
        return run_visit(args)

    import torch

    from imsim_b200 import workload_data as helpers
    from imsim_b200 import OpticsContext, _abi, launch_count
    from imsim_b200.photon_pooling import DevicePhotons, PhotonPool, PinnedPhotons
    from imsim_b200.sensor import Image, SiliconSensor
    from imsim_b200.synthetic import gpu_tracer, make_detector_setup, synthetic_photons

    rank, world, local = dist_info()
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    W, K, P = max(args.warmup, 3), args.steps, args.pool

    # ---- per-rank detector set-up (what telescope_loader / batoid_wcs hand to the ops) ----
    det_name = DETECTORS[rank % len(DETECTORS)]
    stream = torch.cuda.current_stream()
    ctx = OpticsContext(device=local, stream=stream)
    su = make_detector_setup(gpu_tracer(ctx), det_name, rot_tel_pos=np.radians(60.0))
    ctx.set_telescope(su.telescope)
    ctx.set_wcs(su.img_wcs, su.icrf_to_field)
    ctx.set_detector(su.detector)
    ctx.set_diffraction(helpers.default_diffraction())
    cfg, dat, tr, abs_tab = sensor_inputs()
    sensor = SiliconSensor(config=cfg, vertex_data=dat, nrecalc=0, strength=1.0, rng=1234 + rank,
                           treering_func=tr[1], treering_center=tr[0], absorption_table=abs_tab, context=ctx)
    image = Image(np.zeros((su.detector.ny, su.detector.nx), np.float32), 0, 0)
    pool = PhotonPool(ctx, sensor, exptime=30.0, focus_depth=0.0, index_ratio=3.9, seed=99 + rank)

    # ---- synthetic pool: generated once on the host, resident in HBM before timing ----
    hx, hy, hwl, hflux = synthetic_photons(P, su.detector.nx, su.detector.ny, seed=rank, kind="stars")
    pinned = PinnedPhotons(P)
    pinned.x[:], pinned.y[:], pinned.wavelength[:], pinned.flux[:] = hx, hy, hwl, hflux
    src = DevicePhotons(P, device=dev)
    src.upload(pinned)
    torch.cuda.synchronize()
    work = [DevicePhotons(P, device=dev) for _ in range(2)]

    def refill(dp):
        for f in ("x", "y", "wavelength", "flux"):
            getattr(dp, f).copy_(getattr(src, f))

    # first call binds + initialises the image state (untimed set-up, once per image)
    refill(work[0])
    pool.process(work[0], image, resume=False, recalc=False)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps -------------------------------------------------------
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    clocks = ClockSampler(local)
    clocks.start()
    for i in range(W):
        refill(work[i % 2])
        pool.process(work[i % 2], image, resume=True, recalc=True)
    barrier()
    if args.kernel_timing:
        from imsim_b200._lib import timing_report

        timing_report()  # drop warm-up launches
    ctx.kernel_ms()
    ctx.record_kernel_events(True)
    l0 = launch_count()
    step_ms = []
    t_region0_wall = time.time()
    for i in range(K):
        dp = work[i % 2]
        refill(dp)  # untimed input restore (the ops work in place); events bracket only the path
        ev[i][0].record()
        if args.unfused:
            pool.ctx.sample_time_pupil(dp.time, dp.pupil_u, dp.pupil_v, pool.t0, pool.exptime, pool.r_inner,
                                       pool.r_outer, pool.seed, pool.offset)
            pool.opt.photon_offset = pool.offset
            ctx.rubin_optics(dp.x, dp.y, dp.dxdz, dp.dydz, dp.flux, dp.wavelength, dp.pupil_u, dp.pupil_v, dp.time,
                             options=pool.opt, want_stats=False)
            dp._has.update(pupil_u=True, pupil_v=True, time=True, dxdz=True, dydz=True)
            pool.offset += P
            sensor.accumulate(dp, image, resume=True, recalc=True, sync_image=False, want_stats=False)
        else:
            # one kernel per batch: sampler -> DCR/optics/FocusDepth/Refraction -> sensor deposit (b2_pool_step)
            pool.process(dp, image, resume=True, recalc=True, fused=True)
        ev[i][1].record()
    barrier()
    clocks.mark(t_region0_wall, time.time())
    launches = launch_count() - l0
    clk = clocks.stop()
    if args.kernel_timing and rank == 0:
        rep = timing_report()
        sys.stderr.write("per-step kernel ms: " + json.dumps({k: round(v[1] / K, 4) for k, v in rep.items()}) + "\n")
    step_ms = [a.elapsed_time(b) for a, b in ev]
    kern_total_ms, kern_count = ctx.kernel_ms()
    ctx.record_kernel_events(False)
    trace_ms = [kern_total_ms / max(kern_count, 1)]
    dominant = "k_rubin_optics" if args.unfused else "k_pool_step"
    total_ms = float(np.sum(step_ms))
    tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        lt = torch.tensor([launches], dtype=torch.int64, device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    total_ms_max = float(tt.item())
    value = world * P * K / (total_ms_max * 1e-3)

    # ---- end-to-end steps through the public host-facing API: pinned host batches -> H2D (double
    # buffered on a copy stream) -> b2_pool_step -> D2H of the image after every batch ----
    e2e_K = int(os.environ.get("B2_E2E_STEPS", "0")) or max(3, min(K, 20))  # (B2_E2E_STEPS: longer, for PCIe counter sampling)
    host_batches = [pinned] * e2e_K
    pool.run_host_batches(host_batches[:2], image, first_resume=True)  # warm the pipeline
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    h2d, d2h = pool.run_host_batches(host_batches, image, first_resume=True)
    e1.record()
    barrier()
    e_ms = float(e0.elapsed_time(e1))
    et = torch.tensor([e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
    e2e_value = world * P * e2e_K / (float(et.item()) * 1e-3)

    # ---- the same steps through the reference-facing plugin call: LSST_PhotonPoolingImage.buildImage of
    # imsim_b200.galsim_plugin, fed with ordinary (pageable) numpy PhotonArrays by the stamps, image back on the
    # host after every batch (the checkpoint cadence of imsim/photon_pooling.py:167-168) ----
    plug = None
    if not args.no_plugin_e2e:
        try:
            plug = plugin_e2e(su, P, e2e_K, rank, hx, hy, hwl, hflux)
        except Exception as e:  # the headline line must survive a failure of this leg
            plug = {"error": "%s: %s" % (type(e).__name__, e)}
        pt = torch.tensor([plug.get("seconds", 0.0), 1.0 if "error" in plug else 0.0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(pt, op=dist.ReduceOp.MAX)
        if pt[1].item() == 0.0 and pt[0].item() > 0.0:
            plug["value"] = world * P * e2e_K / float(pt[0].item())

    # ---- one short number per BASELINE config (rank 0 of a single-GPU run only) ----
    cfgs = None
    if world == 1 and not args.no_configs:
        try:
            cfgs = config_results(local)
        except Exception as e:
            cfgs = {"error": "%s: %s" % (type(e).__name__, e)}

    # ---- the other half of the metric: LSSTCam visits per hour on these GPUs (every rank takes part) ----
    visit = None if args.no_visit_line else visit_for_line(args, rank, world, local, dev)

    # ---- roofline of the dominant kernel + CPU baseline (rank 0 only) -------------------
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        fp64_peak = ctx.fma_peak(True)
        fp32_peak = ctx.fma_peak(False)
        tr_ms = float(np.mean(trace_ms))
        alg_bytes = ALG_BYTES_TRACE if args.unfused else ALG_BYTES_POOL
        alg_flop = ALG_FLOP_TRACE if args.unfused else ALG_FLOP_POOL
        ach_gbs = alg_bytes * P / (tr_ms * 1e-3) / 1e9
        ach_tf = alg_flop * P / (tr_ms * 1e-3) / 1e12
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": total_ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": workload_config(P),
            "e2e": e2e_line(e2e_value, h2d, d2h, e2e_K, plug),
            "gpu_launches": int(launches),
            "clocks": clk,
            "roofline": {"kernel": dominant, "bound": "hbm", "achieved": ach_gbs, "peak": hbm_peak,
                         "unit": "GB/s", "frac": ach_gbs / hbm_peak,
                         "traffic": None if args.unfused else NCU_TRAFFIC_BYTES_PER_PHOTON_POOL * P,
                         "traffic_source": "ncu --set full capture of the same kernel, per launch scaled to this pool "
                                           "size (profiles/r01_k_pool_step_program_details.csv)", "peak_source": peak_src,
                         "kernel_ms": tr_ms, "share_of_step": tr_ms * K / total_ms,
                         "note": "kernel is FP64-pipe bound, see roofline_fp64"},
            "roofline_fp64": {"kernel": dominant, "bound": "fp64 fma pipe", "achieved": ach_tf,
                              "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach_tf / fp64_peak,
                              "flop_per_photon_algorithmic": alg_flop, "bytes_per_photon_algorithmic": alg_bytes,
                              "peak_source": "b2_fma_peak measured in this run", "fp32_peak": fp32_peak},
            "breakdown_ms": {"step": total_ms / K, dominant: tr_ms, "rest_of_step": total_ms / K - tr_ms},
        }
        if visit is not None:
            line["visit"] = visit
        if cfgs is not None:
            line["configs"] = cfgs
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.cpu_sample, 1)
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
