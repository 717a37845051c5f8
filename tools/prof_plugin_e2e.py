"""cProfile of the e2e leg through the plugin's LSST_PhotonPoolingImage builder (diagnostic)"""
import cProfile, os, pstats, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bench
from imsim_b200 import OpticsContext
from imsim_b200.synthetic import gpu_tracer, make_detector_setup, synthetic_photons
ctx = OpticsContext(device=0)
su = make_detector_setup(gpu_tracer(ctx), "R22_S11", rot_tel_pos=np.radians(60.0))
P = 1 << 25
hx, hy, hwl, hflux = synthetic_photons(P, su.detector.nx, su.detector.ny, seed=0, kind="stars")
if os.environ.get("B2_PLUGIN_PROFILE"):
    # the builder's own phase timers (no profiler distortion) arrive in out["host_phase_seconds"]
    out = bench.plugin_e2e(su, P, 4, 0, hx, hy, hwl, hflux)
    print({k: v for k, v in out.items() if k != "api"})
else:
    pr = cProfile.Profile()
    pr.enable()
    out = bench.plugin_e2e(su, P, 4, 0, hx, hy, hwl, hflux)
    pr.disable()
    print({k: v for k, v in out.items() if k != "api"})
    pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
