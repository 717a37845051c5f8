"""Shared fixtures for the tests: golden data, oracle set-ups, synthetic inputs."""
import os

import numpy as np

from imsim_b200 import _abi
from imsim_b200.detector import lsstcam_like
from imsim_b200.diffraction import diffraction_config
from imsim_b200.sensor import calculate_diff_step, synthetic_absorption_table
from imsim_b200.synthetic import chief_ray_inputs, make_detector_setup
from imsim_b200.treerings import RadialTable, TreeRingRadialFunction

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RUBIN_LAT = np.radians(-30.24463)


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def sensor_model(name="lsst_itl_50_4"):
    """(config dict, vertex_data) from the golden copy of data/sensor_models."""
    g = golden("sensor_models.npz")
    keys = [str(k) for k in g["cfg_keys"]]
    vals = g[name + "_cfg"]
    cfg = {}
    for k, v in zip(keys, vals):
        cfg[k] = int(v) if k in ("NumVertices", "PixelBoundaryNx", "PixelBoundaryNy", "NumPhases",
                                 "CollectingPhases") else float(v)
    return cfg, np.ascontiguousarray(g[name + "_dat"])


def tree_ring_block(det="R22_S11", fn="tree_ring_parameters_2026-04-02.txt"):
    g = golden("tree_rings.npz")
    return [str(s) for s in g["%s|%s" % (fn, det)]]


def tree_ring_table(det="R22_S11", fn="tree_ring_parameters_2026-04-02.txt"):
    block = tree_ring_block(det, fn)
    items = block[1].split()
    center = (float(items[4]) + 2048.5, float(items[5]) + 2048.5)
    func = RadialTable.from_func(TreeRingRadialFunction(block), 0.0, 8000.0, int(8000.0 / 3.0) + 1)
    return center, func


def sensor_pod(cfg, strength=1.0, nrecalc=10000, qdist=3, diffusion_factor=1.0, treering=None, n_abs=0):
    pod = _abi.B2SensorConfig()
    pod.num_vertices = cfg["NumVertices"]
    pod.nx, pod.ny = cfg["PixelBoundaryNx"], cfg["PixelBoundaryNy"]
    pod.qdist = qdist
    pod.num_elec = cfg["CollectedCharge_0_0"] / strength
    pod.nrecalc = nrecalc / strength
    pod.diff_step = calculate_diff_step(cfg) * diffusion_factor
    pod.pixel_size = cfg["PixelSizeX"]
    pod.sensor_thickness = cfg["SensorThickness"]
    if treering is not None:
        (cx, cy), func = treering
        pod.treering_center[0], pod.treering_center[1] = cx, cy
        pod.n_treering = len(func.x)
    pod.n_abs = n_abs
    return pod


def oracle_tracer():
    """Chief-ray tracer backed by the CPU oracle (tests only)."""
    from oracle import oracle as orc

    def trace(tel, thx, thy, wl):
        bt, extras = tel.flatten()
        x, y, z, vx, vy, vz, t, w = chief_ray_inputs(thx, thy, wl, tel.in_medium)
        out = orc.trace_rays(bt, extras, x, y, z, vx, vy, vz, t, w)
        return out[0], out[1]

    return trace


_SETUP_CACHE = {}


def oracle_setup(det_name="R22_S11", rot_tel_pos=np.radians(60.0), **kw):
    key = (det_name, rot_tel_pos, tuple(sorted(kw.items())))
    if key not in _SETUP_CACHE:
        _SETUP_CACHE[key] = make_detector_setup(oracle_tracer(), det_name=det_name, rot_tel_pos=rot_tel_pos, **kw)
    return _SETUP_CACHE[key]


def test_photon_arrays(n=10000, t=0.0, seed=42, wavelength=577.6, center=(0.0, 0.0)):
    """The reference's create_test_photon_array (tests/test_photon_ops.py:45-66):
    numpy seed 42, r_uv in U(2.5, 4.2), r_xy in U(0, 5) px, one wavelength, flux 1."""
    rng = np.random.default_rng(seed=seed)
    r_uv = rng.uniform(2.5, 4.2, n)
    phi_uv = rng.uniform(0.0, 2.0 * np.pi, n)
    u = r_uv * np.cos(phi_uv)
    v = r_uv * np.sin(phi_uv)
    r_xy = rng.uniform(0.0, 5.0, n)
    phi_xy = rng.uniform(0.0, 2.0 * np.pi, n)
    x = r_xy * np.cos(phi_xy) + center[0]
    y = r_xy * np.sin(phi_xy) + center[1]
    return dict(x=x, y=y, wavelength=np.full(n, wavelength), flux=np.ones(n), pupil_u=u, pupil_v=v,
                time=np.full(n, t))


def default_diffraction(enabled=True, field_rotation=True):
    return diffraction_config(latitude=RUBIN_LAT, altitude=np.radians(67.0), azimuth=np.radians(213.0),
                              disable_field_rotation=not field_rotation, enabled=enabled)


def absorption():
    return synthetic_absorption_table()
