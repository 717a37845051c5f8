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
    """Golden vectors of tests/golden/; the two copies of reference data files that the synthetic workloads also
    use live in the package (imsim_b200/data/)."""
    if name in ("sensor_models.npz", "tree_rings.npz"):
        from imsim_b200 import workload_data

        return workload_data.load(name)
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


from imsim_b200.workload_data import (absorption, default_diffraction, sensor_model, tree_ring_block,  # noqa: E402,F401
                                      tree_ring_table)


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




_package_sensor_model = sensor_model


def sensor_model(name="lsst_itl_50_4"):  # noqa: F811
    """The package's four sensor models, plus the reference's two 32-vertex ones from tests/golden."""
    if not name.endswith("_32"):
        return _package_sensor_model(name)
    g = golden("sensor_models_32.npz")
    ints = ("NumVertices", "PixelBoundaryNx", "PixelBoundaryNy", "NumPhases", "CollectingPhases")
    cfg = {str(k): (int(v) if str(k) in ints else float(v)) for k, v in zip(g["cfg_keys"], g[name + "_cfg"])}
    return cfg, np.ascontiguousarray(g[name + "_dat"])
