"""Generate the golden fixtures under tests/golden/ from the reference tree.

Run in the build container (needs /root/reference):
    python tests/golden/make_golden.py

What is pinned:
  * diffraction.npz  -- inputs and outputs of the reference's own imsim/diffraction.py
    (pure numpy, loaded standalone with importlib): apply_diffraction_delta and
    apply_diffraction_delta_field_rot with injected Gaussian draws, directed_dist,
    field_rotation_matrix, e_equatorial.
  * sensor_models.npz -- the .cfg scalars and .dat vertex tables of
    data/sensor_models/lsst_{itl,e2v}_50_{4,8}  (data, not code).
  * tree_rings.npz   -- the parameter blocks of R22_S11 and R34_S22 from
    data/tree_ring_data/tree_ring_parameters_19mar18.txt and of R22_S11 from the
    default 2026-04-02 file, plus the known answers of tests/test_tree_rings.py:16-38.
"""
import importlib.util
import os
import sys

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))


def load_ref_diffraction():
    spec = importlib.util.spec_from_file_location("ref_diffraction", os.path.join(REF, "imsim", "diffraction.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_diffraction"] = mod
    spec.loader.exec_module(mod)
    return mod


def make_diffraction():
    d = load_ref_diffraction()
    rng = np.random.default_rng(20261017)
    n = 4000
    r = rng.uniform(2.3, 4.4, n)  # includes points beyond both circles
    ph = rng.uniform(0, 2 * np.pi, n)
    pos = np.c_[r * np.cos(ph), r * np.sin(ph)]
    # make some points hug the spider vanes
    pos[:200, 1] = -pos[:200, 0] + 0.4 * np.sqrt(2) + rng.normal(0, 0.03, 200)
    th = rng.normal(0, 0.02, (n, 2))
    g = 1 / np.sqrt(1 + (th**2).sum(1))
    nair = 1.000185
    v = np.c_[th[:, 0] * g, th[:, 1] * g, -g] / nair
    wl = rng.uniform(320e-9, 1050e-9, n)
    t = rng.uniform(0, 30, n)
    gauss = rng.standard_normal(n)

    def distribution(phi_star):
        # what RubinDiffraction.diffraction_rng does with GalSim's generate_from_variance:
        # N(0,1) * sqrt(phi^2)   (imsim/photon_ops.py:264-272)
        return gauss * np.sqrt(phi_star**2)

    lat, alt, az = np.radians(-30.24463), np.radians(67.0), np.radians(213.0)
    v_norot = d.apply_diffraction_delta(pos.copy(), v.copy(), wl, d.RUBIN_SPIDER_GEOMETRY, distribution)
    frm = d.prepare_field_rotation_matrix(latitude=lat, azimuth=az, altitude=alt)
    v_rot = d.apply_diffraction_delta_field_rot(pos.copy(), v.copy(), t, wl, frm, d.RUBIN_SPIDER_GEOMETRY,
                                                distribution)
    dist, nrm = d.directed_dist(d.RUBIN_SPIDER_GEOMETRY, pos.copy())
    rotm = frm(t)
    np.savez_compressed(os.path.join(HERE, "diffraction.npz"), pos=pos, v=v, wl=wl, t=t, gauss=gauss, lat=lat,
                        alt=alt, az=az, v_norot=v_norot, v_rot=v_rot, dist=dist, nrm=nrm, rotm=rotm,
                        e_equatorial=d.e_equatorial(latitude=lat, altitude=alt, azimuth=az),
                        lines=d.RUBIN_SPIDER_GEOMETRY.thick_lines, circles=d.RUBIN_SPIDER_GEOMETRY.circles,
                        omega=d.OMEGA_EARTH)
    print("diffraction.npz written")


def make_sensor_models():
    from imsim_b200.sensor import read_config_file

    out = {}
    keys = ['NumVertices', 'PixelBoundaryNx', 'PixelBoundaryNy', 'CollectedCharge_0_0', 'PixelSizeX',
            'SensorThickness', 'NumPhases', 'CollectingPhases', 'ChannelStopWidth', 'FieldOxideTaper', 'Vbb',
            'Vparallel_lo', 'Vparallel_hi', 'CCDTemperature', 'qfh']
    for name in ("lsst_itl_50_4", "lsst_e2v_50_4", "lsst_itl_50_8", "lsst_e2v_50_8"):
        base = os.path.join(REF, "data", "sensor_models", name)
        cfg = read_config_file(base + ".cfg")
        out[name + "_cfg"] = np.array([float(cfg[k]) for k in keys])
        out[name + "_dat"] = np.loadtxt(base + ".dat", skiprows=1).astype(np.float64)
    out["cfg_keys"] = np.array(keys)
    np.savez_compressed(os.path.join(HERE, "..", "..", "imsim_b200", "data", "sensor_models.npz"), **out)
    print("sensor_models.npz written")


def make_tree_rings():
    out = {}
    for fn, dets in (("tree_ring_parameters_19mar18.txt", ("R22_S11", "R34_S22")),
                     ("tree_ring_parameters_2026-04-02.txt", ("R22_S11", "R01_S00"))):
        with open(os.path.join(REF, "data", "tree_ring_data", fn)) as f:
            lines = f.readlines()
        for ib in range(len(lines) // 23):
            block = lines[ib * 23:(ib + 1) * 23]
            items = block[1].split()
            det = "R%s%s_S%s%s" % tuple(items[:4])
            if det in dets:
                out["%s|%s" % (fn, det)] = np.array(block)
    # tests/test_tree_rings.py:16-21
    out["known_r"] = np.array(5280.0)
    out["known_values"] = np.array([.0030205, -.0034135])
    out["known_centers"] = np.array([(-3026.3, -3001.0), (3095.5, -2971.3)])
    np.savez_compressed(os.path.join(HERE, "..", "..", "imsim_b200", "data", "tree_rings.npz"), **out)
    print("tree_rings.npz written")


if __name__ == "__main__":
    make_diffraction()
    make_sensor_models()
    make_tree_rings()
