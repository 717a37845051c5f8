"""Input data of the synthetic workloads (bench.py, tools/, smoke, tests): copies of the reference's own data
files, which a real run reads from the installed imSim (``data/sensor_models/*.cfg|*.dat``,
``data/tree_ring_data/*.txt``) -- the GPU box has no reference tree.  ``imsim_b200/data/*.npz`` are written by
``tests/golden/make_golden.py``."""
from __future__ import annotations

import os

import numpy as np

from .diffraction import diffraction_config
from .sensor import synthetic_absorption_table
from .treerings import RadialTable, TreeRingRadialFunction

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data")
RUBIN_LAT = np.radians(-30.24463)
_INT_KEYS = ("NumVertices", "PixelBoundaryNx", "PixelBoundaryNy", "NumPhases", "CollectingPhases")


def load(name):
    return np.load(os.path.join(DATA, name), allow_pickle=False)


def sensor_model(name="lsst_itl_50_4"):
    """(config dict, vertex table) of ``data/sensor_models/<name>.{cfg,dat}``."""
    g = load("sensor_models.npz")
    cfg = {str(k): (int(v) if str(k) in _INT_KEYS else float(v)) for k, v in zip(g["cfg_keys"], g[name + "_cfg"])}
    return cfg, np.ascontiguousarray(g[name + "_dat"])


def tree_ring_block(det="R22_S11", fn="tree_ring_parameters_2026-04-02.txt"):
    """The detector's 23 text lines of a tree-ring parameter file."""
    return [str(s) for s in load("tree_rings.npz")["%s|%s" % (fn, det)]]


def tree_ring_table(det="R22_S11", fn="tree_ring_parameters_2026-04-02.txt"):
    """(centre, tabulated radial function) as ``TreeRings.get_center`` / ``get_func`` return them."""
    block = tree_ring_block(det, fn)
    tok = block[1].split()
    table = RadialTable.from_func(TreeRingRadialFunction(block), 0.0, 8000.0, int(8000.0 / 3.0) + 1)
    return (float(tok[4]) + 2048.5, float(tok[5]) + 2048.5), table


def absorption():
    """(wavelength [nm], absorption length [um]) -- SYNTHETIC stand-in for GalSim's ``absorption.dat``."""
    return synthetic_absorption_table()


def default_diffraction(enabled=True, field_rotation=True):
    return diffraction_config(latitude=RUBIN_LAT, altitude=np.radians(67.0), azimuth=np.radians(213.0),
                              disable_field_rotation=not field_rotation, enabled=enabled)
