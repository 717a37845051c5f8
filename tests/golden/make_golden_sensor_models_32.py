"""tests/golden/sensor_models_32.npz: the two 32-vertex sensor models of the reference's data directory
(data/sensor_models/lsst_{itl,e2v}_50_32.{cfg,dat}), in the layout of imsim_b200/data/sensor_models.npz
(float32 vertex tables would lose digits the files carry: kept as float64, 4 decimals as in the files).
Used by tests/test_oracle_golden.py::test_sensor_moment_differences_between_models for the 8 -> 32 vertex
differences of the reference's regression moments.  Run in the build container (reads /root/reference):

    python tests/golden/make_golden_sensor_models_32.py
"""
import os
import re

import numpy as np

SRC = "/root/reference/data/sensor_models"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sensor_models_32.npz")
KEYS = ["NumVertices", "PixelBoundaryNx", "PixelBoundaryNy", "CollectedCharge_0_0", "PixelSizeX", "SensorThickness",
        "NumPhases", "CollectingPhases", "ChannelStopWidth", "FieldOxideTaper", "Vbb", "Vparallel_lo", "Vparallel_hi",
        "CCDTemperature", "qfh"]


def read_cfg(path):
    cfg = {}
    for line in open(path):
        m = re.match(r"\s*(\w+)\s*=\s*([-+0-9.eE]+)\s*(#.*)?$", line)
        if m:
            cfg[m.group(1)] = float(m.group(2))
    return cfg


out = {"cfg_keys": np.array(KEYS)}
for name in ("lsst_itl_50_32", "lsst_e2v_50_32"):
    cfg = read_cfg(os.path.join(SRC, name + ".cfg"))
    out[name + "_cfg"] = np.array([cfg[k] for k in KEYS], dtype=np.float64)
    out[name + "_dat"] = np.loadtxt(os.path.join(SRC, name + ".dat"), skiprows=1)
np.savez_compressed(OUT, **out)
print(OUT, os.path.getsize(OUT), {k: v.shape for k, v in out.items()})
