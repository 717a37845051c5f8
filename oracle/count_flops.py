"""Run oracle/_build/count_flops on the C2 workload's optics inputs: instrumented FP64 operation
count per photon of the reference arithmetic (the oracle restatement).  Test infrastructure.

    python oracle/count_flops.py            # prints the JSON line; tests/test_flop_count.py pins it
"""
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:] = [q for q in sys.path if os.path.abspath(q or ".") != HERE]  # "oracle" must be the package
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def build():
    exe = os.path.join(HERE, "_build", "count_flops")
    src = os.path.join(HERE, "count_flops.cpp")
    deps = [src, os.path.join(HERE, "oracle_optics.c"), os.path.join(ROOT, "include", "imsim_b200.h")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        os.makedirs(os.path.dirname(exe), exist_ok=True)
        subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-fpermissive", "-w", "-o", exe, src, "-lm"])
    return exe


def count(n=4000, diffraction=True, refraction=True):
    import helpers
    from imsim_b200 import _abi

    su = helpers.oracle_setup()
    dif = helpers.default_diffraction(enabled=diffraction)
    opt = _abi.B2OpticsOptions()
    if refraction:
        opt.do_refraction, opt.index_ratio = 1, 3.9
    rng = np.random.default_rng(0)
    p = helpers.test_photon_arrays(n=n, center=(2000.0, 1900.0))
    p["time"] = rng.uniform(0, 30, n)
    p["wavelength"] = rng.uniform(550, 690, n)
    gauss = rng.standard_normal(n)
    tel, _ = su.telescope.flatten()
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        for pod in (tel, su.img_wcs.to_pod(), su.icrf_to_field.to_pod(), su.detector.to_pod(), dif, opt):
            f.write(bytes(pod))
        f.write(np.int64(n).tobytes())
        for a in (p["x"], p["y"], p["flux"], p["wavelength"], p["pupil_u"], p["pupil_v"], p["time"], gauss,
                  np.zeros(n)):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
        name = f.name
    try:
        out = subprocess.check_output([build(), name], text=True)
    finally:
        os.unlink(name)
    return json.loads(out)


if __name__ == "__main__":
    print(json.dumps(count()))
