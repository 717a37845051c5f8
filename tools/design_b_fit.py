"""SURVEY section 7 "Design B", measured: how well does a polynomial aberration map reproduce the FP64 trace?

For one CCD the exact trace (CPU oracle, no GPU needed) maps (direction cosines, pupil position, wavelength) of
random rays to (x, y, dxdz, dydz) at the sensor.  After removing nothing but what a polynomial can absorb itself,
least-squares polynomials in the scaled variables A, B (field over the CCD), U, V (pupil), W (wavelength) are
fitted on 150 000 rays and their maximum residual is measured on 200 000 other rays.  The term sets bound the
pupil degree p, the field degree f and the wavelength degree w (plus a mixed-degree budget).

Result (profiles/r02_design_b_fit.json): the position residual falls by about a factor of three per pupil degree --
6e-2 px at p = 8 (599 terms), 6e-3 px at p = 10 (915 terms), 1.3e-3 px at p = 11 (669 terms with a tighter mixed
budget) -- so the 1e-5 px of BASELINE.json's FP32 mode needs p ~ 15-16, i.e. well over 1500 terms per output and
four outputs: more FP32 multiply-adds per photon than the 1700 FP64 instructions of the exact trace cost in issue
slots.  The map is smooth; it is simply not low order: an f/1.2 beam through three aspheres and a corrector.

usage: python tools/design_b_fit.py [R22_S11 ...]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from imsim_b200.detector import lsstcam_like  # noqa: E402
from imsim_b200.telescope import lsst_v33  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def field_center(tel, det, wl=622e-9):
    cfx, cfy = det.center_focal()
    th, h = np.zeros(2), 1e-5

    def tr(thx, thy):
        g = 1 / np.sqrt(1 + thx**2 + thy**2)
        z = np.zeros_like(thx)
        out = orc.trace_rays(*tel.flatten(), z, z, z, thx * g, thy * g, -g, z, wl)
        return out[1] * 1e3, out[0] * 1e3

    for _ in range(8):
        fx, fy = tr(np.array([th[0], th[0] + h, th[0]]), np.array([th[1], th[1], th[1] + h]))
        r = np.array([fx[0] - cfx, fy[0] - cfy])
        J = np.array([[(fx[1] - fx[0]) / h, (fx[2] - fx[0]) / h], [(fy[1] - fy[0]) / h, (fy[2] - fy[0]) / h]])
        th = th - np.linalg.solve(J, r)
    return th


def rays(det_name, n, seed, half_deg=0.13, wl_range=(540.0, 700.0)):
    det = lsstcam_like(det_name)
    tel = lsst_v33("r", rot_tel_pos=np.radians(60.0), detector_z_offset=det.z_offset)
    th0 = field_center(tel, det)
    rng = np.random.default_rng(seed)
    A, B = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    half = np.radians(half_deg)
    thx, thy = th0[0] + A * half, th0[1] + B * half
    g = 1 / np.sqrt(1 + thx**2 + thy**2)
    al, be = thx * g, thy * g
    r, ph = np.sqrt(rng.uniform(2.55**2, 4.18**2, n)), rng.uniform(0, 2 * np.pi, n)
    u, v = r * np.cos(ph), r * np.sin(ph)
    wl = rng.uniform(*wl_range, n)
    z = np.zeros(n)
    out = orc.trace_rays(*tel.flatten(), u, v, z, al, be, -np.sqrt(1 - al**2 - be**2), z, wl * 1e-9)
    x, y, _, vx, vy, vz, _, vig, fail = out
    px, py = det.focal_to_pixel(y * 1e3, x * 1e3)
    J = det.jhat()
    ok = (vig == 0) & (fail == 0)
    return dict(A=A, B=B, U=u / 4.18, V=v / 4.18, W=(wl - 620.0) / 80.0, px=px, py=py,
                dxdz=(J[0, 0] * vx + J[0, 1] * vy) / vz, dydz=(J[1, 0] * vx + J[1, 1] * vy) / vz, ok=ok)


def terms(spec):
    return [(i, j, k, l, m) for k in range(14) for l in range(14 - k) for i in range(6) for j in range(6 - i)
            for m in range(4) if spec(k + l, i + j, m)]


def design(T, d, sel):
    pw = {n: [d[n][sel] ** q for q in range(15)] for n in "ABUVW"}
    M = np.empty((int(sel.sum()) if sel.dtype == bool else sel.size, len(T)))
    for c, (i, j, k, l, m) in enumerate(T):
        M[:, c] = pw["A"][i] * pw["B"][j] * pw["U"][k] * pw["V"][l] * pw["W"][m]
    return M


SPECS = {
    "pupil<=8 field<=2 wave<=2, sum<=9": lambda p, f, w: p <= 8 and f <= 2 and w <= 2 and p + f + w <= 9,
    "pupil<=9 field<=2 wave<=2, p+3f+3w<=11": lambda p, f, w: p <= 9 and f <= 2 and w <= 2 and p + 3 * f + 3 * w <= 11,
    "pupil<=10 field<=3 wave<=2, p+2f+2w<=12": lambda p, f, w: p <= 10 and f <= 3 and w <= 2 and p + 2 * f + 2 * w <= 12,
    "pupil<=11 field<=3 wave<=3, p+3f+3w<=13": lambda p, f, w: p <= 11 and f <= 3 and w <= 3 and p + 3 * f + 3 * w <= 13,
}

if __name__ == "__main__":
    dets = sys.argv[1:] or ["R22_S11"]
    res = {}
    for det in dets:
        tr, va = rays(det, 400000, 1), rays(det, 200000, 2)
        idx = np.nonzero(tr["ok"])[0][:150000]
        res[det] = {"unvignetted_fraction": float(va["ok"].mean()), "fits": []}
        for name, spec in SPECS.items():
            t0 = time.time()
            T = terms(spec)
            M, Mv = design(T, tr, idx), design(T, va, va["ok"])
            row = {"terms": name, "n_terms": len(T)}
            for out in ("px", "py", "dxdz", "dydz"):
                c, *_ = np.linalg.lstsq(M, tr[out][idx], rcond=None)
                row["max_residual_" + out] = float(np.abs(Mv @ c - va[out][va["ok"]]).max())
            res[det]["fits"].append(row)
            print(det, name, len(T), "terms: max residual x %.2e px, y %.2e px, dxdz %.1e  (%.0f s)"
                  % (row["max_residual_px"], row["max_residual_py"], row["max_residual_dxdz"], time.time() - t0), flush=True)
    with open(os.path.join(ROOT, "profiles", "r02_design_b_fit.json"), "w") as f:
        json.dump({"what": __doc__.split("\n\n")[0], "tolerance_px": 1e-5, "results": res}, f, indent=1)
