"""Golden vectors of the atmosphere's parameter draws from the reference's OWN code.

    python tests/golden/make_golden_atmosphere.py      (needs /root/reference)

imsim/atmPSF.py imports galsim (absent here), so the source of ``AtmosphericPSF._vkSeeing``, ``_seeingResid``,
``_r0_500`` and ``_getAtmKwargs`` (atmPSF.py:211-296) is cut out with ``ast`` and executed with a stand-in
``galsim`` that supplies only what those methods touch: ``GaussianDeviate`` / ``UniformDeviate`` replaying
recorded sequences, ``degrees = 1`` (directions are stored in degrees here), and ``Kolmogorov(r0_500, lam).fwhm``
restated as ``0.9758634299 lam / r0`` (galsim/kolmogorov.py ``_fwhm_factor``, recalled -- GalSim itself is not
pinned).  Pinned by this: the reference's draw ORDER (six Gaussian weights, the truncated log-normal outer scale
with its rejection loop, six uniform speeds, six uniform directions), its clipping / renormalisation of the
weights, and the bisection for ``r0_500`` with its bracket.
"""
import ast
import os
import textwrap
import types

import numpy as np
from scipy.optimize import bisect

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
ARCSEC = 206264.80624709636


def reference_class(gauss, unif):
    path = os.path.join(REF, "imsim", "atmPSF.py")
    source = open(path).read()
    tree = ast.parse(source)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "AtmosphericPSF")
    cur = {"g": 0, "u": 0}

    class GaussianDeviate:
        def __init__(self, rng):
            pass

        def __call__(self):
            cur["g"] += 1
            return gauss[cur["g"] - 1]

    class UniformDeviate:
        def __init__(self, rng):
            pass

        def __call__(self):
            cur["u"] += 1
            return unif[cur["u"] - 1]

    class Kolmogorov:
        def __init__(self, r0_500, lam):
            r0 = r0_500 * (lam / 500.0) ** 1.2
            self.fwhm = 0.9758634299 * lam * 1e-9 / r0 * ARCSEC

    galsim = types.SimpleNamespace(GaussianDeviate=GaussianDeviate, UniformDeviate=UniformDeviate,
                                   Kolmogorov=Kolmogorov, degrees=1.0)
    ns = {"np": np, "galsim": galsim, "bisect": bisect}
    body = "class AtmosphericPSF:\n"
    for name in ("_vkSeeing", "_seeingResid", "_r0_500", "_getAtmKwargs"):
        node = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == name)
        seg = ast.get_source_segment(source, node)
        deco = "@staticmethod\n" if node.decorator_list else ""
        body += textwrap.indent(deco + textwrap.dedent(seg), "    ") + "\n"
    exec(body, ns)
    return ns["AtmosphericPSF"], cur


def main():
    rng = np.random.default_rng(20261017)
    cases = []
    wlen = dict(u=365.49, g=480.03, r=622.20, i=754.06, z=868.21, y=991.66)
    out = {}
    for k, (band, airmass, seeing) in enumerate([("r", 1.2, 0.7), ("u", 1.0, 1.1), ("y", 2.0, 0.5), ("i", 1.05, 0.62),
                                                   ("g", 1.4, 0.9), ("z", 1.7, 0.8)]):
        gauss = rng.standard_normal(64)
        if k == 1:
            gauss[6] = 4.0   # exp(4 * 0.6) * 25 > 100: the outer-scale draw is rejected once
        if k == 2:
            gauss[6], gauss[7] = -3.0, 5.0  # < 10, then > 100: rejected twice
        unif = rng.random(16)
        cls, cur = reference_class(gauss, unif)
        self = types.SimpleNamespace(rng=None, logger=None, wlen_eff=wlen[band], airmass=airmass,
                                     targetFWHM=seeing * airmass ** 0.6 * (wlen[band] / 500) ** (-0.3),
                                     screen_size=819.2, screen_scale=0.1)
        kw = cls._getAtmKwargs(self)
        out["case%d_in" % k] = np.array([wlen[band], airmass, seeing])
        out["case%d_gauss" % k] = gauss
        out["case%d_unif" % k] = unif
        out["case%d_used" % k] = np.array([cur["g"], cur["u"]])
        out["case%d_r0_500" % k] = np.array(kw["r0_500"])
        out["case%d_L0" % k] = np.array(kw["L0"])
        out["case%d_speed" % k] = np.array(kw["speed"])
        out["case%d_direction_deg" % k] = np.array(kw["direction"])
        out["case%d_altitude" % k] = np.array(kw["altitude"])
        out["case%d_weights" % k] = np.array(kw["r0_weights"])
        cases.append(k)
    # the seeing relations on a grid
    cls, _ = reference_class([], [])
    grid = [(r0, lam, L0) for r0 in (0.05, 0.12, 0.2, 0.35) for lam in (365.49, 622.2, 991.66) for L0 in (10.0, 25.0, 100.0)]
    out["vk_grid"] = np.array(grid)
    out["vk_seeing"] = np.array([cls._vkSeeing(*g) for g in grid])
    tgt = [(lam, L0, t) for lam in (365.49, 622.2, 991.66) for L0 in (12.0, 25.0, 80.0) for t in (0.5, 0.8, 1.3)]
    out["r0_grid"] = np.array(tgt)
    out["r0_500"] = np.array([cls._r0_500(*g) for g in tgt])
    out["n_cases"] = np.array(len(cases))
    np.savez(os.path.join(HERE, "atmosphere.npz"), **out)
    print("wrote atmosphere.npz:", len(cases), "draw cases,", len(grid), "+", len(tgt), "seeing relations")


if __name__ == "__main__":
    main()
