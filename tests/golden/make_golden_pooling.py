"""Golden vectors of the pooled batching algebra from the reference's OWN code.

    python tests/golden/make_golden_pooling.py      (needs /root/reference)

imsim/photon_pooling.py and imsim/stamp.py import galsim and the LSST stack at module level, so the source of
``ProcessingMode`` / ``ObjectInfo`` (stamp.py:17-34) and of the static methods ``make_batches``,
``make_photon_batches``, ``make_photon_subbatches`` and ``partition_objects`` of
``LSST_PhotonPoolingImageBuilder`` (photon_pooling.py:227-247, 278-331, 355-386) is cut out with ``ast`` and executed
with a stand-in ``galsim`` whose ``UniformDeviate`` replays a recorded sequence.  Each case stores the object list
(index, flux, mode), the uniforms, and what the reference returns, flattened.
"""
import ast
import dataclasses
import itertools
import os
import textwrap
import types
from dataclasses import dataclass
from enum import Enum, auto

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def reference_namespace(uniforms):
    cursor = [0]

    class UniformDeviate:
        def __init__(self, rng):
            pass

        def __call__(self):
            cursor[0] += 1
            return uniforms[cursor[0] - 1]

    galsim = types.SimpleNamespace(UniformDeviate=UniformDeviate,
                                   config=types.SimpleNamespace(GetRNG=lambda *a, **k: None))
    ns = {"np": np, "galsim": galsim, "dataclasses": dataclasses, "dataclass": dataclass, "itertools": itertools,
          "Enum": Enum, "auto": auto}
    stamp_src = open(os.path.join(REF, "imsim", "stamp.py")).read()
    tree = ast.parse(stamp_src)
    for name in ("ProcessingMode", "ObjectInfo"):
        node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == name)
        deco = "@dataclass\n" if node.decorator_list else ""
        exec(deco + ast.get_source_segment(stamp_src, node), ns)
    pool_src = open(os.path.join(REF, "imsim", "photon_pooling.py")).read()
    tree = ast.parse(pool_src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "LSST_PhotonPoolingImageBuilder")
    body = "class Builder:\n"
    for name in ("make_batches", "make_photon_batches", "make_photon_subbatches", "partition_objects"):
        node = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == name)
        body += textwrap.indent("@staticmethod\n" + textwrap.dedent(ast.get_source_segment(pool_src, node)), "    ") + "\n"
    exec(body, ns)
    return ns, cursor


def flat(batches):
    """list of lists of ObjectInfo -> (offsets, index, flux)"""
    off = np.cumsum([0] + [len(b) for b in batches])
    idx = np.array([o.index for b in batches for o in b], dtype=np.int64)
    flux = np.array([o.phot_flux for b in batches for o in b], dtype=np.float64)
    return off, idx, flux


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    cases = [(0, 10), (1, 1), (7, 10), (40, 3), (200, 10), (200, 50), (999, 7), (64, 64)]
    for k, (nobj, nbatch) in enumerate(cases):
        uniforms = rng.random(max(nobj, 1))
        ns, cursor = reference_namespace(uniforms)
        PM, OI, B = ns["ProcessingMode"], ns["ObjectInfo"], ns["Builder"]
        modes = rng.integers(0, 3, nobj)  # 0 FFT, 1 PHOT, 2 FAINT
        flux = np.where(rng.random(nobj) < 0.3, rng.integers(0, 2 * nbatch + 2, nobj),
                        np.round(10 ** rng.uniform(0, 6.5, nobj))).astype(np.int64)
        mode_of = {0: PM.FFT, 1: PM.PHOT, 2: PM.FAINT}
        objs = [OI(i, int(f), mode_of[int(m)]) for i, (f, m) in enumerate(zip(flux, modes))]
        fft, phot, faint = B.partition_objects(objs, nbatch)
        batches = B.make_photon_batches({}, {}, None, phot, faint, nbatch)
        out["c%d_in" % k] = np.array([nobj, nbatch])
        out["c%d_flux" % k], out["c%d_modes" % k], out["c%d_uniforms" % k] = flux, modes, uniforms
        out["c%d_used" % k] = np.array(cursor[0])
        for nm, lst in (("fft", fft), ("phot", phot), ("faint", faint)):
            out["c%d_%s" % (k, nm)] = np.array([o.index for o in lst], dtype=np.int64)
        for nm, arr in zip(("off", "idx", "flux"), flat(batches)):
            out["c%d_batches_%s" % (k, nm)] = arr
        for nm, arr in zip(("off", "idx", "flux"), flat(list(B.make_batches(fft, nbatch)))):
            out["c%d_fftbatches_%s" % (k, nm)] = arr
        if batches:
            for nsub in (1, 4, 7):
                sub = B.make_photon_subbatches(batches[0], nsub)
                for nm, arr in zip(("off", "idx", "flux"), flat(sub)):
                    out["c%d_sub%d_%s" % (k, nsub, nm)] = arr
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "pooling.npz"), **out)
    print("wrote pooling.npz:", len(cases), "cases")


if __name__ == "__main__":
    main()
