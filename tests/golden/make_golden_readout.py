"""Golden vectors of the post-path from the reference's OWN code.

Run in the build container (needs /root/reference):
    python tests/golden/make_golden_readout.py

  * imsim/bleed_trails.py is pure numpy and is loaded standalone with importlib: bleed_eimage on float32
    e-images with saturated blobs (midline stop on / off, blobs touching both ends of a channel).
  * imsim/readout.py cannot be imported here (astropy / galsim / lsst are absent), so the source text of
    cte_matrix, CcdReadout.apply_cte and CcdReadout.apply_crosstalk is cut out of the file with ``ast`` and
    executed against a small stand-in object: the arithmetic that produces the vectors is the reference's,
    line for line.  (numpy 2.3 / NEP 50 promotion rules, scipy.special.binom.)
"""
import ast
import importlib.util
import os
import textwrap
import types

import numpy as np
import scipy.special

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_bleed():
    spec = importlib.util.spec_from_file_location("ref_bleed", os.path.join(REF, "imsim", "bleed_trails.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def cut(source, tree, name, cls=None):
    nodes = tree.body
    if cls is not None:
        nodes = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == cls).body
    node = next(n for n in nodes if isinstance(n, ast.FunctionDef) and n.name == name)
    return textwrap.dedent(ast.get_source_segment(source, node))


def main():
    rng = np.random.default_rng(20261018)
    out = {}
    # ---- bleed trails
    b = load_bleed()
    ny, nx = 96, 40
    for tag, full_well in (("a", 100000.0), ("b", 98765.4)):
        img = rng.poisson(800.0, (ny, nx)).astype(np.float32)
        for _ in range(9):  # saturated blobs, some huge so that trails run off both ends / reach the midline
            cx, cy = rng.integers(0, nx), rng.integers(0, ny)
            amp = 10.0 ** rng.uniform(5.2, 6.8)
            yy, xx = np.mgrid[0:ny, 0:nx]
            img += (amp * np.exp(-0.5 * ((xx - cx) ** 2 + (yy - cy) ** 2) / 1.5 ** 2)).astype(np.float32)
        img[0:3, 5] += 3.0e5      # touching the bottom: charge escapes into the electronics
        img[ny - 2:, 7] += 4.0e5  # touching the top: no escape
        out["bleed_in_" + tag] = img.copy()
        out["bleed_fw_" + tag] = np.float64(full_well)
        out["bleed_mid_" + tag] = b.bleed_eimage(img.copy(), full_well, midline_stop=True)
        out["bleed_nomid_" + tag] = b.bleed_eimage(img.copy(), full_well, midline_stop=False)
    # ---- CTE matrix, apply_cte, apply_crosstalk from readout.py source
    path = os.path.join(REF, "imsim", "readout.py")
    source = open(path).read()
    tree = ast.parse(source)
    ns = {"np": np, "scipy": scipy}
    exec(cut(source, tree, "cte_matrix"), ns)
    exec(cut(source, tree, "apply_cte", "CcdReadout"), ns)
    exec(cut(source, tree, "apply_crosstalk", "CcdReadout"), ns)
    cte_matrix, apply_cte, apply_crosstalk = ns["cte_matrix"], ns["apply_cte"], ns["apply_crosstalk"]
    out["cte_matrix_40_1e-3"] = cte_matrix(40, 1e-3)
    out["cte_matrix_64_1e-6_band"] = np.array([cte_matrix(64, 1e-6)[63, 63 - k] for k in range(24)])
    nry, nrx, namp = 72, 56, 4
    amps = [rng.poisson(1500.0, (nry, nrx)).astype(np.float32) for _ in range(namp)]
    for a in amps:
        a[rng.integers(0, nry, 6), rng.integers(0, nrx, 6)] += rng.uniform(2e4, 1.5e5, 6).astype(np.float32)
    out["amps_in"] = np.array(amps)
    for tag, pcti, scti in (("both", 1e-4, 3e-5), ("p_only", 1e-6, 0.0), ("s_only", 0.0, 1e-6)):
        me = types.SimpleNamespace(pcte_matrix=None if pcti == 0 else cte_matrix(nry, pcti),
                                   scte_matrix=None if scti == 0 else cte_matrix(nrx, scti))
        segs = [types.SimpleNamespace(array=a.copy()) for a in amps]
        res = apply_cte(me, segs)
        out["cte_" + tag] = np.array([s.array for s in res])
        out["cte_" + tag + "_cti"] = np.array([pcti, scti])
    xtalk = rng.normal(0.0, 2e-4, (namp, namp))
    np.fill_diagonal(xtalk, 0.0)
    me = types.SimpleNamespace(ccd=types.SimpleNamespace(xtalk=xtalk))
    res = apply_crosstalk(me, [a.copy() for a in amps])
    out["xtalk"] = xtalk
    out["xtalk_out"] = np.array(res)
    assert out["xtalk_out"].dtype == np.float64
    np.savez_compressed(os.path.join(HERE, "readout.npz"), **out)
    print({k: (v.shape, v.dtype) for k, v in out.items()})


if __name__ == "__main__":
    main()
