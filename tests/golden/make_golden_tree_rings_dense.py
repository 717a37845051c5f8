"""Dense golden table of the tree-ring radial function from the reference's OWN class.

    python tests/golden/make_golden_tree_rings_dense.py      (needs /root/reference)

``TreeRingRadialFunction`` (imsim/treerings.py:14-69: ``__init__``, ``__call__``, ``dfdr``) is cut out with ``ast``
(the module imports galsim at the top) and evaluated on the parameter blocks stored in tree_rings.npz: f(r) on the
nodes of the look-up table the reference builds (r = 0 ... 8000 px, 2667 points, treerings.py:89-92,190-193) and
on off-node radii, plus df/dr.  The two six-decimal known answers of tests/test_tree_rings.py stay in tree_rings.npz.
"""
import ast
import os

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    src = open(os.path.join(REF, "imsim", "treerings.py")).read()
    tree = ast.parse(src)
    node = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "TreeRingRadialFunction")
    ns = {"np": np}
    exec(ast.get_source_segment(src, node), ns)
    cls = ns["TreeRingRadialFunction"]
    blocks = np.load(os.path.join(HERE, "..", "..", "imsim_b200", "data", "tree_rings.npz"))
    rng = np.random.default_rng(20261019)
    r_nodes = np.linspace(0.0, 8000.0, 2667)
    r_off = np.sort(rng.uniform(0.0, 8000.0, 500))
    out = {"r_nodes": r_nodes, "r_off": r_off}
    for key in blocks.files:
        if "|" not in key:
            continue
        func = cls(list(blocks[key]))
        out["f_nodes|" + key] = np.array([func(r) for r in r_nodes])
        out["f_off|" + key] = np.array([func(r) for r in r_off])
        out["dfdr_off|" + key] = np.array([func.dfdr(r) for r in r_off])
    np.savez_compressed(os.path.join(HERE, "tree_rings_dense.npz"), **out)
    print("wrote tree_rings_dense.npz:", [k for k in out if k.startswith("f_nodes")])


if __name__ == "__main__":
    main()
