"""Golden vector of cosmic-ray painting from the reference's OWN code.

    python tests/golden/make_golden_cosmic_rays.py      (needs /root/reference)

imsim/cosmic_rays.py imports astropy and galsim at module level, neither of which exists here, so the source of
``CosmicRays.paint`` and ``CosmicRays.paint_cr`` is cut out with ``ast`` and executed on a plain ``list`` subclass with
a stand-in ``galsim.UniformDeviate`` that replays a recorded sequence of uniforms: the painting arithmetic, the
draw order (index, x, y per cosmic ray) and numpy's indexing rules (negative indices wrap, only indices beyond
the array raise IndexError and are skipped) are the reference's.
"""
import ast
import os
import textwrap
import types
from collections import namedtuple

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    path = os.path.join(REF, "imsim", "cosmic_rays.py")
    source = open(path).read()
    tree = ast.parse(source)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "CosmicRays")
    rng = np.random.default_rng(20261019)
    uniforms = rng.random(4000)
    cursor = [0]

    class UniformDeviate:
        def __init__(self, rng):
            pass

        def __call__(self):
            u = uniforms[cursor[0]]
            cursor[0] += 1
            return u

    ns = {"np": np, "galsim": types.SimpleNamespace(UniformDeviate=UniformDeviate)}
    body = "class CR(list):\n"
    for name in ("paint_cr",):
        node = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == name)
        body += textwrap.indent(textwrap.dedent(ast.get_source_segment(source, node)), "    ") + "\n"
    exec(body, ns)
    CR_Span = namedtuple("CR_Span", "x0 y0 pixel_values".split())
    # a synthetic catalogue: tracks, worms and spots, spans in increasing y, some starting left of the first span
    crs = ns["CR"]()
    fp_id, x0s, y0s, vals = [], [], [], []
    for k in range(40):
        nspan = int(rng.integers(1, 9))
        x, y = int(rng.integers(100, 3900)), int(rng.integers(100, 3900))
        spans = []
        for s in range(nspan):
            ln = int(rng.integers(1, 12))
            spans.append(CR_Span(x + int(rng.integers(-4, 3)) * (s > 0), y + s, rng.integers(30, 4000, ln).astype(np.int32)))
            fp_id.append(k)
            x0s.append(spans[-1].x0)
            y0s.append(spans[-1].y0)
            vals.append(spans[-1].pixel_values)
        crs.append(spans)
    ny, nx = 120, 90
    img = rng.poisson(50.0, (ny, nx)).astype(np.float32)
    out = img.copy()
    ncr = 700  # plenty land near the edges of the small image: wrap-around and skipped pixels
    for _ in range(ncr):
        out = crs.paint_cr(out, None)
    np.savez_compressed(os.path.join(HERE, "cosmic_rays.npz"), image_in=img, image_out=out, uniforms=uniforms[:cursor[0]],
                        fp_id=np.array(fp_id), x0=np.array(x0s), y0=np.array(y0s),
                        pixel_values=np.concatenate(vals), span_len=np.array([len(v) for v in vals]), num_crs=ncr)
    print("painted", ncr, "cosmic rays; uniforms used", cursor[0], "changed pixels", int((out != img).sum()),
          "added", float(out.sum(dtype=np.float64) - img.sum(dtype=np.float64)))


if __name__ == "__main__":
    main()
