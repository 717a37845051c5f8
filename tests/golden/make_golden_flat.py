"""Golden record of the flat builder's control flow from the reference's OWN code.

    python tests/golden/make_golden_flat.py      (needs /root/reference)

``LSST_FlatBuilder.addNoise`` (imsim/flat.py:133-281) is cut out with ``ast`` and run against stand-ins that only
RECORD what the method does: ``galsim.BoundsI`` / ``ImageF`` with just the operations the method uses (integer
bounds algebra, array slicing), a sensor whose ``calculate_pixel_areas`` / ``accumulate`` log the bounds of the
section they are given, a Poisson deviate that returns its mean, a no-op noise builder.  What is pinned: the
section grid (``dx = ncol // nx``, the last section absorbing the remainder, the ``buffer_size`` border, 1-based
bounds), the iteration count ``ceil(counts / max_counts_per_iter)`` and level per iteration, the photon count
``counts_per_iter * bordered_area`` and position ranges of the photon-shot branch, ``resume = it > 0``, and that
only the un-bordered part of a section is added to the image.
"""
import ast
import os
import textwrap
import types

import numpy as np

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


class BoundsI:
    def __init__(self, xmin, xmax, ymin, ymax):
        self.xmin, self.xmax, self.ymin, self.ymax = int(xmin), int(xmax), int(ymin), int(ymax)

    def withBorder(self, b):
        return BoundsI(self.xmin - b, self.xmax + b, self.ymin - b, self.ymax + b)

    def area(self):
        return (self.xmax - self.xmin + 1) * (self.ymax - self.ymin + 1)

    def tup(self):
        return (self.xmin, self.xmax, self.ymin, self.ymax)


class _Wcs:
    def makeSkyImage(self, image, sky_level):
        image.array[:, :] = sky_level


class Image:
    def __init__(self, bounds, wcs=None, array=None):
        self.bounds, self.wcs = bounds, wcs or _Wcs()
        self.array = array if array is not None else np.zeros(
            (bounds.ymax - bounds.ymin + 1, bounds.xmax - bounds.xmin + 1), np.float64)

    def __getitem__(self, b):
        y0, x0 = b.ymin - self.bounds.ymin, b.xmin - self.bounds.xmin
        return Image(b, self.wcs, self.array[y0:y0 + (b.ymax - b.ymin + 1), x0:x0 + (b.xmax - b.xmin + 1)])

    def __setitem__(self, b, value):  # image[b] += x evaluates to image[b] = (image[b] += x): already in place
        pass

    def copy(self):
        return Image(self.bounds, self.wcs, self.array.copy())

    def setZero(self):
        self.array[:, :] = 0

    def __imul__(self, f):
        self.array *= f.array if isinstance(f, Image) else f
        return self

    def __iadd__(self, o):
        self.array += o.array
        return self

    def __truediv__(self, f):
        return Image(self.bounds, self.wcs, self.array / f)


def run(sed, nrow, ncol, nx, ny, buffer_size, counts, max_per_iter):
    src = open(os.path.join(REF, "imsim", "flat.py")).read()
    tree = ast.parse(src)
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "LSST_FlatBuilder")
    node = next(n for n in cls.body if isinstance(n, ast.FunctionDef) and n.name == "addNoise")
    log = []

    class Sensor:
        def updateRNG(self, rng):
            pass

        def calculate_pixel_areas(self, section):
            log.append(("areas",) + section.bounds.tup() + (float(section.array.sum()),))
            return 1.0

        def accumulate(self, photons, section, resume=False):
            log.append(("accumulate",) + section.bounds.tup() + (len(photons.x), int(resume), photons.x.min(),
                                                                  photons.x.max(), photons.y.min(), photons.y.max()))
            section.array += len(photons.x) / section.array.size

    class PhotonArray:
        def __init__(self, n):
            self.x, self.y, self.flux = np.zeros(n), np.zeros(n), None

        def __len__(self):
            return self.x.size

    class UniformDeviate:
        def __init__(self, rng):
            pass

        def generate(self, arr):  # the two extremes and the middle of [0, 1)
            arr[:] = np.resize(np.array([0.0, 0.5, 1.0 - 2.0 ** -53]), arr.shape)

    class PoissonDeviate:
        def __init__(self, rng, mean):
            self.mean = mean

        def __call__(self):
            return self.mean

    def add_noise(base, temp, var, logger):
        log.append(("noise",) + temp.bounds.tup() + (float(temp.array.mean()),))

    logger = types.SimpleNamespace(info=lambda *a: None, debug=lambda *a: None, warning=lambda *a: None)
    galsim = types.SimpleNamespace(
        Sensor=Sensor, ImageF=lambda b, wcs=None: Image(b, wcs), Image=Image, BoundsI=BoundsI,
        PhotonArray=PhotonArray, UniformDeviate=UniformDeviate, PoissonDeviate=PoissonDeviate,
        WavelengthSampler=lambda sed, bp: types.SimpleNamespace(applyTo=lambda photons, rng=None: None),
        config=types.SimpleNamespace(GetRNG=lambda *a, **k: None, AddNoise=add_noise))
    ns = {"np": np, "galsim": galsim, "time": __import__("time")}
    exec("class B:\n" + textwrap.indent(textwrap.dedent(ast.get_source_segment(src, node)), "    "), ns)
    self = types.SimpleNamespace(counts_per_pixel=counts, max_counts_per_iter=max_per_iter, buffer_size=buffer_size,
                                 sed=sed, nx=nx, ny=ny, checkpoint=None)
    image = Image(BoundsI(1, ncol, 1, nrow))
    base = {"current_image": image, "sensor": Sensor(), "bandpass": object()}
    ns["B"].addNoise(self, image, {}, base, 0, 0, 0.0, logger)
    return log, image.array


def main():
    out = {}
    cases = [(None, 40, 64, 8, 2, 5, 2500.0, 1000), (None, 37, 50, 3, 4, 2, 900.0, 1000),
             ("sed", 40, 64, 4, 2, 5, 2500.0, 1000), ("sed", 33, 47, 3, 2, 3, 1000.0, 1000)]
    for k, (sed, nrow, ncol, nx, ny, buf, counts, mx) in enumerate(cases):
        log, img = run(sed, nrow, ncol, nx, ny, buf, counts, mx)
        out["c%d_in" % k] = np.array([0 if sed is None else 1, nrow, ncol, nx, ny, buf, counts, mx], dtype=np.float64)
        kinds = sorted(set(r[0] for r in log))
        for kind in kinds:
            out["c%d_%s" % (k, kind)] = np.array([r[1:] for r in log if r[0] == kind], dtype=np.float64)
        out["c%d_image" % k] = img
    out["n_cases"] = np.array(len(cases))
    np.savez_compressed(os.path.join(HERE, "flat_control_flow.npz"), **out)
    print("wrote flat_control_flow.npz:", {k: v.shape for k, v in out.items() if k.startswith("c0")})


if __name__ == "__main__":
    main()
