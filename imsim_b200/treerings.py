"""Tree-ring radial functions (mirror of imsim/treerings.py).

``TreeRings`` reads the same parameter files (``data/tree_ring_data/*.txt``)
and hands out, per detector, the tree-ring centre and the tabulated radial
function that ``SiliconSensor`` needs.  The reference tabulates with
``galsim.LookupTable.from_func`` (spline, 2667 nodes on [0, 8000] px,
treerings.py:100-103,192-194); here the table is a :class:`RadialTable`, whose
nodes are evaluated with vectorised numpy (the reference loops over 40 terms
per node in Python, "half a minute" for 189 CCDs) and whose spline second
derivatives are what the sensor kernel interpolates with.
"""
from __future__ import annotations

import os
import warnings

import numpy as np

numfreqs = 20  # spatial frequencies per block (treerings.py:100)


class TreeRingsError(Exception):
    pass


class TreeRingRadialFunction:
    """Radial tree-ring function of one CCD (treerings.py:14-68)."""

    def __init__(self, info_block):
        items = info_block[1].split()
        self.A = float(items[6])
        self.B = float(items[7])
        data = np.array([[float(v) for v in line.split()] for line in info_block[3:] if line.strip()])
        self.cfreqs, self.cphases, self.sfreqs, self.sphases = data.T

    def __call__(self, r):
        r = np.asarray(r, dtype=float)
        rr = r[..., None]
        # accumulate term by term, in the reference's order, so the nodes are bit-identical
        shift = np.zeros_like(r)
        for j, fval in enumerate(self.cfreqs):
            shift = shift + np.sin(2 * np.pi * (r / fval) + self.cphases[j]) * fval / (2.0 * np.pi)
        for j, fval in enumerate(self.sfreqs):
            shift = shift + -np.cos(2 * np.pi * (r / fval) + self.sphases[j]) * fval / (2.0 * np.pi)
        del rr
        shift = shift * ((self.A + self.B * r**4) * .01)  # data is in percent
        return shift

    def dfdr(self, r):
        r = np.asarray(r, dtype=float)
        val = np.zeros_like(r)
        for j, fval in enumerate(self.cfreqs):
            val = val + np.cos(2 * np.pi * (r / fval) + self.cphases[j])
        for j, fval in enumerate(self.sfreqs):
            val = val + np.sin(2 * np.pi * (r / fval) + self.sphases[j])
        val = val * ((self.A + self.B * r**4) * .01)
        val = val + self(r) / (self.A + self.B * r**4) * self.B * r**3 / 4.
        return val


def natural_spline_y2(x, f):
    """Second derivatives of the natural cubic spline through (x, f) -- what
    ``galsim.LookupTable(interpolant='spline')`` precomputes."""
    x = np.asarray(x, float)
    f = np.asarray(f, float)
    n = len(x)
    y2 = np.zeros(n)
    if n < 3:
        return y2
    cp = np.zeros(n)
    dp = np.zeros(n)
    for i in range(1, n - 1):
        h0, h1 = x[i] - x[i - 1], x[i + 1] - x[i]
        b = 2.0 * (h0 + h1)
        rhs = 6.0 * ((f[i + 1] - f[i]) / h1 - (f[i] - f[i - 1]) / h0)
        m = b - h0 * cp[i - 1]
        cp[i] = h1 / m
        dp[i] = (rhs - h0 * dp[i - 1]) / m
    for i in range(n - 2, 0, -1):
        y2[i] = dp[i] - cp[i] * y2[i + 1]
    return y2


class RadialTable:
    """Tabulated radial function with GalSim ``LookupTable`` semantics
    (``x``, ``f``, ``interpolant``, callable)."""

    def __init__(self, x, f, interpolant="spline"):
        self.x = np.ascontiguousarray(x, dtype=np.float64)
        self.f = np.ascontiguousarray(f, dtype=np.float64)
        self.interpolant = interpolant
        self.y2 = natural_spline_y2(self.x, self.f) if interpolant == "spline" else None
        self.x_min, self.x_max = float(self.x[0]), float(self.x[-1])

    @classmethod
    def from_func(cls, func, x_min, x_max, npoints=2000, interpolant="spline"):
        xx = np.linspace(x_min, x_max, npoints)
        return cls(xx, func(xx), interpolant)

    def __call__(self, a):
        a = np.asarray(a, dtype=float)
        i = np.clip(np.searchsorted(self.x, a, side="right"), 1, len(self.x) - 1)
        h = self.x[i] - self.x[i - 1]
        aa = self.x[i] - a
        bb = h - aa
        if self.y2 is None:
            ax = aa / h
            return self.f[i] * (1.0 - ax) + self.f[i - 1] * ax
        return (aa * self.f[i - 1] + bb * self.f[i]
                - (1. / 6.) * aa * bb * ((aa + h) * self.y2[i - 1] + (bb + h) * self.y2[i])) / h

    def __len__(self):
        return len(self.x)


class _Position:
    def __init__(self, x, y):
        self.x, self.y = float(x), float(y)

    def __iter__(self):
        return iter((self.x, self.y))

    def __repr__(self):
        return "PositionD(%r,%r)" % (self.x, self.y)


def _position(x, y):
    try:
        import galsim  # noqa: PLC0415

        return galsim.PositionD(x, y)
    except ImportError:
        return _Position(x, y)


class TreeRings:
    """Per-detector tree-ring models read from a parameter file (treerings.py:71-218)."""

    _req_params = {'file_name': str}
    _opt_params = {'only_dets': list, 'defer_load': bool}

    def __init__(self, file_name, only_dets=None, logger=None, defer_load=True, data_dir=None):
        self.file_name = file_name
        if not os.path.isfile(self.file_name):
            for d in filter(None, [data_dir, os.environ.get("IMSIM_DATA_DIR")]):
                cand = os.path.join(d, 'tree_ring_data', file_name)
                if os.path.isfile(cand):
                    self.file_name = cand
                    break
        if not os.path.isfile(self.file_name):
            raise OSError("TreeRing file %s not found" % file_name)
        self.only_dets = only_dets
        self.numfreqs = numfreqs
        self.r_max = 8000.0  # maximum extent of the tree-ring function in pixels
        dr = 3.0  # step size in pixels
        self.npoints = int(self.r_max / dr) + 1
        if logger is not None:
            logger.warning("TreeRing file %s will be used.", self.file_name)
        self._read_info_blocks()
        if only_dets and logger is not None:
            missing_dets = set(only_dets).difference(self.info_blocks)
            if missing_dets:
                logger.info("Requested det_names that are not in the tree ring info file: %s", missing_dets)
        self.info = {}
        if not defer_load:
            self.fill_dict(only_dets=only_dets)

    def _read_info_blocks(self):
        with open(self.file_name, 'r') as fobj:
            lines = fobj.readlines()
        block_size = self.numfreqs + 3
        self.info_blocks = {}
        for iblock in range(len(lines) // block_size):
            block = lines[iblock * block_size:(iblock + 1) * block_size]
            items = block[1].split()
            self.info_blocks["R%s%s_S%s%s" % tuple(items[:4])] = block

    def write(self, outfile, overwrite=False):
        if os.path.isfile(outfile) and not overwrite:
            raise FileExistsError(f"{outfile} already exists.")
        with open(outfile, 'w') as fobj:
            for block in self.info_blocks.values():
                fobj.writelines(block)

    def update_info_block(self, det_name, Cx=None, Cy=None, A=None, B=None):
        keys = ["Rx", "Ry", "Sx", "Sy", "Cx", "Cy", "A", "B"]
        pars = dict(zip(keys, self.info_blocks[det_name][1].split()))
        pars['Cx'] = f"{Cx:.1f}" if Cx is not None else pars['Cx']
        pars['Cy'] = f"{Cy:.1f}" if Cy is not None else pars['Cy']
        pars['A'] = f"{A:.2e}" if A is not None else pars['A']
        pars['B'] = f"{B:.2e}" if B is not None else pars['B']
        self.info_blocks[det_name][1] = "\t".join([pars[k] for k in keys]) + "\n"
        self.info.pop(det_name, None)

    def fill_dict(self, only_dets=None):
        xCenterPix = 2048.5
        yCenterPix = 2048.5
        if only_dets is None:
            only_dets = self.info_blocks.keys()
        for det_name in only_dets:
            if det_name not in self.info_blocks:
                continue
            info_block = self.info_blocks[det_name]
            items = info_block[1].split()
            center = _position(float(items[4]) + xCenterPix, float(items[5]) + yCenterPix)
            func = RadialTable.from_func(TreeRingRadialFunction(info_block), x_min=0.0, x_max=self.r_max,
                                         npoints=self.npoints)
            self.info[det_name] = (center, func)

    def get_dfdr(self, det_name):
        return TreeRingRadialFunction(self.info_blocks[det_name]).dfdr

    def get_center(self, det_name):
        if det_name not in self.info:
            self.fill_dict((det_name,))
        if det_name in self.info:
            return self.info[det_name][0]
        warnings.warn("No treering information available for %s.  Setting treering_center to PositionD(0, 0)." % det_name)
        return _position(0, 0)

    def get_func(self, det_name):
        if det_name not in self.info:
            self.fill_dict((det_name,))
        if det_name in self.info:
            return self.info[det_name][1]
        warnings.warn("No treering information available for %s.  Setting treering_func to None." % det_name)
        return None
