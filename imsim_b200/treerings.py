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
from dataclasses import dataclass

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


@dataclass
class TreeRingRecord:
    """One detector's entry of a tree-ring parameter file, parsed: 23 text lines in the file (a comment line,
    the parameter line ``Rx Ry Sx Sy Cx Cy A B``, a column header and ``numfreqs`` rows
    ``CosFreq CosPhase SinFreq SinPhase``; imsim/treerings.py:120-136)."""

    det_name: str
    comment: str
    columns: str
    raft_slot: tuple          # the four index tokens as written in the file
    cx: str                   # centre offset and amplitudes stay text until asked for: rewriting a file
    cy: str                   # must reproduce untouched entries character by character
    a: str
    b: str
    rows: list                # the numfreqs coefficient lines, verbatim
    original: str = ""        # the parameter line as read (returned while none of its fields was edited)

    @classmethod
    def parse(cls, lines):
        tok = lines[1].split()
        return cls("R%s%s_S%s%s" % tuple(tok[:4]), lines[0], lines[2], tuple(tok[:4]), tok[4], tok[5], tok[6],
                   tok[7], list(lines[3:]), lines[1])

    def parameter_line(self) -> str:
        if self.original and self.original.split()[4:8] == [self.cx, self.cy, self.a, self.b]:
            return self.original
        return "\t".join(self.raft_slot + (self.cx, self.cy, self.a, self.b)) + "\n"

    def lines(self) -> list:
        """The block as text, in the layout ``TreeRingRadialFunction`` reads."""
        return [self.comment, self.parameter_line(), self.columns] + self.rows

    def center(self, x0=2048.5, y0=2048.5):
        return float(self.cx) + x0, float(self.cy) + y0


def read_tree_ring_file(path) -> "dict[str, TreeRingRecord]":
    """All detector records of a parameter file, keyed by ``Rxx_Syy``, in file order."""
    per = numfreqs + 3
    with open(path, "r") as f:
        text = f.readlines()
    recs = (TreeRingRecord.parse(text[k:k + per]) for k in range(0, len(text) - per + 1, per))
    return {r.det_name: r for r in recs}


class _InfoBlocks(dict):
    """``TreeRings.info_blocks`` as the reference exposes it (det_name -> list of text lines), backed by the
    parsed records."""

    def __init__(self, records):
        super().__init__()
        self._records = records

    def __getitem__(self, det_name):
        return self._records[det_name].lines()

    def __contains__(self, det_name):
        return det_name in self._records

    def __iter__(self):
        return iter(self._records)

    def __len__(self):
        return len(self._records)

    def keys(self):
        return self._records.keys()

    def values(self):
        return [r.lines() for r in self._records.values()]

    def items(self):
        return [(k, r.lines()) for k, r in self._records.items()]


class TreeRings:
    """Per-detector tree-ring models from a parameter file: the ``tree_rings`` input object of
    imsim/treerings.py:71-218 (same constructor, ``info`` / ``info_blocks``, ``get_center`` / ``get_func`` /
    ``get_dfdr``, ``update_info_block``, ``write``), over parsed records."""

    _req_params = {'file_name': str}
    _opt_params = {'only_dets': list, 'defer_load': bool}

    R_MAX = 8000.0   # extent of the tabulated function [px]
    DR = 3.0         # node spacing [px]
    PIXEL_CENTER = (2048.5, 2048.5)

    def __init__(self, file_name, only_dets=None, logger=None, defer_load=True, data_dir=None):
        self.file_name = self._locate(file_name, data_dir)
        self.only_dets = only_dets
        self.numfreqs = numfreqs
        self.r_max = self.R_MAX
        self.npoints = int(self.R_MAX / self.DR) + 1
        if logger is not None:
            logger.warning("TreeRing file %s will be used.", self.file_name)
        self.records = read_tree_ring_file(self.file_name)
        self.info_blocks = _InfoBlocks(self.records)
        self.info = {}
        unknown = set(only_dets or ()) - set(self.records)
        if unknown and logger is not None:
            logger.info("Requested det_names that are not in the tree ring info file: %s", unknown)
        if not defer_load:
            self.fill_dict(only_dets=only_dets)

    @staticmethod
    def _locate(file_name, data_dir):
        search = [file_name] + [os.path.join(d, 'tree_ring_data', file_name)
                                for d in (data_dir, os.environ.get("IMSIM_DATA_DIR")) if d]
        for cand in search:
            if os.path.isfile(cand):
                return cand
        raise OSError("TreeRing file %s not found" % file_name)

    def write(self, outfile, overwrite=False):
        if os.path.isfile(outfile) and not overwrite:
            raise FileExistsError(f"{outfile} already exists.")
        with open(outfile, 'w') as f:
            for rec in self.records.values():
                f.writelines(rec.lines())

    def update_info_block(self, det_name, Cx=None, Cy=None, A=None, B=None):
        """New centre offsets / amplitudes for one detector, in the file's number formats."""
        rec = self.records[det_name]
        for attr, value, fmt in (("cx", Cx, "%.1f"), ("cy", Cy, "%.1f"), ("a", A, "%.2e"), ("b", B, "%.2e")):
            if value is not None:
                setattr(rec, attr, fmt % value)
        self.info.pop(det_name, None)  # tabulated again on the next request

    def _tabulate(self, rec: TreeRingRecord):
        table = RadialTable.from_func(TreeRingRadialFunction(rec.lines()), x_min=0.0, x_max=self.r_max,
                                      npoints=self.npoints)
        return _position(*rec.center(*self.PIXEL_CENTER)), table

    def fill_dict(self, only_dets=None):
        for det_name in (self.records if only_dets is None else only_dets):
            rec = self.records.get(det_name)
            if rec is not None:
                self.info[det_name] = self._tabulate(rec)

    def _entry(self, det_name, what, default):
        if det_name not in self.info:
            self.fill_dict((det_name,))
        if det_name not in self.info:
            warnings.warn("No treering information available for %s.  Setting %s." % (det_name, what))
            return default
        return self.info[det_name]

    def get_dfdr(self, det_name):
        return TreeRingRadialFunction(self.records[det_name].lines()).dfdr

    def get_center(self, det_name):
        e = self._entry(det_name, "treering_center to PositionD(0, 0)", None)
        return _position(0, 0) if e is None else e[0]

    def get_func(self, det_name):
        e = self._entry(det_name, "treering_func to None", None)
        return None if e is None else e[1]
