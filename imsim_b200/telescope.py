"""Flat telescope description consumed by the trace kernel.

The reference hands a live ``batoid.Optic`` to ``RubinOptics``
(imsim/photon_ops.py:53-70, built by imsim/telescope_loader.py:210-249 and
specialised per detector at :399-415).  The B200 path consumes the same
information as a flat table of interfaces in trace order
(``B2Telescope`` in include/imsim_b200.h); ``extract.py`` produces that table
from a live batoid object, this module defines the neutral in-between
representation and a built-in Rubin-like prescription for synthetic runs.

Nothing here does per-photon arithmetic.
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

from . import _abi


def rot_z(theta: float) -> np.ndarray:
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def rot_x(theta: float) -> np.ndarray:
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[1.0, 0.0, 0.0], [0.0, c, -s], [0.0, s, c]])


def rot_y(theta: float) -> np.ndarray:
    c, s = np.cos(theta), np.sin(theta)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])


@dataclass
class CoordSys:
    """batoid.CoordSys: ``rot`` columns are the local axes in global coordinates."""

    origin: np.ndarray = field(default_factory=lambda: np.zeros(3))
    rot: np.ndarray = field(default_factory=lambda: np.eye(3))

    def shift_global(self, d) -> "CoordSys":
        return CoordSys(self.origin + np.asarray(d, float), self.rot.copy())

    def shift_local(self, d) -> "CoordSys":
        return CoordSys(self.origin + self.rot @ np.asarray(d, float), self.rot.copy())

    def rotate_local(self, R, center=None) -> "CoordSys":
        """batoid.CoordSys.rotateLocal about a point given in local coordinates."""
        if center is None:
            center = np.zeros(3)
        c_glob = self.origin + self.rot @ np.asarray(center, float)
        new_rot = self.rot @ R
        # keep the rotation centre fixed
        new_origin = c_glob - new_rot @ np.asarray(center, float)
        return CoordSys(new_origin, new_rot)


@dataclass(frozen=True)
class Medium:
    kind: str  # 'const' | 'sellmeier' | 'sumita' | 'air'
    params: tuple

    def key(self):
        return (self.kind, tuple(float(p) for p in self.params))

    def n(self, wavelength_m):
        """Refractive index (host-side helper for set-up code, not the photon path)."""
        wl = np.asarray(wavelength_m, dtype=float)
        p = self.params
        if self.kind == "const":
            return np.full_like(wl, p[0])
        if self.kind == "sellmeier":
            x = (wl * 1e6) ** 2
            return np.sqrt(1 + p[0] * x / (x - p[3]) + p[1] * x / (x - p[4]) + p[2] * x / (x - p[5]))
        if self.kind == "sumita":
            x = (wl * 1e6) ** 2
            y = 1 / x
            return np.sqrt(p[0] + p[1] * x + y * (p[2] + y * (p[3] + y * (p[4] + y * p[5]))))
        if self.kind == "air":
            P, T, W = p[0] * 7.50061683, p[1] - 273.15, p[2] * 7.50061683
            s2 = 1e-12 / (wl * wl)
            nm1 = (64.328 + 29498.1 / (146.0 - s2) + 255.4 / (41.0 - s2)) * 1e-6
            nm1 = nm1 * (P * (1.0 + (1.049 - 0.0157 * T) * 1e-6 * P) / (720.883 * (1.0 + 0.003661 * T)))
            nm1 = nm1 - (0.0624 - 0.000680 * s2) / (1.0 + 0.003661 * T) * W * 1e-6
            return 1 + nm1
        raise ValueError(self.kind)


AIR = Medium("air", (69.328, 293.15, 1.067))  # batoid.Air() defaults
SILICA = Medium(
    "sellmeier",
    (0.6961663, 0.4079426, 0.8974794, 0.0684043**2, 0.1162414**2, 9.896161**2),
)
VACUUM = Medium("const", (1.0,))


@dataclass
class Obscuration:
    kind: str  # 'circle' | 'annulus' | 'rectangle' | 'ray'
    params: Sequence[float]
    negate: bool = False  # Clear* == negated obscuration


@dataclass
class Surface:
    kind: str  # 'plane' | 'sphere' | 'paraboloid' | 'quadric' | 'asphere'
    R: float = 0.0
    conic: float = 0.0
    coefs: Sequence[float] = ()
    # optional summed perturbation (batoid.Sum([base, Zernike|Bicubic]))
    poly: Optional[np.ndarray] = None  # (n, n) xy-polynomial coefficients c[i, j] x^i y^j
    poly_scale: float = 1.0
    bicubic: Optional[dict] = None  # {'xs','ys','zs','dzdxs','dzdys','d2zdxdys'}


@dataclass
class Interface:
    name: str
    surface: Surface
    interact: str  # 'mirror' | 'refract' | 'detector' | 'pass'
    coord_sys: CoordSys
    in_medium: Medium
    out_medium: Medium
    obscurations: List[Obscuration] = field(default_factory=list)
    group: str = ""  # e.g. 'LSSTCamera' for rotator / shifts


_SURF = {"plane": _abi.SURF_PLANE, "sphere": _abi.SURF_SPHERE, "paraboloid": _abi.SURF_PARABOLOID,
         "quadric": _abi.SURF_QUADRIC, "asphere": _abi.SURF_ASPHERE}
_INT = {"detector": _abi.INT_DETECTOR, "mirror": _abi.INT_MIRROR, "refract": _abi.INT_REFRACT, "pass": _abi.INT_PASS}
_OBSC = {"circle": _abi.OBSC_CIRCLE, "annulus": _abi.OBSC_ANNULUS, "rectangle": _abi.OBSC_RECTANGLE,
         "ray": _abi.OBSC_RAY}
_MED = {"const": _abi.MED_CONST, "sellmeier": _abi.MED_SELLMEIER, "sumita": _abi.MED_SUMITA, "air": _abi.MED_AIR}


@dataclass
class Telescope:
    """Sequential optical system in trace order."""

    stop: CoordSys
    items: List[Interface]
    in_medium: Medium = AIR
    name: str = "telescope"

    # -- batoid.Optic.with* equivalents used by imsim/telescope_loader.py ----
    def with_locally_rotated_group(self, group: str, R: np.ndarray, center_item: Optional[str] = None) -> "Telescope":
        """``withLocallyRotatedOptic(group, rot)``: rotate all items of a group about the
        group's origin (telescope_loader.py:242-246 rotates 'LSSTCamera' by RotZ(rotTelPos))."""
        new = copy.deepcopy(self)
        members = [it for it in new.items if it.group == group]
        if not members:
            raise KeyError(group)
        ref = members[0].coord_sys if center_item is None else next(
            it.coord_sys for it in new.items if it.name == center_item)
        ref = CoordSys(ref.origin.copy(), ref.rot.copy())
        Rg = ref.rot @ R @ ref.rot.T  # the same rotation expressed in global axes
        for it in members:
            it.coord_sys = CoordSys(ref.origin + Rg @ (it.coord_sys.origin - ref.origin), Rg @ it.coord_sys.rot)
        return new

    def with_locally_shifted_item(self, name: str, shift) -> "Telescope":
        """``withLocallyShiftedOptic(name, shift)`` (telescope_loader.py:399-405: per-CCD
        ``[0, 0, -z_offset]`` on 'Detector')."""
        new = copy.deepcopy(self)
        for it in new.items:
            if it.name == name:
                it.coord_sys = it.coord_sys.shift_local(shift)
                return new
        raise KeyError(name)

    def with_surface_perturbation(self, name: str, poly=None, poly_scale=1.0, bicubic=None) -> "Telescope":
        new = copy.deepcopy(self)
        for it in new.items:
            if it.name == name:
                it.surface.poly = None if poly is None else np.asarray(poly, float)
                it.surface.poly_scale = float(poly_scale)
                it.surface.bicubic = bicubic
                return new
        raise KeyError(name)

    def with_inserted_screen(self, before: str, poly=None, poly_scale=1.0, bicubic=None, coord_sys=None,
                             obscurations=None, name: str = "Screen") -> "Telescope":
        """``withInsertedOptic(before=..., item=batoid.OPDScreen(surface=Plane(), screen=Zernike | Bicubic,
        coordSys=stopSurface.coordSys, obscuration=...))`` (tests/test_telescope_loader.py:641-653): a thin
        phase plate whose optical path difference [m] is the xy polynomial / bicubic grid."""
        new = copy.deepcopy(self)
        k = next((i for i, it in enumerate(new.items) if it.name == before), None)
        if k is None:
            raise KeyError(before)
        med = new.items[k].in_medium
        surf = Surface("plane", poly=None if poly is None else np.asarray(poly, float), poly_scale=float(poly_scale),
                       bicubic=bicubic)
        cs = coord_sys if coord_sys is not None else CoordSys(new.stop.origin.copy(), new.stop.rot.copy())
        new.items.insert(k, Interface(name, surf, "pass", cs, med, med, list(obscurations or [])))
        return new

    # -- flattening -------------------------------------------------------
    def flatten(self):
        """Return ``(B2Telescope, extras)``; ``extras[i]`` is ``None`` or
        ``(kind, float64 array)`` for surface ``i`` (see ``b2_telescope_set_extra``)."""
        if len(self.items) > _abi.B2_MAX_SURFACES:
            raise ValueError("too many surfaces")
        tel = _abi.B2Telescope()
        media: List[Medium] = []

        def midx(m: Medium) -> int:
            for k, mm in enumerate(media):
                if mm.key() == m.key():
                    return k
            media.append(m)
            return len(media) - 1

        tel.medium_stop = midx(self.in_medium)
        prev = self.stop
        extras = []
        for i, it in enumerate(self.items):
            s = tel.surf[i]
            s.surf_kind = _SURF[it.surface.kind]
            s.interact = _INT[it.interact]
            s.medium_in = midx(it.in_medium)
            s.medium_out = midx(it.out_medium)
            coefs = list(it.surface.coefs)
            if len(coefs) > _abi.B2_MAX_ASPHERE_COEF:
                raise ValueError("too many asphere coefficients")
            s.n_coef = len(coefs)
            for k, c in enumerate(coefs):
                s.coef[k] = c
            s.R = it.surface.R
            s.conic = it.surface.conic
            # batoid CoordTransform(source=prev, dest=this)
            dr = prev.rot.T @ (it.coord_sys.origin - prev.origin)
            drot = prev.rot.T @ it.coord_sys.rot
            # consecutive interfaces of a rigidly rotated group (the camera behind the rotator)
            # share their frame: snap round-off (|.| < 4 ulp) to the exact identity
            if np.abs(drot - np.eye(3)).max() < 1e-15:
                drot = np.eye(3)
            s.rot_identity = int(np.array_equal(drot, np.eye(3)))
            for k in range(3):
                s.dr[k] = dr[k]
            for k, val in enumerate(drot.ravel()):
                s.drot[k] = val
            if len(it.obscurations) > _abi.B2_MAX_OBSC:
                raise ValueError("too many obscurations on %s" % it.name)
            s.n_obsc = len(it.obscurations)
            for k, ob in enumerate(it.obscurations):
                o = s.obsc[k]
                o.kind = _OBSC[ob.kind]
                o.negate = int(ob.negate)
                p = list(ob.params)
                if ob.kind == "rectangle":  # width height x0 y0 theta
                    p = [p[0], p[1], p[2], p[3], np.cos(p[4]), np.sin(p[4])]
                elif ob.kind == "ray":  # width theta x0 y0 -> width x0 y0 cos sin
                    p = [p[0], p[2], p[3], np.cos(p[1]), np.sin(p[1])]
                for j, val in enumerate(p):
                    o.p[j] = val
            s.extra_kind = _abi.EXTRA_NONE
            s.extra_slot = -1
            s.poly_n = 0
            s.poly_scale = 1.0
            extra = None
            if it.surface.poly is not None:
                c = np.ascontiguousarray(it.surface.poly, dtype=np.float64)
                if c.ndim != 2 or c.shape[0] != c.shape[1] or c.shape[0] > _abi.B2_MAX_POLY_ORDER:
                    raise ValueError("poly must be square with order < %d" % _abi.B2_MAX_POLY_ORDER)
                s.extra_kind = _abi.EXTRA_POLY2D
                s.poly_n = c.shape[0]
                s.poly_scale = it.surface.poly_scale
                extra = (_abi.EXTRA_POLY2D, c.ravel().copy())
            elif it.surface.bicubic is not None:
                b = it.surface.bicubic
                xs, ys = np.asarray(b["xs"], float), np.asarray(b["ys"], float)
                hdr = np.array([xs[0], xs[1] - xs[0], len(xs), ys[0], ys[1] - ys[0], len(ys)], float)
                grids = [np.asarray(b[k], float).reshape(len(ys), len(xs)).ravel()
                         for k in ("zs", "dzdxs", "dzdys", "d2zdxdys")]
                s.extra_kind = _abi.EXTRA_BICUBIC
                extra = (_abi.EXTRA_BICUBIC, np.concatenate([hdr] + grids))
            extras.append(extra)
            prev = it.coord_sys
        tel.n_surfaces = len(self.items)
        if len(media) > _abi.B2_MAX_MEDIA:
            raise ValueError("too many media")
        tel.n_media = len(media)
        for k, m in enumerate(media):
            tel.media[k].kind = _MED[m.kind]
            for j, val in enumerate(m.params):
                tel.media[k].p[j] = val
        return tel, extras


def _clear_annulus(inner, outer):
    return [Obscuration("annulus", (inner, outer, 0.0, 0.0), negate=True)]


def _clear_circle(radius):
    return [Obscuration("circle", (radius, 0.0, 0.0), negate=True)]


#: per-band filter (R1, R2, thickness) [m]; the L3-side gap keeps the total track fixed.  Only the r band is
#: pinned (against the Zemax wavefront the reference's tests hold, see ``lsst_v33``); the other bands are recalled.
_FILTER = {
    "u": (5.624, 5.564, 0.0262),
    "g": (5.624, 5.594, 0.0215),
    "r": (5.632, 5.606, 0.0179),
    "i": (5.632, 5.612, 0.0156),
    "z": (5.632, 5.617, 0.0141),
    "y": (5.632, 5.618, 0.0135),
}

#: the medium the rays start in.  The Zemax design (and the batoid model that reproduces it to 0.01 nm,
#: /root/reference/tests/test_opd.py:16-95) treats air as n = 1 with the glass indices relative to it.
UNIT_AIR = Medium("const", (1.0,))


def lsst_v33(band: str = "r", rot_tel_pos: float = 0.0, detector_z_offset: float = 0.0,
             air: Optional[Medium] = None) -> Telescope:
    """The LSST Ver. 3.3 baseline optical design (three mirrors, three lenses, filter), metres.

    What batoid's ``LSST_<band>.yaml`` describes (the file itself is not in the reference tree; at run time
    the telescope comes from the live ``batoid.Optic`` via ``extract.py``).  PINNED for the r band: traced with
    M2 decentred by 100 um at field (1.121, 1.231) deg, 694 nm, the optical path differences of this
    prescription reproduce the Zemax wavefront map that the reference's own test holds
    (/root/reference/tests/data/LSST_WF_v3.3_c3_f6_w3_M2_dx_100um.txt, test_opd.py:16-95) to 0.003 nm RMS /
    0.014 nm max over 28 610 pupil points, and its 28 annular Zernike coefficients to 0.08 nm -- inside that
    test's own tolerances (tests/test_oracle_golden.py::test_opd_zemax, tests/test_gpu_optics.py).
    Asphere coefficients multiply r^4, r^6, ...
    """
    R1, R2, tf = _FILTER[band]
    air = UNIT_AIR if air is None else air
    cam_z = 3.3974725882045593
    items: List[Interface] = []
    I = np.eye(3)

    def cs(z):
        return CoordSys(np.array([0.0, 0.0, z]), I.copy())

    items.append(Interface("M1", Surface("asphere", 19.835, -1.215, (0.0, -1.381e-9)), "mirror", cs(0.0), air, air,
                           _clear_annulus(2.558, 4.18)))
    items.append(Interface("M2", Surface("asphere", 6.788, -0.222, (0.0, 1.274e-5, 9.68e-7)), "mirror",
                           cs(6.1562006), air, air, _clear_annulus(0.9, 1.71)))
    items.append(Interface("M3", Surface("asphere", 8.3445, 0.155, (0.0, 4.5e-7, 8.15e-9)), "mirror",
                           cs(-0.2338), air, air, _clear_annulus(0.55, 2.508)))
    cam = "LSSTCamera"
    z = cam_z
    items.append(Interface("L1_entrance", Surface("sphere", 2.824), "refract", cs(z), air, SILICA,
                           _clear_circle(0.775), cam))
    items.append(Interface("L1_exit", Surface("sphere", 5.021), "refract", cs(z + 0.08223), SILICA, air,
                           _clear_circle(0.775), cam))
    z2 = z + 0.08223 + 0.41264202
    items.append(Interface("L2_entrance", Surface("plane"), "refract", cs(z2), air, SILICA,
                           _clear_circle(0.551), cam))
    items.append(Interface("L2_exit", Surface("asphere", 2.529, -1.57, (0.0, -1.656e-3)), "refract", cs(z2 + 0.030),
                           SILICA, air, _clear_circle(0.551), cam))
    zf = z2 + 0.030 + 0.34958
    items.append(Interface("Filter_entrance", Surface("sphere", R1), "refract", cs(zf), air, SILICA,
                           _clear_circle(0.375), cam))
    items.append(Interface("Filter_exit", Surface("sphere", R2), "refract", cs(zf + tf), SILICA, air,
                           _clear_circle(0.375), cam))
    z3 = zf + 0.0179 + 0.0511
    items.append(Interface("L3_entrance", Surface("quadric", 3.169, -0.962), "refract", cs(z3), air, SILICA,
                           _clear_circle(0.361), cam))
    items.append(Interface("L3_exit", Surface("sphere", -13.36), "refract", cs(z3 + 0.060), SILICA, air,
                           _clear_circle(0.361), cam))
    items.append(Interface("Detector", Surface("plane"), "detector", cs(z3 + 0.060 + 0.0285), air, air,
                           _clear_circle(0.4), cam))
    tel = Telescope(stop=cs(0.4393899), items=items, in_medium=air, name="LSST_" + band)
    if rot_tel_pos != 0.0:
        tel = tel.with_locally_rotated_group(cam, rot_z(rot_tel_pos))
    if detector_z_offset != 0.0:
        tel = tel.with_locally_shifted_item("Detector", [0.0, 0.0, -detector_z_offset])
    return tel


#: pupil geometry of the design (batoid: ``pupilSize``, ``pupilObscuration``, ``sphereRadius``)
LSST_PUPIL_SIZE = 8.36
LSST_PUPIL_OBSCURATION = 0.612
LSST_SPHERE_RADIUS = 5.0

rubin_like = lsst_v33  # round-1 name


def paraboloid_test_telescope(focal_length: float = 10.0) -> Telescope:
    """Single paraboloid mirror with the detector at its focus (physics unit tests)."""
    I = np.eye(3)
    R = 2 * focal_length
    items = [
        Interface("M", Surface("paraboloid", R), "mirror", CoordSys(np.zeros(3), I.copy()), VACUUM, VACUUM, []),
        Interface("D", Surface("plane"), "detector", CoordSys(np.array([0, 0, focal_length], float), I.copy()),
                  VACUUM, VACUUM, []),
    ]
    return Telescope(stop=CoordSys(np.array([0, 0, 2 * focal_length], float), I.copy()), items=items,
                     in_medium=VACUUM, name="paraboloid")
