"""Extractors: flatten live reference objects into the POD descriptions of the
B200 path.

At run time this package sits beside an installed GalSim + batoid + LSST stack
(it is a GalSim config plugin), and the reference hands *objects* across the
photon-op boundary: a ``batoid.Optic`` (``base['det_telescope']``), two
``galsim.GSFitsWCS`` and an ``lsst.afw.cameraGeom.Detector``
(imsim/photon_ops.py:400-451).  Each function here reads only public attributes
of those objects, once per detector, and never does per-photon work.

None of these libraries exists in the build container, so this module is
exercised only by duck-typed fakes in tests/test_extract.py; UNVERIFIED against
the live libraries.
"""
from __future__ import annotations

import numpy as np

from .detector import DetectorGeometry
from .telescope import CoordSys, Interface, Medium, Obscuration, Surface, Telescope
from .wcs import TanSipWCS


class ExtractError(TypeError):
    pass


def _cls(obj) -> str:
    return type(obj).__name__


def medium_from_batoid(m) -> Medium:
    name = _cls(m)
    if name == "ConstMedium":
        return Medium("const", (float(m.n),))
    if name == "SellmeierMedium":
        return Medium("sellmeier", tuple(float(c) for c in m.coefs))
    if name == "SumitaMedium":
        return Medium("sumita", tuple(float(c) for c in m.coefs))
    if name == "Air":
        return Medium("air", (float(m.pressure), float(m.temperature), float(m.h2o_pressure)))
    raise ExtractError("unsupported batoid medium %s" % name)


def _obsc_list(ob, negate=False):
    """Flatten a batoid obscuration into OR-ed primitives (with negation flags)."""
    if ob is None:
        return []
    name = _cls(ob)
    if name in ("ObscNegation", "ClearCircle", "ClearAnnulus", "ClearRectangle") and hasattr(ob, "original"):
        inner = _obsc_list(ob.original, not negate)
        if len(inner) != 1:
            raise ExtractError("negation of a compound obscuration is not supported")
        return inner
    if name == "ObscCircle":
        return [Obscuration("circle", (float(ob.radius), float(ob.x), float(ob.y)), negate)]
    if name == "ObscAnnulus":
        return [Obscuration("annulus", (float(ob.inner), float(ob.outer), float(ob.x), float(ob.y)), negate)]
    if name == "ObscRectangle":
        return [Obscuration("rectangle", (float(ob.width), float(ob.height), float(ob.x), float(ob.y),
                                          float(ob.theta)), negate)]
    if name == "ObscRay":
        return [Obscuration("ray", (float(ob.width), float(ob.theta), float(ob.x), float(ob.y)), negate)]
    if name == "ObscUnion" and not negate:
        out = []
        for it in ob.items:
            out.extend(_obsc_list(it, False))
        return out
    raise ExtractError("unsupported batoid obscuration %s%s" % ("negated " if negate else "", name))


def _zernike_xy(z):
    """xy-polynomial coefficient array of a batoid.Zernike surface."""
    for attr in ("_xycoef", "xycoef"):
        if hasattr(z, attr):
            return np.array(getattr(z, attr), dtype=float), 1.0
    import galsim.zernike  # noqa: PLC0415

    gz = galsim.zernike.Zernike(z.coef, R_outer=z.R_outer, R_inner=z.R_inner)
    return np.array(gz._coef_array_xy, dtype=float), 1.0


def surface_from_batoid(s) -> Surface:
    name = _cls(s)
    if name == "Plane":
        return Surface("plane")
    if name == "Sphere":
        return Surface("sphere", float(s.R))
    if name == "Paraboloid":
        return Surface("paraboloid", float(s.R))
    if name == "Quadric":
        return Surface("quadric", float(s.R), float(s.conic))
    if name == "Asphere":
        return Surface("asphere", float(s.R), float(s.conic), tuple(float(c) for c in s.coefs))
    if name == "Sum":
        parts = list(s.surfaces)
        base = surface_from_batoid(parts[0])
        if base.poly is not None or base.bicubic is not None:
            raise ExtractError("first member of a Sum surface must be a plain conic/asphere")
        for extra in parts[1:]:
            en = _cls(extra)
            if en == "Zernike":
                c, scale = _zernike_xy(extra)
                n = max(c.shape)
                sq = np.zeros((n, n))
                sq[: c.shape[0], : c.shape[1]] = c
                if base.poly is None:
                    base.poly, base.poly_scale = sq, scale
                else:
                    m = max(n, base.poly.shape[0])
                    tot = np.zeros((m, m))
                    tot[: base.poly.shape[0], : base.poly.shape[1]] += base.poly
                    tot[:n, :n] += sq
                    base.poly = tot
            elif en == "Bicubic":
                if base.bicubic is not None:
                    raise ExtractError("only one Bicubic term per surface is supported")
                base.bicubic = dict(xs=np.array(extra.xs), ys=np.array(extra.ys), zs=np.array(extra.zs),
                                    dzdxs=np.array(extra.dzdxs), dzdys=np.array(extra.dzdys),
                                    d2zdxdys=np.array(extra.d2zdxdys))
            else:
                raise ExtractError("unsupported Sum member %s" % en)
        if base.poly is not None and base.bicubic is not None:
            raise ExtractError("Zernike and Bicubic on the same surface are not supported yet")
        return base
    raise ExtractError("unsupported batoid surface %s" % name)


def _coordsys(cs) -> CoordSys:
    return CoordSys(np.array(cs.origin, dtype=float), np.array(cs.rot, dtype=float))


def _walk(optic, out):
    if getattr(optic, "skip", False):
        return
    if hasattr(optic, "items"):
        for it in optic.items:
            _walk(it, out)
        return
    name = _cls(optic)
    if name == "OPDScreen":
        # batoid.OPDScreen(surface=Plane(), screen=Zernike | Bicubic, ...) as telescope_loader inserts it
        # (tests/test_telescope_loader.py:641-653): a 'pass' interface whose summed term is the OPD
        if _cls(optic.surface) != "Plane":
            raise ExtractError("OPDScreen is supported on a Plane surface only")
        screen, sn = optic.screen, _cls(optic.screen)
        surf = Surface("plane")
        if sn == "Zernike":
            c, scale = _zernike_xy(screen)
            n = max(c.shape)
            surf.poly = np.zeros((n, n))
            surf.poly[: c.shape[0], : c.shape[1]] = c
            surf.poly_scale = scale
        elif sn == "Bicubic":
            surf.bicubic = dict(xs=np.array(screen.xs), ys=np.array(screen.ys), zs=np.array(screen.zs),
                                dzdxs=np.array(screen.dzdxs), dzdys=np.array(screen.dzdys),
                                d2zdxdys=np.array(screen.d2zdxdys))
        elif sn != "Plane":
            raise ExtractError("unsupported OPDScreen screen %s" % sn)
        out.append(Interface(name=str(optic.name), surface=surf, interact="pass", coord_sys=_coordsys(optic.coordSys),
                             in_medium=medium_from_batoid(optic.inMedium), out_medium=medium_from_batoid(optic.outMedium),
                             obscurations=_obsc_list(getattr(optic, "obscuration", None))))
        return
    kinds = {"Mirror": "mirror", "RefractiveInterface": "refract", "Detector": "detector", "Baffle": "detector"}
    if name not in kinds:
        raise ExtractError("unsupported batoid optic %s (%s)" % (name, getattr(optic, "name", "?")))
    out.append(Interface(
        name=str(optic.name), surface=surface_from_batoid(optic.surface), interact=kinds[name],
        coord_sys=_coordsys(optic.coordSys), in_medium=medium_from_batoid(optic.inMedium),
        out_medium=medium_from_batoid(optic.outMedium), obscurations=_obsc_list(getattr(optic, "obscuration", None))))


def telescope_from_batoid(optic) -> Telescope:
    """Flatten a ``batoid.CompoundOptic`` (after all ``with*`` perturbations, e.g.
    ``base['det_telescope']``, imsim/telescope_loader.py:399-415,463)."""
    stop = optic.stopSurface
    if _cls(stop.surface) != "Plane":
        raise ExtractError("stop surface must be a Plane (imsim/photon_ops.py:108 evaluates its sag)")
    items = []
    _walk(optic, items)
    if not items or items[-1].interact != "detector":
        raise ExtractError("the last interface must be the Detector")
    return Telescope(stop=_coordsys(stop.coordSys), items=items, in_medium=medium_from_batoid(optic.inMedium),
                     name=str(getattr(optic, "name", "telescope")))


def tansip_from_galsim(wcs) -> TanSipWCS:
    """``galsim.GSFitsWCS`` / ``FittedSIPWCS`` -> TanSipWCS."""
    if getattr(wcs, "wcs_type", "TAN") not in ("TAN", "TAN-SIP"):
        raise ExtractError("only TAN / TAN-SIP WCS are supported, got %s" % wcs.wcs_type)
    if getattr(wcs, "pv", None) is not None:
        raise ExtractError("TPV distortions are not supported")
    ab = getattr(wcs, "ab", None)
    order = 0
    abm = np.zeros((2, 4, 4))
    if ab is not None:
        ab = np.array(ab, dtype=float)
        order = ab.shape[1] - 1
        if order > 3:
            raise ExtractError("SIP order %d > 3" % order)
        abm[:, : order + 1, : order + 1] = ab
    c = wcs.center
    return TanSipWCS(crpix=np.array(wcs.crpix, dtype=float), cd=np.array(wcs.cd, dtype=float),
                     center=(float(c.ra.rad), float(c.dec.rad)), ab=abm, order=order)


def detector_from_lsst(det, z_offset=None) -> DetectorGeometry:
    """Probe the FOCAL_PLANE -> PIXELS transform of an ``lsst.afw.cameraGeom.Detector``
    with three points (it is affine; imsim/utils.py:42-78)."""
    from lsst.afw import cameraGeom  # noqa: PLC0415

    tx = det.getTransform(cameraGeom.FOCAL_PLANE, cameraGeom.PIXELS).getMapping()
    pts = np.array([[0.0, 1.0, 0.0], [0.0, 0.0, 1.0]])
    x, y = tx.applyForward(pts)
    b = np.array([x[0], y[0]])
    A = np.array([[x[1] - x[0], x[2] - x[0]], [y[1] - y[0], y[2] - y[0]]])
    bbox = det.getBBox()
    return DetectorGeometry(det.getName(), A, b, nx=bbox.getWidth(), ny=bbox.getHeight(),
                            z_offset=0.0 if z_offset is None else float(z_offset), xmin=bbox.getMinX(),
                            ymin=bbox.getMinY())
