"""ctypes front-end of the CPU oracle (oracle/_build/liboracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  Never by imsim_b200/.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from imsim_b200 import _abi  # noqa: E402  (POD struct mirror only; no CUDA)

_LIB = None
dp = C.POINTER(C.c_double)
u8p = C.POINTER(C.c_uint8)


def build(force=False):
    so = os.path.join(_HERE, "_build", "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("oracle_optics.c", "oracle_sensor.c")]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "imsim_b200.h"))
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.orc_medium_n.restype = C.c_double
        _LIB.orc_medium_n.argtypes = [C.POINTER(_abi.B2Medium), C.c_double]
        _LIB.orc_sensor_create.restype = C.c_void_p
        _LIB.orc_sensor_accumulate.restype = C.c_double
        _LIB.orc_plain_accumulate.restype = C.c_double
        _LIB.orc_table_spline.restype = C.c_double
    return _LIB


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(dp)


def _opt(a):
    if a is None:
        return None, None
    return _d(a)


def medium_n(medium: _abi.B2Medium, wl_m: float) -> float:
    return lib().orc_medium_n(C.byref(medium), float(wl_m))


def _extras_ptrs(extras):
    """extras: list (per surface) of None | (kind, array) -> two arrays of pointers"""
    n = _abi.B2_MAX_SURFACES
    poly = (dp * n)()
    bic = (dp * n)()
    keep = []
    if extras:
        for i, e in enumerate(extras):
            if e is None:
                continue
            kind, arr = e
            arr, p = _d(arr)
            keep.append(arr)
            if kind == _abi.EXTRA_POLY2D:
                poly[i] = p
            else:
                bic[i] = p
    return poly, bic, keep


def xy_to_v(img, field, x, y):
    x, px = _d(x)
    y, py = _d(y)
    n = x.size
    vx, vy, vz = np.empty(n), np.empty(n), np.empty(n)
    lib().orc_xy_to_v(C.byref(img), C.byref(field), C.c_int64(n), px, py, vx.ctypes.data_as(dp),
                      vy.ctypes.data_as(dp), vz.ctypes.data_as(dp))
    return vx, vy, vz


def v_to_xy(img, field, vx, vy, vz):
    vx, p0 = _d(vx)
    vy, p1 = _d(vy)
    vz, p2 = _d(vz)
    n = vx.size
    x, y = np.empty(n), np.empty(n)
    lib().orc_v_to_xy(C.byref(img), C.byref(field), C.c_int64(n), p0, p1, p2, x.ctypes.data_as(dp),
                      y.ctypes.data_as(dp))
    return x, y


def tansip_fwd(w, x, y):
    ra, dec = C.c_double(), C.c_double()
    lib().orc_tansip_fwd(C.byref(w), C.c_double(x), C.c_double(y), C.byref(ra), C.byref(dec))
    return ra.value, dec.value


def tansip_inv(w, ra, dec):
    x, y = C.c_double(), C.c_double()
    lib().orc_tansip_inv(C.byref(w), C.c_double(ra), C.c_double(dec), C.byref(x), C.byref(y))
    return x.value, y.value


def diffraction(cfg, pu, pv, t, wl_m, gauss, vx, vy, vz):
    """apply_diffraction_delta[_field_rot] on arrays; returns new (vx, vy, vz)."""
    pu, p0 = _d(pu)
    pv, p1 = _d(pv)
    t, p2 = _d(t)
    wl_m, p3 = _d(wl_m)
    gauss, p4 = _d(gauss)
    vx, vy, vz = (np.array(a, dtype=np.float64, copy=True) for a in (vx, vy, vz))
    lib().orc_diffraction(C.byref(cfg), C.c_int64(pu.size), p0, p1, p2, p3, p4, vx.ctypes.data_as(dp),
                          vy.ctypes.data_as(dp), vz.ctypes.data_as(dp))
    return vx, vy, vz


def trace_rays(tel, extras, x, y, z, vx, vy, vz, t, wl_m, vignetted=None, failed=None):
    """batoid Optic.trace restatement; returns new arrays."""
    arrs = [np.array(a, dtype=np.float64, copy=True) for a in (x, y, z, vx, vy, vz, t)]
    n = arrs[0].size
    wl_m, pw = _d(np.broadcast_to(wl_m, (n,)))
    vig = np.zeros(n, np.uint8) if vignetted is None else np.array(vignetted, np.uint8, copy=True)
    fail = np.zeros(n, np.uint8) if failed is None else np.array(failed, np.uint8, copy=True)
    poly, bic, keep = _extras_ptrs(extras)
    lib().orc_trace_rays(C.byref(tel), poly, bic, C.c_int64(n), *[a.ctypes.data_as(dp) for a in arrs], pw,
                         vig.ctypes.data_as(u8p), fail.ctypes.data_as(u8p))
    return (*arrs, vig, fail)


def rubin_optics(tel, extras, img, field, det, dif, opt, x, y, flux, wavelength_nm, pupil_u, pupil_v, time,
                 gauss=None, want_time=False):
    """RubinOptics / RubinDiffractionOptics.applyTo restatement.
    Returns dict(x, y, dxdz, dydz, flux, time_out, stats)."""
    x = np.array(x, np.float64, copy=True)
    y = np.array(y, np.float64, copy=True)
    flux = np.array(flux, np.float64, copy=True)
    n = x.size
    dxdz, dydz = np.empty(n), np.empty(n)
    wl, pwl = _d(wavelength_nm)
    pu, ppu = _d(pupil_u)
    pv, ppv = _d(pupil_v)
    tm, ptm = _d(time)
    g, pg = _opt(gauss)
    tout = np.empty(n) if want_time else None
    stats = _abi.B2OpticsStats()
    poly, bic, keep = _extras_ptrs(extras)
    lib().orc_rubin_optics(C.byref(tel), poly, bic, C.byref(img), C.byref(field), C.byref(det),
                           C.byref(dif) if dif is not None else None, C.byref(opt), C.c_int64(n),
                           x.ctypes.data_as(dp), y.ctypes.data_as(dp), dxdz.ctypes.data_as(dp),
                           dydz.ctypes.data_as(dp), flux.ctypes.data_as(dp), pwl, ppu, ppv, ptm, pg,
                           tout.ctypes.data_as(dp) if want_time else None, C.byref(stats))
    return dict(x=x, y=y, dxdz=dxdz, dydz=dydz, flux=flux, time_out=tout, stats=stats)


def rubin_diffraction(tel, img, field, dif, opt, x, y, wavelength_nm, pupil_u, pupil_v, time, gauss):
    x = np.array(x, np.float64, copy=True)
    y = np.array(y, np.float64, copy=True)
    wl, pwl = _d(wavelength_nm)
    pu, ppu = _d(pupil_u)
    pv, ppv = _d(pupil_v)
    tm, ptm = _d(time)
    g, pg = _d(gauss)
    lib().orc_rubin_diffraction(C.byref(tel), C.byref(img), C.byref(field), C.byref(dif), C.byref(opt),
                                C.c_int64(x.size), x.ctypes.data_as(dp), y.ctypes.data_as(dp), pwl, ppu, ppv, ptm,
                                pg)
    return x, y


def treering_func(A, B, cfreqs, cphases, sfreqs, sphases, r):
    r, pr = _d(r)
    out = np.empty_like(r)
    arrs = [_d(a) for a in (cfreqs, cphases, sfreqs, sphases)]
    lib().orc_treering_func(C.c_double(A), C.c_double(B), C.c_int(len(arrs[0][0])), arrs[0][1], arrs[1][1],
                            arrs[2][1], arrs[3][1], C.c_int64(r.size), pr, out.ctypes.data_as(dp))
    return out


def spline_y2(x, f):
    x, px = _d(x)
    f, pf = _d(f)
    y2 = np.empty_like(x)
    lib().orc_spline_y2(C.c_int(x.size), px, pf, y2.ctypes.data_as(dp))
    return y2


def table_spline(x, f, y2, a):
    x, px = _d(x)
    f, pf = _d(f)
    y2, py2 = _d(y2)
    return lib().orc_table_spline(C.c_int(x.size), px, pf, py2, C.c_double(a))


class Sensor:
    """galsim.SiliconSensor restatement bound to one numpy image."""

    def __init__(self, cfg: _abi.B2SensorConfig, vertex_data, tr_r=None, tr_f=None, tr_spline=True, abs_w=None,
                 abs_l=None):
        self.cfg = cfg
        v, pv = _d(vertex_data)
        self._keep = [v]
        ptr, pf = (None, None), (None, None)
        if tr_r is not None and cfg.n_treering > 2:
            ptr = _d(tr_r)
            pf = _d(tr_f)
        pw, pl = (None, None), (None, None)
        if abs_w is not None:
            pw = _d(abs_w)
            pl = _d(abs_l)
        self._keep += [ptr[0], pf[0], pw[0], pl[0]]
        self._h = C.c_void_p(lib().orc_sensor_create(C.byref(cfg), pv, ptr[1], pf[1], C.c_int(int(tr_spline)),
                                                     pw[1], pl[1]))
        self.image = None

    def bind_image(self, image: np.ndarray, xmin=0, ymin=0):
        assert image.flags.c_contiguous and image.dtype in (np.float32, np.float64)
        self.image = image
        ny, nx = image.shape
        lib().orc_sensor_bind_image(self._h, C.c_int(xmin), C.c_int(ymin), C.c_int(nx), C.c_int(ny),
                                    C.c_int(image.dtype.itemsize), image.ctypes.data_as(C.c_void_p))

    def accumulate(self, x, y, flux, rand4, dxdz=None, dydz=None, wavelength=None, orig_center=(0, 0),
                   resume=False, recalc=False):
        x, px = _d(x)
        y, py = _d(y)
        flux, pf = _d(flux)
        rand4, pr = _d(rand4)
        assert rand4.size == 4 * x.size
        a, pa = _opt(dxdz)
        b, pb = _opt(dydz)
        w, pw = _opt(wavelength)
        st = _abi.B2AccumStats()
        added = lib().orc_sensor_accumulate(self._h, C.c_int64(x.size), px, py, pa, pb, pw, pf, pr,
                                            C.c_int(orig_center[0]), C.c_int(orig_center[1]), C.c_int(int(resume)),
                                            C.c_int(int(recalc)), C.byref(st))
        return added, st

    def pixel_areas(self, orig_center=(0, 0), use_flux=True):
        ny, nx = self.image.shape
        areas = np.empty((ny, nx))
        lib().orc_sensor_pixel_areas(self._h, C.c_int(orig_center[0]), C.c_int(orig_center[1]),
                                     C.c_int(int(use_flux)), areas.ctypes.data_as(dp))
        return areas

    def get_pixel(self, ix, iy):
        npoly = 4 * self.cfg.num_vertices + 4
        poly = np.empty((npoly, 2))
        bounds = np.empty(8)
        lib().orc_sensor_get_pixel(self._h, C.c_int(ix), C.c_int(iy), poly.ctypes.data_as(dp),
                                   bounds.ctypes.data_as(dp))
        return poly, bounds

    def __del__(self):
        try:
            lib().orc_sensor_destroy(self._h)
        except Exception:
            pass


def plain_accumulate(image, x, y, flux, xmin=0, ymin=0):
    x, px = _d(x)
    y, py = _d(y)
    flux, pf = _d(flux)
    ny, nx = image.shape
    return lib().orc_plain_accumulate(C.c_int(xmin), C.c_int(ymin), C.c_int(nx), C.c_int(ny),
                                      C.c_int(image.dtype.itemsize), image.ctypes.data_as(C.c_void_p),
                                      C.c_int64(x.size), px, py, pf)


def set_threads(n: int):
    lib().orc_set_threads(C.c_int(int(n)))


def max_threads() -> int:
    return int(lib().orc_max_threads())
