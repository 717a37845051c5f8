"""B200 silicon sensor with the ``galsim.SiliconSensor`` interface.

imSim builds its sensor through GalSim's config (``image.sensor.type: Silicon``,
config/imsim-config.yaml:230-235) with the vendor-specific model name chosen at
imsim/lsst_image.py:93-103, and then calls
``sensor.accumulate(photons, image, orig_center, resume, recalc)``
(imsim/photon_pooling.py:210, imsim/stamp.py:562-572, imsim/flat.py:261),
``sensor.calculate_pixel_areas(image)`` (imsim/flat.py:223) and
``sensor.updateRNG(rng)`` (imsim/photon_pooling.py:71).  This class keeps those
signatures and runs the arithmetic in csrc/sensor.cu.

Random numbers: GalSim draws from one serial mt19937 stream; the device path
uses counter-based Philox keyed on (seed, global photon index), so realisations
differ from GalSim's while the per-photon arithmetic is identical (the parity
tests inject identical draws on both sides through ``rand4``).
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

from . import _abi, _lib
from .context import OpticsContext
from .treerings import RadialTable

#: numpy layout of ``B2StampJob`` (include/imsim_b200.h)
STAMP_JOB_DTYPE = np.dtype([("p0", "<i8"), ("n", "<i8"), ("xmin", "<i4"), ("ymin", "<i4"), ("nx", "<i4"),
                            ("ny", "<i4"), ("plain", "<i4"), ("pad", "<i4")])
assert STAMP_JOB_DTYPE.itemsize == C.sizeof(_abi.B2StampJob)


class Image:
    """Minimal ``galsim.Image`` stand-in: ``array`` (ny, nx) + integer origin."""

    def __init__(self, array: np.ndarray, xmin: int = 1, ymin: int = 1):
        self.array = array
        self.xmin, self.ymin = int(xmin), int(ymin)

    @property
    def dtype(self):
        return self.array.dtype

    @property
    def bounds(self):
        return self

    @property
    def xmax(self):
        return self.xmin + self.array.shape[1] - 1

    @property
    def ymax(self):
        return self.ymin + self.array.shape[0] - 1


def _image_parts(image):
    """(array, xmin, ymin) of a galsim.Image or our Image."""
    arr = image.array
    b = image.bounds
    return arr, int(b.xmin), int(b.ymin)


def read_config_file(filename: str) -> dict:
    """Parse a Poisson_CCD ``.cfg`` file: ``key = value  # comment``
    (galsim/sensor.py ``_read_config_file``)."""
    config = {}
    with open(filename, 'r') as f:
        for line in f:
            line = line.split('#', 1)[0].strip()
            if '=' not in line:
                continue
            key, val = (t.strip() for t in line.split('=', 1))
            toks = val.split()
            if not toks:
                continue

            def conv(t):
                try:
                    return int(t)
                except ValueError:
                    try:
                        return float(t)
                    except ValueError:
                        return t

            vals = [conv(t) for t in toks]
            config[key] = vals[0] if len(vals) == 1 else vals
    return config


def calculate_diff_step(config: dict) -> float:
    """Diffusion step [microns] for conversion at the entrance surface
    (galsim/sensor.py ``_calculate_diff_step`` as quoted in
    doc/validation/diffusion.rst "Since Aug 8")."""
    NumPhases = config['NumPhases']
    CollectingPhases = config['CollectingPhases']
    PixelSize = config['PixelSizeX']
    SensorThickness = config['SensorThickness']
    ChannelStopWidth = config['ChannelStopWidth']
    FieldOxideTaper = config['FieldOxideTaper']
    Vbb = config['Vbb']
    Vparallel_lo = config['Vparallel_lo']
    Vparallel_hi = config['Vparallel_hi']
    CCDTemperature = config['CCDTemperature']
    qfh = config['qfh']
    VChannelStop = qfh  # near zero
    VCollect = Vparallel_hi + 12.0  # estimate from simulation
    VBarrier = Vparallel_lo + 15.0  # estimate from simulation
    ChannelStopRegionWidth = 2.0 * (ChannelStopWidth / 2.0 + FieldOxideTaper)
    ChannelStopRegionArea = ChannelStopRegionWidth * PixelSize
    CollectArea = (PixelSize - ChannelStopRegionWidth) * PixelSize * CollectingPhases / NumPhases
    BarrierArea = (PixelSize - ChannelStopRegionWidth) * PixelSize * (NumPhases - CollectingPhases) / NumPhases
    Vfront = (ChannelStopRegionArea * VChannelStop + CollectArea * VCollect + BarrierArea * VBarrier) / (PixelSize**2)
    Vdiff = max(Vfront - Vbb, 1.0)
    MobilityFactor = 0.27  # Green et al.
    # 0.026 is kT/q at room temperature (298 K)
    return float(np.sqrt(2 * 0.026 * CCDTemperature / 298.0 / Vdiff / MobilityFactor) * SensorThickness)


def synthetic_absorption_table():
    """SYNTHETIC silicon absorption lengths [nm -> microns] (room temperature, after
    Green 2008), log-interpolated to a 5 nm grid.  GalSim's own table
    (share/sensors/absorption.dat) is not part of the reference tree; at run time
    beside GalSim the real one is used (``find_absorption_table``)."""
    wl = np.array([250, 300, 350, 400, 450, 500, 550, 600, 650, 700, 750, 800, 850, 900, 950, 1000, 1050, 1100,
                   1150, 1200, 1450], float)
    alpha_cm = np.array([1.84e6, 1.73e6, 1.04e6, 9.52e4, 2.55e4, 1.11e4, 6.39e3, 4.14e3, 2.81e3, 1.90e3, 1.30e3,
                         8.50e2, 5.35e2, 3.06e2, 1.57e2, 6.4e1, 1.63e1, 3.5, 0.68, 0.22, 3.2e-8])
    grid = np.arange(255.0, 1450.0 + 1e-9, 5.0)
    length_um = 1e4 / np.exp(np.interp(grid, wl, np.log(alpha_cm)))
    return grid, length_um


def find_absorption_table():
    """(wavelength_nm, abs_length_um, provenance)."""
    try:
        import galsim  # noqa: PLC0415

        fn = os.path.join(galsim.meta_data.share_dir, 'sensors', 'absorption.dat')
        data = np.loadtxt(fn, skiprows=1)
        return np.ascontiguousarray(data[:, 0]), np.ascontiguousarray(data[:, 1]), fn
    except Exception:
        w, l_ = synthetic_absorption_table()
        return w, l_, "synthetic"


def find_sensor_files(name: str, data_dir: Optional[str] = None):
    """Locate ``<name>.cfg`` / ``<name>.dat`` like galsim/sensor.py: as given, then
    in the imSim data dir, then in GalSim's share dir."""
    cands = [name]
    for d in filter(None, [data_dir, os.environ.get("IMSIM_DATA_DIR")]):
        cands.append(os.path.join(d, 'sensor_models', name))
    try:
        import galsim  # noqa: PLC0415

        cands.append(os.path.join(galsim.meta_data.share_dir, 'sensors', name))
    except (ImportError, AttributeError):
        pass
    cands.append(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data', name))
    for c in cands:
        if os.path.isfile(c + '.cfg') and os.path.isfile(c + '.dat'):
            return c + '.cfg', c + '.dat'
    raise OSError("Cannot locate sensor model files for %s" % name)


def _seed_from_rng(rng) -> int:
    if rng is None:
        return int.from_bytes(os.urandom(8), 'little')
    if isinstance(rng, (int, np.integer)):
        return int(rng)
    if hasattr(rng, 'raw'):  # galsim.BaseDeviate
        return (int(rng.raw()) << 32) | int(rng.raw())
    raise TypeError("rng must be None, an int seed or a galsim.BaseDeviate")


class SiliconSensor:
    """Drop-in for ``galsim.SiliconSensor`` running on one B200.

    Parameters are GalSim's (name, strength, rng, diffusion_factor, qdist, nrecalc,
    treering_func, treering_center, transpose).  Extra keyword-only arguments select
    the device context and data locations.
    """

    def __init__(self, name='lsst_itl_50_8', strength=1.0, rng=None, diffusion_factor=1.0, qdist=3, nrecalc=10000,
                 treering_func=None, treering_center=(0.0, 0.0), transpose=False, *, context: OpticsContext = None,
                 device: int = 0, stream=None, data_dir: Optional[str] = None, absorption_table=None,
                 config: Optional[dict] = None, vertex_data: Optional[np.ndarray] = None):
        self.name = name
        self.strength = float(strength)
        self.diffusion_factor = float(diffusion_factor)
        self.qdist = int(qdist)
        self.nrecalc = float(nrecalc)
        self.treering_func = treering_func
        self.treering_center = treering_center
        self.transpose = bool(transpose)
        self._last_image = None
        self._seed = _seed_from_rng(rng)
        self._photon_offset = 0

        if config is None or vertex_data is None:
            self.config_file, self.vertex_file = find_sensor_files(name, data_dir)
            self.config = read_config_file(self.config_file)
            vertex_data = np.loadtxt(self.vertex_file, skiprows=1)
        else:
            self.config = config
        cfgd = self.config
        nv = int(cfgd['NumVertices'])
        Nx, Ny = int(cfgd['PixelBoundaryNx']), int(cfgd['PixelBoundaryNy'])
        vertex_data = np.ascontiguousarray(vertex_data, dtype=np.float64)
        if vertex_data.shape != (Nx * Ny * (4 * nv + 4), 5):
            raise OSError("Vertex file %s does not match config file" % name)
        self.vertex_data = vertex_data
        self.diff_step = calculate_diff_step(cfgd) * self.diffusion_factor

        if absorption_table is None:
            aw, al, self.absorption_provenance = find_absorption_table()
        else:
            aw, al = absorption_table
            self.absorption_provenance = "user"
        self.abs_wave = np.ascontiguousarray(aw, dtype=np.float64)
        self.abs_len = np.ascontiguousarray(al, dtype=np.float64)

        tr = self._treering_arrays(treering_func)
        cx, cy = (treering_center.x, treering_center.y) if hasattr(treering_center, 'x') else treering_center

        pod = _abi.B2SensorConfig()
        pod.num_vertices = nv
        pod.nx, pod.ny = Nx, Ny
        pod.qdist = self.qdist
        pod.num_elec = float(cfgd['CollectedCharge_0_0']) / self.strength
        pod.nrecalc = self.nrecalc / self.strength  # scaled like GalSim (matters for strength >> 1)
        pod.diff_step = self.diff_step
        pod.pixel_size = float(cfgd['PixelSizeX'])
        pod.sensor_thickness = float(cfgd['SensorThickness'])
        pod.treering_center[0], pod.treering_center[1] = float(cx), float(cy)
        pod.n_treering = 0 if tr is None else len(tr[0])
        pod.n_abs = len(self.abs_wave)
        pod.transpose = int(self.transpose)
        self.pod = pod
        self._tr = tr

        self._own_ctx = context is None
        self.ctx = context if context is not None else OpticsContext(device=device, stream=stream)
        self._lib = _lib.load()
        self._h = C.c_void_p()
        self._create_handle()
        self.last_stats = None

    def _create_handle(self):
        tr = self._tr
        tr_r = tr[0].ctypes.data if tr is not None else None
        tr_f = tr[1].ctypes.data if tr is not None else None
        tr_y2 = tr[2].ctypes.data if (tr is not None and tr[2] is not None) else None
        _lib.check(self._lib.b2_sensor_create(self.ctx.handle, C.byref(self.pod), self.vertex_data.ctypes.data, tr_r,
                                              tr_f, tr_y2, self.abs_wave.ctypes.data, self.abs_len.ctypes.data,
                                              C.byref(self._h)))
        self._bound_shape = None
        self._last_image = None

    def move_to(self, context: OpticsContext):
        """Re-create the device side of this sensor on another context (the fused pooled step needs the sensor
        and the optics on one context / stream).  Tables are uploaded again; a bound image is dropped."""
        if context is self.ctx:
            return self
        self.close()
        self.ctx, self._own_ctx = context, False
        self._create_handle()
        return self

    @staticmethod
    def _treering_arrays(func):
        """(r, f, y2|None) of a tree-ring lookup table, or None for "no tree rings"
        (GalSim's sentinel is a 2-point table)."""
        if func is None:
            return None
        if isinstance(func, RadialTable):
            x, f, y2 = func.x, func.f, func.y2
        elif hasattr(func, 'x') and hasattr(func, 'f'):  # galsim.LookupTable
            x = np.ascontiguousarray(func.x, dtype=np.float64)
            f = np.ascontiguousarray(func.f, dtype=np.float64)
            interp = getattr(func, 'interpolant', 'spline')
            if interp == 'spline':
                from .treerings import natural_spline_y2

                y2 = natural_spline_y2(x, f)
            elif interp == 'linear':
                y2 = None
            else:
                raise ValueError("treering_func interpolant must be 'spline' or 'linear'")
        else:
            raise TypeError("treering_func must be a LookupTable-like object with x and f")
        if len(x) <= 2:
            return None
        return (np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(f, dtype=np.float64),
                None if y2 is None else np.ascontiguousarray(y2, dtype=np.float64))

    def set_treerings(self, treering_func, treering_center=(0.0, 0.0)):
        """Switch to another detector's tree rings, keeping the device allocations."""
        tr = self._treering_arrays(treering_func)
        cx, cy = (treering_center.x, treering_center.y) if hasattr(treering_center, 'x') else treering_center
        self.treering_func, self.treering_center, self._tr = treering_func, treering_center, tr
        n = 0 if tr is None else len(tr[0])
        _lib.check(self._lib.b2_sensor_set_treerings(
            self._h, float(cx), float(cy), tr[0].ctypes.data if tr else None, tr[1].ctypes.data if tr else None,
            tr[2].ctypes.data if (tr and tr[2] is not None) else None, n))
        self.pod.treering_center[0], self.pod.treering_center[1] = float(cx), float(cy)
        self.pod.n_treering = n
        self._last_image = None

    @classmethod
    def simple_treerings(cls, amplitude=0.5, period=100., r_max=8000., dr=None):
        """``galsim.SiliconSensor.simple_treerings``: a cosine tree-ring table (used by
        tests/test_flats.py:135 of the reference)."""
        k = 2. * np.pi / float(period)
        if dr is None:
            dr = period / 100.
        npoints = int(r_max / dr) + 1
        return RadialTable.from_func(lambda r: amplitude * np.cos(k * r), x_min=0., x_max=r_max, npoints=npoints)

    # -- GalSim API ---------------------------------------------------------
    def updateRNG(self, rng):
        self._seed = _seed_from_rng(rng)
        self._photon_offset = 0

    def _bind(self, image):
        arr, xmin, ymin = _image_parts(image)
        if arr.dtype not in (np.float32, np.float64) or not arr.flags.c_contiguous:
            raise _lib.B2Error("SiliconSensor needs a C-contiguous float32/float64 image")
        ny, nx = arr.shape
        _lib.check(self._lib.b2_sensor_bind_image(self._h, xmin, ymin, nx, ny, arr.dtype.itemsize, arr.ctypes.data,
                                                  _abi.B2_HOST))
        self._bound_shape = (ny, nx, arr.dtype)

    def bind_stamp(self, xmin: int, ymin: int, nx: int, ny: int, dtype=np.float32):
        """Bind an all-zero image of the given bounds that lives on the device only (per-object stamps of the
        classic pipeline, imsim/stamp.py:562-572): no host array, no synchronisation.  Follow with
        ``accumulate(..., image=<any token>, prebound=True)`` and ``snapshot_image``."""
        _lib.check(self._lib.b2_sensor_bind_image(self._h, int(xmin), int(ymin), int(nx), int(ny),
                                                  np.dtype(dtype).itemsize, None, _abi.B2_DEVICE))
        self._bound_shape = (ny, nx, np.dtype(dtype))
        self._last_image = None

    def plain_accumulate_bound(self, photons):
        """``sensor=None`` drawing (galsim.Sensor: bin by nominal pixel) onto the bound device image -- what the
        reference does for objects below ``max_flux_simple`` (imsim/stamp.py:534-537,555-556)."""
        added = C.c_double(0.0)
        _lib.check(self._lib.b2_plain_accumulate(self._h, len(photons), _lib.ptr(photons.x), _lib.ptr(photons.y),
                                                 _lib.ptr(photons.flux), _lib.where_of(photons.x), C.byref(added)))
        return added.value

    def accumulate_stamps(self, jobs, photons, full, full_xmin=0, full_ymin=0, orig_center=(0, 0), rand4=None,
                          want_stats=True, want_added=False):
        """The object loop of the classic pipeline in one call (``b2_sensor_accumulate_stamps``): every job
        ``(p0, n, xmin, ymin, nx, ny, plain)`` gets its own zero stamp with fresh boundaries, its photons
        ``photons[p0:p0+n]`` are accumulated there at this sensor's ``nrecalc`` cadence -- the loop runs on the
        device, one thread block per stamp -- and the stamp is added to ``full`` (a 2-d CUDA tensor, float32 or
        float64, whose pixel (0, 0) is image pixel (full_xmin, full_ymin)).  ``photons``: ``DevicePhotons``.
        Returns the stats (``added_flux`` summed over the stamps) and, if asked, the flux per job."""
        import torch

        if isinstance(jobs, np.ndarray) and jobs.dtype == STAMP_JOB_DTYPE:
            arr = np.ascontiguousarray(jobs)
        else:
            jobs = list(jobs)
            arr = np.zeros(len(jobs), dtype=STAMP_JOB_DTYPE)
            for k, j in enumerate(jobs):
                arr[k] = tuple(int(v) for v in j[:6]) + (int(bool(j[6])) if len(j) > 6 else 0, 0)
        jobs = arr
        if not (full.is_cuda and full.dim() == 2 and full.is_contiguous() and full.dtype in (torch.float32, torch.float64)):
            raise _lib.B2Error("accumulate_stamps needs a contiguous 2-d float32 / float64 CUDA tensor as full image")
        n = len(photons)
        has_ang = photons.hasAllocatedAngles()
        has_wl = photons.hasAllocatedWavelengths()
        stats = _abi.B2AccumStats() if want_stats else None
        added = np.zeros(max(len(jobs), 1)) if want_added else None
        _lib.check(self._lib.b2_sensor_accumulate_stamps(
            self._h, len(jobs), C.c_void_p(arr.ctypes.data if len(jobs) else None), n, _lib.ptr(photons.x), _lib.ptr(photons.y),
            _lib.ptr(photons.dxdz) if has_ang else None, _lib.ptr(photons.dydz) if has_ang else None,
            _lib.ptr(photons.wavelength) if has_wl else None, _lib.ptr(photons.flux), _lib.ptr(rand4),
            self._seed & 0xFFFFFFFFFFFFFFFF, self._photon_offset, int(orig_center[0]), int(orig_center[1]),
            C.c_void_p(full.data_ptr()), int(full_xmin), int(full_ymin), int(full.shape[1]), int(full.shape[0]),
            full.element_size(), C.byref(stats) if want_stats else None,
            added.ctypes.data if want_added else None))
        self._photon_offset += n
        self.last_stats = stats
        return (stats, added[:len(jobs)]) if want_added else stats

    def accumulate(self, photons, image, orig_center=None, resume=False, recalc=False, rand4=None,
                   sync_image=True, want_stats=True, prebound=False):
        """Accumulate photons on the image; returns the flux that landed on it.

        ``photons``: a (GalSim or imsim_b200) ``PhotonArray`` on the host, or a
        ``DevicePhotons`` whose fields are CUDA tensors.  ``rand4`` (testing):
        injected draws ``[g1, g2, u_notfound, u_depth]``, shape (4, N)."""
        if resume and image is not self._last_image:
            raise _lib.B2Error("image must be the same as used for the last accumulate call if resume is True")
        ocx, ocy = (0, 0) if orig_center is None else (int(orig_center.x), int(orig_center.y)) \
            if hasattr(orig_center, 'x') else (int(orig_center[0]), int(orig_center[1]))
        n = photons.size() if hasattr(photons, 'size') and callable(photons.size) else len(photons)
        if not resume and not prebound:
            self._bind(image)
        # only a successful bind makes this the image a later resume=True may continue on; an empty first
        # batch still initialises the boundaries from the image (GalSim has no early-out either)
        self._last_image = image
        if n == 0:
            if resume and not recalc:
                return 0.0
            _lib.check(self._lib.b2_sensor_accumulate(
                self._h, 0, None, None, None, None, None, None, None, self._seed & 0xFFFFFFFFFFFFFFFF,
                self._photon_offset, ocx, ocy, int(bool(resume)), int(bool(recalc)), _abi.B2_HOST, None))
            if sync_image:
                self.read_image(image)
            return 0.0
        x, y, flux = photons.x, photons.y, photons.flux
        where = _lib.where_of(x)
        dxdz = photons.dxdz if photons.hasAllocatedAngles() else None
        dydz = photons.dydz if photons.hasAllocatedAngles() else None
        wl = photons.wavelength if photons.hasAllocatedWavelengths() else None
        if rand4 is not None:
            rand4 = np.ascontiguousarray(rand4, dtype=np.float64) if where == _abi.B2_HOST else rand4
        stats = _abi.B2AccumStats() if want_stats else None
        _lib.check(self._lib.b2_sensor_accumulate(
            self._h, n, _lib.ptr(x), _lib.ptr(y), _lib.ptr(dxdz), _lib.ptr(dydz), _lib.ptr(wl), _lib.ptr(flux),
            _lib.ptr(rand4), self._seed & 0xFFFFFFFFFFFFFFFF, self._photon_offset, ocx, ocy, int(bool(resume)),
            int(bool(recalc)), where, C.byref(stats) if want_stats else None))
        self._photon_offset += n
        self.last_stats = stats
        if sync_image:
            self.read_image(image)
        return stats.added_flux if want_stats else None

    def read_image(self, image):
        """Copy the device image into ``image.array``."""
        arr, _, _ = _image_parts(image)
        _lib.check(self._lib.b2_sensor_read_image(self._h, arr.ctypes.data, _abi.B2_HOST))

    def snapshot_image(self, device_tensor):
        """Device-to-device copy of the current image (target + pending charge) into a CUDA tensor
        of the image's dtype, on the context's stream (no host synchronisation)."""
        _lib.check(self._lib.b2_sensor_read_image(self._h, C.c_void_p(device_tensor.data_ptr()), _abi.B2_DEVICE))

    def calculate_pixel_areas(self, image, orig_center=(0, 0), use_flux=True):
        """Areas of the (tree-ring and, if ``use_flux``, charge-) distorted pixels.
        Returns 1.0 when trivially undistorted, like GalSim (imsim/flat.py:228 checks)."""
        arr, xmin, ymin = _image_parts(image)
        if self._tr is None and (not use_flux or not np.any(arr)):
            return 1.0
        ocx, ocy = (int(orig_center.x), int(orig_center.y)) if hasattr(orig_center, 'x') else \
            (int(orig_center[0]), int(orig_center[1]))
        self._bind(image)
        self._last_image = None
        areas = np.empty(arr.shape, dtype=np.float64)
        _lib.check(self._lib.b2_sensor_pixel_areas(self._h, ocx, ocy, int(bool(use_flux)), areas.ctypes.data,
                                                   _abi.B2_HOST))
        try:
            import galsim  # noqa: PLC0415

            return galsim.ImageD(areas, xmin=xmin, ymin=ymin)
        except ImportError:
            return Image(areas, xmin, ymin)

    def get_pixel(self, ix, iy):
        """(polygon[4nv+4, 2], bounds[8]) of one pixel's current boundary (debug / tests)."""
        npoly = 4 * self.pod.num_vertices + 4
        poly = np.empty((npoly, 2))
        bounds = np.empty(8)
        _lib.check(self._lib.b2_sensor_get_pixel(self._h, ix, iy, poly.ctypes.data, bounds.ctypes.data))
        return poly, bounds

    def close(self):
        if getattr(self, '_h', None):
            self._lib.b2_sensor_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Sensor:
    """``galsim.Sensor``: photons land in the pixel they hit (``PhotonArray.addTo``);
    used by imsim/photon_pooling.py:139-140 when no sensor is configured."""

    def __init__(self, *, context: OpticsContext = None, device: int = 0, stream=None):
        # a plain sensor is a silicon sensor handle whose boundary state is never touched
        self._sil = None
        self._ctx_args = dict(context=context, device=device, stream=stream)

    def _impl(self):
        if self._sil is None:
            nv = 2
            npoly = 4 * nv + 4
            cfg = dict(NumVertices=nv, PixelBoundaryNx=9, PixelBoundaryNy=9, CollectedCharge_0_0=1e5,
                       PixelSizeX=10.0, SensorThickness=100.0, NumPhases=3, CollectingPhases=2, ChannelStopWidth=2.0,
                       FieldOxideTaper=0.0, Vbb=-50.0, Vparallel_lo=-8.0, Vparallel_hi=2.0, CCDTemperature=173.0,
                       qfh=0.0)
            self._sil = SiliconSensor(config=cfg, vertex_data=np.zeros((81 * npoly, 5)), rng=0,
                                      absorption_table=(np.array([300., 1200.]), np.array([1., 1.])),
                                      **self._ctx_args)
        return self._sil

    def updateRNG(self, rng):
        pass

    def accumulate(self, photons, image, orig_center=None, resume=False):
        s = self._impl()
        n = photons.size() if callable(getattr(photons, 'size', None)) else len(photons)
        if n == 0:
            return 0.0
        s._bind(image)
        added = C.c_double(0.0)
        x = photons.x
        _lib.check(s._lib.b2_plain_accumulate(s._h, n, _lib.ptr(x), _lib.ptr(photons.y), _lib.ptr(photons.flux),
                                              _lib.where_of(x), C.byref(added)))
        s.read_image(image)
        return added.value
