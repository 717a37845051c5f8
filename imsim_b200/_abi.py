"""ctypes mirror of ``include/imsim_b200.h`` (the POD structs and constants).

Kept free of any CUDA dependency so that host logic and the oracle tests can use
the same flattened descriptions.  ``_lib.py`` cross-checks ``ctypes.sizeof`` of
every struct against ``b2_sizeof`` of the loaded library.
"""
import ctypes as C

B2_ABI_VERSION = 2
B2_HOST = 0
B2_DEVICE = 1

B2_MAX_SURFACES = 24
B2_MAX_MEDIA = 8
B2_MAX_ASPHERE_COEF = 8
B2_MAX_OBSC = 4
B2_MAX_POLY_ORDER = 12

SURF_PLANE, SURF_SPHERE, SURF_PARABOLOID, SURF_QUADRIC, SURF_ASPHERE = range(5)
INT_DETECTOR, INT_MIRROR, INT_REFRACT, INT_PASS = range(4)
EXTRA_NONE, EXTRA_POLY2D, EXTRA_BICUBIC = range(3)
OBSC_CIRCLE, OBSC_ANNULUS, OBSC_RECTANGLE, OBSC_RAY = range(4)
MED_CONST, MED_SELLMEIER, MED_SUMITA, MED_AIR = range(4)


class B2Obsc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("negate", C.c_int32), ("p", C.c_double * 6)]


class B2Medium(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("p", C.c_double * 6)]


class B2Surface(C.Structure):
    _fields_ = [
        ("surf_kind", C.c_int32),
        ("interact", C.c_int32),
        ("medium_in", C.c_int32),
        ("medium_out", C.c_int32),
        ("n_coef", C.c_int32),
        ("rot_identity", C.c_int32),
        ("n_obsc", C.c_int32),
        ("extra_kind", C.c_int32),
        ("R", C.c_double),
        ("conic", C.c_double),
        ("coef", C.c_double * B2_MAX_ASPHERE_COEF),
        ("dr", C.c_double * 3),
        ("drot", C.c_double * 9),
        ("obsc", B2Obsc * B2_MAX_OBSC),
        ("poly_n", C.c_int32),
        ("extra_slot", C.c_int32),
        ("poly_scale", C.c_double),
    ]


class B2Telescope(C.Structure):
    _fields_ = [
        ("n_surfaces", C.c_int32),
        ("n_media", C.c_int32),
        ("medium_stop", C.c_int32),
        ("pad", C.c_int32),
        ("surf", B2Surface * B2_MAX_SURFACES),
        ("media", B2Medium * B2_MAX_MEDIA),
    ]


class B2TanSip(C.Structure):
    _fields_ = [
        ("crpix", C.c_double * 2),
        ("cd", C.c_double * 4),
        ("ab", ((C.c_double * 4) * 4) * 2),
        ("ra0", C.c_double),
        ("dec0", C.c_double),
        ("order", C.c_int32),
        ("pad", C.c_int32),
    ]


class B2Detector(C.Structure):
    _fields_ = [("A", C.c_double * 4), ("b", C.c_double * 2), ("Jhat", C.c_double * 4)]


class B2Diffraction(C.Structure):
    _fields_ = [
        ("enabled", C.c_int32),
        ("field_rotation", C.c_int32),
        ("n_lines", C.c_int32),
        ("n_circles", C.c_int32),
        ("lines", (C.c_double * 4) * 8),
        ("circles", (C.c_double * 3) * 4),
        ("e_z_0", C.c_double * 3),
        ("e_focal", C.c_double * 3),
        ("cos_lat", C.c_double),
        ("sin_lat", C.c_double),
        ("omega", C.c_double),
    ]


class B2OpticsOptions(C.Structure):
    _fields_ = [
        ("shift_in", C.c_int32),
        ("shift_out", C.c_int32),
        ("stamp_center", C.c_double * 2),
        ("do_focus_depth", C.c_int32),
        ("do_refraction", C.c_int32),
        ("focus_depth", C.c_double),
        ("index_ratio", C.c_double),
        ("seed", C.c_uint64),
        ("photon_offset", C.c_uint64),
        ("do_dcr", C.c_int32),
        ("pad", C.c_int32),
        ("dcr_base_wavelength", C.c_double),
        ("dcr_alpha", C.c_double),
        ("dcr_center", C.c_double * 2),
        ("dcr_base_refraction", C.c_double),
        ("dcr_tanz", C.c_double),
        ("dcr_pth", C.c_double * 3),
        ("dcr_m", C.c_double * 2),
    ]


class B2OpticsStats(C.Structure):
    _fields_ = [("n_vignetted", C.c_uint64), ("n_failed", C.c_uint64), ("n_offdetector_z", C.c_uint64)]


class B2SensorConfig(C.Structure):
    _fields_ = [
        ("num_vertices", C.c_int32),
        ("nx", C.c_int32),
        ("ny", C.c_int32),
        ("qdist", C.c_int32),
        ("num_elec", C.c_double),
        ("nrecalc", C.c_double),
        ("diff_step", C.c_double),
        ("pixel_size", C.c_double),
        ("sensor_thickness", C.c_double),
        ("treering_center", C.c_double * 2),
        ("n_treering", C.c_int32),
        ("n_abs", C.c_int32),
        ("transpose", C.c_int32),
        ("pad", C.c_int32),
    ]


class B2AccumStats(C.Structure):
    _fields_ = [
        ("added_flux", C.c_double),
        ("n_polygon_tests", C.c_uint64),
        ("n_neighbor_search", C.c_uint64),
        ("n_not_found", C.c_uint64),
        ("n_boundary_1e9", C.c_uint64),
        ("n_updates", C.c_uint64),
        ("n_dropped_bottom", C.c_uint64),
    ]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class B2StampJob(C.Structure):
    _fields_ = [("p0", C.c_int64), ("n", C.c_int64), ("xmin", C.c_int32), ("ymin", C.c_int32), ("nx", C.c_int32),
                ("ny", C.c_int32), ("plain", C.c_int32), ("pad", C.c_int32)]


PROF_DELTA, PROF_GAUSSIAN, PROF_RADIAL, PROF_KNOTS, PROF_BOX = range(5)
B2_MAX_SCREENS = 8
B2_STAGE1_NRAND = 12


class B2Object(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("sed", C.c_int32), ("lut", C.c_int32), ("n_knots", C.c_int32),
        ("x", C.c_double), ("y", C.c_double),
        ("m", C.c_double * 4),
        ("p0", C.c_double), ("p1", C.c_double),
        ("tanx", C.c_double), ("tany", C.c_double),
        ("knot_seed", C.c_uint64),
    ]


class B2Psf(C.Structure):
    _fields_ = [
        ("n_screens", C.c_int32), ("npix", C.c_int32), ("screen_f32", C.c_int32), ("n_kick", C.c_int32),
        ("screen_scale", C.c_double),
        ("altitude", C.c_double * B2_MAX_SCREENS),
        ("vx", C.c_double * B2_MAX_SCREENS), ("vy", C.c_double * B2_MAX_SCREENS),
        ("t0", C.c_double), ("exptime", C.c_double),
        ("r_inner", C.c_double), ("r_outer", C.c_double),
        ("base_wavelength", C.c_double), ("exponent", C.c_double),
        ("kick_delta_prob", C.c_double), ("kick_tmax", C.c_double),
        ("gauss_sigma", C.c_double),
        ("arcsec_to_pix", C.c_double * 4),
    ]


class B2Amp(C.Structure):
    _fields_ = [
        ("x0", C.c_int32), ("y0", C.c_int32), ("nx", C.c_int32), ("ny", C.c_int32),
        ("raw_nx", C.c_int32), ("raw_ny", C.c_int32), ("data_x0", C.c_int32), ("data_y0", C.c_int32),
        ("flip_x", C.c_int32), ("flip_y", C.c_int32),
        ("gain", C.c_double), ("bias_level", C.c_double), ("read_noise", C.c_double),
    ]


# numpy structured dtype with the layout of B2Object (object tables are built vectorised)
import numpy as _np  # noqa: E402

OBJECT_DTYPE = _np.dtype([("kind", "<i4"), ("sed", "<i4"), ("lut", "<i4"), ("n_knots", "<i4"), ("x", "<f8"), ("y", "<f8"),
                          ("m", "<f8", (4,)), ("p0", "<f8"), ("p1", "<f8"), ("tanx", "<f8"), ("tany", "<f8"),
                          ("knot_seed", "<u8")])
assert OBJECT_DTYPE.itemsize == C.sizeof(B2Object)

# order of b2_sizeof(which)
SIZEOF_ORDER = [
    B2Telescope, B2Surface, B2TanSip, B2Detector, B2Diffraction, B2OpticsOptions,
    B2OpticsStats, B2SensorConfig, B2AccumStats, B2Obsc, B2Medium, B2Object, B2Psf, B2Amp, B2StampJob,
]
