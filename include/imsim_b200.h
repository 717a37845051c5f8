/*
 * imsim_b200.h -- C ABI of the B200-native photon-shooting hot path.
 *
 * This is the drop-in boundary.  Everything above it is
 * host Python (imsim_b200/ *.py, mirroring imsim's GalSim plugin classes);
 * everything below it is hand-written sm_100a CUDA (imsim_b200/csrc).  The
 * signatures use only plain pointers, sizes and POD structs so that the
 * reference can bind them with ctypes.
 *
 * Each entry point names the reference interface it replaces, as
 * `imsim/<file>.py:<line>` (relative to the LSSTDESC/imSim tree) or the
 * third-party call made from that line (GalSim / batoid are not vendored in
 * the reference; see DESIGN.md "Oracle provenance").
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure;
 *     b2_last_error() gives the message (thread-local).
 *   - `where` says where the array pointers of a call live:
 *       B2_HOST   : host memory (numpy).  The library stages the arrays
 *                   through device scratch; copies are part of the call.
 *                   b2_rubin_optics and b2_sensor_accumulate move calls of
 *                   B2_PIPE_MIN (default 2^18) photons or more through a
 *                   ring of pinned slots in chunks of B2_PIPE_CHUNK (2^19)
 *                   photons, copied by B2_HOST_THREADS (min(8, cores)) host
 *                   threads while the previous chunks are transferred and
 *                   computed; results are identical to the single copy.
 *       B2_DEVICE : device memory (e.g. torch.empty(..., device='cuda')).
 *   - all photon arrays are SoA, float64, length n (GalSim PhotonArray layout).
 *   - all calls are asynchronous on the context's stream for B2_DEVICE and
 *     synchronous for B2_HOST.
 */
#ifndef IMSIM_B200_H
#define IMSIM_B200_H

#include <stdint.h>

#if defined(__cplusplus) && !defined(B2_STRUCTS_ONLY)
extern "C" {
#endif

#define B2_ABI_VERSION 2

#define B2_HOST 0
#define B2_DEVICE 1

/* ------------------------------------------------------------------ */
/* Telescope description (flattened batoid.Optic)                      */
/* ------------------------------------------------------------------ */

#define B2_MAX_SURFACES 24
#define B2_MAX_MEDIA 8
#define B2_MAX_ASPHERE_COEF 8
#define B2_MAX_OBSC 4
#define B2_MAX_POLY_ORDER 12 /* Poly2D extra sag: degree <= 11 */

/* surface kinds (batoid.Plane/Sphere/Paraboloid/Quadric/Asphere) */
enum { B2_SURF_PLANE = 0, B2_SURF_SPHERE = 1, B2_SURF_PARABOLOID = 2, B2_SURF_QUADRIC = 3, B2_SURF_ASPHERE = 4 };
/* interaction kinds (batoid.Detector/Mirror/RefractiveInterface/OPDScreen); PASS = OPDScreen on a Plane: the
   surface's extra term is the optical path difference W(x, y) [m] of a thin phase plate, not sag */
enum { B2_INT_DETECTOR = 0, B2_INT_MIRROR = 1, B2_INT_REFRACT = 2, B2_INT_PASS = 3 };
/* extra (summed) sag term: batoid.Sum([base, Zernike]) or Sum([base, Bicubic]) */
enum { B2_EXTRA_NONE = 0, B2_EXTRA_POLY2D = 1, B2_EXTRA_BICUBIC = 2 };
/* obscuration primitives (batoid.ObscCircle/ObscAnnulus/ObscRectangle/ObscRay) */
enum { B2_OBSC_CIRCLE = 0, B2_OBSC_ANNULUS = 1, B2_OBSC_RECTANGLE = 2, B2_OBSC_RAY = 3 };
/* media (batoid.ConstMedium/SellmeierMedium/SumitaMedium/Air) */
enum { B2_MED_CONST = 0, B2_MED_SELLMEIER = 1, B2_MED_SUMITA = 2, B2_MED_AIR = 3 };

typedef struct {
    int32_t kind;   /* B2_OBSC_* */
    int32_t negate; /* 1: batoid.ObscNegation(primitive), i.e. Clear* */
    /* circle: p0=radius p1=x0 p2=y0
       annulus: p0=inner p1=outer p2=x0 p3=y0
       rectangle: p0=width p1=height p2=x0 p3=y0 p4=cos(theta) p5=sin(theta)
       ray: p0=width p1=x0 p2=y0 p3=cos(theta) p4=sin(theta) */
    double p[6];
} B2Obsc;

typedef struct {
    int32_t kind; /* B2_MED_* */
    int32_t pad;
    /* const: p0=n
       sellmeier: p0..2=B1..B3, p3..5=C1..C3 (um^2)
       sumita: p0..5=A0..A5
       air: p0=pressure[kPa] p1=temperature[K] p2=h2o_pressure[kPa] */
    double p[6];
} B2Medium;

typedef struct {
    int32_t surf_kind;      /* B2_SURF_* */
    int32_t interact;       /* B2_INT_* */
    int32_t medium_in;      /* index into B2Telescope.media */
    int32_t medium_out;
    int32_t n_coef;         /* asphere: number of even coefs (r^4, r^6, ...) */
    int32_t rot_identity;   /* 1: drot is the identity (skip the 3x3 products) */
    int32_t n_obsc;         /* obscurations OR-ed together */
    int32_t extra_kind;     /* B2_EXTRA_* */
    double R;               /* radius of curvature (sphere/paraboloid/quadric/asphere) */
    double conic;
    double coef[B2_MAX_ASPHERE_COEF];
    /* transform from the previous interface's coordSys (or the stop surface's
       for the first one) to this one: r' = drot^T (r - dr), v' = drot^T v
       (batoid CoordTransform.applyForward) ; drot row-major */
    double dr[3];
    double drot[9];
    B2Obsc obsc[B2_MAX_OBSC];
    /* Poly2D extra: sag += sum_{i,j} c[i*poly_n + j] X^i Y^j with X=x*poly_scale
       (batoid.Zernike's xy-coefficient array; scale = 1/R_outer).
       Bicubic extra: uniform grid, see b2_telescope_set_extra(). */
    int32_t poly_n;         /* poly order+1 (<= B2_MAX_POLY_ORDER) */
    int32_t extra_slot;     /* slot in the context's extra table, -1 if none */
    double poly_scale;
} B2Surface;

typedef struct {
    int32_t n_surfaces;
    int32_t n_media;
    int32_t medium_stop; /* medium the rays start in (telescope.inMedium) */
    int32_t pad;
    B2Surface surf[B2_MAX_SURFACES];
    B2Medium media[B2_MAX_MEDIA];
} B2Telescope;

/* ------------------------------------------------------------------ */
/* WCS pair, detector, diffraction                                     */
/* ------------------------------------------------------------------ */

/* galsim.GSFitsWCS / FittedSIPWCS, TAN-SIP of order <= 3:
   imsim/photon_ops.py:471-473,482-483 call xyToradec / radecToxy on two of them */
typedef struct {
    double crpix[2];
    double cd[4];        /* row-major 2x2, degrees per pixel unit */
    double ab[2][4][4];  /* ab[k][i][j] multiplies u^i v^j, identity folded in; zero if unused */
    double ra0, dec0;    /* tangent point, radians */
    int32_t order;       /* 0: no SIP (ab ignored); 1..3 */
    int32_t pad;
} B2TanSip;

/* imsim/photon_ops.py:486-503 + imsim/utils.py:42-78 (constant per detector) */
typedef struct {
    /* (x_pix, y_pix) = A (fpx, fpy) + b with fpx = ray.y*1e3, fpy = ray.x*1e3 [mm] */
    double A[4];
    double b[2];
    /* (dxdz, dydz) = Jhat (vx, vy) / vz */
    double Jhat[4];
} B2Detector;

/* imsim/diffraction.py:32-42 (geometry), :284-415 (field rotation) */
typedef struct {
    int32_t enabled;          /* 0: plain RubinOptics */
    int32_t field_rotation;   /* 0: disable_field_rotation=True */
    int32_t n_lines, n_circles;
    double lines[8][4];       /* nx ny d thickness */
    double circles[4][3];     /* x y r */
    double e_z_0[3];          /* zenith at t=0, equatorial frame */
    double e_focal[3];        /* pointing, equatorial frame */
    double cos_lat, sin_lat;
    double omega;             /* earth rotation rate [rad/s] */
} B2Diffraction;

/* options of the fused optics photon-op */
typedef struct {
    int32_t shift_in;    /* add stamp_center before (RubinOptics.shift_photons) */
    int32_t shift_out;   /* subtract stamp_center after (stamp_center is not None) */
    double stamp_center[2];
    int32_t do_focus_depth; /* galsim.FocusDepth fused as epilogue */
    int32_t do_refraction;  /* galsim.Refraction fused as epilogue */
    double focus_depth;     /* pixels */
    double index_ratio;
    uint64_t seed;          /* Philox key when gauss == NULL */
    uint64_t photon_offset; /* global index of photon 0 (Philox counter) */
    /* galsim.PhotonDCR fused as prologue (config/imsim-config.yaml:290-296): applied to x, y
       before anything else, like the op that precedes RubinDiffractionOptics in the list */
    int32_t do_dcr;
    int32_t pad;
    double dcr_base_wavelength; /* nm */
    double dcr_alpha;           /* (w / base)^alpha scaling about dcr_center; 0: none */
    double dcr_center[2];       /* local_wcs.origin */
    double dcr_base_refraction; /* get_refraction(base_wavelength, zenith) [rad] */
    double dcr_tanz;            /* tan(zenith angle) */
    double dcr_pth[3];          /* pressure [kPa], temperature [K], H2O pressure [kPa] */
    double dcr_m[2];            /* pixel shift per radian of refraction: J^-1 (-sin q, cos q) * rad->scale_unit */
} B2OpticsOptions;

typedef struct {
    uint64_t n_vignetted;
    uint64_t n_failed;
    uint64_t n_offdetector_z; /* |z| >= 1e-15 on the detector (reference asserts) */
} B2OpticsStats;

/* ------------------------------------------------------------------ */
/* Silicon sensor                                                      */
/* ------------------------------------------------------------------ */

/* galsim.SiliconSensor.__init__ / _init_silicon arguments, already reduced
   to what the C++ Silicon constructor receives */
typedef struct {
    int32_t num_vertices;   /* NumVertices per edge */
    int32_t nx, ny;         /* PixelBoundaryNx/Ny of the vertex table (9) */
    int32_t qdist;
    double num_elec;        /* CollectedCharge_0_0 / strength */
    double nrecalc;         /* nrecalc / strength ; 0 = only on recalc=True */
    double diff_step;       /* microns, already times diffusion_factor */
    double pixel_size;      /* microns */
    double sensor_thickness;/* microns */
    double treering_center[2];
    int32_t n_treering;     /* tree-ring table points (<=2: no tree rings) */
    int32_t n_abs;          /* absorption table points */
    int32_t transpose;
    int32_t pad;
} B2SensorConfig;

typedef struct {
    double added_flux;
    uint64_t n_polygon_tests;   /* photons that needed the full polygon test */
    uint64_t n_neighbor_search; /* photons not in their nominal pixel */
    uint64_t n_not_found;       /* photons resolved by the coin flip */
    uint64_t n_boundary_1e9;    /* photons within 1e-9 px of a nominal pixel edge */
    uint64_t n_updates;         /* boundary updates performed in this call */
    uint64_t n_dropped_bottom;  /* zconv < 0 */
} B2AccumStats;

/* One object of the classic per-object pipeline for b2_sensor_accumulate_stamps: photons [p0, p0 + n) of the
   arrays handed to the call are accumulated on the object's own zero stamp image (imsim/stamp.py:562-572) */
typedef struct {
    int64_t p0, n;
    int32_t xmin, ymin, nx, ny; /* stamp bounds in image coordinates */
    int32_t plain;              /* != 0: galsim.Sensor semantics (objects below max_flux_simple, stamp.py:534-537) */
    int32_t pad;
} B2StampJob;

/* ------------------------------------------------------------------ */
/* Stage 1: photons of catalogue objects generated in HBM              */
/* ------------------------------------------------------------------ */

/* unit profiles (before the object's 2x2 matrix):
   DELTA    galsim.DeltaFunction                                   (imsim/instcat.py:483-484)
   GAUSSIAN unit-sigma Gaussian
   RADIAL   radial inverse-CDF table row `lut` (Sersic(n), Exponential ...; radius in units of the half-light
            radius, tabulated on t = -log(1 - u))                   (instcat.py:496-520)
   KNOTS    galsim.RandomKnots: one of n_knots points, each a unit-half-light-radius Gaussian deviate derived
            from knot_seed                                          (instcat.py:522-545)
   BOX      galsim.Box(p0, p1): uniform over p0 x p1                (instcat.py:486-494) */
enum { B2_PROF_DELTA = 0, B2_PROF_GAUSSIAN = 1, B2_PROF_RADIAL = 2, B2_PROF_KNOTS = 3, B2_PROF_BOX = 4 };

typedef struct {
    int32_t kind;     /* B2_PROF_* */
    int32_t sed;      /* row of the wavelength inverse-CDF table (SED x bandpass of this object) */
    int32_t lut;      /* row of the radial table (B2_PROF_RADIAL) */
    int32_t n_knots;  /* B2_PROF_KNOTS */
    double x, y;      /* centre on the image [px] */
    double m[4];      /* unit-profile offset -> pixels, row-major: size x shear(q, beta) x lens(g1, g2, mu) x
                         local WCS (arcsec -> px) */
    double p0, p1;    /* BOX: length, width in the units m maps from */
    double tanx, tany;  /* tangents of the object's field angle: the screens are looked up at altitude * tan(theta) */
    uint64_t knot_seed;
} B2Object;

#define B2_MAX_SCREENS 8
/* imsim.atmPSF.AtmosphericPSF.getPSF (imsim/atmPSF.py:298-336) as a photon kick: frozen-flow phase screens
   (first kick, geometric), the SecondKick radial sampler, the optics Gaussian of config/imsim-config.yaml:239-256,
   chromatic dilation (lambda / base_wavelength)^exponent applied to the atmospheric part. */
typedef struct {
    int32_t n_screens;            /* 0: no atmosphere */
    int32_t npix;                 /* each screen npix x npix, periodic */
    int32_t screen_f32;           /* 0: float64 [npix][npix]; 1: float32 [npix][npix]; 2: float32 quads [npix][npix][4] =
                                     (f[y][x], f[y][x+1], f[y+1][x], f[y+1][x+1]) with periodic neighbours, so that a
                                     bilinear gradient is one 16-byte load per screen */
    int32_t n_kick;               /* entries of the second-kick radial table, 0: none */
    double screen_scale;          /* [m] */
    double altitude[B2_MAX_SCREENS];  /* [m] */
    double vx[B2_MAX_SCREENS], vy[B2_MAX_SCREENS];  /* wind [m/s] */
    double t0, exptime;           /* the shooter's own time draw (overwritten later by TimeSampler: SURVEY Q1) */
    double r_inner, r_outer;      /* its own pupil draw: annulus radii [m] (diam 8.36, obscuration 0.61) */
    double base_wavelength;       /* [nm] */
    double exponent;              /* -0.3 */
    double kick_delta_prob;       /* probability that the second kick leaves the photon unmoved */
    double kick_tmax;             /* radial table abscissa: t = -log(1 - u) in [0, kick_tmax] */
    double gauss_sigma;           /* optics Gaussian sigma [arcsec], 0: none */
    double arcsec_to_pix[4];      /* inverse local WCS Jacobian, row-major (one per pool: SURVEY Q2) */
} B2Psf;

/* ------------------------------------------------------------------ */
/* Post-path electronics (readout)                                     */
/* ------------------------------------------------------------------ */

/* one amplifier segment (imsim/camera.py:19-92 Amp: bounds, raw_data_bounds, raw_bounds, gain, raw_flip_x/y,
   bias_level, read_noise), 0-based pixel indices */
typedef struct {
    int32_t x0, y0, nx, ny;     /* imaging area of this amp in the e-image */
    int32_t raw_nx, raw_ny;     /* full segment with prescan / overscan */
    int32_t data_x0, data_y0;   /* position of the imaging area inside the raw segment */
    int32_t flip_x, flip_y;     /* readout order flips */
    double gain;                /* e- / ADU */
    double bias_level;          /* ADU */
    double read_noise;          /* ADU rms */
} B2Amp;

#ifndef B2_STRUCTS_ONLY /* (the oracle's FLOP counter re-reads only the POD structs above) */
typedef struct b2_ctx b2_ctx;
typedef struct b2_sensor b2_sensor;

/* ---- library -------------------------------------------------------- */
const char* b2_last_error(void);
int b2_abi_version(void);
/* number of kernels launched by this library since load (bench.py gpu_launches) */
uint64_t b2_launch_count(void);
/* sizeof() of the POD structs, for binding self-checks:
   0 B2Telescope, 1 B2Surface, 2 B2TanSip, 3 B2Detector, 4 B2Diffraction,
   5 B2OpticsOptions, 6 B2OpticsStats, 7 B2SensorConfig, 8 B2AccumStats, 9 B2Obsc, 10 B2Medium,
   11 B2Object, 12 B2Psf, 13 B2Amp, 14 B2StampJob */
int64_t b2_sizeof(int32_t which);

/* ---- context: one per (process, detector) --------------------------- */
int b2_ctx_create(int device, void* cuda_stream, b2_ctx** out);
int b2_ctx_destroy(b2_ctx* ctx);
int b2_ctx_set_stream(b2_ctx* ctx, void* cuda_stream);
int b2_ctx_synchronize(b2_ctx* ctx);
/* measurement aid: register-resident FMA chains on every SM; returns the achieved
   non-tensor FMA rate (fp64 != 0: double, else float) in TFLOP/s.  The roofline
   denominator of the compute-bound trace kernel (MEASURED_PEAKS.json has none). */
int b2_fma_peak(b2_ctx* ctx, int32_t fp64, double* tflops);
/* test aid: out[i] = the device's round_to_f32(in[i]) -- the FP64-adder rounding to float precision the boundary
   update uses instead of double -> float -> double conversions; must equal (double)(float)in[i] (host arrays) */
int b2_test_round_f32(b2_ctx* ctx, int64_t n, const double* in, double* out);
/* measured atomicAdd throughput [atomics/s] on an nx x ny image (fp64: double, else float): pattern 0 uniform random
   pixels, 1 one hot pixel, 2 a thousand Gaussian star images -- the ceiling of the charge deposit */
int b2_atomic_peak(b2_ctx* ctx, int32_t fp64, int32_t pattern, int32_t nx, int32_t ny, int64_t n, double* atomics_per_s);
/* measurement aid: with B2_TIMING=1 in the environment every kernel launch is bracketed by CUDA
   events on its stream; this returns {"kernel": [launches, total_ms], ...} as JSON and clears the log */
int b2_timing_report(char* buf, int64_t cap);
/* measurement aid for the roofline of the dominant kernel: when switched on, b2_pool_step and
   b2_rubin_optics bracket their main kernel with CUDA events on the launching stream;
   b2_ctx_kernel_ms returns the summed durations and the number of launches, and clears the list */
int b2_ctx_record_kernel_events(b2_ctx* ctx, int32_t on);
int b2_ctx_kernel_ms(b2_ctx* ctx, double* total_ms, int64_t* count);

/* replaces: base['det_telescope'] (imsim/telescope_loader.py:463) as consumed by
   imsim/photon_ops.py:108-123 */
int b2_telescope_upload(b2_ctx* ctx, const B2Telescope* tel);
/* extra sag tables.  poly2d: data = poly_n*poly_n coefficients.
   bicubic: data = [x0, dx, nx, y0, dy, ny] followed by 4 grids (z, dzdx, dzdy, d2zdxdy),
   each ny*nx row-major (batoid.Bicubic). */
int b2_telescope_set_extra(b2_ctx* ctx, int surface_index, int extra_kind, const double* data, int64_t n);
/* Surface program the trace kernels run for the uploaded telescope: 0 = generic interpreter over the
   surface list, 1 = the Rubin layout compiled as straight-line code (3 aspheric mirrors, 4 two-surface
   lenses with an aspheric L2 exit, detector; no summed perturbation terms).  Chosen by
   b2_telescope_upload / b2_telescope_set_extra from the telescope's signature; environment B2_PROGRAM=0
   forces the interpreter.  Both run the same arithmetic. */
int b2_telescope_program(b2_ctx* ctx);
/* replaces: base['current_image'].wcs, base['_icrf_to_field'] (imsim/photon_ops.py:407-408) */
int b2_wcs_upload(b2_ctx* ctx, const B2TanSip* img_wcs, const B2TanSip* icrf_to_field);
/* XyToV (imsim/photon_ops.py:454-475) compiled for one detector: over the pixel box the chain img_wcs.xyToradec
   -> icrf_to_field.radecToxy is replaced by one degree 5 x 5 polynomial of the field-angle tangents, fitted to the
   exact chain at Chebyshev points and adopted only if it reproduces the exact chain on a check grid to tol_px
   pixels (1e-9 is typical: ~3e-10 px is reached, 1e-13 of the coordinate); max_resid_px reports the check.
   Positions outside the box always take the exact chain; b2_wcs_upload discards the compiled form. */
int b2_xytov_compile(b2_ctx* ctx, double xlo, double xhi, double ylo, double yhi, double tol_px,
                     double* max_resid_px);
/* replaces: camera[det_name] as used by imsim/photon_ops.py:495-500 */
int b2_detector_upload(b2_ctx* ctx, const B2Detector* det);
/* replaces: RubinDiffraction.__init__ (imsim/photon_ops.py:233-262) */
int b2_diffraction_config(b2_ctx* ctx, const B2Diffraction* cfg);

/* ---- photon ops ------------------------------------------------------- */
/* XyToV.__call__ (imsim/photon_ops.py:469-475): v is 3 SoA arrays */
int b2_xy_to_v(b2_ctx* ctx, int64_t n, const double* x, const double* y, double* vx, double* vy, double* vz, int where);
/* XyToV.inverse (imsim/photon_ops.py:477-483) */
int b2_v_to_xy(b2_ctx* ctx, int64_t n, const double* vx, const double* vy, const double* vz, double* x, double* y, int where);
/* batoid Optic.trace on a RayVector in the stop surface's coordSys
   (imsim/photon_ops.py:109-123).  All arrays in/out.  vignetted/failed: uint8. */
int b2_trace_rays(b2_ctx* ctx, int64_t n, double* x, double* y, double* z, double* vx, double* vy, double* vz,
                  double* t, const double* wavelength_m, uint8_t* vignetted, uint8_t* failed, int where);
/* RubinOptics.applyTo / RubinDiffractionOptics.applyTo (imsim/photon_ops.py:81-127,203-208),
   optionally fused with galsim.FocusDepth and galsim.Refraction
   (config/imsim-config.yaml:304-320).  x,y,flux in/out; dxdz,dydz out.
   gauss: one standard-normal draw per photon (injected, reference order) or NULL (Philox).
   time_out: optional, ray time at the detector (batoid RayVector.t), may be NULL. */
int b2_rubin_optics(b2_ctx* ctx, int64_t n, double* x, double* y, double* dxdz, double* dydz, double* flux,
                    const double* wavelength_nm, const double* pupil_u, const double* pupil_v,
                    const double* time, const double* gauss, double* time_out, const B2OpticsOptions* opt,
                    int where, B2OpticsStats* stats);
/* RubinDiffraction.applyTo (imsim/photon_ops.py:304-352): x,y in/out */
int b2_rubin_diffraction(b2_ctx* ctx, int64_t n, double* x, double* y, const double* wavelength_nm,
                         const double* pupil_u, const double* pupil_v, const double* time,
                         const double* gauss, const B2OpticsOptions* opt, int where);
/* galsim.TimeSampler + galsim.PupilAnnulusSampler (config/imsim-config.yaml:281-289), Philox */
int b2_sample_time_pupil(b2_ctx* ctx, int64_t n, double* time, double* pupil_u, double* pupil_v,
                         double t0, double exptime, double r_inner, double r_outer,
                         uint64_t seed, uint64_t photon_offset, int where);

/* uniform unit-flux photons over [xlo,xhi) x [ylo,yhi) and, if wl != NULL, wavelengths from a
   tabulated inverse CDF (cdf[k] = P(wave <= cdf_wave[k]), k < ncdf): the photon generation of the
   photon-shot flat, imsim/flat.py:239-259.  DEVICE pointers only (the pool never leaves HBM). */
int b2_flat_photons(b2_ctx* ctx, int64_t n, double* x, double* y, double* flux, double* wl, double xlo, double xhi,
                    double ylo, double yhi, const double* cdf, const double* cdf_wave, int32_t ncdf, uint64_t seed,
                    uint64_t photon_offset);

/* photons of a table of point-like objects behind a Gaussian PSF, generated on the device: object j
   owns photons [obj_cum[j], obj_cum[j+1]) of the pool (obj_cum has nobj+1 entries, obj_cum[nobj] = n),
   i.e. build_stamps + merge_photon_arrays (imsim/photon_pooling.py:151-152) for DeltaFunction objects
   with their per-batch photon counts (photon_pooling.py:300-304).  DEVICE pointers only. */
int b2_object_photons(b2_ctx* ctx, int64_t n, double* x, double* y, double* flux, double* wl, const double* obj_x,
                      const double* obj_y, const double* obj_sigma, const int64_t* obj_cum, int32_t nobj,
                      const double* cdf, const double* cdf_wave, int32_t ncdf, uint64_t seed, uint64_t photon_offset);

/* Stage 1 of the north star on the device: catalogue objects -> pooled photons (x, y, flux, wavelength), i.e.
   build_stamps + drawImage(method='phot') + merge_photon_arrays (imsim/photon_pooling.py:151-152, stamp.py:727-743)
   without the host.  Object j owns photons [obj_cum[j], obj_cum[j+1]).  Per photon: profile offset through the
   object's matrix, wavelength from row `sed` of the inverse-CDF table (ncdf columns, shared abscissa cdf_wave
   per row), then the PSF kicks of B2Psf (screens / second kick / Gaussian).  rand: NULL (Philox streams 7-10 of
   `seed`, counted from photon_offset) or B2_STAGE1_NRAND uniforms per photon, rand[k * n + i], for parity tests.
   DEVICE pointers only; screens[l] npix*npix row-major [y][x]. */
#define B2_STAGE1_NRAND 12
int b2_psf_upload(b2_ctx* ctx, const B2Psf* psf, const void* const* screens, const double* kick_table);
int b2_radial_luts_upload(b2_ctx* ctx, const double* lut, int32_t n_lut, int32_t n_entries, double tmax);
int b2_stage1_photons(b2_ctx* ctx, int64_t n, double* x, double* y, double* flux, double* wl, const B2Object* objects,
                      const int64_t* obj_cum, int32_t nobj, const double* cdf, const double* cdf_wave, int32_t n_sed,
                      int32_t ncdf, const double* rand, uint64_t seed, uint64_t photon_offset);

/* imsim.bleed_trails.bleed_eimage (imsim/bleed_trails.py:26-60) on a float32 e-image [ny][nx], in place:
   every channel (column; the two halves separately if midline_stop, e2v sensors, readout.py:427-431) spills
   the charge above full_well alternately down and up the column.  Bit-identical to the reference. */
int b2_bleed_trails(b2_ctx* ctx, float* eimage, int32_t nx, int32_t ny, double full_well, int32_t midline_stop,
                    int where);
/* CcdReadout.build_amp_images (imsim/readout.py:414-480) on a DEVICE e-image (modified in place by the bleed
   trails, if full_well > 0, and the dark current, if dark_mean > 0): amp split / gain / flips, crosstalk
   (xtalk: HOST namp x namp or NULL, readout.py:403-412), prescan / overscan, CTI (pband / sband: HOST band
   form of cte_matrix, [raw_ny or raw_nx][ntransfers + 1], band[i][k] = matrix[i][i - k], or NULL;
   readout.py:153-203,391-401), then bias + read noise -> int32.  segments (DEVICE float32 [namp][raw_ny][raw_nx],
   before bias / noise) and raw (DEVICE int32, same shape) are optional outputs. */
int b2_readout(b2_ctx* ctx, float* eimage, int32_t nx, int32_t ny, const B2Amp* amps, int32_t namp, const double* xtalk,
               const double* pband, const double* sband, int32_t ntransfers, double full_well, int32_t midline_stop,
               double dark_mean, uint64_t seed, float* segments, int32_t* raw);
/* LSST_ImageBuilderBase.addNoise (imsim/lsst_image.py:128-199) for a photon-shot e-image on the DEVICE:
   image[i] += Poisson(sky_level * areas[i] * modulation[i]) with exact Poisson deviates (Philox).  areas: DEVICE
   float64 pixel areas from b2_sensor_pixel_areas (tree rings / brighter-fatter) or NULL; modulation: DEVICE
   float32 map (sky gradient x vignetting x fringing) or NULL. */
int b2_add_sky(b2_ctx* ctx, void* image, int32_t dtype_bytes, int64_t npix, double sky_level, const double* areas,
               const float* modulation, uint64_t seed);
/* CosmicRays.paint_cr (imsim/cosmic_rays.py:74-111) for a flattened list of span pixels: image[iy][ix] += value on a
   DEVICE image with numpy's indexing rules (negative indices wrap, indices beyond the array are skipped).
   iy, ix, values: DEVICE arrays of n entries. */
int b2_scatter_add(b2_ctx* ctx, void* image, int32_t dtype_bytes, int32_t nx, int32_t ny, int64_t n, const int32_t* iy,
                   const int32_t* ix, const float* values);

/* ---- silicon sensor ---------------------------------------------------- */
/* replaces galsim.SiliconSensor.__init__ (imsim/lsst_image.py:93-103,
   config/imsim-config.yaml:230-235).  vertex_data: the .dat table,
   nx*ny*(4*num_vertices+4) rows of 5 doubles.  treering_r/f/y2: tree-ring lookup
   table (imsim/treerings.py:192-194) with the second derivatives of its natural
   cubic spline (galsim.LookupTable default interpolant), y2 == NULL: linear.
   abs_wave/abs_len: absorption length table, nm -> microns (linear). */
int b2_sensor_create(b2_ctx* ctx, const B2SensorConfig* cfg, const double* vertex_data,
                     const double* treering_r, const double* treering_f, const double* treering_y2,
                     const double* abs_wave, const double* abs_len, b2_sensor** out);
int b2_sensor_destroy(b2_sensor* s);
/* new tree-ring table / centre for the next detector of the same vendor (config/imsim-config.yaml:233-235
   rebuilds the sensor per image; this keeps the per-image boundary arrays allocated) */
int b2_sensor_set_treerings(b2_sensor* s, double center_x, double center_y, const double* treering_r,
                            const double* treering_f, const double* treering_y2, int32_t n);
/* bind the image the next accumulate calls add to: bounds (xmin,ymin,nx,ny),
   dtype_bytes 4 (float32) or 8 (float64); pixels: row-major ny*nx, host or device per `where`.
   (galsim.Image passed to SiliconSensor.accumulate, imsim/photon_pooling.py:210) */
int b2_sensor_bind_image(b2_sensor* s, int32_t xmin, int32_t ymin, int32_t nx, int32_t ny,
                         int32_t dtype_bytes, const void* pixels, int where);
/* copy the current device image (target + pending delta) out */
int b2_sensor_read_image(b2_sensor* s, void* pixels, int where);
/* galsim.SiliconSensor.accumulate(photons, image, orig_center, resume, recalc)
   (imsim/photon_pooling.py:210, imsim/stamp.py:562-572, imsim/flat.py:261).
   dxdz/dydz/wavelength_nm may be NULL (not allocated in the PhotonArray).
   rand4: injected randoms [g1[n], g2[n], u_notfound[n], u_depth[n]] or NULL (Philox). */
int b2_sensor_accumulate(b2_sensor* s, int64_t n, const double* x, const double* y,
                         const double* dxdz, const double* dydz, const double* wavelength_nm,
                         const double* flux, const double* rand4, uint64_t seed, uint64_t photon_offset,
                         int32_t orig_center_x, int32_t orig_center_y, int32_t resume, int32_t recalc,
                         int where, B2AccumStats* stats);
/* The object loop of the classic pipeline (imsim/lsst_image.py:342-389 around imsim/stamp.py:562-572) in one
   call: for every job a zero stamp is bound (fresh tree-ring boundaries), its photons are accumulated with the
   sensor's nrecalc cadence inside the stamp -- the whole loop runs on the device, one thread block per stamp --
   and the stamp is added to the full image (full_image[bounds] += stamp[bounds]).
   jobs: HOST; photon arrays, rand4 ([4][n], optional) and full_pixels (full_ny x full_nx, float32 / float64): DEVICE.
   Draws: Philox(seed) at photon_offset + index in the arrays.  Stamp states live in an arena
   (B2_STAMP_ARENA_MB, default 6144; larger job lists run as several launches).  The sensor's own bound image is
   not touched.  added_per_job (HOST, optional): flux that landed on each stamp.  Bit-identical, stamp by stamp,
   to b2_sensor_bind_image(zeros) + b2_sensor_accumulate on the same photons. */
int b2_sensor_accumulate_stamps(b2_sensor* s, int32_t njobs, const B2StampJob* jobs, int64_t n, const double* x,
                                const double* y, const double* dxdz, const double* dydz, const double* wavelength_nm,
                                const double* flux, const double* rand4, uint64_t seed, uint64_t photon_offset,
                                int32_t orig_center_x, int32_t orig_center_y, void* full_pixels, int32_t full_xmin,
                                int32_t full_ymin, int32_t full_nx, int32_t full_ny, int32_t dtype_bytes,
                                B2AccumStats* stats, double* added_per_job);
/* galsim.SiliconSensor.calculate_pixel_areas(image, orig_center, use_flux)
   (imsim/flat.py:223).  Uses the bound image as the charge; areas: ny*nx float64. */
int b2_sensor_pixel_areas(b2_sensor* s, int32_t orig_center_x, int32_t orig_center_y, int32_t use_flux,
                          double* areas, int where);
/* galsim.Sensor.accumulate (= PhotonArray.addTo), used when sensor is None
   (imsim/photon_pooling.py:139-140,212) */
int b2_plain_accumulate(b2_sensor* s, int64_t n, const double* x, const double* y, const double* flux,
                        int where, double* added_flux);
/* replaces: merge_photon_arrays (imsim/photon_pooling.py:177-192) + the upload of the pool.  The photon arrays of
   many stamps (pageable host memory, as GalSim leaves them) are gathered into one device array per field: the
   merge happens inside a ring of pinned chunks filled by host copy threads (B2_HOST_THREADS) while the copy engine
   moves the previous chunk.  seg: nfields * nseg HOST pointers, field-major (seg[f * nseg + g] = field f of stamp g);
   seg_len[g]: photons of stamp g; dst[f]: DEVICE array of sum(seg_len) doubles.  Returns once the host arrays have
   been read; the device copies are ordered before later work on the context's stream. */
int b2_photons_upload(b2_ctx* ctx, int32_t nfields, int64_t nseg, const double* const* seg, const int64_t* seg_len,
                      double* const* dst);
/* a large array between pageable HOST memory and DEVICE memory through the pinned ring (to_device != 0: host -> device):
   what b2_sensor_bind_image / b2_sensor_read_image do for a full CCD, for arrays the caller keeps on the device itself
   (the full image of the classic pipeline, imsim/lsst_image.py:359-368).  bytes must be a multiple of 8.  Returns once
   the host array has been read / written. */
int b2_copy_through_ring(b2_ctx* ctx, void* host, void* device, int64_t bytes, int32_t to_device);
/* A pooled field that is one number per stamp (GalSim's shooters give every photon of an object the same flux) need
   not cross PCIe: b2_segments_constant reads the stamps' HOST arrays once on the copy threads (value[g] = first element
   of segment g, *all_constant = 1 iff every segment is constant, compared bit for bit), b2_fill_segments writes
   dst[sum(seg_len[:g]) ...] = value[g] on the DEVICE in stream order.  Same bits as uploading the arrays. */
int b2_segments_constant(int64_t nseg, const double* const* seg, const int64_t* seg_len, double* value,
                         int32_t* all_constant);
int b2_fill_segments(b2_ctx* ctx, int64_t nseg, const int64_t* seg_len, const double* value, double* dst);
/* host-to-host copy on the library's copy threads (B2_HOST_THREADS): used to hand a pinned snapshot of the image to
   the caller's pageable array when a checkpoint is written while the next batch is already uploading */
int b2_host_memcpy(void* dst, const void* src, int64_t bytes);
/* One photon batch of the pooled pipeline as a single kernel, on DEVICE arrays
   (imsim/photon_pooling.py:149-160 with the op list of config/imsim-config.yaml:281-320):
   TimeSampler + PupilAnnulusSampler -> [PhotonDCR] -> RubinDiffractionOptics -> FocusDepth -> Refraction ->
   SiliconSensor.accumulate(photons, image, resume, recalc).  Needs nrecalc == 0 (pooled cadence).
   x, y, flux in (pixel positions of the pooled photons), wl_nm in; if write_back != 0 the traced
   photons (x, y, dxdz, dydz, flux) are stored like the separate ops would leave them.
   Draws: sampler Philox(sampler_seed) counted from photon_offset, kick Philox(opt->seed) from
   opt->photon_offset, sensor Philox(sensor_seed) from sensor_offset (the sensor's own photon counter,
   what b2_sensor_accumulate is given) -- identical to the unfused calls. */
int b2_pool_step(b2_ctx* ctx, b2_sensor* sensor, int64_t n, double* x, double* y, double* dxdz, double* dydz,
                 double* flux, const double* wl_nm, const B2OpticsOptions* opt, double t0, double exptime,
                 double r_inner, double r_outer, uint64_t sampler_seed, uint64_t sensor_seed,
                 uint64_t photon_offset, uint64_t sensor_offset, int32_t resume, int32_t recalc,
                 int32_t write_back, B2OpticsStats* ostats, B2AccumStats* astats);
/* One iteration of the photon-shot flat (imsim/flat.py:239-264) on the sensor's bound image, fused:
   photons are generated tile by tile (tile x tile pixels) straight into the charge deposit.
   tile_cum: HOST int64[tiles+1], tile_cum[0] = 0, cumulative per-tile photon counts of this iteration (Poisson
   counts drawn by the caller: the same distribution as n_total uniform photons; tiles are processed in passes
   of <= 2^27 photons so scratch stays bounded); cdf / cdf_wave: DEVICE; wavelengths from the inverse CDF if
   ncdf >= 2 (else conversion at 1 micron, like a PhotonArray without wavelengths).
   resume as in accumulate(); update_after != 0 performs the boundary update (nrecalc reached) before return. */
int b2_flat_step(b2_ctx* ctx, b2_sensor* sensor, const int64_t* tile_cum, int64_t n_total, int32_t tile,
                 const double* cdf, const double* cdf_wave, int32_t ncdf, uint64_t seed, uint64_t sensor_seed,
                 uint64_t photon_offset, int32_t resume, int32_t update_after, B2AccumStats* astats);
/* debug/inspection: copy boundary state of pixel (ix,iy) in image coords:
   poly: (4*nv+4)*2 doubles in polygon order; bounds: inner[4], outer[4] (xmin,xmax,ymin,ymax) */
int b2_sensor_get_pixel(b2_sensor* s, int32_t ix, int32_t iy, double* poly, double* bounds);

#endif /* B2_STRUCTS_ONLY */
#if defined(__cplusplus) && !defined(B2_STRUCTS_ONLY)
}
#endif
#endif /* IMSIM_B200_H */
