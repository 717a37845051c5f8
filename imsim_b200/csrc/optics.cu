// optics.cu -- photon ray-trace kernels (sm_100a).
//
// One thread per photon, photon state in registers, all arrays SoA float64 so a
// warp touches 256 contiguous bytes per array.  The telescope / WCS / detector
// description travels as a __grid_constant__ kernel parameter (constant bank,
// uniform broadcast loads), so concurrent detectors on different streams never
// share mutable global state.
//
// Replaces (per photon): imsim/photon_ops.py:81-148,274-302,454-503 and the
// batoid / GalSim C++ those lines call; see include/imsim_b200.h.
#include <cstdlib>

#include "optics_device.cuh"

// ------------------------------------------------------------------ kernels
__global__ void __launch_bounds__(256)
k_xy_to_v(const __grid_constant__ DevOptics o, int64_t n, const double* __restrict__ x, const double* __restrict__ y,
          double* __restrict__ vx, double* __restrict__ vy, double* __restrict__ vz) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a, b, c;
    xy_to_v(o, x[i], y[i], a, b, c);
    vx[i] = a;
    vy[i] = b;
    vz[i] = c;
}

__global__ void __launch_bounds__(256)
k_v_to_xy(const __grid_constant__ DevOptics o, int64_t n, const double* __restrict__ vx, const double* __restrict__ vy,
          const double* __restrict__ vz, double* __restrict__ x, double* __restrict__ y) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double a, b;
    v_to_xy(o, vx[i], vy[i], vz[i], a, b);
    x[i] = a;
    y[i] = b;
}

template <int PROG>
__global__ void __launch_bounds__(256)
k_trace_rays(const __grid_constant__ DevOptics o, int64_t n, double* x, double* y, double* z, double* vx, double* vy,
             double* vz, double* t, const double* __restrict__ wl, uint8_t* vig, uint8_t* fail) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ray r{x[i], y[i], z[i], vx[i], vy[i], vz[i], t[i], vig[i] != 0, fail[i] != 0};
    trace_ray<PROG>(o, media_of<PROG>(o, wl[i]), r);
    x[i] = r.x; y[i] = r.y; z[i] = r.z;
    vx[i] = r.vx; vy[i] = r.vy; vz[i] = r.vz;
    t[i] = r.t;
    vig[i] = r.vignetted;
    fail[i] = r.failed;
}

// RubinOptics / RubinDiffractionOptics.applyTo fused with FocusDepth + Refraction
template <int PROG>
__device__ __forceinline__ void
rubin_optics_body(const DevOptics& o, const B2OpticsOptions& opt, int64_t n,
               double* __restrict__ x, double* __restrict__ y, double* __restrict__ dxdz, double* __restrict__ dydz,
               double* __restrict__ flux, const double* __restrict__ wl_nm, const double* __restrict__ pu,
               const double* __restrict__ pv, const double* __restrict__ time, const double* __restrict__ gauss,
               double* __restrict__ time_out, unsigned long long* __restrict__ stats) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool active = i < n;
    bool vig = false, fail = false, offz = false;
    if (active) {
        double g = 0.0;
        if (o.dif.enabled) g = gauss ? gauss[i] : philox_normal(opt.seed, opt.photon_offset + (uint64_t)i, 0u);
        OpticsOut r = optics_photon<PROG>(o, opt, x[i], y[i], wl_nm[i], pu[i], pv[i], time[i], flux[i], g);
        vig = r.vig;
        fail = r.fail;
        offz = r.offz;
        x[i] = r.x;
        y[i] = r.y;
        dxdz[i] = r.dxdz;
        dydz[i] = r.dydz;
        flux[i] = r.flux;
        if (time_out) time_out[i] = r.t;
    }
    if (stats) {
        unsigned nv = __popc(__ballot_sync(0xffffffffu, vig));
        unsigned nf = __popc(__ballot_sync(0xffffffffu, fail));
        unsigned nz = __popc(__ballot_sync(0xffffffffu, offz));
        if ((threadIdx.x & 31) == 0) {
            if (nv) atomicAdd(&stats[0], (unsigned long long)nv);
            if (nf) atomicAdd(&stats[1], (unsigned long long)nf);
            if (nz) atomicAdd(&stats[2], (unsigned long long)nz);
        }
    }
}

#define B2_OPTICS_ARGS                                                                                         \
    const __grid_constant__ DevOptics o, const __grid_constant__ B2OpticsOptions opt, int64_t n,                   \
        double *__restrict__ x, double *__restrict__ y, double *__restrict__ dxdz, double *__restrict__ dydz,     \
        double *__restrict__ flux, const double *__restrict__ wl_nm, const double *__restrict__ pu,               \
        const double *__restrict__ pv, const double *__restrict__ time, const double *__restrict__ gauss,         \
        double *__restrict__ time_out, unsigned long long *__restrict__ stats
// The trace is latency bound on dependent FP64 chains (ncu: stall "wait" dominates at 4 warps per
// scheduler), so resident warps matter more than a few spilled registers: builds of the same body at
// 2 / 3 / 4 blocks per SM (B2_OPTICS_OCC picks one; default set from measurements), each as the generic
// interpreter and as the LSST surface program.
template <int MINB, int PROG>
__global__ void __launch_bounds__(256, MINB) k_rubin_optics(B2_OPTICS_ARGS) {
    rubin_optics_body<PROG>(o, opt, n, x, y, dxdz, dydz, flux, wl_nm, pu, pv, time, gauss, time_out, stats);
}

// RubinDiffraction.applyTo
__global__ void __launch_bounds__(256)
k_rubin_diffraction(const __grid_constant__ DevOptics o, const __grid_constant__ B2OpticsOptions opt, int64_t n,
                    double* __restrict__ x, double* __restrict__ y, const double* __restrict__ wl_nm,
                    const double* __restrict__ pu, const double* __restrict__ pv, const double* __restrict__ time,
                    const double* __restrict__ gauss) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double xi = x[i], yi = y[i];
    if (opt.shift_in) {
        xi += opt.stamp_center[0];
        yi += opt.stamp_center[1];
    }
    double wl = wl_nm[i] * 1e-9;
    double vx, vy, vz;
    xy_to_v(o, xi, yi, vx, vy, vz);
    double inair = b2rcp(medium_n(o.media[o.medium_stop], wl));
    vx *= inair;
    vy *= inair;
    vz *= inair;
    double g = gauss ? gauss[i] : philox_normal(opt.seed, opt.photon_offset + (uint64_t)i, 0u);
    diffraction_kick(o.dif, pu[i], pv[i], time[i], wl, g, vx, vy, vz);
    v_to_xy(o, vx, vy, vz, xi, yi);
    if (opt.shift_in) {
        xi -= opt.stamp_center[0];
        yi -= opt.stamp_center[1];
    }
    x[i] = xi;
    y[i] = yi;
}

// galsim.TimeSampler + galsim.PupilAnnulusSampler
__global__ void __launch_bounds__(256)
k_sample_time_pupil(int64_t n, double* __restrict__ time, double* __restrict__ pu, double* __restrict__ pv, double t0,
                    double exptime, double r_in, double r_out, uint64_t seed, uint64_t offset) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double t, u, v;
    sample_time_pupil(seed, offset + (uint64_t)i, t0, exptime, r_in, r_out, t, u, v);
    if (time) time[i] = t;
    if (pu) {
        pu[i] = u;
        pv[i] = v;
    }
}

// uniform photons over a rectangle with unit flux (imsim/flat.py:246-257) and wavelengths drawn from
// a tabulated inverse CDF (the role of galsim.WavelengthSampler, flat.py:177,259)
__global__ void __launch_bounds__(256)
k_flat_photons(int64_t n, double* __restrict__ x, double* __restrict__ y, double* __restrict__ flux,
               double* __restrict__ wl, double xlo, double xhi, double ylo, double yhi,
               const double* __restrict__ cdf, const double* __restrict__ cdf_wave, int ncdf, uint64_t seed,
               uint64_t offset) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t r[4], q[4];
    philox4(seed, offset + (uint64_t)i, 5u, r);
    x[i] = xlo + (xhi - xlo) * u01(r[0], r[1]);
    y[i] = ylo + (yhi - ylo) * u01(r[2], r[3]);
    flux[i] = 1.0;
    if (wl) {
        philox4(seed, offset + (uint64_t)i, 6u, q);
        double u = u01(q[0], q[1]);
        // largest k with cdf[k] <= u, then linear in the bin (piecewise-constant pdf)
        int lo = 0, hi = ncdf - 1;
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (__ldg(cdf + mid) <= u) lo = mid; else hi = mid;
        }
        double c0 = __ldg(cdf + lo), c1 = __ldg(cdf + hi);
        double f = (c1 > c0) ? (u - c0) / (c1 - c0) : 0.0;
        wl[i] = __ldg(cdf_wave + lo) + f * (__ldg(cdf_wave + hi) - __ldg(cdf_wave + lo));
    }
}

// Photons of a table of point-like objects seen through a Gaussian PSF, generated in HBM: thread k
// finds its object by binary search in the cumulative photon counts (the device form of
// merge_photon_arrays o build_stamps, imsim/photon_pooling.py:151-152), draws the PSF offset and
// the wavelength.  First step of SURVEY section 8 f1 (stage-1 generation on device): DeltaFunction
// objects x Gaussian PSF only; flux 1 per photon.
__global__ void __launch_bounds__(256)
k_object_photons(int64_t n, double* __restrict__ x, double* __restrict__ y, double* __restrict__ flux,
                 double* __restrict__ wl, const double* __restrict__ obj_x, const double* __restrict__ obj_y,
                 const double* __restrict__ obj_sigma, const int64_t* __restrict__ obj_cum, int32_t nobj,
                 const double* __restrict__ cdf, const double* __restrict__ cdf_wave, int ncdf, uint64_t seed,
                 uint64_t offset) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // object j owns photons [obj_cum[j], obj_cum[j+1])
    int lo = 0, hi = nobj;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(obj_cum + mid) <= i) lo = mid; else hi = mid;
    }
    uint32_t r[4];
    philox4(seed, offset + (uint64_t)i, 7u, r);
    float u1 = ((float)(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float u2 = ((float)(r[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    float rad = sqrtf(-2.0f * logf(u1));
    double sg = __ldg(obj_sigma + lo);
    x[i] = __ldg(obj_x + lo) + sg * (double)(rad * cs);
    y[i] = __ldg(obj_y + lo) + sg * (double)(rad * sn);
    flux[i] = 1.0;
    if (wl) {
        double u = ((double)r[2] * 4294967296.0 + (double)r[3] + 0.5) * (1.0 / 18446744073709551616.0);
        int a = 0, b = ncdf - 1;
        while (b - a > 1) {
            int mid = (a + b) >> 1;
            if (__ldg(cdf + mid) <= u) a = mid; else b = mid;
        }
        double c0 = __ldg(cdf + a), c1 = __ldg(cdf + b);
        double f = (c1 > c0) ? (u - c0) / (c1 - c0) : 0.0;
        wl[i] = __ldg(cdf_wave + a) + f * (__ldg(cdf_wave + b) - __ldg(cdf_wave + a));
    }
}

// ------------------------------------------------------------------ host side
static inline int nblocks(int64_t n, int bs = 256) { return (int)((n + bs - 1) / bs); }

static int sip_index(int i, int j) {
    // 00 01 02 03 10 11 12 20 21 30
    static const int base[4] = {0, 4, 7, 9};
    return base[i] + j;
}

static void fill_devwcs(const B2TanSip* w, DevWcs* d) {
    memset(d, 0, sizeof(*d));
    d->crpix[0] = w->crpix[0];
    d->crpix[1] = w->crpix[1];
    for (int k = 0; k < 4; ++k) d->cd[k] = w->cd[k];
    double det = w->cd[0] * w->cd[3] - w->cd[1] * w->cd[2];
    d->cdinv[0] = w->cd[3] / det;
    d->cdinv[1] = -w->cd[1] / det;
    d->cdinv[2] = -w->cd[2] / det;
    d->cdinv[3] = w->cd[0] / det;
    d->order = w->order;
    for (int k = 0; k < 2; ++k)
        for (int i = 0; i <= 3; ++i)
            for (int j = 0; j <= 3 - i; ++j) d->ab[k][sip_index(i, j)] = (i + j <= w->order) ? w->ab[k][i][j] : 0.0;
    // characteristic size of one degree in pixel units; Newton stops when the
    // step is below 1e-9 of it (quadratic convergence => next error ~1e-18)
    d->newton_tol = 1e-9 / sqrt(fabs(det));
}

static void tangent_basis(double ra, double dec, double e[3], double nn[3], double r[3]) {
    double sa = sin(ra), ca = cos(ra), sd = sin(dec), cd = cos(dec);
    e[0] = -sa; e[1] = ca; e[2] = 0.0;
    nn[0] = -sd * ca; nn[1] = -sd * sa; nn[2] = cd;
    r[0] = cd * ca; r[1] = cd * sa; r[2] = sd;
}

extern "C" int b2_wcs_upload(b2_ctx* ctx, const B2TanSip* img, const B2TanSip* field) {
    B2_REQUIRE(ctx && img && field, "b2_wcs_upload: null argument");
    B2_REQUIRE(img->order >= 0 && img->order <= 3 && field->order >= 0 && field->order <= 3,
               "b2_wcs_upload: SIP order must be 0..3");
    fill_devwcs(img, &ctx->opt.img);
    fill_devwcs(field, &ctx->opt.field);
    double e0[3], n0[3], r0[3], e1[3], n1[3], r1[3];
    tangent_basis(img->ra0, img->dec0, e0, n0, r0);
    tangent_basis(field->ra0, field->dec0, e1, n1, r1);
    const double* rows[3] = {e1, n1, r1};
    const double* cols[3] = {e0, n0, r0};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j)
            ctx->opt.M_if[3 * i + j] = rows[i][0] * cols[j][0] + rows[i][1] * cols[j][1] + rows[i][2] * cols[j][2];
    ctx->img_host = *img;
    ctx->field_host = *field;
    ctx->have_wcs = true;
    ctx->opt.xyv.enabled = 0;  // a compiled XyToV belongs to the previous WCS pair
    return 0;
}

// XyToV compiled for one detector.  The field tangents (thx, thy) as functions of the pixel position are
// interpolated at the 12 x 12 Chebyshev points of the box [xlo, xhi] x [ylo, yhi] by a tensor Chebyshev series
// (discrete orthogonality: a projection, no linear solve), truncated to degree 5 x 5, converted to monomials of
// the scaled coordinates, and checked against the exact chain on a 33 x 33 grid of other points.  The
// polynomial is adopted only if it reproduces the exact chain to tol_px pixels; max_resid_px reports it either way.
extern "C" int b2_xytov_compile(b2_ctx* ctx, double xlo, double xhi, double ylo, double yhi, double tol_px,
                                double* max_resid_px) {
    B2_REQUIRE(ctx && ctx->have_wcs, "b2_xytov_compile: upload the WCS pair first");
    B2_REQUIRE(xhi > xlo && yhi > ylo, "b2_xytov_compile: empty box");
    DevOptics& o = ctx->opt;
    o.xyv.enabled = 0;
    constexpr int NC = 12, ND = B2_XYPOLY_N;
    const double cx0 = 0.5 * (xlo + xhi), cy0 = 0.5 * (ylo + yhi), hx = 0.5 * (xhi - xlo), hy = 0.5 * (yhi - ylo);
    double node[NC], T[NC][ND];
    for (int k = 0; k < NC; ++k) {
        node[k] = cos(M_PI * (k + 0.5) / NC);
        for (int d = 0; d < ND; ++d) T[k][d] = cos(d * M_PI * (k + 0.5) / NC);
    }
    double fx[NC][NC], fy[NC][NC];
    for (int a = 0; a < NC; ++a)
        for (int b = 0; b < NC; ++b) xy_to_field_exact(o, cx0 + hx * node[a], cy0 + hy * node[b], fx[a][b], fy[a][b]);
    // Chebyshev coefficients c[i][j] of T_i(X) T_j(Y)
    double chx[ND][ND], chy[ND][ND];
    for (int i = 0; i < ND; ++i)
        for (int j = 0; j < ND; ++j) {
            double sx = 0.0, sy = 0.0;
            for (int a = 0; a < NC; ++a)
                for (int b = 0; b < NC; ++b) {
                    sx += fx[a][b] * T[a][i] * T[b][j];
                    sy += fy[a][b] * T[a][i] * T[b][j];
                }
            const double w = (i ? 2.0 : 1.0) * (j ? 2.0 : 1.0) / (NC * NC);
            chx[i][j] = w * sx;
            chy[i][j] = w * sy;
        }
    // T_d(x) = sum_m tm[d][m] x^m
    double tm[ND][ND] = {};
    tm[0][0] = 1.0;
    if (ND > 1) tm[1][1] = 1.0;
    for (int d = 2; d < ND; ++d)
        for (int m = 0; m < ND; ++m) tm[d][m] = (m ? 2.0 * tm[d - 1][m - 1] : 0.0) - tm[d - 2][m];
    DevXyPoly p;
    memset(&p, 0, sizeof(p));
    for (int i = 0; i < ND; ++i)
        for (int j = 0; j < ND; ++j)
            for (int a = 0; a < ND; ++a)
                for (int b = 0; b < ND; ++b) {
                    p.cx[a][b] += chx[i][j] * tm[i][a] * tm[j][b];
                    p.cy[a][b] += chy[i][j] * tm[i][a] * tm[j][b];
                }
    p.box[0] = xlo; p.box[1] = xhi; p.box[2] = ylo; p.box[3] = yhi;
    p.c0[0] = cx0; p.c0[1] = cy0;
    p.sc[0] = 1.0 / hx; p.sc[1] = 1.0 / hy;
    // validation against the exact chain, in pixels: radians of field angle per pixel from the box diagonal
    double t00x, t00y, t11x, t11y;
    xy_to_field_exact(o, xlo, ylo, t00x, t00y);
    xy_to_field_exact(o, xhi, yhi, t11x, t11y);
    const double rad_per_px = hypot(t11x - t00x, t11y - t00y) / hypot(xhi - xlo, yhi - ylo);
    B2_REQUIRE(rad_per_px > 0.0 && std::isfinite(rad_per_px), "b2_xytov_compile: degenerate WCS over the box");
    double worst = 0.0;
    constexpr int NV = 33;
    for (int a = 0; a < NV; ++a)
        for (int b = 0; b < NV; ++b) {
            const double X = -1.0 + 2.0 * a / (NV - 1), Y = -1.0 + 2.0 * b / (NV - 1);
            double ex, ey;
            xy_to_field_exact(o, cx0 + hx * X, cy0 + hy * Y, ex, ey);
            double gx = 0.0, gy = 0.0;
            for (int i = ND - 1; i >= 0; --i) {
                double rx = p.cx[i][ND - 1], ry = p.cy[i][ND - 1];
                for (int j = ND - 2; j >= 0; --j) {
                    rx = fma(rx, Y, p.cx[i][j]);
                    ry = fma(ry, Y, p.cy[i][j]);
                }
                gx = fma(gx, X, rx);
                gy = fma(gy, X, ry);
            }
            worst = std::max(worst, std::max(fabs(gx - ex), fabs(gy - ey)) / rad_per_px);
        }
    if (max_resid_px) *max_resid_px = worst;
    if (worst <= tol_px) {
        p.enabled = 1;
        o.xyv = p;
    }
    return 0;
}

extern "C" int b2_detector_upload(b2_ctx* ctx, const B2Detector* det) {
    B2_REQUIRE(ctx && det, "b2_detector_upload: null argument");
    ctx->opt.det = *det;
    ctx->have_det = true;
    return 0;
}

extern "C" int b2_diffraction_config(b2_ctx* ctx, const B2Diffraction* cfg) {
    B2_REQUIRE(ctx && cfg, "b2_diffraction_config: null argument");
    B2_REQUIRE(cfg->n_lines <= 8 && cfg->n_circles <= 4, "b2_diffraction_config: too many primitives");
    ctx->opt.dif = *cfg;
    return 0;
}

static void select_program(b2_ctx* ctx);

extern "C" int b2_telescope_upload(b2_ctx* ctx, const B2Telescope* tel) {
    B2_REQUIRE(ctx && tel, "b2_telescope_upload: null argument");
    B2_REQUIRE(tel->n_surfaces >= 1 && tel->n_surfaces <= B2_DEV_MAX_SURF, "b2_telescope_upload: 1..16 surfaces supported");
    B2_REQUIRE(tel->n_media >= 1 && tel->n_media <= B2_DEV_MAX_MEDIA, "b2_telescope_upload: 1..4 media supported");
    DevOptics& o = ctx->opt;
    o.n_surf = tel->n_surfaces;
    o.n_media = tel->n_media;
    o.medium_stop = tel->medium_stop;
    for (int m = 0; m < tel->n_media; ++m) {
        o.media[m] = tel->media[m];
        if (o.media[m].kind == B2_MED_AIR) {  // photon-independent factors of batoid.Air, see medium_n()
            const double P = o.media[m].p[0] * 7.50061683, T = o.media[m].p[1] - 273.15, W = o.media[m].p[2] * 7.50061683;
            o.media[m].p[3] = P * (1.0 + (1.049 - 0.0157 * T) * 1.e-6 * P) / (720.883 * (1.0 + 0.003661 * T));
            o.media[m].p[4] = W * 1.e-6 / (1.0 + 0.003661 * T);
        }
    }
    for (int i = 0; i < tel->n_surfaces; ++i) {
        const B2Surface& s = tel->surf[i];
        DevSurf& d = o.surf[i];
        memset(&d, 0, sizeof(d));
        d.kind = s.surf_kind;
        d.interact = s.interact;
        B2_REQUIRE(s.interact != B2_INT_PASS || s.surf_kind == B2_SURF_PLANE,
                   "b2_telescope_upload: OPDScreen interfaces are supported on a Plane surface only");
        d.med_in = s.medium_in;
        d.med_out = s.medium_out;
        d.n_coef = s.n_coef;
        d.rot_identity = s.rot_identity;
        d.n_obsc = s.n_obsc;
        d.extra_kind = B2_EXTRA_NONE;  // armed by b2_telescope_set_extra
        d.R = s.R;
        d.invR = (s.surf_kind == B2_SURF_PLANE) ? 0.0 : 1.0 / s.R;
        d.k1 = (s.surf_kind == B2_SURF_PARABOLOID) ? 0.0 : (s.surf_kind == B2_SURF_SPHERE ? 1.0 : 1.0 + s.conic);
        if (s.surf_kind == B2_SURF_PLANE) d.k1 = 0.0;
        for (int k = 0; k < B2_MAX_ASPHERE_COEF; ++k) d.coef[k] = s.coef[k];
        for (int k = 0; k < 3; ++k) d.dr[k] = s.dr[k];
        for (int k = 0; k < 9; ++k) d.drot[k] = s.drot[k];
        for (int k = 0; k < s.n_obsc; ++k) {
            const B2Obsc& ob = s.obsc[k];
            DevObsc& dob = d.obsc[k];
            dob.kind = ob.kind;
            dob.negate = ob.negate;
            for (int j = 0; j < 6; ++j) dob.p[j] = ob.p[j];
            if (ob.kind == B2_OBSC_CIRCLE) dob.p[0] = ob.p[0] * ob.p[0];
            if (ob.kind == B2_OBSC_ANNULUS) { dob.p[0] = ob.p[0] * ob.p[0]; dob.p[1] = ob.p[1] * ob.p[1]; }
            if (ob.kind == B2_OBSC_RECTANGLE) { dob.p[0] = ob.p[0] / 2; dob.p[1] = ob.p[1] / 2; }
            if (ob.kind == B2_OBSC_RAY) dob.p[0] = ob.p[0] / 2;
        }
        d.simple_clear = 0;
        if (s.n_obsc == 1 && s.obsc[0].negate) {
            const B2Obsc& ob = s.obsc[0];
            if (ob.kind == B2_OBSC_CIRCLE && ob.p[1] == 0.0 && ob.p[2] == 0.0) {
                d.simple_clear = 1;
                d.clr_in2 = -1.0;
                d.clr_out2 = ob.p[0] * ob.p[0];
            } else if (ob.kind == B2_OBSC_ANNULUS && ob.p[2] == 0.0 && ob.p[3] == 0.0) {
                d.simple_clear = 1;
                d.clr_in2 = ob.p[0] * ob.p[0];
                d.clr_out2 = ob.p[1] * ob.p[1];
            }
        }
        d.poly_n = s.poly_n;
        d.poly_scale = s.poly_scale;
        d.extra = nullptr;
        if (s.extra_kind != B2_EXTRA_NONE) d.pad = s.extra_kind;  // remembered until the table arrives
    }
    ctx->have_tel = true;
    select_program(ctx);
    return 0;
}

extern "C" int b2_telescope_set_extra(b2_ctx* ctx, int is, int kind, const double* data, int64_t n) {
    B2_REQUIRE(ctx && ctx->have_tel, "b2_telescope_set_extra: upload the telescope first");
    B2_REQUIRE(is >= 0 && is < ctx->opt.n_surf, "b2_telescope_set_extra: bad surface index");
    B2_REQUIRE(kind == B2_EXTRA_POLY2D || kind == B2_EXTRA_BICUBIC, "b2_telescope_set_extra: bad kind");
    DevSurf& d = ctx->opt.surf[is];
    if (kind == B2_EXTRA_POLY2D)
        B2_REQUIRE(n == (int64_t)d.poly_n * d.poly_n && d.poly_n <= B2_MAX_POLY_ORDER, "b2_telescope_set_extra: poly size mismatch");
    B2_CUDA(cudaSetDevice(ctx->device));
    b2_ctx::ExtraTable& e = ctx->extras[is];
    const size_t bytes = (size_t)n * sizeof(double);
    if (b2_scratch_reserve(ctx, e.dev, bytes)) return 1;  // grow-only, reused by every later upload
    const int k = e.next;
    e.next ^= 1;
    if (e.pin_bytes[k] < bytes) {
        if (e.ev[k]) B2_CUDA(cudaEventSynchronize(e.ev[k]));
        if (e.pin[k]) B2_CUDA(cudaFreeHost(e.pin[k]));
        e.pin[k] = nullptr;
        e.pin_bytes[k] = 0;
        B2_CUDA(cudaMallocHost(&e.pin[k], bytes));
        e.pin_bytes[k] = bytes;
    }
    if (!e.ev[k]) B2_CUDA(cudaEventCreateWithFlags(&e.ev[k], cudaEventDisableTiming));
    else B2_CUDA(cudaEventSynchronize(e.ev[k]));  // the copy that last read this slot (two uploads ago)
    memcpy(e.pin[k], data, bytes);
    B2_CUDA(cudaMemcpyAsync(e.dev.ptr, e.pin[k], bytes, cudaMemcpyHostToDevice, ctx->stream));
    B2_CUDA(cudaEventRecord(e.ev[k], ctx->stream));
    d.extra = (const double*)e.dev.ptr;
    d.extra_kind = kind;
    select_program(ctx);
    return 0;
}

static int check_ready(b2_ctx* ctx, bool need_tel, bool need_wcs, bool need_det) {
    B2_REQUIRE(ctx, "null context");
    if (need_tel) {
        B2_REQUIRE(ctx->have_tel, "telescope not uploaded");
        for (int i = 0; i < ctx->opt.n_surf; ++i)
            B2_REQUIRE(ctx->opt.surf[i].pad == 0 || ctx->opt.surf[i].extra != nullptr,
                       "a surface declares an extra sag term but b2_telescope_set_extra was not called");
    }
    if (need_wcs) B2_REQUIRE(ctx->have_wcs, "wcs not uploaded");
    if (need_det) B2_REQUIRE(ctx->have_det, "detector not uploaded");
    B2_CUDA(cudaSetDevice(ctx->device));
    return 0;
}


extern "C" int b2_xy_to_v(b2_ctx* ctx, int64_t n, const double* x, const double* y, double* vx, double* vy, double* vz,
                          int where) {
    if (check_ready(ctx, false, true, false)) return 1;
    if (n <= 0) return 0;
    if (where == B2_HOST) {
        Stager st{ctx};
        if (st.init(5 * pad256(n * 8))) return 1;
        double *dx = st.take<double>(n), *dy = st.take<double>(n), *a = st.take<double>(n), *b = st.take<double>(n),
               *c = st.take<double>(n);
        H2D(dx, x, n);
        H2D(dy, y, n);
        k_xy_to_v<<<nblocks(n), 256, 0, ctx->stream>>>(ctx->opt, n, dx, dy, a, b, c);
        B2_CHECK_LAUNCH();
        D2H(vx, a, n);
        D2H(vy, b, n);
        D2H(vz, c, n);
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
        k_xy_to_v<<<nblocks(n), 256, 0, ctx->stream>>>(ctx->opt, n, x, y, vx, vy, vz);
        B2_CHECK_LAUNCH();
    }
    return 0;
}

extern "C" int b2_v_to_xy(b2_ctx* ctx, int64_t n, const double* vx, const double* vy, const double* vz, double* x,
                          double* y, int where) {
    if (check_ready(ctx, false, true, false)) return 1;
    if (n <= 0) return 0;
    if (where == B2_HOST) {
        Stager st{ctx};
        if (st.init(5 * pad256(n * 8))) return 1;
        double *a = st.take<double>(n), *b = st.take<double>(n), *c = st.take<double>(n), *dx = st.take<double>(n),
               *dy = st.take<double>(n);
        H2D(a, vx, n);
        H2D(b, vy, n);
        H2D(c, vz, n);
        k_v_to_xy<<<nblocks(n), 256, 0, ctx->stream>>>(ctx->opt, n, a, b, c, dx, dy);
        B2_CHECK_LAUNCH();
        D2H(x, dx, n);
        D2H(y, dy, n);
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
        k_v_to_xy<<<nblocks(n), 256, 0, ctx->stream>>>(ctx->opt, n, vx, vy, vz, x, y);
        B2_CHECK_LAUNCH();
    }
    return 0;
}

extern "C" int b2_trace_rays(b2_ctx* ctx, int64_t n, double* x, double* y, double* z, double* vx, double* vy,
                             double* vz, double* t, const double* wl, uint8_t* vig, uint8_t* fail, int where) {
    if (check_ready(ctx, true, false, false)) return 1;
    if (n <= 0) return 0;
    if (where == B2_HOST) {
        Stager st{ctx};
        if (st.init(8 * pad256(n * 8) + 2 * pad256(n))) return 1;
        double* d[8];
        for (int k = 0; k < 8; ++k) d[k] = st.take<double>(n);
        uint8_t *dv = st.take<uint8_t>(n), *df = st.take<uint8_t>(n);
        double* h[8] = {x, y, z, vx, vy, vz, t, (double*)wl};
        for (int k = 0; k < 8; ++k) H2D(d[k], h[k], n);
        B2_CUDA(cudaMemcpyAsync(dv, vig, n, cudaMemcpyHostToDevice, ctx->stream));
        B2_CUDA(cudaMemcpyAsync(df, fail, n, cudaMemcpyHostToDevice, ctx->stream));
        if (ctx->program == B2_PROG_LSST)
            k_trace_rays<B2_PROG_LSST><<<nblocks(n), 256, 0, ctx->stream>>>(ctx->opt, n, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], dv, df);
        else
            k_trace_rays<B2_PROG_GENERIC><<<nblocks(n), 256, 0, ctx->stream>>>(ctx->opt, n, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], dv, df);
        B2_CHECK_LAUNCH();
        for (int k = 0; k < 7; ++k) D2H(h[k], d[k], n);
        B2_CUDA(cudaMemcpyAsync(vig, dv, n, cudaMemcpyDeviceToHost, ctx->stream));
        B2_CUDA(cudaMemcpyAsync(fail, df, n, cudaMemcpyDeviceToHost, ctx->stream));
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
        if (ctx->program == B2_PROG_LSST)
            k_trace_rays<B2_PROG_LSST><<<nblocks(n), 256, 0, ctx->stream>>>(ctx->opt, n, x, y, z, vx, vy, vz, t, wl, vig, fail);
        else
            k_trace_rays<B2_PROG_GENERIC><<<nblocks(n), 256, 0, ctx->stream>>>(ctx->opt, n, x, y, z, vx, vy, vz, t, wl, vig, fail);
        B2_CHECK_LAUNCH();
    }
    return 0;
}

extern "C" int b2_telescope_program(b2_ctx* ctx) { return (ctx && ctx->have_tel) ? ctx->program : -1; }

static int optics_occ(int program) {
    static int occ = -1;
    if (occ < 0) {
        const char* e = getenv("B2_OPTICS_OCC");
        occ = e ? atoi(e) : 0;
        if (occ < 2 || occ > 4) occ = 0;
    }
    if (occ) return occ;
    return program == B2_PROG_LSST ? 4 : 3;  // k_rubin_optics per 2^24 photons: program 2.50 / 2.37 ms at 3 / 4; interpreter 2.95 at 3
}

// Surface program of the uploaded telescope (optics_device.cuh); B2_PROGRAM=0 forces the interpreter.
static void select_program(b2_ctx* ctx) {
    const DevOptics& o = ctx->opt;
    ctx->program = B2_PROG_GENERIC;
    const char* e = getenv("B2_PROGRAM");
    if (e && e[0] == '0') return;
    if (o.n_surf != B2_PROG_LSST_LEN || o.n_media != 2) return;
    for (int i = 0; i < o.n_surf; ++i) {
        const DevSurf& d = o.surf[i];
        const SurfSpec sp = lsst_spec(i);
        if ((d.kind == B2_SURF_ASPHERE) != sp.asphere || d.interact != sp.interact) return;
        if (sp.interact == B2_INT_REFRACT && (d.med_in != sp.med_in || d.med_out != sp.med_out)) return;
        if (d.extra_kind != B2_EXTRA_NONE || d.pad != 0 || !d.simple_clear || d.n_coef > 4) return;
    }
    ctx->program = B2_PROG_LSST;
}

static void launch_rubin_optics(b2_ctx* ctx, cudaStream_t st, const B2OpticsOptions& opt, int64_t n, double* x,
                                double* y, double* dxdz, double* dydz, double* flux, const double* wl, const double* pu,
                                const double* pv, const double* time, const double* gauss, double* time_out,
                                unsigned long long* stats) {
    B2_TIMED("k_rubin_optics", st);
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx->record_events) {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
    }
#define B2_LAUNCH_OPTICS(MINB, PROG)                                                                      \
    k_rubin_optics<MINB, PROG><<<nblocks(n), 256, 0, st>>>(ctx->opt, opt, n, x, y, dxdz, dydz, flux, wl, pu, pv, \
                                                           time, gauss, time_out, stats)
    const int occ = optics_occ(ctx->program);
    if (ctx->program == B2_PROG_LSST) {
        if (occ == 2) B2_LAUNCH_OPTICS(2, B2_PROG_LSST);
        else if (occ == 4) B2_LAUNCH_OPTICS(4, B2_PROG_LSST);
        else B2_LAUNCH_OPTICS(3, B2_PROG_LSST);
    } else {
        if (occ == 2) B2_LAUNCH_OPTICS(2, B2_PROG_GENERIC);
        else if (occ == 4) B2_LAUNCH_OPTICS(4, B2_PROG_GENERIC);
        else B2_LAUNCH_OPTICS(3, B2_PROG_GENERIC);
    }
#undef B2_LAUNCH_OPTICS
    if (ctx->record_events) {
        cudaEventRecord(e1, st);
        ctx->events.emplace_back(e0, e1);
    }
}

extern "C" int b2_rubin_optics(b2_ctx* ctx, int64_t n, double* x, double* y, double* dxdz, double* dydz, double* flux,
                               const double* wl_nm, const double* pu, const double* pv, const double* time,
                               const double* gauss, double* time_out, const B2OpticsOptions* opt, int where,
                               B2OpticsStats* stats) {
    if (check_ready(ctx, true, true, true)) return 1;
    B2_REQUIRE(opt, "b2_rubin_optics: null options");
    B2_REQUIRE(x && y && dxdz && dydz && flux && wl_nm, "b2_rubin_optics: null photon array");
    // imsim/photon_ops.py:139-140 asserts the pupil and time arrays are allocated
    B2_REQUIRE(pu && pv, "b2_rubin_optics: photon array has no pupil coordinates (hasAllocatedPupil)");
    B2_REQUIRE(time, "b2_rubin_optics: photon array has no time stamps (hasAllocatedTimes)");
    if (stats) memset(stats, 0, sizeof(*stats));
    if (n <= 0) return 0;
    unsigned long long* dstats = nullptr;
    if (stats) {
        if (b2_scratch_reserve(ctx, ctx->stats, 64)) return 1;
        dstats = (unsigned long long*)ctx->stats.ptr;
        B2_CUDA(cudaMemsetAsync(dstats, 0, 64, ctx->stream));
    }
    if (where == B2_HOST) {
        Stager st{ctx};
        if (st.init(11 * pad256(n * 8))) return 1;
        double *dx = st.take<double>(n), *dy = st.take<double>(n), *da = st.take<double>(n), *db = st.take<double>(n),
               *df = st.take<double>(n), *dw = st.take<double>(n), *du = st.take<double>(n), *dv = st.take<double>(n),
               *dt = st.take<double>(n), *dg = st.take<double>(n), *dto = st.take<double>(n);
        if (b2_pipe_enabled(n)) {
            // large pageable arrays: chunks through the pinned ring, copies and kernel overlapped (hostpipe.cu)
            const double* hin[8] = {x, y, flux, wl_nm, pu, pv, time, gauss};
            double* din[8] = {dx, dy, df, dw, du, dv, dt, dg};
            double* hout[6] = {x, y, dxdz, dydz, flux, time_out};
            const double* dout[6] = {dx, dy, da, db, df, dto};
            const std::function<int(int64_t, int64_t, cudaStream_t)> chunk_kernel = [&](int64_t off, int64_t cnt,
                                                                                       cudaStream_t cs) -> int {
                B2OpticsOptions o = *opt;
                o.photon_offset += (uint64_t)off;  // the kick's Philox counter is the photon's index in the call
                launch_rubin_optics(ctx, cs, o, cnt, dx + off, dy + off, da + off, db + off, df + off, dw + off, du + off,
                                    dv + off, dt + off, gauss ? dg + off : nullptr, time_out ? dto + off : nullptr,
                                    dstats);
                B2_CHECK_LAUNCH();
                return 0;
            };
            if (b2_pipe_run(ctx, n, gauss ? 8 : 7, hin, din, time_out ? 6 : 5, hout, dout, &chunk_kernel)) return 1;
            if (stats) {
                B2_CUDA(cudaMemcpyAsync(stats, dstats, sizeof(*stats), cudaMemcpyDeviceToHost, ctx->stream));
                B2_CUDA(cudaStreamSynchronize(ctx->stream));
            }
            return 0;
        }
        H2D(dx, x, n);
        H2D(dy, y, n);
        H2D(df, flux, n);
        H2D(dw, wl_nm, n);
        H2D(du, pu, n);
        H2D(dv, pv, n);
        H2D(dt, time, n);
        if (gauss) H2D(dg, gauss, n);
        launch_rubin_optics(ctx, ctx->stream, *opt, n, dx, dy, da, db, df, dw, du, dv, dt, gauss ? dg : nullptr,
                            time_out ? dto : nullptr, dstats);
        B2_CHECK_LAUNCH();
        D2H(x, dx, n);
        D2H(y, dy, n);
        D2H(dxdz, da, n);
        D2H(dydz, db, n);
        D2H(flux, df, n);
        if (time_out) D2H(time_out, dto, n);
        if (stats) B2_CUDA(cudaMemcpyAsync(stats, dstats, sizeof(*stats), cudaMemcpyDeviceToHost, ctx->stream));
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
        launch_rubin_optics(ctx, ctx->stream, *opt, n, x, y, dxdz, dydz, flux, wl_nm, pu, pv, time, gauss, time_out,
                            dstats);
        B2_CHECK_LAUNCH();
        if (stats) {
            B2_CUDA(cudaMemcpyAsync(stats, dstats, sizeof(*stats), cudaMemcpyDeviceToHost, ctx->stream));
            B2_CUDA(cudaStreamSynchronize(ctx->stream));
        }
    }
    return 0;
}

extern "C" int b2_rubin_diffraction(b2_ctx* ctx, int64_t n, double* x, double* y, const double* wl_nm, const double* pu,
                                    const double* pv, const double* time, const double* gauss,
                                    const B2OpticsOptions* opt, int where) {
    if (check_ready(ctx, true, true, false)) return 1;
    B2_REQUIRE(opt && x && y && wl_nm, "b2_rubin_diffraction: null argument");
    B2_REQUIRE(pu && pv, "b2_rubin_diffraction: photon array has no pupil coordinates (hasAllocatedPupil)");
    B2_REQUIRE(time, "b2_rubin_diffraction: photon array has no time stamps (hasAllocatedTimes)");
    B2_REQUIRE(ctx->opt.dif.enabled, "b2_rubin_diffraction: diffraction not configured");
    if (n <= 0) return 0;
    if (where == B2_HOST) {
        Stager st{ctx};
        if (st.init(7 * pad256(n * 8))) return 1;
        double *dx = st.take<double>(n), *dy = st.take<double>(n), *dw = st.take<double>(n), *du = st.take<double>(n),
               *dv = st.take<double>(n), *dt = st.take<double>(n), *dg = st.take<double>(n);
        H2D(dx, x, n);
        H2D(dy, y, n);
        H2D(dw, wl_nm, n);
        H2D(du, pu, n);
        H2D(dv, pv, n);
        H2D(dt, time, n);
        if (gauss) H2D(dg, gauss, n);
        k_rubin_diffraction<<<nblocks(n), 256, 0, ctx->stream>>>(ctx->opt, *opt, n, dx, dy, dw, du, dv, dt,
                                                                 gauss ? dg : nullptr);
        B2_CHECK_LAUNCH();
        D2H(x, dx, n);
        D2H(y, dy, n);
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
        k_rubin_diffraction<<<nblocks(n), 256, 0, ctx->stream>>>(ctx->opt, *opt, n, x, y, wl_nm, pu, pv, time, gauss);
        B2_CHECK_LAUNCH();
    }
    return 0;
}

extern "C" int b2_sample_time_pupil(b2_ctx* ctx, int64_t n, double* time, double* pu, double* pv, double t0,
                                    double exptime, double r_in, double r_out, uint64_t seed, uint64_t offset,
                                    int where) {
    B2_REQUIRE(ctx, "null context");
    B2_REQUIRE((pu == nullptr) == (pv == nullptr), "b2_sample_time_pupil: pupil_u and pupil_v go together");
    B2_CUDA(cudaSetDevice(ctx->device));
    if (n <= 0) return 0;
    if (where == B2_HOST) {
        Stager st{ctx};
        if (st.init(3 * pad256(n * 8))) return 1;
        double *dt = st.take<double>(n), *du = st.take<double>(n), *dv = st.take<double>(n);
        k_sample_time_pupil<<<nblocks(n), 256, 0, ctx->stream>>>(n, time ? dt : nullptr, pu ? du : nullptr,
                                                                 pv ? dv : nullptr, t0, exptime, r_in, r_out, seed, offset);
        B2_CHECK_LAUNCH();
        if (time) D2H(time, dt, n);
        if (pu) {
            D2H(pu, du, n);
            D2H(pv, dv, n);
        }
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
        B2_TIMED("k_sample_time_pupil", ctx->stream);
        k_sample_time_pupil<<<nblocks(n), 256, 0, ctx->stream>>>(n, time, pu, pv, t0, exptime, r_in, r_out, seed, offset);
        B2_CHECK_LAUNCH();
    }
    return 0;
}

// ------------------------------------------------------------------ FMA peak probe
// Register-resident chains of independent FMAs: the FP64 / FP32 pipe ceilings used
// as roofline denominators for the (compute-bound) trace kernel.
template <typename T>
__global__ void __launch_bounds__(256) k_fma_peak(T* out, int iters, T a, T b) {
    T x0 = a + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
    for (int i = 0; i < iters; ++i) {
        x0 = x0 * a + b; x1 = x1 * a + b; x2 = x2 * a + b; x3 = x3 * a + b;
        x4 = x4 * a + b; x5 = x5 * a + b; x6 = x6 * a + b; x7 = x7 * a + b;
    }
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" int b2_fma_peak(b2_ctx* ctx, int32_t fp64, double* tflops) {
    B2_REQUIRE(ctx && tflops, "b2_fma_peak: null argument");
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaDeviceProp prop;
    B2_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = fp64 ? 8192 : 16384;
    if (b2_scratch_reserve(ctx, ctx->scratch, (size_t)blocks * threads * 8)) return 1;
    cudaEvent_t e0, e1;
    B2_CUDA(cudaEventCreate(&e0));
    B2_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        B2_CUDA(cudaEventRecord(e0, ctx->stream));
        if (fp64) k_fma_peak<double><<<blocks, threads, 0, ctx->stream>>>((double*)ctx->scratch.ptr, iters, 0.999999, 1e-7);
        else k_fma_peak<float><<<blocks, threads, 0, ctx->stream>>>((float*)ctx->scratch.ptr, iters, 0.999999f, 1e-7f);
        B2_CHECK_LAUNCH();
        B2_CUDA(cudaEventRecord(e1, ctx->stream));
        B2_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        B2_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    *tflops = flops / (best * 1e-3) / 1e12;
    return 0;
}

// atomicAdd throughput on an image-sized buffer: the ceiling of the charge deposit (roofline denominator of
// the accumulate kernels).  pattern 0: uniform random pixels, 1: every thread hits the same pixel,
// 2: Gaussian blobs (1000 stars, sigma 1.5 px) like the bright-star workload.
template <typename T>
__global__ void __launch_bounds__(256) k_atomic_peak(T* img, int nx, int ny, int64_t n, int pattern, uint64_t seed) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t r[4];
    philox4(seed, (uint64_t)i, 30u, r);
    size_t k;
    if (pattern == 0) {
        k = (size_t)(r[0] % (uint32_t)nx) + (size_t)(r[1] % (uint32_t)ny) * nx;
    } else if (pattern == 1) {
        k = (size_t)(ny / 2) * nx + nx / 2;
    } else {
        uint32_t star = r[2] % 1000u;
        uint32_t q[4];
        philox4(seed + 1, (uint64_t)star, 31u, q);
        float u1 = ((float)(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f), u2 = ((float)(r[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        float rad = 1.5f * sqrtf(-2.0f * logf(u1));
        int px = (int)(q[0] % (uint32_t)(nx - 40)) + 20 + (int)lrintf(rad * cs);
        int py = (int)(q[1] % (uint32_t)(ny - 40)) + 20 + (int)lrintf(rad * sn);
        k = (size_t)py * nx + px;
    }
    atomicAdd(&img[k], (T)1);
}

extern "C" int b2_atomic_peak(b2_ctx* ctx, int32_t fp64, int32_t pattern, int32_t nx, int32_t ny, int64_t n,
                              double* atomics_per_s) {
    B2_REQUIRE(ctx && atomics_per_s && nx > 64 && ny > 64 && n > 0, "b2_atomic_peak: bad argument");
    B2_CUDA(cudaSetDevice(ctx->device));
    if (b2_scratch_reserve(ctx, ctx->scratch, (size_t)nx * ny * 8)) return 1;
    B2_CUDA(cudaMemsetAsync(ctx->scratch.ptr, 0, (size_t)nx * ny * 8, ctx->stream));
    cudaEvent_t e0, e1;
    B2_CUDA(cudaEventCreate(&e0));
    B2_CUDA(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        B2_CUDA(cudaEventRecord(e0, ctx->stream));
        if (fp64) k_atomic_peak<double><<<nblocks(n), 256, 0, ctx->stream>>>((double*)ctx->scratch.ptr, nx, ny, n, pattern, 7 + rep);
        else k_atomic_peak<float><<<nblocks(n), 256, 0, ctx->stream>>>((float*)ctx->scratch.ptr, nx, ny, n, pattern, 7 + rep);
        B2_CHECK_LAUNCH();
        B2_CUDA(cudaEventRecord(e1, ctx->stream));
        B2_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        B2_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *atomics_per_s = (double)n / (best * 1e-3);
    return 0;
}

extern "C" int b2_flat_photons(b2_ctx* ctx, int64_t n, double* x, double* y, double* flux, double* wl, double xlo,
                               double xhi, double ylo, double yhi, const double* cdf, const double* cdf_wave,
                               int32_t ncdf, uint64_t seed, uint64_t offset) {
    B2_REQUIRE(ctx && x && y && flux, "b2_flat_photons: null argument");
    B2_REQUIRE(!wl || (cdf && cdf_wave && ncdf >= 2), "b2_flat_photons: wavelength sampling needs a CDF table");
    B2_CUDA(cudaSetDevice(ctx->device));
    if (n <= 0) return 0;
    B2_TIMED("k_flat_photons", ctx->stream);
    k_flat_photons<<<nblocks(n), 256, 0, ctx->stream>>>(n, x, y, flux, wl, xlo, xhi, ylo, yhi, cdf, cdf_wave, ncdf, seed, offset);
    B2_CHECK_LAUNCH();
    return 0;
}

extern "C" int b2_object_photons(b2_ctx* ctx, int64_t n, double* x, double* y, double* flux, double* wl,
                                 const double* obj_x, const double* obj_y, const double* obj_sigma,
                                 const int64_t* obj_cum, int32_t nobj, const double* cdf, const double* cdf_wave,
                                 int32_t ncdf, uint64_t seed, uint64_t offset) {
    B2_REQUIRE(ctx && x && y && flux && obj_x && obj_y && obj_sigma && obj_cum && nobj > 0, "b2_object_photons: null argument");
    B2_REQUIRE(!wl || (cdf && cdf_wave && ncdf >= 2), "b2_object_photons: wavelength sampling needs a CDF table");
    B2_CUDA(cudaSetDevice(ctx->device));
    if (n <= 0) return 0;
    B2_TIMED("k_object_photons", ctx->stream);
    k_object_photons<<<nblocks(n), 256, 0, ctx->stream>>>(n, x, y, flux, wl, obj_x, obj_y, obj_sigma, obj_cum, nobj, cdf,
                                                          cdf_wave, ncdf, seed, offset);
    B2_CHECK_LAUNCH();
    return 0;
}
