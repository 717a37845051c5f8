/*
 * oracle_optics.c -- CPU restatement of the optics half of the photon path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under imsim_b200/ may import, link or
 * execute this file; it is the checker for the CUDA path (tests/, smoke(),
 * bench.py's cpu_baseline / --impl reference legs).
 *
 * Provenance of each function (paths relative to the LSSTDESC/imSim tree):
 *   - diffraction kick: imsim/diffraction.py (source in tree; PINNED against the
 *     reference module itself, tests/golden/make_diffraction_golden.py).
 *   - photon_velocity / applyTo sequencing, ray->pixel: imsim/photon_ops.py
 *     (source in tree; ray->pixel PINNED by tests/test_photon_ops.py:668-691).
 *   - TAN-SIP WCS: GalSim >= 2.7.2 galsim/fitswcs.py + coord/celestial.py
 *     (third-party, NOT in the reference tree, NOT importable here):
 *     restated from the published algorithm -- PARITY UNPINNED.
 *   - sequential ray trace: batoid (unpinned version, third-party, absent):
 *     restated from the published algorithm (batoid src/{plane,sphere,
 *     paraboloid,quadric,asphere,surface,sum,bicubic,medium,obscuration}.cpp,
 *     batoid.cpp intersect/reflect/refract) -- PARITY UNPINNED; pinned only
 *     through physics invariants (tests/test_oracle_physics.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/imsim_b200.h"

#define PI 3.14159265358979323846

/* ------------------------------------------------------------------ */
/* media: batoid src/medium.cpp                                        */
/* ------------------------------------------------------------------ */
double orc_medium_n(const B2Medium* m, double wl /* metres */) {
    switch (m->kind) {
        case B2_MED_CONST:
            return m->p[0];
        case B2_MED_SELLMEIER: {
            /* Sellmeier: wavelength in microns */
            double x = wl * 1e6;
            x *= x;
            return sqrt(1.0 + m->p[0] * x / (x - m->p[3]) + m->p[1] * x / (x - m->p[4]) + m->p[2] * x / (x - m->p[5]));
        }
        case B2_MED_SUMITA: {
            double x = wl * 1e6;
            x *= x;
            double y = 1.0 / x;
            return sqrt(m->p[0] + m->p[1] * x + y * (m->p[2] + y * (m->p[3] + y * (m->p[4] + y * m->p[5]))));
        }
        case B2_MED_AIR: {
            /* Filippenko 1982 / Edlen 1953, as batoid.Air and galsim.dcr */
            double P = m->p[0] * 7.50061683; /* kPa -> mmHg */
            double T = m->p[1] - 273.15;     /* K -> C */
            double W = m->p[2] * 7.50061683;
            double sigma_squared = 1e-12 / (wl * wl); /* inverse wavenumber^2, um^-2 */
            double n_minus_one = (64.328 + (29498.1 / (146.0 - sigma_squared)) + (255.4 / (41.0 - sigma_squared))) * 1.e-6;
            n_minus_one *= P * (1.0 + (1.049 - 0.0157 * T) * 1.e-6 * P) / (720.883 * (1.0 + 0.003661 * T));
            n_minus_one -= (0.0624 - 0.000680 * sigma_squared) / (1.0 + 0.003661 * T) * W * 1.e-6;
            return 1.0 + n_minus_one;
        }
    }
    return NAN;
}

/* ------------------------------------------------------------------ */
/* TAN-SIP WCS: galsim/fitswcs.py GSFitsWCS._radec / _xy               */
/* ------------------------------------------------------------------ */
static void sip_eval(const B2TanSip* w, double u, double v, double* u1, double* v1) {
    if (w->order <= 0) {
        *u1 = u;
        *v1 = v;
        return;
    }
    /* horner2d(u, v, ab[k], triangle=True): sum_{i+j<=order} ab[k][i][j] u^i v^j */
    double r[2];
    for (int k = 0; k < 2; ++k) {
        double acc = 0.0;
        for (int i = w->order; i >= 0; --i) {
            double row = 0.0;
            for (int j = w->order - i; j >= 0; --j) row = row * v + w->ab[k][i][j];
            acc = acc * u + row;
        }
        r[k] = acc;
    }
    *u1 = r[0];
    *v1 = r[1];
}

static void sip_eval_jac(const B2TanSip* w, double u, double v, double* f, double* g, double J[4]) {
    /* value and jacobian of the SIP polynomial, used by the Newton inversion
       (GalSim src/WCS.cpp InvertAB) */
    double val[2] = {0, 0}, du[2] = {0, 0}, dv[2] = {0, 0};
    int o = w->order;
    for (int k = 0; k < 2; ++k) {
        for (int i = 0; i <= o; ++i) {
            for (int j = 0; j <= o - i; ++j) {
                double c = w->ab[k][i][j];
                if (c == 0.0) continue;
                double ui = pow(u, i), vj = pow(v, j);
                val[k] += c * ui * vj;
                if (i > 0) du[k] += c * i * pow(u, i - 1) * vj;
                if (j > 0) dv[k] += c * j * ui * pow(v, j - 1);
            }
        }
    }
    *f = val[0];
    *g = val[1];
    J[0] = du[0];
    J[1] = dv[0];
    J[2] = du[1];
    J[3] = dv[1];
}

/* pixel -> (ra, dec) radians */
void orc_tansip_fwd(const B2TanSip* w, double x, double y, double* ra, double* dec) {
    double u = x - w->crpix[0];
    double v = y - w->crpix[1];
    double u1, v1;
    sip_eval(w, u, v, &u1, &v1);
    /* apply CD: intermediate world coordinates in degrees */
    double xi = w->cd[0] * u1 + w->cd[1] * v1;
    double eta = w->cd[2] * u1 + w->cd[3] * v1;
    /* degrees -> radians; FITS +x is east, coord's +u is west */
    double factor = PI / 180.0;
    double uu = -xi * factor;
    double vv = eta * factor;
    /* coord.CelestialCoord.deproject_rad, gnomonic */
    double sindec0 = sin(w->dec0), cosdec0 = cos(w->dec0);
    double rsq = uu * uu + vv * vv;
    double cosc = 1.0 / sqrt(1.0 + rsq);
    double sinc_over_r = cosc;
    double sindec = vv * sinc_over_r * cosdec0 + cosc * sindec0;
    double tandra_num = -uu * sinc_over_r;
    double tandra_denom = cosc * cosdec0 - vv * sinc_over_r * sindec0;
    *dec = asin(sindec);
    *ra = w->ra0 + atan2(tandra_num, tandra_denom);
}

/* (ra, dec) radians -> pixel */
void orc_tansip_inv(const B2TanSip* w, double ra, double dec, double* x, double* y) {
    /* coord.CelestialCoord.project_rad, gnomonic */
    double sindec0 = sin(w->dec0), cosdec0 = cos(w->dec0);
    double sinra0 = sin(w->ra0), cosra0 = cos(w->ra0);
    double cosra = cos(ra), sinra = sin(ra), cosdec = cos(dec), sindec = sin(dec);
    double cosdra = cosra0 * cosra + sinra0 * sinra;
    double sindra = sinra0 * cosra - cosra0 * sinra; /* = -sin(ra - ra0): +u is west */
    double cosc = cosdec * cosdra * cosdec0 + sindec0 * sindec;
    double k = 1.0 / cosc;
    double uu = k * cosdec * sindra;
    double vv = k * (cosdec0 * sindec - sindec0 * cosdec * cosdra);
    /* radians -> degrees, flip u back to FITS convention */
    double factor = 180.0 / PI;
    double xi = -uu * factor;
    double eta = vv * factor;
    /* CD^-1 */
    double det = w->cd[0] * w->cd[3] - w->cd[1] * w->cd[2];
    double u1 = (w->cd[3] * xi - w->cd[1] * eta) / det;
    double v1 = (-w->cd[2] * xi + w->cd[0] * eta) / det;
    double u = u1, v = v1;
    if (w->order > 0) {
        /* Newton-Raphson on the SIP polynomial, start at (u1, v1) */
        for (int iter = 0; iter < 30; ++iter) {
            double f, g, J[4];
            sip_eval_jac(w, u, v, &f, &g, J);
            double df = f - u1, dg = g - v1;
            double d = J[0] * J[3] - J[1] * J[2];
            double du = -(df * J[3] - dg * J[1]) / d;
            double dv = -(-df * J[2] + dg * J[0]) / d;
            u += du;
            v += dv;
            if (fabs(du) <= 1e-15 * (fabs(u) + 1e-300) + 1e-300 && fabs(dv) <= 1e-15 * (fabs(v) + 1e-300) + 1e-300) break;
            if (fabs(du) < 1e-17 * (1.0 + fabs(u)) && fabs(dv) < 1e-17 * (1.0 + fabs(v))) break;
        }
    }
    *x = u + w->crpix[0];
    *y = v + w->crpix[1];
}

/* XyToV.__call__: imsim/photon_ops.py:469-475 */
void orc_xy_to_v(const B2TanSip* img, const B2TanSip* field, int64_t n, const double* x, const double* y, double* vx,
                 double* vy, double* vz) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double ra, dec, thx, thy;
        orc_tansip_fwd(img, x[i], y[i], &ra, &dec);
        orc_tansip_inv(field, ra, dec, &thx, &thy);
        /* batoid.utils.gnomonicToDirCos */
        double gamma = 1.0 / sqrt(1.0 + thx * thx + thy * thy);
        vx[i] = thx * gamma;
        vy[i] = thy * gamma;
        vz[i] = -gamma;
    }
}

/* XyToV.inverse: imsim/photon_ops.py:477-483 */
void orc_v_to_xy(const B2TanSip* img, const B2TanSip* field, int64_t n, const double* vx, const double* vy,
                 const double* vz, double* x, double* y) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        /* batoid.utils.dirCosToGnomonic */
        double thx = -vx[i] / vz[i];
        double thy = -vy[i] / vz[i];
        double ra, dec;
        orc_tansip_fwd(field, thx, thy, &ra, &dec);
        orc_tansip_inv(img, ra, dec, &x[i], &y[i]);
    }
}

/* ------------------------------------------------------------------ */
/* spider diffraction: imsim/diffraction.py                            */
/* ------------------------------------------------------------------ */
/* directed_dist (diffraction.py:192-224) for one point */
static void directed_dist(const B2Diffraction* c, double px, double py, double* dist, double* nx, double* ny) {
    /* dist_thick_line (:227-252): | |n.p - d| - thickness | ; np.argmin = first minimum */
    double min_line = INFINITY;
    int il = 0;
    for (int k = 0; k < c->n_lines; ++k) {
        double d = fabs(fabs(c->lines[k][0] * px + c->lines[k][1] * py - c->lines[k][2]) - c->lines[k][3]);
        if (d < min_line) {
            min_line = d;
            il = k;
        }
    }
    /* dist_circle (:255-276): | |p - c| - r | */
    double min_circ = INFINITY;
    int ic = 0;
    for (int k = 0; k < c->n_circles; ++k) {
        double dx = px - c->circles[k][0], dy = py - c->circles[k][1];
        double d = fabs(sqrt(dx * dx + dy * dy) - c->circles[k][2]);
        if (d < min_circ) {
            min_circ = d;
            ic = k;
        }
    }
    if (min_line < min_circ) { /* line_mask */
        *dist = min_line;
        *nx = c->lines[il][0];
        *ny = c->lines[il][1];
    } else {
        *dist = min_circ;
        double dx = c->circles[ic][0] - px, dy = c->circles[ic][1] - py;
        double nrm = sqrt(dx * dx + dy * dy);
        *nx = dx / nrm;
        *ny = dy / nrm;
    }
}

/* field_rotation_sin_cos (diffraction.py:318-351) */
static void field_rot(const B2Diffraction* c, double t, double* cs, double* sn) {
    double ez[3] = {c->cos_lat * cos(c->omega * t), c->cos_lat * sin(c->omega * t), c->sin_lat};
    const double* ef = c->e_focal;
    const double* e0 = c->e_z_0;
    double eh[3] = {ef[1] * ez[2] - ef[2] * ez[1], ef[2] * ez[0] - ef[0] * ez[2], ef[0] * ez[1] - ef[1] * ez[0]};
    double eh0[3] = {ef[1] * e0[2] - ef[2] * e0[1], ef[2] * e0[0] - ef[0] * e0[2], ef[0] * e0[1] - ef[1] * e0[0]};
    double nrm = sqrt(eh[0] * eh[0] + eh[1] * eh[1] + eh[2] * eh[2]) * sqrt(eh0[0] * eh0[0] + eh0[1] * eh0[1] + eh0[2] * eh0[2]);
    *cs = (eh[0] * eh0[0] + eh[1] * eh0[1] + eh[2] * eh0[2]) / nrm;
    *sn = (ez[0] * eh0[0] + ez[1] * eh0[1] + ez[2] * eh0[2]) / nrm;
}

/* apply_diffraction_delta[_field_rot] (diffraction.py:63-128) on one photon.
   gauss: standard normal draw (the reference's GaussianDeviate.generate_from_variance
   returns gauss*sqrt(phi*^2), photon_ops.py:264-272) */
void orc_diffraction_kick(const B2Diffraction* c, double pu, double pv, double t, double wl_m, double gauss, double v[3]) {
    double cs = 1.0, sn = 0.0;
    double px = pu, py = pv;
    if (c->field_rotation) {
        field_rot(c, t, &cs, &sn);
        /* rot_inv: R^T pos with R = [[c, s], [-s, c]] */
        px = cs * pu - sn * pv;
        py = sn * pu + cs * pv;
    }
    double d, nx, ny;
    directed_dist(c, px, py, &d, &nx, &ny);
    /* phi_star (:182-189) */
    double k = 2.0 * PI / wl_m;
    double phi = atan(1.0 / (2.0 * k * d));
    double d_tan_phi = gauss * sqrt(phi * phi);
    double v_z = -v[2];
    double sx = d_tan_phi * v_z * nx;
    double sy = d_tan_phi * v_z * ny;
    if (c->field_rotation) {
        double rx = cs * sx + sn * sy;
        double ry = -sn * sx + cs * sy;
        sx = rx;
        sy = ry;
    }
    /* apply_delta_v (:45-60) */
    double v_before = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    v[0] += sx;
    v[1] += sy;
    double v_after = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    double f = v_before / v_after;
    v[0] *= f;
    v[1] *= f;
    v[2] *= f;
}

/* array front-end used by the golden test against the reference's diffraction.py */
void orc_diffraction(const B2Diffraction* c, int64_t n, const double* pu, const double* pv, const double* t,
                     const double* wl_m, const double* gauss, double* vx, double* vy, double* vz) {
    for (int64_t i = 0; i < n; ++i) {
        double v[3] = {vx[i], vy[i], vz[i]};
        orc_diffraction_kick(c, pu[i], pv[i], t ? t[i] : 0.0, wl_m[i], gauss[i], v);
        vx[i] = v[0];
        vy[i] = v[1];
        vz[i] = v[2];
    }
}

/* ------------------------------------------------------------------ */
/* surfaces: batoid src/{plane,sphere,paraboloid,quadric,asphere}.cpp   */
/* ------------------------------------------------------------------ */
typedef struct {
    const double* poly;    /* poly2d coefficients or NULL */
    const double* bicubic; /* bicubic block or NULL */
} OrcExtra;

static double base_sag(const B2Surface* s, double x, double y) {
    double r2 = x * x + y * y;
    switch (s->surf_kind) {
        case B2_SURF_PLANE:
            return 0.0;
        case B2_SURF_PARABOLOID:
            return r2 / (2.0 * s->R);
        case B2_SURF_SPHERE:
            return r2 / (s->R * (1.0 + sqrt(1.0 - r2 / s->R / s->R)));
        case B2_SURF_QUADRIC:
        case B2_SURF_ASPHERE: {
            double z = r2 / (s->R * (1.0 + sqrt(1.0 - (1.0 + s->conic) * r2 / s->R / s->R)));
            if (s->surf_kind == B2_SURF_ASPHERE) {
                double rr = r2;
                for (int k = 0; k < s->n_coef; ++k) {
                    rr *= r2;
                    z += s->coef[k] * rr;
                }
            }
            return z;
        }
    }
    return NAN;
}

/* dz/dr divided by r (so that gradient = x*g, y*g); finite at r=0 */
static double base_dzdr_over_r(const B2Surface* s, double x, double y) {
    double r2 = x * x + y * y;
    switch (s->surf_kind) {
        case B2_SURF_PLANE:
            return 0.0;
        case B2_SURF_PARABOLOID:
            return 1.0 / s->R;
        case B2_SURF_SPHERE:
            return 1.0 / (s->R * sqrt(1.0 - r2 / s->R / s->R));
        case B2_SURF_QUADRIC:
        case B2_SURF_ASPHERE: {
            double g = 1.0 / (s->R * sqrt(1.0 - (1.0 + s->conic) * r2 / s->R / s->R));
            if (s->surf_kind == B2_SURF_ASPHERE) {
                double rr = r2; /* r^(2k+2) */
                for (int k = 0; k < s->n_coef; ++k) {
                    g += (4.0 + 2.0 * k) * s->coef[k] * rr;
                    rr *= r2;
                }
            }
            return g;
        }
    }
    return NAN;
}

static void poly2d_eval(const B2Surface* s, const double* c, double x, double y, double* f, double* fx, double* fy) {
    int n = s->poly_n;
    double X = x * s->poly_scale, Y = y * s->poly_scale;
    double val = 0, dx = 0, dy = 0;
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < n; ++j) {
            double cij = c[i * n + j];
            if (cij == 0.0) continue;
            double xi = pow(X, i), yj = pow(Y, j);
            val += cij * xi * yj;
            if (i > 0) dx += cij * i * pow(X, i - 1) * yj;
            if (j > 0) dy += cij * j * xi * pow(Y, j - 1);
        }
    }
    *f = val;
    *fx = dx * s->poly_scale;
    *fy = dy * s->poly_scale;
}

/* cubic Hermite on [0,1]: batoid bicubic.cpp oneDSpline / oneDGrad */
static double h1(double x, double v0, double v1, double d0, double d1) {
    double a = 2 * (v0 - v1) + d0 + d1;
    double b = 3 * (v1 - v0) - 2 * d0 - d1;
    return v0 + x * (d0 + x * (b + x * a));
}
static double h1g(double x, double v0, double v1, double d0, double d1) {
    double a = 2 * (v0 - v1) + d0 + d1;
    double b = 3 * (v1 - v0) - 2 * d0 - d1;
    return d0 + x * (2 * b + x * 3 * a);
}

static void bicubic_eval(const double* blk, double x, double y, double* f, double* fx, double* fy) {
    double x0 = blk[0], dx = blk[1];
    int nx = (int)blk[2];
    double y0 = blk[3], dy = blk[4];
    int ny = (int)blk[5];
    const double* z = blk + 6;
    const double* zx = z + (size_t)nx * ny;
    const double* zy = zx + (size_t)nx * ny;
    const double* zxy = zy + (size_t)nx * ny;
    int ix = (int)floor((x - x0) / dx);
    int iy = (int)floor((y - y0) / dy);
    if (ix < 0 || ix >= nx - 1 || iy < 0 || iy >= ny - 1) {
        *f = *fx = *fy = NAN;
        return;
    }
    double xf = (x - (x0 + ix * dx)) / dx;
    double yf = (y - (y0 + iy * dy)) / dy;
    size_t i00 = (size_t)iy * nx + ix, i01 = i00 + 1, i10 = i00 + nx, i11 = i10 + 1;
    double val0 = h1(xf, z[i00], z[i01], zx[i00] * dx, zx[i01] * dx);
    double val1 = h1(xf, z[i10], z[i11], zx[i10] * dx, zx[i11] * dx);
    double der0 = h1(xf, zy[i00], zy[i01], zxy[i00] * dx, zxy[i01] * dx);
    double der1 = h1(xf, zy[i10], zy[i11], zxy[i10] * dx, zxy[i11] * dx);
    *f = h1(yf, val0, val1, der0 * dy, der1 * dy);
    *fy = h1g(yf, val0, val1, der0 * dy, der1 * dy) / dy;
    double gx0 = h1g(xf, z[i00], z[i01], zx[i00] * dx, zx[i01] * dx);
    double gx1 = h1g(xf, z[i10], z[i11], zx[i10] * dx, zx[i11] * dx);
    double gd0 = h1g(xf, zy[i00], zy[i01], zxy[i00] * dx, zxy[i01] * dx);
    double gd1 = h1g(xf, zy[i10], zy[i11], zxy[i10] * dx, zxy[i11] * dx);
    *fx = h1(yf, gx0, gx1, gd0 * dy, gd1 * dy) / dx;
}

static void surf_sag_grad(const B2Surface* s, const OrcExtra* e, double x, double y, double* z, double* zx, double* zy) {
    double g = base_dzdr_over_r(s, x, y);
    *z = base_sag(s, x, y);
    *zx = x * g;
    *zy = y * g;
    if (s->extra_kind == B2_EXTRA_POLY2D && e && e->poly) {
        double f, fx, fy;
        poly2d_eval(s, e->poly, x, y, &f, &fx, &fy);
        *z += f;
        *zx += fx;
        *zy += fy;
    } else if (s->extra_kind == B2_EXTRA_BICUBIC && e && e->bicubic) {
        double f, fx, fy;
        bicubic_eval(e->bicubic, x, y, &f, &fx, &fy);
        *z += f;
        *zx += fx;
        *zy += fy;
    }
}

/* closed-form time to the base conic r^2 - 2 R z + (1+k) z^2 = 0 (plane if
   R == inf encoded as surf_kind PLANE); root closest to the vertex plane */
static int conic_time(const B2Surface* s, double x, double y, double z, double vx, double vy, double vz, double* dt) {
    if (s->surf_kind == B2_SURF_PLANE) {
        if (vz == 0.0) return 0;
        *dt = -z / vz;
        return 1;
    }
    double k1 = (s->surf_kind == B2_SURF_PARABOLOID) ? 0.0 : (s->surf_kind == B2_SURF_SPHERE ? 1.0 : 1.0 + s->conic);
    double R = s->R;
    double a = vx * vx + vy * vy + k1 * vz * vz;
    double b = 2.0 * (x * vx + y * vy - R * vz + k1 * z * vz);
    double c = x * x + y * y - 2.0 * R * z + k1 * z * z;
    if (a == 0.0) {
        if (b == 0.0) return 0;
        *dt = -c / b;
        return 1;
    }
    double disc = b * b - 4.0 * a * c;
    if (disc < 0.0) return 0;
    double sq = sqrt(disc);
    double q = -0.5 * (b + (b >= 0 ? sq : -sq));
    double t1 = q / a;
    double t2 = (q != 0.0) ? c / q : t1;
    double z1 = fabs(z + vz * t1), z2 = fabs(z + vz * t2);
    *dt = (z1 <= z2) ? t1 : t2;
    return 1;
}

/* batoid Surface::timeToIntersect: fixed 5 Newton steps on the tangent plane */
static int surf_time(const B2Surface* s, const OrcExtra* e, double x, double y, double z, double vx, double vy,
                     double vz, double* dt_out) {
    double dt;
    if (!conic_time(s, x, y, z, vx, vy, vz, &dt)) return 0;
    int needs_newton = (s->surf_kind == B2_SURF_ASPHERE) || (s->extra_kind != B2_EXTRA_NONE);
    if (needs_newton) {
        double rPx = x + vx * dt, rPy = y + vy * dt, rPz = z + vz * dt;
        double sz, zx, zy;
        surf_sag_grad(s, e, rPx, rPy, &sz, &zx, &zy);
        for (int iter = 0; iter < 5; ++iter) {
            /* intersect plane tangent to the surface at (rPx, rPy, sz) with the ray;
               normal direction (-zx, -zy, 1) (normalisation cancels) */
            double nx = -zx, ny = -zy, nz = 1.0;
            dt = (rPx - x) * nx + (rPy - y) * ny + (sz - z) * nz;
            dt /= (nx * vx + ny * vy + nz * vz);
            rPx = x + vx * dt;
            rPy = y + vy * dt;
            rPz = z + vz * dt;
            surf_sag_grad(s, e, rPx, rPy, &sz, &zx, &zy);
        }
        if (!(fabs(sz - rPz) < 1e-14)) return 0;
    }
    *dt_out = dt;
    return 1;
}

/* batoid src/obscuration.cpp */
static int obsc_contains(const B2Obsc* o, double x, double y) {
    int in = 0;
    switch (o->kind) {
        case B2_OBSC_CIRCLE: {
            double dx = x - o->p[1], dy = y - o->p[2];
            in = sqrt(dx * dx + dy * dy) < o->p[0];
            break;
        }
        case B2_OBSC_ANNULUS: {
            double dx = x - o->p[2], dy = y - o->p[3];
            double h = sqrt(dx * dx + dy * dy);
            in = (o->p[0] <= h) && (h < o->p[1]);
            break;
        }
        case B2_OBSC_RECTANGLE: {
            double dx = x - o->p[2], dy = y - o->p[3];
            double xp = dx * o->p[4] + dy * o->p[5];
            double yp = -dx * o->p[5] + dy * o->p[4];
            in = (xp > -o->p[0] / 2 && xp < o->p[0] / 2 && yp > -o->p[1] / 2 && yp < o->p[1] / 2);
            break;
        }
        case B2_OBSC_RAY: {
            double dx = x - o->p[1], dy = y - o->p[2];
            double xp = dx * o->p[3] + dy * o->p[4];
            double yp = -dx * o->p[4] + dy * o->p[3];
            in = (xp > 0.0 && yp > -o->p[0] / 2 && yp < o->p[0] / 2);
            break;
        }
    }
    return o->negate ? !in : in;
}

/* One ray through the whole telescope: batoid CompoundOptic.trace ->
   Interface.trace (coordinate transform, intersect, reflect/refract, obscure).
   In: ray in the stop surface's coordSys; out: ray in the last surface's. */
void orc_trace_one(const B2Telescope* tel, const OrcExtra* extras, double r[3], double v[3], double* t, double wl_m,
                   int* vignetted, int* failed) {
    double nmed[B2_MAX_MEDIA];
    for (int m = 0; m < tel->n_media; ++m) nmed[m] = orc_medium_n(&tel->media[m], wl_m);
    for (int is = 0; is < tel->n_surfaces; ++is) {
        const B2Surface* s = &tel->surf[is];
        const OrcExtra* e = extras ? &extras[is] : NULL;
        /* coordinate transformation (batoid.cpp intersect/reflect/refract prologue) */
        double dx = r[0] - s->dr[0], dy = r[1] - s->dr[1], dz = r[2] - s->dr[2];
        const double* M = s->drot;
        double x = dx * M[0] + dy * M[3] + dz * M[6];
        double y = dx * M[1] + dy * M[4] + dz * M[7];
        double z = dx * M[2] + dy * M[5] + dz * M[8];
        double vx = v[0] * M[0] + v[1] * M[3] + v[2] * M[6];
        double vy = v[0] * M[1] + v[1] * M[4] + v[2] * M[7];
        double vz = v[0] * M[2] + v[1] * M[5] + v[2] * M[8];
        double dt;
        B2Surface base;
        const B2Surface* geom = s;
        if (s->interact == B2_INT_PASS) { /* batoid.OPDScreen: the summed term is the screen, the surface stays bare */
            base = *s;
            base.extra_kind = B2_EXTRA_NONE;
            geom = &base;
        }
        if (!surf_time(geom, e, x, y, z, vx, vy, vz, &dt)) {
            *failed = 1;
            *vignetted = 1;
            r[0] = x, r[1] = y, r[2] = z;
            v[0] = vx, v[1] = vy, v[2] = vz;
            continue;
        }
        x += vx * dt;
        y += vy * dt;
        z += vz * dt;
        *t += dt;
        if (s->interact == B2_INT_PASS) {
            /* thin phase screen on a plane (batoid.OPDScreen with surface=Plane, tests/test_telescope_loader.py:
               641-653): the eikonal gains W(x, y), so the tangential part of the unit direction n v gains grad W
               and the path length gains W; restated from the definition, batoid's refractScreen is not in the tree */
            double W = 0.0, Wx = 0.0, Wy = 0.0;
            if (s->extra_kind == B2_EXTRA_POLY2D && e && e->poly) poly2d_eval(s, e->poly, x, y, &W, &Wx, &Wy);
            else if (s->extra_kind == B2_EXTRA_BICUBIC && e && e->bicubic) bicubic_eval(e->bicubic, x, y, &W, &Wx, &Wy);
            double n1 = nmed[s->medium_in];
            double ux = vx * n1 + Wx, uy = vy * n1 + Wy, uz = vz * n1;
            double w2 = 1.0 - ux * ux - uy * uy;
            if (w2 <= 0.0) {
                *failed = 1;
                *vignetted = 1;
            } else {
                uz = (uz < 0.0 ? -1.0 : 1.0) * sqrt(w2);
                vx = ux / n1;
                vy = uy / n1;
                vz = uz / n1;
                *t += W;
            }
        }
        if (s->interact == B2_INT_MIRROR || s->interact == B2_INT_REFRACT) {
            double sz, zx, zy;
            surf_sag_grad(s, e, x, y, &sz, &zx, &zy);
            double nz = 1.0 / sqrt(1.0 + zx * zx + zy * zy);
            double nx = -zx * nz, ny = -zy * nz;
            if (s->interact == B2_INT_MIRROR) {
                double alpha = vx * nx + vy * ny + vz * nz;
                vx -= 2 * alpha * nx;
                vy -= 2 * alpha * ny;
                vz -= 2 * alpha * nz;
            } else {
                double n1 = nmed[s->medium_in], n2 = nmed[s->medium_out];
                /* unit direction */
                double ux = vx * n1, uy = vy * n1, uz = vz * n1;
                double alpha = ux * nx + uy * ny + uz * nz;
                if (alpha > 0.0) {
                    nx = -nx, ny = -ny, nz = -nz;
                    alpha = -alpha;
                }
                double eta = n1 / n2;
                double sinsqr = eta * eta * (1.0 - alpha * alpha);
                double nfactor = eta * alpha + sqrt(1.0 - sinsqr);
                vx = (eta * ux - nfactor * nx) / n2;
                vy = (eta * uy - nfactor * ny) / n2;
                vz = (eta * uz - nfactor * nz) / n2;
            }
        }
        for (int k = 0; k < s->n_obsc; ++k)
            if (obsc_contains(&s->obsc[k], x, y)) *vignetted = 1;
        r[0] = x, r[1] = y, r[2] = z;
        v[0] = vx, v[1] = vy, v[2] = vz;
    }
}

/* extras: per-surface pointers, arrays of length n_surfaces (entries may be NULL) */
void orc_trace_rays(const B2Telescope* tel, const double* const* poly, const double* const* bicubic, int64_t n,
                    double* x, double* y, double* z, double* vx, double* vy, double* vz, double* t,
                    const double* wl_m, uint8_t* vignetted, uint8_t* failed) {
    OrcExtra ex[B2_MAX_SURFACES];
    for (int i = 0; i < B2_MAX_SURFACES; ++i) {
        ex[i].poly = poly ? poly[i] : NULL;
        ex[i].bicubic = bicubic ? bicubic[i] : NULL;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double r[3] = {x[i], y[i], z[i]}, v[3] = {vx[i], vy[i], vz[i]};
        double tt = t[i];
        int vg = vignetted[i], fl = failed[i];
        orc_trace_one(tel, ex, r, v, &tt, wl_m[i], &vg, &fl);
        x[i] = r[0], y[i] = r[1], z[i] = r[2];
        vx[i] = v[0], vy[i] = v[1], vz[i] = v[2];
        t[i] = tt;
        vignetted[i] = (uint8_t)vg;
        failed[i] = (uint8_t)fl;
    }
}

/* galsim/dcr.py: air_refractive_index_minus_one + get_refraction (wave in nm); third-party,
   restated from the published formula (Filippenko 1982), parity unpinned */
double orc_dcr_refraction(double wave_nm, const double pth[3], double tanz) {
    double P = pth[0] * 7.50061683;
    double T = pth[1] - 273.15;
    double W = pth[2] * 7.50061683;
    double sigma_squared = 1.0 / ((wave_nm * 1.e-3) * (wave_nm * 1.e-3));
    double n_minus_one = (64.328 + (29498.1 / (146.0 - sigma_squared)) + (255.4 / (41.0 - sigma_squared))) * 1.e-6;
    n_minus_one *= P * (1.0 + (1.049 - 0.0157 * T) * 1.e-6 * P) / (720.883 * (1.0 + 0.003661 * T));
    n_minus_one -= (0.0624 - 0.000680 * sigma_squared) / (1.0 + 0.003661 * T) * W * 1.e-6;
    double r0 = n_minus_one * (n_minus_one + 2) / 2.0 / (n_minus_one * n_minus_one + 2 * n_minus_one + 1);
    return r0 * tanz;
}

/* ------------------------------------------------------------------ */
/* RubinOptics.applyTo / RubinDiffractionOptics.applyTo                 */
/* imsim/photon_ops.py:81-127, 136-148, 274-302, 486-503               */
/* ------------------------------------------------------------------ */
void orc_rubin_optics(const B2Telescope* tel, const double* const* poly, const double* const* bicubic,
                      const B2TanSip* img, const B2TanSip* field, const B2Detector* det, const B2Diffraction* dif,
                      const B2OpticsOptions* opt, int64_t n, double* x, double* y, double* dxdz, double* dydz,
                      double* flux, const double* wavelength_nm, const double* pupil_u, const double* pupil_v,
                      const double* time, const double* gauss, double* time_out, B2OpticsStats* stats) {
    OrcExtra ex[B2_MAX_SURFACES];
    for (int i = 0; i < B2_MAX_SURFACES; ++i) {
        ex[i].poly = poly ? poly[i] : NULL;
        ex[i].bicubic = bicubic ? bicubic[i] : NULL;
    }
    uint64_t nvig = 0, nfail = 0, nz = 0;
#pragma omp parallel for schedule(static) reduction(+ : nvig, nfail, nz)
    for (int64_t i = 0; i < n; ++i) {
        double xi = x[i], yi = y[i];
        if (opt->do_dcr) { /* galsim.PhotonDCR.applyTo, the op before the optics in the list */
            if (opt->dcr_alpha != 0.0) {
                double sc = pow(wavelength_nm[i] / opt->dcr_base_wavelength, opt->dcr_alpha);
                xi = sc * (xi - opt->dcr_center[0]) + opt->dcr_center[0];
                yi = sc * (yi - opt->dcr_center[1]) + opt->dcr_center[1];
            }
            double shift = orc_dcr_refraction(wavelength_nm[i], opt->dcr_pth, opt->dcr_tanz) - opt->dcr_base_refraction;
            xi += shift * opt->dcr_m[0];
            yi += shift * opt->dcr_m[1];
        }
        if (opt->shift_in) { /* photon_ops.py:100-102 */
            xi += opt->stamp_center[0];
            yi += opt->stamp_center[1];
        }
        /* photon_velocity (:136-148) */
        double v[3];
        orc_xy_to_v(img, field, 1, &xi, &yi, &v[0], &v[1], &v[2]);
        double wl = wavelength_nm[i] * 1e-9;
        double nair = orc_medium_n(&tel->media[tel->medium_stop], wl);
        v[0] /= nair;
        v[1] /= nair;
        v[2] /= nair;
        if (dif && dif->enabled) /* :294-301 */
            orc_diffraction_kick(dif, pupil_u[i], pupil_v[i], time[i], wl, gauss[i], v);
        /* :106-122: ray on the stop surface (a plane: sag 0), t = 0 */
        double r[3] = {pupil_u[i], pupil_v[i], 0.0};
        double tt = 0.0;
        int vg = 0, fl = 0;
        orc_trace_one(tel, ex, r, v, &tt, wl, &vg, &fl);
        /* ray_vector_to_photon_array (:486-503) */
        if (!vg && !(fabs(r[2]) < 1.0e-15)) nz++;
        double fpx = r[1] * 1e3, fpy = r[0] * 1e3;
        double xo = det->A[0] * fpx + det->A[1] * fpy + det->b[0];
        double yo = det->A[2] * fpx + det->A[3] * fpy + det->b[1];
        double dx = (det->Jhat[0] * v[0] + det->Jhat[1] * v[1]) / v[2];
        double dy = (det->Jhat[2] * v[0] + det->Jhat[3] * v[1]) / v[2];
        if (vg) {
            flux[i] = 0.0;
            nvig++;
        }
        if (fl) nfail++;
        if (opt->shift_out) { /* :125-127 */
            xo -= opt->stamp_center[0];
            yo -= opt->stamp_center[1];
        }
        if (opt->do_focus_depth) { /* galsim.FocusDepth.applyTo */
            xo += dx * opt->focus_depth;
            yo += dy * opt->focus_depth;
        }
        if (opt->do_refraction) { /* galsim.Refraction.applyTo */
            double n2 = opt->index_ratio * opt->index_ratio;
            double normsqr = 1.0 + dx * dx + dy * dy; /* (n3)^-2 with n = (dxdz, dydz, 1)/norm */
            /* galsim: x = dxdz, y = dydz ; factor = 1/sqrt(n^2 + (n^2 - 1)(x^2 + y^2)) */
            double f = 1.0 / sqrt(n2 + (n2 - 1.0) * (normsqr - 1.0));
            dx *= f;
            dy *= f;
            if (isnan(dx) || isnan(dy)) {
                dx = dy = 0.0;
                flux[i] = 0.0;
            }
        }
        x[i] = xo;
        y[i] = yo;
        dxdz[i] = dx;
        dydz[i] = dy;
        if (time_out) time_out[i] = tt;
    }
    if (stats) {
        stats->n_vignetted = nvig;
        stats->n_failed = nfail;
        stats->n_offdetector_z = nz;
    }
}

/* RubinDiffraction.applyTo: imsim/photon_ops.py:304-352 */
void orc_rubin_diffraction(const B2Telescope* tel, const B2TanSip* img, const B2TanSip* field, const B2Diffraction* dif,
                           const B2OpticsOptions* opt, int64_t n, double* x, double* y, const double* wavelength_nm,
                           const double* pupil_u, const double* pupil_v, const double* time, const double* gauss) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < n; ++i) {
        double xi = x[i], yi = y[i];
        if (opt->shift_in) {
            xi += opt->stamp_center[0];
            yi += opt->stamp_center[1];
        }
        double v[3];
        orc_xy_to_v(img, field, 1, &xi, &yi, &v[0], &v[1], &v[2]);
        double wl = wavelength_nm[i] * 1e-9;
        double nair = orc_medium_n(&tel->media[tel->medium_stop], wl);
        v[0] /= nair;
        v[1] /= nair;
        v[2] /= nair;
        orc_diffraction_kick(dif, pupil_u[i], pupil_v[i], time[i], wl, gauss[i], v);
        orc_v_to_xy(img, field, 1, &v[0], &v[1], &v[2], &xi, &yi);
        if (opt->shift_in) { /* symmetric in RubinDiffraction (:350-352) */
            xi -= opt->stamp_center[0];
            yi -= opt->stamp_center[1];
        }
        x[i] = xi;
        y[i] = yi;
    }
}

/* imsim/treerings.py:31-48 TreeRingRadialFunction.__call__ on an array */
void orc_treering_func(double A, double B, int nfreq, const double* cfreqs, const double* cphases, const double* sfreqs,
                       const double* sphases, int64_t n, const double* r, double* out) {
    for (int64_t i = 0; i < n; ++i) {
        double cs = 0.0;
        for (int j = 0; j < nfreq; ++j) cs += sin(2 * PI * (r[i] / cfreqs[j]) + cphases[j]) * cfreqs[j] / (2.0 * PI);
        for (int j = 0; j < nfreq; ++j) cs += -cos(2 * PI * (r[i] / sfreqs[j]) + sphases[j]) * sfreqs[j] / (2.0 * PI);
        cs *= (A + B * r[i] * r[i] * r[i] * r[i]) * .01;
        out[i] = cs;
    }
}
