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

#include "b2_common.cuh"

#define PI_D 3.14159265358979323846

// ------------------------------------------------------------------ fast FP64 reciprocal / rsqrt
// MUFU seed (rcp.approx / rsqrt.approx, ~2^-20) + two Newton steps, branch free.  Measured on
// B200 (tools/rcp_test.cu): b2rcp equals 1.0/x on 4 M random inputs, b2rsqrt is within 2.7e-16.
// The library division / sqrt cost ~27 issue slots each (special-case branches); these cost 5-9.
// Only the optics path uses them (1e-10 tolerance); the sensor path keeps IEEE operations.
__device__ __forceinline__ double b2rcp(double x) {
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
__device__ __forceinline__ double b2rsqrt(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double t = x * y;
    double e = fma(-t, y, 1.0);
    y = fma(0.5 * y, e, y);
    t = x * y;
    e = fma(-t, y, 1.0);
    y = fma(0.5 * y, e, y);
    return y;
}
// sqrt(x) for x >= 0 (x == 0 -> 0); negative x gives NaN like sqrt
__device__ __forceinline__ double b2sqrt(double x) {
    double y = b2rsqrt(x);
    double sq = x * y;
    double r = fma(-sq, sq, x);
    sq = fma(r, 0.5 * y, sq);
    return x == 0.0 ? 0.0 : sq;
}

// ------------------------------------------------------------------ media
__device__ __forceinline__ double medium_n(const B2Medium& m, double wl) {
    switch (m.kind) {
        case B2_MED_CONST:
            return m.p[0];
        case B2_MED_SELLMEIER: {
            double x = wl * 1e6;
            x *= x;
            return b2sqrt(1.0 + m.p[0] * x * b2rcp(x - m.p[3]) + m.p[1] * x * b2rcp(x - m.p[4]) + m.p[2] * x * b2rcp(x - m.p[5]));
        }
        case B2_MED_SUMITA: {
            double x = wl * 1e6;
            x *= x;
            double y = b2rcp(x);
            return b2sqrt(m.p[0] + m.p[1] * x + y * (m.p[2] + y * (m.p[3] + y * (m.p[4] + y * m.p[5]))));
        }
        default: {  // B2_MED_AIR; the pressure / temperature factors are uniform and hoisted by the compiler
            double P = m.p[0] * 7.50061683;
            double T = m.p[1] - 273.15;
            double W = m.p[2] * 7.50061683;
            double s2 = 1e-12 * b2rcp(wl * wl);
            double nm1 = (64.328 + 29498.1 * b2rcp(146.0 - s2) + 255.4 * b2rcp(41.0 - s2)) * 1.e-6;
            nm1 *= P * (1.0 + (1.049 - 0.0157 * T) * 1.e-6 * P) / (720.883 * (1.0 + 0.003661 * T));
            nm1 -= (0.0624 - 0.000680 * s2) / (1.0 + 0.003661 * T) * W * 1.e-6;
            return 1.0 + nm1;
        }
    }
}

// galsim.dcr.air_refractive_index_minus_one / get_refraction (wave in nm), used by PhotonDCR
__device__ __forceinline__ double dcr_refraction(double wave_nm, const double pth[3], double tanz) {
    double P = pth[0] * 7.50061683;
    double T = pth[1] - 273.15;
    double W = pth[2] * 7.50061683;
    double wu = wave_nm * 1.e-3;
    double s2 = b2rcp(wu * wu);
    double nm1 = (64.328 + 29498.1 * b2rcp(146.0 - s2) + 255.4 * b2rcp(41.0 - s2)) * 1.e-6;
    nm1 *= P * (1.0 + (1.049 - 0.0157 * T) * 1.e-6 * P) / (720.883 * (1.0 + 0.003661 * T));
    nm1 -= (0.0624 - 0.000680 * s2) / (1.0 + 0.003661 * T) * W * 1.e-6;
    double r0 = nm1 * (nm1 + 2.0) * 0.5 * b2rcp(nm1 * nm1 + 2.0 * nm1 + 1.0);
    return r0 * tanz;
}

// ------------------------------------------------------------------ TAN-SIP
// packed triangle index for (i,j), i+j<=3, order: 00 01 02 03 10 11 12 20 21 30
//   f(u,v) = sum ab[i][j] u^i v^j
__device__ __forceinline__ void sip_fwd(const DevWcs& w, double u, double v, double& f, double& g) {
    if (w.order <= 0) {
        f = u;
        g = v;
        return;
    }
    const double* a = w.ab[0];
    const double* b = w.ab[1];
    // Horner in v inside Horner in u
    f = ((a[9] * u + (a[7] + a[8] * v)) * u + (a[4] + v * (a[5] + v * a[6]))) * u + (a[0] + v * (a[1] + v * (a[2] + v * a[3])));
    g = ((b[9] * u + (b[7] + b[8] * v)) * u + (b[4] + v * (b[5] + v * b[6]))) * u + (b[0] + v * (b[1] + v * (b[2] + v * b[3])));
}

__device__ __forceinline__ void sip_jac(const double* a, double u, double v, double& f, double& fu, double& fv) {
    double r0 = a[0] + v * (a[1] + v * (a[2] + v * a[3]));
    double r1 = a[4] + v * (a[5] + v * a[6]);
    double r2 = a[7] + a[8] * v;
    double r3 = a[9];
    f = ((r3 * u + r2) * u + r1) * u + r0;
    fu = (3.0 * r3 * u + 2.0 * r2) * u + r1;
    double d0 = a[1] + v * (2.0 * a[2] + 3.0 * v * a[3]);
    double d1 = a[5] + 2.0 * v * a[6];
    double d2 = a[8];
    fv = (d2 * u + d1) * u + d0;
}

// Newton inversion of the SIP polynomial (GalSim src/WCS.cpp InvertAB)
__device__ __forceinline__ void sip_inv(const DevWcs& w, double u1, double v1, double& u, double& v) {
    u = u1;
    v = v1;
    if (w.order <= 0) return;
#pragma unroll 1
    for (int it = 0; it < 12; ++it) {
        double f, fu, fv, g, gu, gv;
        sip_jac(w.ab[0], u, v, f, fu, fv);
        sip_jac(w.ab[1], u, v, g, gu, gv);
        double df = f - u1, dg = g - v1;
        double idet = b2rcp(fu * gv - fv * gu);
        double du = -(df * gv - dg * fv) * idet;
        double dv = -(-df * gu + dg * fu) * idet;
        u += du;
        v += dv;
        // quadratic convergence: once the step is below tol the next error is ~tol^2
        if (fabs(du) < w.newton_tol && fabs(dv) < w.newton_tol) break;
    }
}

// pixel -> tangent-plane (xi, eta) in radians, east/north positive
__device__ __forceinline__ void wcs_pix_to_tan(const DevWcs& w, double x, double y, double& xi, double& eta) {
    double u = x - w.crpix[0], v = y - w.crpix[1];
    double f, g;
    sip_fwd(w, u, v, f, g);
    const double d2r = PI_D / 180.0;
    xi = (w.cd[0] * f + w.cd[1] * g) * d2r;
    eta = (w.cd[2] * f + w.cd[3] * g) * d2r;
}

__device__ __forceinline__ void wcs_tan_to_pix(const DevWcs& w, double xi, double eta, double& x, double& y) {
    const double r2d = 180.0 / PI_D;
    double xd = xi * r2d, ed = eta * r2d;
    double u1 = w.cdinv[0] * xd + w.cdinv[1] * ed;
    double v1 = w.cdinv[2] * xd + w.cdinv[3] * ed;
    double u, v;
    sip_inv(w, u1, v1, u, v);
    x = u + w.crpix[0];
    y = v + w.crpix[1];
}

// XyToV.__call__: the deproject(img centre) o project(field centre) pair of
// galsim/coord is a rotation of the unit sphere, i.e. a homography between the
// two tangent planes: (a,b,c) = M (xi, eta, 1), (xi', eta') = (a/c, b/c).
__device__ __forceinline__ void xy_to_v(const DevOptics& o, double x, double y, double& vx, double& vy, double& vz) {
    double xi, eta;
    wcs_pix_to_tan(o.img, x, y, xi, eta);
    const double* M = o.M_if;
    double a = M[0] * xi + M[1] * eta + M[2];
    double b = M[3] * xi + M[4] * eta + M[5];
    double c = M[6] * xi + M[7] * eta + M[8];
    double ic = b2rcp(c);
    double thx, thy;
    wcs_tan_to_pix(o.field, a * ic, b * ic, thx, thy);
    // batoid.utils.gnomonicToDirCos
    double gamma = b2rsqrt(1.0 + thx * thx + thy * thy);
    vx = thx * gamma;
    vy = thy * gamma;
    vz = -gamma;
}

// XyToV.inverse
__device__ __forceinline__ void v_to_xy(const DevOptics& o, double vx, double vy, double vz, double& x, double& y) {
    double iz = b2rcp(vz);
    double thx = -vx * iz, thy = -vy * iz;
    double xi, eta;
    wcs_pix_to_tan(o.field, thx, thy, xi, eta);
    const double* M = o.M_if;  // inverse rotation = transpose
    double a = M[0] * xi + M[3] * eta + M[6];
    double b = M[1] * xi + M[4] * eta + M[7];
    double c = M[2] * xi + M[5] * eta + M[8];
    double ic = b2rcp(c);
    wcs_tan_to_pix(o.img, a * ic, b * ic, x, y);
}

// ------------------------------------------------------------------ diffraction
// imsim/diffraction.py: directed_dist, phi_star, diffraction_delta[_field_rot], apply_delta_v
__device__ __forceinline__ void sincos_small(double a, double& sn, double& cs) {
    // omega * t stays below 0.05 rad for any exposure shorter than 11 minutes: Taylor to 1e-19
    if (fabs(a) < 0.05) {
        double a2 = a * a;
        sn = a * (1.0 - a2 * (1.0 / 6.0) * (1.0 - a2 * (1.0 / 20.0) * (1.0 - a2 * (1.0 / 42.0) * (1.0 - a2 * (1.0 / 72.0)))));
        cs = 1.0 - a2 * 0.5 * (1.0 - a2 * (1.0 / 12.0) * (1.0 - a2 * (1.0 / 30.0) * (1.0 - a2 * (1.0 / 56.0))));
    } else {
        sincos(a, &sn, &cs);
    }
}

__device__ __forceinline__ double atan_small(double a) {
    // phi* = atan(lambda / (4 pi delta)) is ~1e-6 except for photons grazing an edge
    if (a < 0.01) {
        double a2 = a * a;
        return a * (1.0 - a2 * (1.0 / 3.0 - a2 * (1.0 / 5.0 - a2 * (1.0 / 7.0 - a2 * (1.0 / 9.0)))));
    }
    return atan(a);
}

__device__ __forceinline__ void diffraction_kick(const B2Diffraction& c, double pu, double pv, double t, double wl,
                                                 double gauss, double& vx, double& vy, double& vz) {
    double cs = 1.0, sn = 0.0, px = pu, py = pv;
    if (c.field_rotation) {
        double so, co;
        sincos_small(c.omega * t, so, co);
        double ez0 = c.cos_lat * co, ez1 = c.cos_lat * so, ez2 = c.sin_lat;
        const double* ef = c.e_focal;
        const double* e0 = c.e_z_0;
        double eh0 = ef[1] * ez2 - ef[2] * ez1, eh1 = ef[2] * ez0 - ef[0] * ez2, eh2 = ef[0] * ez1 - ef[1] * ez0;
        double h0 = ef[1] * e0[2] - ef[2] * e0[1], h1 = ef[2] * e0[0] - ef[0] * e0[2], h2 = ef[0] * e0[1] - ef[1] * e0[0];
        double inrm = b2rsqrt((eh0 * eh0 + eh1 * eh1 + eh2 * eh2) * (h0 * h0 + h1 * h1 + h2 * h2));
        cs = (eh0 * h0 + eh1 * h1 + eh2 * h2) * inrm;
        sn = (ez0 * h0 + ez1 * h1 + ez2 * h2) * inrm;
        px = cs * pu - sn * pv;  // R^T pos
        py = sn * pu + cs * pv;
    }
    double min_line = INFINITY, lnx = 0.0, lny = 0.0;
    for (int k = 0; k < c.n_lines; ++k) {
        double d = fabs(fabs(c.lines[k][0] * px + c.lines[k][1] * py - c.lines[k][2]) - c.lines[k][3]);
        if (d < min_line) {
            min_line = d;
            lnx = c.lines[k][0];
            lny = c.lines[k][1];
        }
    }
    double min_circ = INFINITY, cdx = 0.0, cdy = 0.0, icnrm = 1.0;
    for (int k = 0; k < c.n_circles; ++k) {
        double dx = px - c.circles[k][0], dy = py - c.circles[k][1];
        double r2 = dx * dx + dy * dy;
        double inr = b2rsqrt(r2);
        double d = fabs(r2 * inr - c.circles[k][2]);
        if (d < min_circ) {
            min_circ = d;
            cdx = -dx;
            cdy = -dy;
            icnrm = inr;
        }
    }
    double dist, nx, ny;
    if (min_line < min_circ) {
        dist = min_line;
        nx = lnx;
        ny = lny;
    } else {
        dist = min_circ;
        nx = cdx * icnrm;
        ny = cdy * icnrm;
    }
    // phi* = atan(1 / (2 k dist)), k = 2 pi / lambda
    double phi = atan_small(wl * b2rcp(4.0 * PI_D * dist));
    double d_tan_phi = gauss * fabs(phi);
    double v_z = -vz;
    double sx = d_tan_phi * v_z * nx, sy = d_tan_phi * v_z * ny;
    if (c.field_rotation) {
        double rx = cs * sx + sn * sy, ry = -sn * sx + cs * sy;
        sx = rx;
        sy = ry;
    }
    double vb2 = vx * vx + vy * vy + vz * vz;
    vx += sx;
    vy += sy;
    double f = b2sqrt(vb2) * b2rsqrt(vx * vx + vy * vy + vz * vz);
    vx *= f;
    vy *= f;
    vz *= f;
}

// ------------------------------------------------------------------ surfaces
// extra (summed) sag terms
__device__ __forceinline__ void poly2d_eval(const DevSurf& s, double x, double y, double& f, double& fx, double& fy) {
    const int n = s.poly_n;
    const double* c = s.extra;
    double X = x * s.poly_scale, Y = y * s.poly_scale;
    // f = sum_i X^i * row_i(Y); Horner in X over rows, rows Horner in Y
    double val = 0.0, dX = 0.0, dY = 0.0;
    for (int i = n - 1; i >= 0; --i) {
        double row = 0.0, drow = 0.0;
        for (int j = n - 1; j >= 0; --j) {
            drow = drow * Y + row;
            row = row * Y + __ldg(&c[i * n + j]);
        }
        dX = dX * X + val;
        val = val * X + row;
        dY = dY * X + drow;
    }
    f = val;
    fx = dX * s.poly_scale;
    fy = dY * s.poly_scale;
}

__device__ __forceinline__ double h1(double x, double v0, double v1, double d0, double d1) {
    double a = 2 * (v0 - v1) + d0 + d1;
    double b = 3 * (v1 - v0) - 2 * d0 - d1;
    return v0 + x * (d0 + x * (b + x * a));
}
__device__ __forceinline__ double h1g(double x, double v0, double v1, double d0, double d1) {
    double a = 2 * (v0 - v1) + d0 + d1;
    double b = 3 * (v1 - v0) - 2 * d0 - d1;
    return d0 + x * (2 * b + x * 3 * a);
}

__device__ __forceinline__ void bicubic_eval(const double* blk, double x, double y, double& f, double& fx, double& fy) {
    double x0 = __ldg(blk + 0), dx = __ldg(blk + 1);
    int nx = (int)__ldg(blk + 2);
    double y0 = __ldg(blk + 3), dy = __ldg(blk + 4);
    int ny = (int)__ldg(blk + 5);
    const double* z = blk + 6;
    size_t npts = (size_t)nx * ny;
    const double* zx = z + npts;
    const double* zy = zx + npts;
    const double* zxy = zy + npts;
    int ix = (int)floor((x - x0) / dx);
    int iy = (int)floor((y - y0) / dy);
    if (ix < 0 || ix >= nx - 1 || iy < 0 || iy >= ny - 1) {
        f = fx = fy = nan("");
        return;
    }
    double xf = (x - (x0 + ix * dx)) / dx;
    double yf = (y - (y0 + iy * dy)) / dy;
    size_t i00 = (size_t)iy * nx + ix, i01 = i00 + 1, i10 = i00 + nx, i11 = i10 + 1;
    double z00 = __ldg(z + i00), z01 = __ldg(z + i01), z10 = __ldg(z + i10), z11 = __ldg(z + i11);
    double a00 = __ldg(zx + i00) * dx, a01 = __ldg(zx + i01) * dx, a10 = __ldg(zx + i10) * dx, a11 = __ldg(zx + i11) * dx;
    double b00 = __ldg(zy + i00), b01 = __ldg(zy + i01), b10 = __ldg(zy + i10), b11 = __ldg(zy + i11);
    double c00 = __ldg(zxy + i00) * dx, c01 = __ldg(zxy + i01) * dx, c10 = __ldg(zxy + i10) * dx, c11 = __ldg(zxy + i11) * dx;
    double val0 = h1(xf, z00, z01, a00, a01);
    double val1 = h1(xf, z10, z11, a10, a11);
    double der0 = h1(xf, b00, b01, c00, c01);
    double der1 = h1(xf, b10, b11, c10, c11);
    f = h1(yf, val0, val1, der0 * dy, der1 * dy);
    fy = h1g(yf, val0, val1, der0 * dy, der1 * dy) / dy;
    double gx0 = h1g(xf, z00, z01, a00, a01);
    double gx1 = h1g(xf, z10, z11, a10, a11);
    double gd0 = h1g(xf, b00, b01, c00, c01);
    double gd1 = h1g(xf, b10, b11, c10, c11);
    fx = h1(yf, gx0, gx1, gd0 * dy, gd1 * dy) / dx;
}

// even-asphere polynomial P(r^2) = sum coef[k] r^(4+2k) with dP/d(r^2) and d2P/d(r^2)^2, plus the
// summed extra term E(x, y) with its gradient: everything on the surface that is not the base conic
__device__ __forceinline__ void departure(const DevSurf& s, double x, double y, double r2, double& P, double& dP,
                                          double& ddP, double& E, double& Ex, double& Ey) {
    P = dP = ddP = 0.0;
    if (s.kind == B2_SURF_ASPHERE) {
        // P = r2^2 h(r2); Horner for h, h', h'' from the highest coefficient
        double h = 0.0, dh = 0.0, ddh = 0.0;
        if (s.n_coef <= 4) {  // the usual case, fully unrolled (unused coefficients are zero)
#pragma unroll
            for (int k = 3; k >= 0; --k) {
                ddh = ddh * r2 + 2.0 * dh;
                dh = dh * r2 + h;
                h = h * r2 + s.coef[k];
            }
        } else {
            for (int k = s.n_coef - 1; k >= 0; --k) {
                ddh = ddh * r2 + 2.0 * dh;
                dh = dh * r2 + h;
                h = h * r2 + s.coef[k];
            }
        }
        P = r2 * r2 * h;
        dP = r2 * (2.0 * h + r2 * dh);
        ddP = 2.0 * h + r2 * (4.0 * dh + r2 * ddh);
    }
    E = Ex = Ey = 0.0;
    if (s.extra_kind == B2_EXTRA_POLY2D) poly2d_eval(s, x, y, E, Ex, Ey);
    else if (s.extra_kind == B2_EXTRA_BICUBIC) bicubic_eval(s.extra, x, y, E, Ex, Ey);
}

__device__ __forceinline__ bool obscured(const DevObsc& o, double x, double y) {
    bool in;
    switch (o.kind) {
        case B2_OBSC_CIRCLE: {
            double dx = x - o.p[1], dy = y - o.p[2];
            in = (dx * dx + dy * dy) < o.p[0];  // p0 = radius^2
            break;
        }
        case B2_OBSC_ANNULUS: {
            double dx = x - o.p[2], dy = y - o.p[3];
            double h2 = dx * dx + dy * dy;
            in = (o.p[0] <= h2) && (h2 < o.p[1]);  // squared radii
            break;
        }
        case B2_OBSC_RECTANGLE: {
            double dx = x - o.p[2], dy = y - o.p[3];
            double xp = dx * o.p[4] + dy * o.p[5];
            double yp = -dx * o.p[5] + dy * o.p[4];
            in = (xp > -o.p[0] && xp < o.p[0] && yp > -o.p[1] && yp < o.p[1]);  // half sizes
            break;
        }
        default: {  // ray
            double dx = x - o.p[1], dy = y - o.p[2];
            double xp = dx * o.p[3] + dy * o.p[4];
            double yp = -dx * o.p[4] + dy * o.p[3];
            in = (xp > 0.0 && yp > -o.p[0] && yp < o.p[0]);  // half width
            break;
        }
    }
    return o.negate ? !in : in;
}

struct Ray {
    double x, y, z, vx, vy, vz, t;
    bool vignetted, failed;
};

// batoid CompoundOptic.trace: sequential interfaces
__device__ __forceinline__ void trace_ray(const DevOptics& o, Ray& r, double wl) {
    // refractive indices and their inverses, once per photon per medium
    double n0 = medium_n(o.media[0], wl);
    double n1 = o.n_media > 1 ? medium_n(o.media[1], wl) : 1.0;
    double n2 = o.n_media > 2 ? medium_n(o.media[2], wl) : 1.0;
    double n3 = o.n_media > 3 ? medium_n(o.media[3], wl) : 1.0;
    double i0 = b2rcp(n0), i1 = b2rcp(n1), i2 = b2rcp(n2), i3 = b2rcp(n3);
#pragma unroll 1
    for (int is = 0; is < o.n_surf; ++is) {
        const DevSurf& s = o.surf[is];
        // coordinate transformation: r' = drot^T (r - dr)
        double dx = r.x - s.dr[0], dy = r.y - s.dr[1], dz = r.z - s.dr[2];
        double x, y, z, vx, vy, vz;
        if (s.rot_identity) {
            x = dx; y = dy; z = dz;
            vx = r.vx; vy = r.vy; vz = r.vz;
        } else {
            const double* M = s.drot;
            x = dx * M[0] + dy * M[3] + dz * M[6];
            y = dx * M[1] + dy * M[4] + dz * M[7];
            z = dx * M[2] + dy * M[5] + dz * M[8];
            vx = r.vx * M[0] + r.vy * M[3] + r.vz * M[6];
            vy = r.vx * M[1] + r.vy * M[4] + r.vz * M[7];
            vz = r.vx * M[2] + r.vy * M[5] + r.vz * M[8];
        }
        // intersection: go to the vertex plane first, then the near-vertex (small) root of the
        // base conic x^2 + y^2 - 2 R z + k1 z^2 = 0
        bool ok = (vz != 0.0);
        double dt = -z * b2rcp(vz);
        double px = x + vx * dt, py = y + vy * dt, pz = 0.0;
        const bool curved = (s.kind != B2_SURF_PLANE);
        if (curved) {
            double A = vx * vx + vy * vy + s.k1 * vz * vz;
            double B = 2.0 * (px * vx + py * vy - s.R * vz);
            double C = px * px + py * py;
            double disc = B * B - 4.0 * A * C;
            ok = ok && (disc >= 0.0);
            double q = -0.5 * (B + copysign(b2sqrt(disc), B));
            double t1 = C * b2rcp(q);
            dt += t1;
            px += vx * t1;
            py += vy * t1;
            pz = vz * t1;
        }
        // zc: height of the base conic under the hit point (= pz unless the surface departs from it)
        double zc = pz, gP = 0.0, Ex = 0.0, Ey = 0.0;
        if (s.kind == B2_SURF_ASPHERE || s.extra_kind != B2_EXTRA_NONE) {
            // Newton on the implicit form G(t) = r^2 - 2 R zc + k1 zc^2 with zc = z - P(r^2) - E(x, y):
            // polynomial in the ray parameter, no square root; quadratic convergence from the conic hit
            bool conv = false;
            const bool pure = (s.extra_kind == B2_EXTRA_NONE);
#pragma unroll 1
            for (int it = 0; it < 8; ++it) {
                double r2 = px * px + py * py;
                double P, dP, ddP, E;
                departure(s, px, py, r2, P, dP, ddP, E, Ex, Ey);
                zc = pz - P - E;
                double rv = px * vx + py * vy;
                double dzc = vz - 2.0 * dP * rv - (Ex * vx + Ey * vy);
                double G, dG;
                if (curved) {
                    G = r2 - 2.0 * s.R * zc + s.k1 * zc * zc;
                    dG = 2.0 * rv - 2.0 * (s.R - s.k1 * zc) * dzc;
                } else {
                    G = zc;
                    dG = dzc;
                }
                double step = -G * b2rcp(dG);
                dt += step;
                px += vx * step;
                py += vy * step;
                pz += vz * step;
                zc += dzc * step;
                gP = dP;
                // The departure gradient was evaluated one step back; the normal needs it at the hit
                // point to ~3e-14 rad (1e-10 px at the focal plane ~ 4e-13 rad).  Pure aspheres:
                // refresh dP to first order with d2P once the step is small (second evaluation from
                // the conic seed); summed Zernike / bicubic terms: iterate until the step is < 1e-13 m.
                if (pure && fabs(step) < 1e-6) {
                    gP = dP + ddP * (2.0 * rv * step);
                    conv = true;
                    break;
                }
                if (fabs(step) < 1e-13) {
                    conv = true;
                    break;
                }
            }
            ok = ok && conv;
        }
        if (!ok) {
            r.failed = true;
            r.vignetted = true;
            r.x = x; r.y = y; r.z = z;
            r.vx = vx; r.vy = vy; r.vz = vz;
            continue;
        }
        r.t += dt;
        if (s.interact == B2_INT_MIRROR || s.interact == B2_INT_REFRACT) {
            // surface gradient: conic part from grad F = (x, y, k1 z - R) (no square root), plus departure
            double zx = 2.0 * gP * px + Ex, zy = 2.0 * gP * py + Ey;
            if (curved) {
                double ig = b2rcp(s.R - s.k1 * zc);
                zx += px * ig;
                zy += py * ig;
            }
            double NN = 1.0 + zx * zx + zy * zy;
            double iNN = b2rcp(NN);
            double vn = -zx * vx - zy * vy + vz;  // v.N with N = (-zx, -zy, 1) unnormalised
            if (s.interact == B2_INT_MIRROR) {
                double f = 2.0 * vn * iNN;
                vx += f * zx;
                vy += f * zy;
                vz -= f;
            } else {
                int mi = s.med_in, mo = s.med_out;
                double na = mi == 0 ? n0 : (mi == 1 ? n1 : (mi == 2 ? n2 : n3));
                double nb = mo == 0 ? n0 : (mo == 1 ? n1 : (mo == 2 ? n2 : n3));
                double inb = mo == 0 ? i0 : (mo == 1 ? i1 : (mo == 2 ? i2 : i3));
                // u = na v is the unit direction; orient N against u
                double uN = na * vn;
                double sgn = uN > 0.0 ? -1.0 : 1.0;
                uN *= sgn;
                double eta = na * inb;
                // v' = (eta u - [eta uN + sqrt((1-eta^2) NN + eta^2 uN^2)]/NN N) / nb
                double fac = (eta * uN + b2sqrt((1.0 - eta * eta) * NN + eta * eta * uN * uN)) * iNN * sgn;
                double e2 = eta * na;
                vx = (e2 * vx + fac * zx) * inb;
                vy = (e2 * vy + fac * zy) * inb;
                vz = (e2 * vz - fac) * inb;
            }
        }
        if (s.simple_clear) {
            double r2 = px * px + py * py;
            if (!(s.clr_in2 <= r2 && r2 < s.clr_out2)) r.vignetted = true;
        } else {
            for (int k = 0; k < s.n_obsc; ++k)
                if (obscured(s.obsc[k], px, py)) r.vignetted = true;
        }
        r.x = px; r.y = py; r.z = pz;
        r.vx = vx; r.vy = vy; r.vz = vz;
    }
}

// standard normal from Philox (Box-Muller).  The transcendental part runs in FP32: a deviate with
// 1e-7 relative granularity is statistically indistinguishable, and parity tests inject the draws.
__device__ __forceinline__ double philox_normal(uint64_t seed, uint64_t idx, uint32_t stream) {
    uint32_t r[4];
    philox4(seed, idx, stream, r);
    float u1 = ((float)(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float u2 = ((float)(r[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    return (double)(sqrtf(-2.0f * logf(u1)) * cs);
}

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

__global__ void __launch_bounds__(256)
k_trace_rays(const __grid_constant__ DevOptics o, int64_t n, double* x, double* y, double* z, double* vx, double* vy,
             double* vz, double* t, const double* __restrict__ wl, uint8_t* vig, uint8_t* fail) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Ray r{x[i], y[i], z[i], vx[i], vy[i], vz[i], t[i], vig[i] != 0, fail[i] != 0};
    trace_ray(o, r, wl[i]);
    x[i] = r.x; y[i] = r.y; z[i] = r.z;
    vx[i] = r.vx; vy[i] = r.vy; vz[i] = r.vz;
    t[i] = r.t;
    vig[i] = r.vignetted;
    fail[i] = r.failed;
}

// RubinOptics / RubinDiffractionOptics.applyTo fused with FocusDepth + Refraction
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
        double xi = x[i], yi = y[i];
        double wl = wl_nm[i] * 1e-9;
        double u = pu[i], v = pv[i];
        if (opt.do_dcr) {  // galsim.PhotonDCR.applyTo
            if (opt.dcr_alpha != 0.0) {
                double sc = pow(wl_nm[i] / opt.dcr_base_wavelength, opt.dcr_alpha);
                xi = sc * (xi - opt.dcr_center[0]) + opt.dcr_center[0];
                yi = sc * (yi - opt.dcr_center[1]) + opt.dcr_center[1];
            }
            double shift = dcr_refraction(wl_nm[i], opt.dcr_pth, opt.dcr_tanz) - opt.dcr_base_refraction;
            xi += shift * opt.dcr_m[0];
            yi += shift * opt.dcr_m[1];
        }
        if (opt.shift_in) {
            xi += opt.stamp_center[0];
            yi += opt.stamp_center[1];
        }
        Ray r;
        xy_to_v(o, xi, yi, r.vx, r.vy, r.vz);
        double inair = b2rcp(medium_n(o.media[o.medium_stop], wl));
        r.vx *= inair;
        r.vy *= inair;
        r.vz *= inair;
        if (o.dif.enabled) {
            double g = gauss ? gauss[i] : philox_normal(opt.seed, opt.photon_offset + (uint64_t)i, 0u);
            diffraction_kick(o.dif, u, v, time[i], wl, g, r.vx, r.vy, r.vz);
        }
        r.x = u;
        r.y = v;
        r.z = 0.0;
        r.t = 0.0;
        r.vignetted = false;
        r.failed = false;
        trace_ray(o, r, wl);
        vig = r.vignetted;
        fail = r.failed;
        offz = !vig && !(fabs(r.z) < 1.0e-15);
        // ray_vector_to_photon_array
        double fpx = r.y * 1e3, fpy = r.x * 1e3;
        double xo = o.det.A[0] * fpx + o.det.A[1] * fpy + o.det.b[0];
        double yo = o.det.A[2] * fpx + o.det.A[3] * fpy + o.det.b[1];
        double iz = b2rcp(r.vz);
        double dx = (o.det.Jhat[0] * r.vx + o.det.Jhat[1] * r.vy) * iz;
        double dy = (o.det.Jhat[2] * r.vx + o.det.Jhat[3] * r.vy) * iz;
        double fl = flux[i];
        if (vig) fl = 0.0;
        if (opt.shift_out) {
            xo -= opt.stamp_center[0];
            yo -= opt.stamp_center[1];
        }
        if (opt.do_focus_depth) {
            xo += dx * opt.focus_depth;
            yo += dy * opt.focus_depth;
        }
        if (opt.do_refraction) {
            double n2 = opt.index_ratio * opt.index_ratio;
            double f = b2rsqrt(n2 + (n2 - 1.0) * (dx * dx + dy * dy));
            dx *= f;
            dy *= f;
            if (isnan(dx) || isnan(dy)) {
                dx = dy = 0.0;
                fl = 0.0;
            }
        }
        x[i] = xo;
        y[i] = yo;
        dxdz[i] = dx;
        dydz[i] = dy;
        flux[i] = fl;
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
#define B2_OPTICS_CALL rubin_optics_body(o, opt, n, x, y, dxdz, dydz, flux, wl_nm, pu, pv, time, gauss, time_out, stats)
// The trace is latency bound on dependent FP64 chains (ncu: stall "wait" dominates at 4 warps per
// scheduler), so resident warps matter more than a few spilled registers: three builds of the same
// body at 2 / 3 / 4 blocks per SM; B2_OPTICS_OCC picks one (default set from measurements).
__global__ void __launch_bounds__(256, 2) k_rubin_optics(B2_OPTICS_ARGS) { B2_OPTICS_CALL; }
__global__ void __launch_bounds__(256, 3) k_rubin_optics_occ3(B2_OPTICS_ARGS) { B2_OPTICS_CALL; }
__global__ void __launch_bounds__(256, 4) k_rubin_optics_occ4(B2_OPTICS_ARGS) { B2_OPTICS_CALL; }

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
    uint32_t r[4], q[4];
    philox4(seed, offset + (uint64_t)i, 1u, r);
    philox4(seed, offset + (uint64_t)i, 2u, q);
    double ut = u01(r[0], r[1]), ur = u01(r[2], r[3]), uphi = u01(q[0], q[1]);
    if (time) time[i] = t0 + exptime * ut;
    if (pu) {
        double rr = sqrt(r_in * r_in + (r_out * r_out - r_in * r_in) * ur);
        double s, c;
        sincospi(2.0 * uphi, &s, &c);
        pu[i] = rr * c;
        pv[i] = rr * s;
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

extern "C" int b2_telescope_upload(b2_ctx* ctx, const B2Telescope* tel) {
    B2_REQUIRE(ctx && tel, "b2_telescope_upload: null argument");
    B2_REQUIRE(tel->n_surfaces >= 1 && tel->n_surfaces <= B2_DEV_MAX_SURF, "b2_telescope_upload: 1..16 surfaces supported");
    B2_REQUIRE(tel->n_media >= 1 && tel->n_media <= B2_DEV_MAX_MEDIA, "b2_telescope_upload: 1..4 media supported");
    DevOptics& o = ctx->opt;
    o.n_surf = tel->n_surfaces;
    o.n_media = tel->n_media;
    o.medium_stop = tel->medium_stop;
    for (int m = 0; m < tel->n_media; ++m) o.media[m] = tel->media[m];
    for (int i = 0; i < tel->n_surfaces; ++i) {
        const B2Surface& s = tel->surf[i];
        DevSurf& d = o.surf[i];
        memset(&d, 0, sizeof(d));
        d.kind = s.surf_kind;
        d.interact = s.interact;
        B2_REQUIRE(s.interact != B2_INT_PASS, "b2_telescope_upload: OPDScreen-like interfaces are not supported yet");
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
    void* p = nullptr;
    B2_CUDA(cudaMalloc(&p, n * sizeof(double)));
    ctx->extras.push_back(p);
    B2_CUDA(cudaMemcpyAsync(p, data, n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    d.extra = (const double*)p;
    d.extra_kind = kind;
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
        k_trace_rays<<<nblocks(n), 256, 0, ctx->stream>>>(ctx->opt, n, d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], dv, df);
        B2_CHECK_LAUNCH();
        for (int k = 0; k < 7; ++k) D2H(h[k], d[k], n);
        B2_CUDA(cudaMemcpyAsync(vig, dv, n, cudaMemcpyDeviceToHost, ctx->stream));
        B2_CUDA(cudaMemcpyAsync(fail, df, n, cudaMemcpyDeviceToHost, ctx->stream));
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
    } else {
        k_trace_rays<<<nblocks(n), 256, 0, ctx->stream>>>(ctx->opt, n, x, y, z, vx, vy, vz, t, wl, vig, fail);
        B2_CHECK_LAUNCH();
    }
    return 0;
}

static int optics_occ() {
    static int occ = -1;
    if (occ < 0) {
        const char* e = getenv("B2_OPTICS_OCC");
        occ = e ? atoi(e) : 3;
        if (occ < 2 || occ > 4) occ = 3;
    }
    return occ;
}

static void launch_rubin_optics(b2_ctx* ctx, const B2OpticsOptions& opt, int64_t n, double* x, double* y, double* dxdz,
                                double* dydz, double* flux, const double* wl, const double* pu, const double* pv,
                                const double* time, const double* gauss, double* time_out, unsigned long long* stats) {
    B2_TIMED("k_rubin_optics", ctx->stream);
    switch (optics_occ()) {
        case 2:
            k_rubin_optics<<<nblocks(n), 256, 0, ctx->stream>>>(ctx->opt, opt, n, x, y, dxdz, dydz, flux, wl, pu, pv, time, gauss, time_out, stats);
            break;
        case 4:
            k_rubin_optics_occ4<<<nblocks(n), 256, 0, ctx->stream>>>(ctx->opt, opt, n, x, y, dxdz, dydz, flux, wl, pu, pv, time, gauss, time_out, stats);
            break;
        default:
            k_rubin_optics_occ3<<<nblocks(n), 256, 0, ctx->stream>>>(ctx->opt, opt, n, x, y, dxdz, dydz, flux, wl, pu, pv, time, gauss, time_out, stats);
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
        H2D(dx, x, n);
        H2D(dy, y, n);
        H2D(df, flux, n);
        H2D(dw, wl_nm, n);
        H2D(du, pu, n);
        H2D(dv, pv, n);
        H2D(dt, time, n);
        if (gauss) H2D(dg, gauss, n);
        launch_rubin_optics(ctx, *opt, n, dx, dy, da, db, df, dw, du, dv, dt, gauss ? dg : nullptr,
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
        launch_rubin_optics(ctx, *opt, n, x, y, dxdz, dydz, flux, wl_nm, pu, pv, time, gauss, time_out, dstats);
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
