// optics_device.cuh -- device functions of the photon ray trace (shared by optics.cu and pool.cu)
#pragma once
#include "b2_common.cuh"

#define PI_D 3.14159265358979323846

// ------------------------------------------------------------------ fast FP64 reciprocal / rsqrt
// MUFU seed (rcp.approx / rsqrt.approx, 2^-20) + ONE cubically convergent correction, branch free.
// Measured on B200 (tools/rcp_test.cu, 4 M random inputs): b2rcp within 2.3e-16 of 1.0/x, b2rsqrt
// within 2.8e-16, b2sqrt equal to sqrt().  The library division / sqrt cost ~27 issue slots each
// (special-case branches); these cost 4-8.  Only the optics path uses them (1e-10 tolerance); the
// sensor path keeps IEEE operations.
__host__ __device__ __forceinline__ double b2rcp(double x) {
#ifndef __CUDA_ARCH__
    return 1.0 / x;  // host copy: set-up code only (the per-detector XyToV fit)
#else
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);  // r (1 + e + e^2) = 1/x (1 - e^3)
    double t = fma(e, e, e);
    return fma(r, t, r);
#endif
}
__host__ __device__ __forceinline__ double b2rsqrt(double x) {
#ifndef __CUDA_ARCH__
    return 1.0 / sqrt(x);
#else
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double t = x * y;
    double e = fma(-t, y, 1.0);  // y (1 + e/2 + 3 e^2 / 8) = x^-1/2 (1 + O(e^3))
    double p = fma(0.375, e, 0.5);
    p *= e;
    return fma(y, p, y);
#endif
}
// sqrt(x) for x >= 0 (x == 0 -> 0); negative x gives NaN like sqrt
__device__ __forceinline__ double b2sqrt(double x) {
    double y = b2rsqrt(x);
    double sq = x * y;
    double r = fma(-sq, sq, x);
    sq = fma(r, 0.5 * y, sq);
    return x == 0.0 ? 0.0 : sq;
}

// sqrt(x) = x * rsqrt(x) without the final correction step of b2sqrt: relative error <= ~4e-16 (rsqrt's
// 2.8e-16 plus one rounding), far inside the 1e-10 parity bar of the trace; x == 0 -> 0.
__device__ __forceinline__ double b2sqrt_fast(double x) {
    double sq = x * b2rsqrt(x);
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
        default: {  // B2_MED_AIR; the pressure / temperature factors do not depend on the photon: p[3], p[4] hold
                    // P (1 + (1.049 - 0.0157 T) 1e-6 P) / (720.883 (1 + 0.003661 T)) and W 1e-6 / (1 + 0.003661 T),
                    // filled in by b2_telescope_upload
            double s2 = 1e-12 * b2rcp(wl * wl);
            double nm1 = (64.328 + 29498.1 * b2rcp(146.0 - s2) + 255.4 * b2rcp(41.0 - s2)) * 1.e-6;
            nm1 *= m.p[3];
            nm1 -= (0.0624 - 0.000680 * s2) * m.p[4];
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
__host__ __device__ __forceinline__ void sip_fwd(const DevWcs& w, double u, double v, double& f, double& g) {
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

__host__ __device__ __forceinline__ void sip_jac(const double* a, double u, double v, double& f, double& fu, double& fv) {
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
__host__ __device__ __forceinline__ void sip_inv(const DevWcs& w, double u1, double v1, double& u, double& v) {
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
__host__ __device__ __forceinline__ void wcs_pix_to_tan(const DevWcs& w, double x, double y, double& xi, double& eta) {
    double u = x - w.crpix[0], v = y - w.crpix[1];
    double f, g;
    sip_fwd(w, u, v, f, g);
    const double d2r = PI_D / 180.0;
    xi = (w.cd[0] * f + w.cd[1] * g) * d2r;
    eta = (w.cd[2] * f + w.cd[3] * g) * d2r;
}

__host__ __device__ __forceinline__ void wcs_tan_to_pix(const DevWcs& w, double xi, double eta, double& x, double& y) {
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
__host__ __device__ __forceinline__ void xy_to_field_exact(const DevOptics& o, double x, double y, double& thx,
                                                           double& thy) {
    double xi, eta;
    wcs_pix_to_tan(o.img, x, y, xi, eta);
    const double* M = o.M_if;
    double a = M[0] * xi + M[1] * eta + M[2];
    double b = M[3] * xi + M[4] * eta + M[5];
    double c = M[6] * xi + M[7] * eta + M[8];
    double ic = b2rcp(c);
    wcs_tan_to_pix(o.field, a * ic, b * ic, thx, thy);
}

// The same map compiled per detector (b2_xytov_compile): over one CCD the field tangents are a smooth,
// almost affine function of the pixel position, and a tensor polynomial of degree 5 x 5 in the scaled
// coordinates reproduces the exact chain (SIP forward, homography, Newton inversion of the field SIP) to
// a few 1e-10 px -- 1e-13 of the coordinate, three orders inside the 1e-10 parity bar -- for 70 fused
// multiply-adds.  Positions outside the fitted box take the exact chain.
__device__ __forceinline__ void xy_to_field(const DevOptics& o, double x, double y, double& thx, double& thy) {
    const DevXyPoly& p = o.xyv;
    if (p.enabled && x >= p.box[0] && x <= p.box[1] && y >= p.box[2] && y <= p.box[3]) {
        const double X = (x - p.c0[0]) * p.sc[0], Y = (y - p.c0[1]) * p.sc[1];
        double fx = 0.0, fy = 0.0;
#pragma unroll
        for (int i = B2_XYPOLY_N - 1; i >= 0; --i) {
            double rx = p.cx[i][B2_XYPOLY_N - 1], ry = p.cy[i][B2_XYPOLY_N - 1];
#pragma unroll
            for (int j = B2_XYPOLY_N - 2; j >= 0; --j) {
                rx = fma(rx, Y, p.cx[i][j]);
                ry = fma(ry, Y, p.cy[i][j]);
            }
            fx = fma(fx, X, rx);
            fy = fma(fy, X, ry);
        }
        thx = fx;
        thy = fy;
    } else {
        xy_to_field_exact(o, x, y, thx, thy);
    }
}

__device__ __forceinline__ void xy_to_v(const DevOptics& o, double x, double y, double& vx, double& vy, double& vz) {
    double thx, thy;
    xy_to_field(o, x, y, thx, thy);
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
    // fixed trip counts (B2Diffraction holds at most 8 lines and 4 circles; Rubin has 4 and 2): unrolled, so the
    // distances are independent instruction streams instead of a loop with a branch per vane
    double min_line = INFINITY, lnx = 0.0, lny = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        if (k >= c.n_lines) break;
        double d = fabs(fabs(c.lines[k][0] * px + c.lines[k][1] * py - c.lines[k][2]) - c.lines[k][3]);
        if (d < min_line) {
            min_line = d;
            lnx = c.lines[k][0];
            lny = c.lines[k][1];
        }
    }
    double min_circ = INFINITY, cdx = 0.0, cdy = 0.0, icnrm = 1.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (k >= c.n_circles) break;
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
__device__ __forceinline__ void departure(const DevSurf& s, const int kind, const int extra_kind, const bool small_poly,
                                          double x, double y, double r2, double& P, double& dP, double& ddP, double& E,
                                          double& Ex, double& Ey) {
    P = dP = ddP = 0.0;
    if (kind == B2_SURF_ASPHERE) {
        // P = r2^2 h(r2); Horner for h, h', h'' from the highest coefficient
        double h = 0.0, dh = 0.0, ddh = 0.0;
        if (small_poly) {  // n_coef <= 4: the usual case, fully unrolled (unused coefficients are zero)
#pragma unroll
            for (int k = 3; k >= 0; --k) {
                ddh = ddh * r2 + 2.0 * dh;
                dh = dh * r2 + h;
                h = h * r2 + s.coef[k];
            }
        } else {
#pragma unroll 1
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
    if (extra_kind == B2_EXTRA_POLY2D) poly2d_eval(s, x, y, E, Ex, Ey);
    else if (extra_kind == B2_EXTRA_BICUBIC) bicubic_eval(s.extra, x, y, E, Ex, Ey);
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

// One interface of batoid's CompoundOptic.trace: transform into the surface frame, intersect, interact
// (reflect / refract / detect), clear-aperture test.  The shape flags arrive as arguments: run-time
// fields of the DevSurf in the generic loop, compile-time constants in a surface program (below), where
// the branches fold away and every coefficient is a constant-bank operand at a fixed offset.
struct MediaN {
    double n[4], inv[4];
};

__device__ __forceinline__ double media_pick(const double v[4], int m) {
    return m == 0 ? v[0] : (m == 1 ? v[1] : (m == 2 ? v[2] : v[3]));
}
__device__ __forceinline__ double media_pick_n(const MediaN& mn, int m) { return media_pick(mn.n, m); }
__device__ __forceinline__ double media_pick_inv(const MediaN& mn, int m) { return media_pick(mn.inv, m); }

// what Snell's law needs of a (medium_in, medium_out) pair, per photon: na, 1/nb, eta = na/nb, eta^2, 1 - eta^2
struct Refr {
    double na, inb, eta, eta2, om;
};
__device__ __forceinline__ Refr make_refr(const MediaN& mn, int mi, int mo) {
    Refr f;
    f.na = media_pick_n(mn, mi);
    f.inb = media_pick_inv(mn, mo);
    f.eta = f.na * f.inb;
    f.eta2 = f.eta * f.eta;
    f.om = 1.0 - f.eta2;
    return f;
}



__device__ __forceinline__ void surface_step(const DevSurf& s, const int kind, const int interact, const Refr& rf,
                                             const int extra_kind, const bool simple_clear, const bool small_poly,
                                             Ray& r) {
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
    const bool curved = (kind != B2_SURF_PLANE);
    if (curved) {
        double A = vx * vx + vy * vy + s.k1 * vz * vz;
        double B = 2.0 * (px * vx + py * vy - s.R * vz);
        double C = px * px + py * py;
        double disc = B * B - 4.0 * A * C;
        ok = ok && (disc >= 0.0);
        double q = -0.5 * (B + copysign(b2sqrt_fast(disc), B));
        double t1 = C * b2rcp(q);
        dt += t1;
        px += vx * t1;
        py += vy * t1;
        pz = vz * t1;
    }
    // zc: height of the base conic under the hit point (= pz unless the surface departs from it)
    double zc = pz, gP = 0.0, Ex = 0.0, Ey = 0.0;
    const bool screen = (interact == B2_INT_PASS);  // batoid.OPDScreen: the summed term is an OPD, not sag
    if (!screen && (kind == B2_SURF_ASPHERE || extra_kind != B2_EXTRA_NONE)) {
        // Newton on the implicit form G(t) = r^2 - 2 R zc + k1 zc^2 with zc = z - P(r^2) - E(x, y):
        // polynomial in the ray parameter, no square root; quadratic convergence from the conic hit
        bool conv = false;
        const bool pure = (extra_kind == B2_EXTRA_NONE);
#pragma unroll 1
        for (int it = 0; it < 8; ++it) {
            double r2 = px * px + py * py;
            double P, dP, ddP, E;
            departure(s, kind, extra_kind, small_poly, px, py, r2, P, dP, ddP, E, Ex, Ey);
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
        return;
    }
    r.t += dt;
    if (screen) {
        // thin phase screen on a plane: the tangential part of the unit direction n v gains grad W, the path W
        double P, dP, ddP, W = 0.0, Wx = 0.0, Wy = 0.0;
        departure(s, B2_SURF_PLANE, extra_kind, true, px, py, px * px + py * py, P, dP, ddP, W, Wx, Wy);
        const double ux = vx * rf.na + Wx, uy = vy * rf.na + Wy;
        const double w2 = 1.0 - ux * ux - uy * uy;
        if (w2 <= 0.0) {
            r.failed = true;
            r.vignetted = true;
        } else {
            const double inv = b2rcp(rf.na);
            vx = ux * inv;
            vy = uy * inv;
            vz = copysign(b2sqrt(w2), vz) * inv;
            r.t += W;
        }
    }
    if (interact == B2_INT_MIRROR || interact == B2_INT_REFRACT) {
        // un-normalised normal N = (-Zx, -Zy, g): the surface gradient scaled by g = R - k1 zc, so the
        // conic part grad F = (x, y, k1 z - R) needs no division; the departure gradient is scaled to match
        const bool departs = (kind == B2_SURF_ASPHERE || extra_kind != B2_EXTRA_NONE);
        double g = 1.0, Zx = 0.0, Zy = 0.0;
        if (departs) {
            Zx = 2.0 * gP * px + Ex;
            Zy = 2.0 * gP * py + Ey;
        }
        if (curved) {
            g = s.R - s.k1 * zc;
            Zx = departs ? Zx * g + px : px;  // a pure conic needs no multiplication at all
            Zy = departs ? Zy * g + py : py;
        }
        double NN = g * g + Zx * Zx + Zy * Zy;
        double iNN = b2rcp(NN);
        double vn = -Zx * vx - Zy * vy + g * vz;  // v.N
        if (interact == B2_INT_MIRROR) {
            double f = 2.0 * vn * iNN;
            vx += f * Zx;
            vy += f * Zy;
            vz -= f * g;
        } else {
            // u = na v is the unit direction; N is oriented against u by flipping the sign of u.N (and of the
            // N term) instead of N itself
            const double uN = -fabs(rf.na * vn);  // u.N' <= 0
            // v' = (eta u - [eta uN + sqrt((1-eta^2) NN + eta^2 uN^2)]/NN N) / nb, with eta na / nb = eta^2
            double fac = (rf.eta * uN + b2sqrt_fast(rf.om * NN + rf.eta2 * uN * uN)) * (iNN * rf.inb);
            fac = vn > 0.0 ? -fac : fac;
            vx = rf.eta2 * vx + fac * Zx;
            vy = rf.eta2 * vy + fac * Zy;
            vz = rf.eta2 * vz - fac * g;
        }
    }
    if (simple_clear) {
        double r2 = px * px + py * py;
        if (!(s.clr_in2 <= r2 && r2 < s.clr_out2)) r.vignetted = true;
    } else {
        for (int k = 0; k < s.n_obsc; ++k)
            if (obscured(s.obsc[k], px, py)) r.vignetted = true;
    }
    r.x = px; r.y = py; r.z = pz;
    r.vx = vx; r.vy = vy; r.vz = vz;
}

// ---- surface programs: interface sequences known at compile time -----------------------------------
// Program 0 is the generic interpreter over DevOptics::surf.  Program B2_PROG_LSST is the Rubin
// layout (batoid LSST_[ugrizy].yaml): M1 M2 M3 aspheric mirrors, L1, L2 (exit aspheric), filter, L3
// (each an entrance and an exit conic or plane), detector plane; media 0 = air, 1 = glass; one centred
// annular / circular clear aperture per surface; no summed perturbation terms.  The host selects it
// when the uploaded telescope has exactly this signature (rotations and decentres are free: camera
// rotator, detector heights and rigid-body perturbations keep it); anything else runs program 0.
struct SurfSpec {
    bool asphere;  // base conic + even polynomial (Newton); otherwise plane / sphere / paraboloid / quadric,
                   // told apart at run time by the uniform `kind` and k1 of the DevSurf
    int interact, med_in, med_out;
};
#define B2_PROG_GENERIC 0
#define B2_PROG_LSST 1
#define B2_PROG_LSST_LEN 12

__host__ __device__ constexpr SurfSpec lsst_spec(int i) {
    switch (i) {
        case 0: case 1: case 2: return SurfSpec{true, B2_INT_MIRROR, 0, 0};  // M1, M2, M3
        case 3: return SurfSpec{false, B2_INT_REFRACT, 0, 1};                // L1 entrance
        case 4: return SurfSpec{false, B2_INT_REFRACT, 1, 0};                // L1 exit
        case 5: return SurfSpec{false, B2_INT_REFRACT, 0, 1};                // L2 entrance
        case 6: return SurfSpec{true, B2_INT_REFRACT, 1, 0};                 // L2 exit (asphere)
        case 7: return SurfSpec{false, B2_INT_REFRACT, 0, 1};                // filter entrance
        case 8: return SurfSpec{false, B2_INT_REFRACT, 1, 0};                // filter exit
        case 9: return SurfSpec{false, B2_INT_REFRACT, 0, 1};                // L3 entrance
        case 10: return SurfSpec{false, B2_INT_REFRACT, 1, 0};               // L3 exit
        default: return SurfSpec{false, B2_INT_DETECTOR, 0, 0};              // detector
    }
}

template <int IS>
__device__ __forceinline__ void lsst_steps(const DevOptics& o, const Refr& air_glass, const Refr& glass_air, Ray& r) {
    if constexpr (IS < B2_PROG_LSST_LEN) {
        constexpr SurfSpec sp = lsst_spec(IS);
        const DevSurf& s = o.surf[IS];
        // two media: the Snell constants of the two interface directions are computed once per photon
        surface_step(s, sp.asphere ? B2_SURF_ASPHERE : s.kind, sp.interact, sp.med_in == 0 ? air_glass : glass_air,
                     B2_EXTRA_NONE, true, true, r);
        lsst_steps<IS + 1>(o, air_glass, glass_air, r);
    }
}

// refractive indices and their inverses, once per photon per medium
template <int PROG>
__device__ __forceinline__ MediaN media_of(const DevOptics& o, double wl) {
    MediaN mn;
    if constexpr (PROG == B2_PROG_LSST) {
        mn.n[0] = medium_n(o.media[0], wl);
        mn.n[1] = medium_n(o.media[1], wl);
        mn.n[2] = mn.n[3] = 1.0;
        mn.inv[0] = b2rcp(mn.n[0]);
        mn.inv[1] = b2rcp(mn.n[1]);
        mn.inv[2] = mn.inv[3] = 1.0;
    } else {
        mn.n[0] = medium_n(o.media[0], wl);
        mn.n[1] = o.n_media > 1 ? medium_n(o.media[1], wl) : 1.0;
        mn.n[2] = o.n_media > 2 ? medium_n(o.media[2], wl) : 1.0;
        mn.n[3] = o.n_media > 3 ? medium_n(o.media[3], wl) : 1.0;
#pragma unroll
        for (int k = 0; k < 4; ++k) mn.inv[k] = b2rcp(mn.n[k]);
    }
    return mn;
}

// batoid CompoundOptic.trace: sequential interfaces
template <int PROG>
__device__ __forceinline__ void trace_ray(const DevOptics& o, const MediaN& mn, Ray& r) {
    if constexpr (PROG == B2_PROG_LSST) {
        const Refr ag = make_refr(mn, 0, 1), ga = make_refr(mn, 1, 0);
        lsst_steps<0>(o, ag, ga, r);
    } else {
#pragma unroll 1
        for (int is = 0; is < o.n_surf; ++is) {
            const DevSurf& s = o.surf[is];
            surface_step(s, s.kind, s.interact, make_refr(mn, s.med_in, s.med_out), s.extra_kind, s.simple_clear != 0,
                         s.n_coef <= 4, r);
        }
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

// One photon through the fused optics op: PhotonDCR -> shift -> xy->v -> n_air -> spider kick ->
// ray trace -> ray->pixel -> shift back -> FocusDepth -> Refraction.  In: pixel position, wavelength,
// pupil position, time, flux (+ the standard-normal draw of the kick).  Out: sensor-surface
// position, slopes, flux (0 if vignetted) and the flags.
struct OpticsOut {
    double x, y, dxdz, dydz, flux, t;
    bool vig, fail, offz;
};

template <int PROG>
__device__ __forceinline__ OpticsOut optics_photon(const DevOptics& o, const B2OpticsOptions& opt, double xi, double yi,
                                                   double wl_nm, double u, double v, double time, double flux,
                                                   double g) {
    OpticsOut out;
    double wl = wl_nm * 1e-9;
    if (opt.do_dcr) {  // galsim.PhotonDCR.applyTo
        if (opt.dcr_alpha != 0.0) {
            double sc = pow(wl_nm / opt.dcr_base_wavelength, opt.dcr_alpha);
            xi = sc * (xi - opt.dcr_center[0]) + opt.dcr_center[0];
            yi = sc * (yi - opt.dcr_center[1]) + opt.dcr_center[1];
        }
        double shift = dcr_refraction(wl_nm, opt.dcr_pth, opt.dcr_tanz) - opt.dcr_base_refraction;
        xi += shift * opt.dcr_m[0];
        yi += shift * opt.dcr_m[1];
    }
    if (opt.shift_in) {
        xi += opt.stamp_center[0];
        yi += opt.stamp_center[1];
    }
    Ray r;
    xy_to_v(o, xi, yi, r.vx, r.vy, r.vz);
    const MediaN mn = media_of<PROG>(o, wl);  // the stop medium's index is one of these: computed once
    const double inair = media_pick(mn.inv, o.medium_stop);
    r.vx *= inair;
    r.vy *= inair;
    r.vz *= inair;
    if (o.dif.enabled) diffraction_kick(o.dif, u, v, time, wl, g, r.vx, r.vy, r.vz);
    r.x = u;
    r.y = v;
    r.z = 0.0;
    r.t = 0.0;
    r.vignetted = false;
    r.failed = false;
    trace_ray<PROG>(o, mn, r);
    out.vig = r.vignetted;
    out.fail = r.failed;
    out.offz = !out.vig && !(fabs(r.z) < 1.0e-15);
    // ray_vector_to_photon_array
    double fpx = r.y * 1e3, fpy = r.x * 1e3;
    double xo = o.det.A[0] * fpx + o.det.A[1] * fpy + o.det.b[0];
    double yo = o.det.A[2] * fpx + o.det.A[3] * fpy + o.det.b[1];
    double iz = b2rcp(r.vz);
    double dx = (o.det.Jhat[0] * r.vx + o.det.Jhat[1] * r.vy) * iz;
    double dy = (o.det.Jhat[2] * r.vx + o.det.Jhat[3] * r.vy) * iz;
    double fl = out.vig ? 0.0 : flux;
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
    out.x = xo;
    out.y = yo;
    out.dxdz = dx;
    out.dydz = dy;
    out.flux = fl;
    out.t = r.t;
    return out;
}

// galsim.TimeSampler + galsim.PupilAnnulusSampler draws of photon `idx` (Philox streams 1, 2)
__device__ __forceinline__ void sample_time_pupil(uint64_t seed, uint64_t idx, double t0, double exptime, double r_in,
                                                  double r_out, double& time, double& pu, double& pv) {
    uint32_t r[4], q[4];
    philox4(seed, idx, 1u, r);
    philox4(seed, idx, 2u, q);
    double ut = u01(r[0], r[1]), ur = u01(r[2], r[3]), uphi = u01(q[0], q[1]);
    // explicit fused operations: the same bits whether this is inlined in the fused pool step or runs as the
    // sampler kernel (a contraction left to the compiler may differ between the two, and a one-ulp change of the
    // pupil position is amplified by the spider kick of a photon grazing a vane)
    time = fma(exptime, ut, t0);
    const double rin2 = r_in * r_in;
    double rr = b2sqrt_fast(fma(__dsub_rn(r_out * r_out, rin2), ur, rin2));
    double sn, cs;
    sincospi(2.0 * uphi, &sn, &cs);
    pu = rr * cs;
    pv = rr * sn;
}

