// sensor_device.cuh -- device state and functions of the silicon sensor (shared by sensor.cu and pool.cu)
#pragma once
#include "b2_common.cuh"

#ifndef PI_D
#define PI_D 3.14159265358979323846
#endif
#define B2_MAX_NV 32

struct DevSensor {
    int nv, nx9, ny9, qdist;
    int xmin, ymin, nx, ny;
    int ntr, nabs, tr_spline, pad;
    double diff_step, pixel_size, thickness;
    double inv_pixel_size, diff_step_pixel_z;  // 1 / pixel_size and diff_step / (thickness * pixel_size), from the host
    double trc[2];
    double tr_max;
    const double *tr_r, *tr_f, *tr_y2;
    const double *abs_w, *abs_l;
    double abs_x0, abs_inv_dx;  // uniform absorption table: direct index (inv_dx = 0: binary search)
    const float2 *KH, *KV;
    const double2 *KHd, *KVd;  // the same tables widened to double (the boundary update adds in double)
    float2 *H, *V;
    double *inner, *outer;
    double* delta;
    void* target;
    int dtype_bytes, pad2;
    double frac[B2_MAX_NV];  // (tan(theta_k)+1)/2, k = 0..nv-1, double (GalSim _emptypoly)
};

struct b2_sensor {
    b2_ctx* ctx = nullptr;
    B2SensorConfig cfg;
    DevSensor d;
    std::vector<void*> owned;        // tables
    std::vector<void*> image_owned;  // per-image state
    bool bound = false, initialized = false;
    int sm_count = 148;
    double accum_flux = 0.0;
    uint8_t* changed = nullptr;
    uint8_t* tiles = nullptr;  // charge occupancy per 32x32 tile
    int tnx = 0, tny = 0;
    unsigned long long* dstats = nullptr;  // device counters
    double* dadded = nullptr;
    Scratch cum;  // cumulative flux scratch
    size_t cap_H = 0, cap_V = 0, cap_pix_bytes = 0, cap_tiles = 0;  // capacities of the per-image arrays
    Scratch slow;  // compact list of photons that need the full polygon / neighbour treatment
    double* tr_buf[3] = {nullptr, nullptr, nullptr};  // tree-ring tables of b2_sensor_set_treerings (reused)
    size_t tr_cap = 0;
    unsigned long long* dnslow = nullptr;
    Scratch stamp_meta, stamp_arena;  // stamps.cu: job tables + per-block lists; boundary state of the stamps in flight
    cudaStream_t stamp_aux = nullptr;  // stamps.cu: the cluster launch runs beside the one-block-per-stamp launch
    cudaEvent_t stamp_ev[2] = {nullptr, nullptr};
};

// x rounded to the nearest float (ties to even, float denormals included) and kept as a double: (double)(float)x
// without the two trips through the conversion unit (16 lanes / clock / SM).  Adding 1.5 * 2^(e+29), e the exponent
// of x, pushes the bits below float precision out of the double's significand under the FP64 adder's own
// round-to-nearest-even (the sum stays inside the magic's binade for either sign of x); subtracting it again is
// exact.  Below 2^-126 the magic is pinned so that the spacing is the float denormals' 2^-149.  (No float overflow
// handling: boundary points are pixel fractions; -0.0 returns +0.0.)
__device__ __forceinline__ double round_to_f32(double x) {
    const int e = max(__double2hiint(x) & 0x7ff00000, 897 << 20);
    const double m = __hiloint2double(e + ((29 << 20) | 0x00080000), 0);
    return __dsub_rn(__dadd_rn(x, m), m);
}

// one term of Silicon::updatePixelDistortions on a boundary coordinate held as a float-valued double:
// p = float(double(p) + double(d) * c), GalSim's float += float * double
__device__ __forceinline__ void bf_term(double& p, double d, double c) {
    p = round_to_f32(__dadd_rn(p, __dmul_rn(d, c)));
}

enum { ST_POLY = 0, ST_NEIGH = 1, ST_NOTFOUND = 2, ST_B9 = 3, ST_DROP = 4, ST_N = 8 };

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ size_t Hidx(const DevSensor& s, int x, int y) {
    return ((size_t)y * s.nx + x) * (s.nv + 2);
}
__device__ __forceinline__ size_t Vidx(const DevSensor& s, int x, int y) {
    return ((size_t)y * (s.nx + 1) + x) * s.nv;
}

__device__ __forceinline__ int table_index(int n, const double* __restrict__ x, double a) {
    if (a <= __ldg(x)) return 1;
    if (a >= __ldg(x + n - 1)) return n - 1;
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(x + mid) <= a) lo = mid; else hi = mid;
    }
    return hi;
}

// the same index for a uniformly spaced table without the search: a guess from the spacing, corrected against
// the table itself so that it is exactly the index the search returns (first node above a)
__device__ __forceinline__ int table_index_uniform(int n, const double* __restrict__ x, double a, double x0,
                                                   double inv_dx) {
    if (a <= __ldg(x)) return 1;
    if (a >= __ldg(x + n - 1)) return n - 1;
    int i = (int)((a - x0) * inv_dx) + 1;
    i = min(max(i, 1), n - 1);
    if (a < __ldg(x + i - 1)) --i;
    else if (a >= __ldg(x + i)) ++i;
    return min(max(i, 1), n - 1);
}

// GalSim Table.cpp linear / spline interpolation; explicit _rn ops: no FMA contraction,
// so the values match the host (non-FMA) evaluation bit for bit
__device__ __forceinline__ double table_linear(int n, const double* __restrict__ x, const double* __restrict__ f, double a,
                                               double x0 = 0.0, double inv_dx = 0.0) {
    a = fmin(fmax(a, __ldg(x)), __ldg(x + n - 1));
    int i = inv_dx != 0.0 ? table_index_uniform(n, x, a, x0, inv_dx) : table_index(n, x, a);
    double xi = __ldg(x + i), xm = __ldg(x + i - 1);
    double ax = __ddiv_rn(__dsub_rn(xi, a), __dsub_rn(xi, xm));
    double bx = __dsub_rn(1.0, ax);
    return __dadd_rn(__dmul_rn(__ldg(f + i), bx), __dmul_rn(__ldg(f + i - 1), ax));
}

__device__ __forceinline__ double table_spline(int n, const double* __restrict__ x, const double* __restrict__ f,
                                               const double* __restrict__ y2, double a) {
    int i = table_index(n, x, a);
    double xi = __ldg(x + i), xm = __ldg(x + i - 1);
    double h = __dsub_rn(xi, xm);
    double aa = __dsub_rn(xi, a);
    double bb = __dsub_rn(h, aa);
    double t1 = __dadd_rn(__dmul_rn(aa, __ldg(f + i - 1)), __dmul_rn(bb, __ldg(f + i)));
    double t2 = __dadd_rn(__dmul_rn(__dadd_rn(aa, h), __ldg(y2 + i - 1)), __dmul_rn(__dadd_rn(bb, h), __ldg(y2 + i)));
    double t3 = __dmul_rn(__dmul_rn(__dmul_rn(1. / 6., aa), bb), t2);
    return __ddiv_rn(__dsub_rn(t1, t3), h);
}

// walk the polygon of pixel (ax, ay) counter-clockwise starting at the BL corner and
// call f(n_is, px, py, ex, ey) with the stored point (pixel frame) and the undistorted one
// (NVT > 0: vertex count known at compile time, loops unrolled and the loads independent)
template <int NVT = 0, typename F>
__device__ __forceinline__ void walk_polygon(const DevSensor& s, int ax, int ay, F&& f) {
    const int nv = NVT > 0 ? NVT : s.nv;
    const float2* hb = s.H + Hidx(s, ax, ay);
    const float2* ht = s.H + Hidx(s, ax, ay + 1);
    const float2* vl = s.V + Vidx(s, ax, ay);
    const float2* vr = vl + nv;
    // bottom edge: BL corner, nv points, BR corner (all in this pixel's frame)
    {
        float2 p = hb[0];
        f((double)p.x, (double)p.y, 0.0, 0.0);
    }
#pragma unroll
    for (int k = 0; k < nv; ++k) {
        float2 p = hb[k + 1];
        f((double)p.x, (double)p.y, s.frac[k], 0.0);
    }
    {
        float2 p = hb[nv + 1];
        f((double)p.x, (double)p.y, 1.0, 0.0);
    }
#pragma unroll
    for (int k = 0; k < nv; ++k) {
        float2 p = vr[k];
        f((double)p.x + 1.0, (double)p.y, 1.0, s.frac[k]);
    }
    {
        float2 p = ht[nv + 1];
        f((double)p.x, (double)p.y + 1.0, 1.0, 1.0);
    }
#pragma unroll
    for (int k = nv - 1; k >= 0; --k) {
        float2 p = ht[k + 1];
        f((double)p.x, (double)p.y + 1.0, s.frac[k], 1.0);
    }
    {
        float2 p = ht[0];
        f((double)p.x, (double)p.y + 1.0, 0.0, 1.0);
    }
#pragma unroll
    for (int k = nv - 1; k >= 0; --k) {
        float2 p = vl[k];
        f((double)p.x, (double)p.y, 0.0, s.frac[k]);
    }
}

// Silicon::insidePixel.  ix, iy: image coordinates.  Returns inside; sets *off_edge like GalSim.
template <int NVT = 0>
__device__ __forceinline__ bool inside_pixel(const DevSensor& s, int ix, int iy, double x, double y, double zconv,
                                             bool* off_edge, unsigned& npoly) {
    int ax = ix - s.xmin, ay = iy - s.ymin;
    if (ax < 0 || ax >= s.nx || ay < 0 || ay >= s.ny) {
        if (off_edge) *off_edge = true;
        return false;
    }
    size_t k = ((size_t)ay * s.nx + ax) * 4;
    const double4 in = *reinterpret_cast<const double4*>(s.inner + k);  // xmin xmax ymin ymax
    bool inside;
    if (x >= in.x && x <= in.y && y >= in.z && y <= in.w) {
        inside = true;
    } else {
        const double4 out = *reinterpret_cast<const double4*>(s.outer + k);
        if (!(x >= out.x && x <= out.y && y >= out.z && y <= out.w)) {
            inside = false;
        } else {
            const double zfactor = tanh(zconv / 12.0);
            // Polygon::contains crossing test over consecutive vertices
            bool in_poly = false;
            double x1 = 0.0, y1 = 0.0, xf = 0.0, yf = 0.0;
            bool first = true;
            auto edge = [&](double xa, double ya, double xb, double yb) {
                if (y > fmin(ya, yb)) {
                    if (y <= fmax(ya, yb)) {
                        if (x <= fmax(xa, xb)) {
                            double xinters = 0.0;
                            bool have = (ya != yb);
                            if (have) xinters = __dadd_rn(__ddiv_rn(__dmul_rn(__dsub_rn(y, ya), __dsub_rn(xb, xa)), __dsub_rn(yb, ya)), xa);
                            // GalSim keeps the previous xinters when ya == yb; with y > min and
                            // y <= max that case is impossible (min == max), so it is never read
                            if ((xa == xb) || (x <= xinters)) in_poly = !in_poly;
                        }
                    }
                }
            };
            walk_polygon<NVT>(s, ax, ay, [&](double px, double py, double ex, double ey) {
                double qx = __dadd_rn(ex, __dmul_rn(__dsub_rn(px, ex), zfactor));
                double qy = __dadd_rn(ey, __dmul_rn(__dsub_rn(py, ey), zfactor));
                if (first) {
                    xf = qx; yf = qy;
                    first = false;
                } else {
                    edge(x1, y1, qx, qy);
                }
                x1 = qx; y1 = qy;
            });
            edge(x1, y1, xf, yf);
            inside = in_poly;
            npoly++;
        }
    }
    if (!inside && off_edge) {
        *off_edge = false;
        if (ax == 0 && x < in.x) *off_edge = true;
        if (ax == s.nx - 1 && x > in.y) *off_edge = true;
        if (ay == 0 && y < in.z) *off_edge = true;
        if (ay == s.ny - 1 && y > in.w) *off_edge = true;
    }
    return inside;
}

static __constant__ int c_xoff[9] = {0, 1, 1, 0, -1, -1, -1, 0, 1};
static __constant__ int c_yoff[9] = {0, 0, 1, 1, 1, 0, -1, -1, -1};

__device__ __forceinline__ unsigned long long warp_sum(unsigned v) {
    return (unsigned long long)__reduce_add_sync(0xffffffffu, v);
}

struct SlowRec {
    int ix, iy;
    double x, y, zconv, flux;
    int coin, pad;  // unf > 0.5
};

// The first (converged) phase of Silicon::accumulate for one photon: conversion depth, drift to the
// conversion point, diffusion, nominal pixel, inner-box test.  Deposits decided photons into `delta`
// and returns true (filling `rec`) for the ones that need the polygon / neighbour treatment.
// (dax, day): array coordinates of the pixel a decided photon was deposited in, (-1, -1) if none
__device__ __forceinline__ bool sensor_fast_path_ex(const DevSensor& s, double x0, double y0, bool has_angles, double a,
                                                    double b, bool has_wl, double wl_nm, double flux, double g1,
                                                    double g2, double unf, double udep, SlowRec& rec, double& my_added,
                                                    unsigned& nb9, unsigned& ndrop, int& dax, int& day) {
    dax = day = -1;
    const double T = s.thickness;
    const double invPixelSize = s.inv_pixel_size;         // the same IEEE quotients the reference forms per call,
    const double diffStep_pixel_z = s.diff_step_pixel_z;  // computed once on the host
    // calculateConversionDepth
    double dz;
    if (has_wl) {
        double abs_length = table_linear(s.nabs, s.abs_w, s.abs_l, wl_nm, s.abs_x0, s.abs_inv_dx);
        double si_length = __dmul_rn(-abs_length, log(__dsub_rn(1.0, udep)));
        if (has_angles) {
            double nrm = sqrt(__dadd_rn(__dadd_rn(1.0, __dmul_rn(a, a)), __dmul_rn(b, b)));
            dz = fmin(T - 1.0, __ddiv_rn(si_length, nrm));
        } else {
            dz = si_length;
        }
    } else {
        dz = 1.0;
    }
    if (has_angles) {
        double dz_pixel = __dmul_rn(dz, invPixelSize);
        x0 = __dadd_rn(x0, __dmul_rn(a, dz_pixel));
        y0 = __dadd_rn(y0, __dmul_rn(b, dz_pixel));
    }
    double zconv = __dsub_rn(T, dz);
    if (zconv < 0.0) {
        ndrop = 1;
        return false;
    }
    if (s.diff_step != 0.) {
        double diffStep = fmax(0.0, __dmul_rn(diffStep_pixel_z, sqrt(__dmul_rn(zconv, T))));
        x0 = __dadd_rn(x0, __dmul_rn(diffStep, g1));
        y0 = __dadd_rn(y0, __dmul_rn(diffStep, g2));
    }
    int ix = (int)floor(x0 + 0.5);
    int iy = (int)floor(y0 + 0.5);
    double x = __dadd_rn(__dsub_rn(x0, (double)ix), 0.5);
    double y = __dadd_rn(__dsub_rn(y0, (double)iy), 0.5);
    if (fabs(x) < 1e-9 || fabs(x - 1.0) < 1e-9 || fabs(y) < 1e-9 || fabs(y - 1.0) < 1e-9) nb9 = 1;
    int ax = ix - s.xmin, ay = iy - s.ymin;
    // nominal pixel off the image: insidePixel() fails with off_edge set -> the photon is lost
    if (!(ax >= 0 && ax < s.nx && ay >= 0 && ay < s.ny)) return false;
    size_t k = (size_t)ay * s.nx + ax;
    const double4 in = *reinterpret_cast<const double4*>(s.inner + k * 4);
    if (x >= in.x && x <= in.y && y >= in.z && y <= in.w) {
        atomicAdd(&s.delta[k], flux);
        my_added = flux;
        dax = ax;
        day = ay;
        return false;
    }
    rec.ix = ix; rec.iy = iy;
    rec.x = x; rec.y = y;
    rec.zconv = zconv; rec.flux = flux;
    rec.coin = (unf > 0.5) ? 1 : 0;
    rec.pad = 0;
    return true;
}

__device__ __forceinline__ bool sensor_fast_path(const DevSensor& s, double x0, double y0, bool has_angles, double a,
                                                 double b, bool has_wl, double wl_nm, double flux, double g1, double g2,
                                                 double unf, double udep, SlowRec& rec, double& my_added, unsigned& nb9,
                                                 unsigned& ndrop) {
    int dax, day;
    return sensor_fast_path_ex(s, x0, y0, has_angles, a, b, has_wl, wl_nm, flux, g1, g2, unf, udep, rec, my_added, nb9,
                               ndrop, dax, day);
}

// The second phase for one listed photon: outer box, polygon test, neighbour search, coin flip -- the
// reference's sequence (Silicon::accumulate / searchNeighbors).  Returns the flux deposited (0 if lost).
template <int NVT>
__device__ __forceinline__ double slow_photon(const DevSensor& s, const SlowRec& r, unsigned& npoly, unsigned& nneigh,
                                              unsigned& nnf, int& dax, int& day) {
    dax = day = -1;
    int ix = r.ix, iy = r.iy;
    const double x = r.x, y = r.y, zconv = r.zconv;
    bool off_edge = false;
    bool found = inside_pixel<NVT>(s, ix, iy, x, y, zconv, &off_edge, npoly);
    if (!found && off_edge) return 0.0;
    int step = 0;
    if (!found) {
        nneigh++;
        if ((x > y) && (x > 1.0 - y)) step = 1;
        else if ((x > y) && (x < 1.0 - y)) step = 7;
        else if ((x < y) && (x > 1.0 - y)) step = 3;
        else step = 5;
        int nn = step;
#pragma unroll 1
        for (int m = 1; m < 9; ++m) {
            int ix_off = ix + c_xoff[nn], iy_off = iy + c_yoff[nn];
            double x_off = x - c_xoff[nn], y_off = y - c_yoff[nn];
            if (inside_pixel<NVT>(s, ix_off, iy_off, x_off, y_off, zconv, nullptr, npoly)) {
                ix = ix_off;
                iy = iy_off;
                found = true;
                break;
            }
            nn = ((nn - 1) + step) % 8 + 1;
        }
    }
    if (!found) {
        nnf++;
        int nn = r.coin ? 0 : step;
        ix += c_xoff[nn];
        iy += c_yoff[nn];
    }
    int ax = ix - s.xmin, ay = iy - s.ymin;
    if (ax >= 0 && ax < s.nx && ay >= 0 && ay < s.ny) {
        atomicAdd(&s.delta[(size_t)ay * s.nx + ax], r.flux);
        dax = ax;
        day = ay;
        return r.flux;
    }
    return 0.0;
}

// Silicon::calculateTreeRingDistortion on one stored boundary point of the slot (i, j) (image coordinates)
__device__ __forceinline__ void treering_point(const DevSensor& s, float2& pt, int i, int j, int ocx, int ocy) {
    double tx = (double)i + (double)pt.x - s.trc[0] + (double)ocx;
    double ty = (double)j + (double)pt.y - s.trc[1] + (double)ocy;
    double r = sqrt(__dadd_rn(__dmul_rn(tx, tx), __dmul_rn(ty, ty)));
    if (r > 0 && r < s.tr_max) {
        double shift = s.tr_spline ? table_spline(s.ntr, s.tr_r, s.tr_f, s.tr_y2, r) : table_linear(s.ntr, s.tr_r, s.tr_f, r);
        double dx = __ddiv_rn(__dmul_rn(shift, tx), r);
        double dy = __ddiv_rn(__dmul_rn(shift, ty), r);
        pt.x = (float)((double)pt.x + dx);
        pt.y = (float)((double)pt.y + dy);
    }
}


// the four sensor draws of photon `idx` (Philox stream 3): one Philox block per photon; two 24-bit
// uniforms -> Box-Muller in FP32 (a diffusion step with 1e-7 relative granularity is statistically
// exact), two 32-bit uniforms
__device__ __forceinline__ void sensor_draws(uint64_t seed, uint64_t idx, double& g1, double& g2, double& unf,
                                             double& udep) {
    uint32_t r[4];
    philox4(seed, idx, 3u, r);
    float u1 = ((float)(r[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float u2 = ((float)(r[1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
    float sn, cs;
    sincospif(2.0f * u2, &sn, &cs);
    float rad = sqrtf(-2.0f * logf(u1));
    g1 = (double)(rad * cs);
    g2 = (double)(rad * sn);
    udep = ((double)r[2] + 0.5) * (1.0 / 4294967296.0);
    unf = ((double)r[3] + 0.5) * (1.0 / 4294967296.0);
}

// warp-aggregated append of the lanes with to_slow set
__device__ __forceinline__ void slow_append(bool to_slow, const SlowRec& rec, SlowRec* __restrict__ slow,
                                            unsigned long long* __restrict__ nslow) {
    unsigned m = __ballot_sync(0xffffffffu, to_slow);
    if (m) {
        int lane = threadIdx.x & 31;
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(nslow, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (to_slow) slow[base + __popc(m & ((1u << lane) - 1u))] = rec;
    }
}

// internal host entry points of sensor.cu used by the fused pool step (pool.cu)
int b2_sensor_begin_accumulate(b2_sensor* s, int32_t ocx, int32_t ocy, int32_t resume, int32_t recalc, int64_t n,
                               uint64_t* n_updates);
int b2_sensor_run_slow(b2_sensor* s, int64_t n);
int b2_sensor_end_accumulate(b2_sensor* s);
int b2_sensor_update_now(b2_sensor* s);
