/*
 * oracle_sensor.c -- CPU restatement of galsim.SiliconSensor.accumulate /
 * calculate_pixel_areas (GalSim >= 2.7.2: galsim/sensor.py, src/Silicon.cpp,
 * src/Polygon.cpp, src/Table.cpp).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle_optics.c header).
 *
 * GalSim is a third-party dependency of the reference, not vendored in
 * /root/reference and not importable in the build container.  This file
 * restates its published algorithm from the reference's call sites
 * (imsim/photon_pooling.py:195-225, imsim/flat.py:220-264, imsim/stamp.py:562-572),
 * the reference's validation docs (doc/validation/{brighter-fatter,diffusion,
 * tree-ring}.rst) and the sensor-model data format (data/sensor_models/ *.dat).
 * PARITY UNPINNED at photon granularity: the reference's own tests pin this
 * path only through image moments (tests/test_sensor_models.py:13-37) and flat
 * statistics (tests/test_flats.py:69-165), which tests/test_sensor_stats.py
 * reproduces statistically.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../include/imsim_b200.h"
#ifdef _OPENMP
#include <omp.h>
#endif

#define PI 3.14159265358979323846

typedef struct {
    float x, y;
} f2;

typedef struct {
    B2SensorConfig cfg;
    int nv;   /* vertices per edge */
    int npoly; /* 4*nv+4 */
    double* emptypoly; /* npoly*2 */
    /* per-electron distortion kernels of the nx9*ny9 neighbourhood:
       KH[(ky*nx9+kx)*(nv+2)+k]: bottom edge of kernel pixel (kx,ky): BL corner (k=0), nv
                                 points, BR corner (k=nv+1) -- corners belong to the
                                 horizontal arrays, one pair per pixel
       KV[(ky*nx9+kx)*nv+k]: left-edge points, bottom -> top */
    f2* KH;
    f2* KV;
    /* tree-ring table (natural cubic spline if y2 != NULL else linear) */
    int ntr;
    double *tr_r, *tr_f, *tr_y2;
    int nabs;
    double *abs_w, *abs_l;
    /* bound image */
    int xmin, ymin, nx, ny, dtype_bytes;
    void* target; /* caller's pixel buffer, row-major ny*nx */
    double* delta;
    /* boundary state: H[(y*nx+x)*(nv+2)+k],     y in [0,ny], x in [0,nx): bottom edge of pixel
                                                 (x,y) with both its corners, pixel frame;
                       V[(y*(nx+1)+x)*nv+k],     y in [0,ny), x in [0,nx]: left edge of pixel (x,y) */
    f2* H;
    f2* V;
    double* inner; /* nx*ny*4: xmin xmax ymin ymax */
    double* outer;
    double accum_flux; /* flux since the last boundary update */
    int initialized;
} OrcSensor;

static double edge_frac(int nv, int k /* 0..nv-1 */) {
    double theta0 = -PI / 4.0;
    double dtheta = PI / (2.0 * (nv + 1.0));
    double theta = theta0 + (k + 1.0) * dtheta;
    return (tan(theta) + 1.0) / 2.0;
}

/* polygon order = order of the rows in the .dat file = sorted by angle from
   just past -pi: lower half of left edge (going down), BL corner, bottom edge
   (left->right), BR corner, right edge (up), TR corner, top edge (right->left),
   TL corner, upper half of left edge (going down). */
static void build_emptypoly(int nv, double* p) {
    int n = 0;
    for (int k = nv / 2 - 1; k >= 0; --k) { p[2 * n] = 0.0; p[2 * n + 1] = edge_frac(nv, k); n++; }
    p[2 * n] = 0.0; p[2 * n + 1] = 0.0; n++;
    for (int k = 0; k < nv; ++k) { p[2 * n] = edge_frac(nv, k); p[2 * n + 1] = 0.0; n++; }
    p[2 * n] = 1.0; p[2 * n + 1] = 0.0; n++;
    for (int k = 0; k < nv; ++k) { p[2 * n] = 1.0; p[2 * n + 1] = edge_frac(nv, k); n++; }
    p[2 * n] = 1.0; p[2 * n + 1] = 1.0; n++;
    for (int k = nv - 1; k >= 0; --k) { p[2 * n] = edge_frac(nv, k); p[2 * n + 1] = 1.0; n++; }
    p[2 * n] = 0.0; p[2 * n + 1] = 1.0; n++;
    for (int k = nv - 1; k >= nv / 2; --k) { p[2 * n] = 0.0; p[2 * n + 1] = edge_frac(nv, k); n++; }
}

/* natural cubic spline second derivatives (GalSim Table.cpp TSpline::setupSpline) */
static void spline_y2(int n, const double* x, const double* f, double* y2) {
    if (n < 3) {
        for (int i = 0; i < n; ++i) y2[i] = 0.0;
        return;
    }
    double* cp = (double*)malloc(sizeof(double) * n);
    double* dp = (double*)malloc(sizeof(double) * n);
    y2[0] = y2[n - 1] = 0.0;
    /* tridiagonal: h[i-1] y2[i-1] + 2(h[i-1]+h[i]) y2[i] + h[i] y2[i+1] = 6((f[i+1]-f[i])/h[i] - (f[i]-f[i-1])/h[i-1]) */
    cp[0] = 0.0;
    dp[0] = 0.0;
    for (int i = 1; i < n - 1; ++i) {
        double h0 = x[i] - x[i - 1], h1 = x[i + 1] - x[i];
        double b = 2.0 * (h0 + h1);
        double rhs = 6.0 * ((f[i + 1] - f[i]) / h1 - (f[i] - f[i - 1]) / h0);
        double m = b - h0 * cp[i - 1];
        cp[i] = h1 / m;
        dp[i] = (rhs - h0 * dp[i - 1]) / m;
    }
    for (int i = n - 2; i >= 1; --i) y2[i] = dp[i] - cp[i] * y2[i + 1];
    free(cp);
    free(dp);
}

static int table_index(int n, const double* x, double a) {
    /* upper index i such that x[i-1] <= a <= x[i] (GalSim ArgVec::upperIndex) */
    if (a <= x[0]) return 1;
    if (a >= x[n - 1]) return n - 1;
    int lo = 0, hi = n - 1;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (x[mid] <= a) lo = mid; else hi = mid;
    }
    return hi;
}

static double table_linear(int n, const double* x, const double* f, double a) {
    if (a < x[0]) a = x[0];
    if (a > x[n - 1]) a = x[n - 1];
    int i = table_index(n, x, a);
    double ax = (x[i] - a) / (x[i] - x[i - 1]);
    double bx = 1.0 - ax;
    return f[i] * bx + f[i - 1] * ax;
}

static double table_spline(int n, const double* x, const double* f, const double* y2, double a) {
    int i = table_index(n, x, a);
    double h = x[i] - x[i - 1];
    double aa = x[i] - a;
    double bb = h - aa;
    return (aa * f[i - 1] + bb * f[i] - (1. / 6.) * aa * bb * ((aa + h) * y2[i - 1] + (bb + h) * y2[i])) / h;
}

OrcSensor* orc_sensor_create(const B2SensorConfig* cfg, const double* vertex_data, const double* tr_r,
                             const double* tr_f, int tr_spline, const double* abs_w, const double* abs_l) {
    OrcSensor* s = (OrcSensor*)calloc(1, sizeof(OrcSensor));
    s->cfg = *cfg;
    int nv = s->nv = cfg->num_vertices;
    s->npoly = 4 * nv + 4;
    s->emptypoly = (double*)malloc(sizeof(double) * 2 * s->npoly);
    build_emptypoly(nv, s->emptypoly);
    int nx9 = cfg->nx, ny9 = cfg->ny;
    s->KH = (f2*)calloc((size_t)nx9 * ny9 * (nv + 2), sizeof(f2));
    s->KV = (f2*)calloc((size_t)nx9 * ny9 * nv, sizeof(f2));
    /* .dat rows: pixel index x-major (i = index/(ny*npoly)), then j, then vertex n.
       Per-electron displacement of vertex n of neighbour pixel (i,j) caused by
       num_elec electrons in the central pixel. */
    for (int i = 0; i < nx9; ++i)
        for (int j = 0; j < ny9; ++j)
            for (int n = 0; n < s->npoly; ++n) {
                const double* row = vertex_data + 5 * (((size_t)i * ny9 + j) * s->npoly + n);
                double x0 = row[0], y0 = row[1], x1 = row[3], y1 = row[4];
                double px = (x1 - x0) / cfg->pixel_size + 0.5;
                double py = (y1 - y0) / cfg->pixel_size + 0.5;
                f2 d;
                d.x = (float)((px - s->emptypoly[2 * n]) / cfg->num_elec);
                d.y = (float)((py - s->emptypoly[2 * n + 1]) / cfg->num_elec);
                /* own bottom edge (with both corners) and own left edge of each kernel pixel */
                if (n >= nv / 2 && n <= nv / 2 + nv + 1) {
                    s->KH[((size_t)j * nx9 + i) * (nv + 2) + (n - nv / 2)] = d;
                } else if (n < nv / 2) {
                    s->KV[((size_t)j * nx9 + i) * nv + (nv / 2 - 1 - n)] = d;
                } else if (n >= 7 * nv / 2 + 4) {
                    s->KV[((size_t)j * nx9 + i) * nv + (nv - 1 - (n - (7 * nv / 2 + 4)))] = d;
                }
            }
    s->ntr = cfg->n_treering;
    if (s->ntr > 2) {
        s->tr_r = (double*)malloc(sizeof(double) * s->ntr);
        s->tr_f = (double*)malloc(sizeof(double) * s->ntr);
        memcpy(s->tr_r, tr_r, sizeof(double) * s->ntr);
        memcpy(s->tr_f, tr_f, sizeof(double) * s->ntr);
        if (tr_spline) {
            s->tr_y2 = (double*)malloc(sizeof(double) * s->ntr);
            spline_y2(s->ntr, s->tr_r, s->tr_f, s->tr_y2);
        }
    }
    s->nabs = cfg->n_abs;
    if (s->nabs > 0) {
        s->abs_w = (double*)malloc(sizeof(double) * s->nabs);
        s->abs_l = (double*)malloc(sizeof(double) * s->nabs);
        memcpy(s->abs_w, abs_w, sizeof(double) * s->nabs);
        memcpy(s->abs_l, abs_l, sizeof(double) * s->nabs);
    }
    return s;
}

static void free_image_state(OrcSensor* s) {
    free(s->delta); free(s->H); free(s->V); free(s->inner); free(s->outer);
    s->delta = NULL; s->H = s->V = NULL; s->inner = s->outer = NULL;
}

void orc_sensor_destroy(OrcSensor* s) {
    if (!s) return;
    free_image_state(s);
    free(s->emptypoly); free(s->KH); free(s->KV);
    free(s->tr_r); free(s->tr_f); free(s->tr_y2); free(s->abs_w); free(s->abs_l);
    free(s);
}

/* spline second derivatives, exposed so the host side can be checked against it */
void orc_spline_y2(int n, const double* x, const double* f, double* y2) { spline_y2(n, x, f, y2); }
double orc_table_spline(int n, const double* x, const double* f, const double* y2, double a) {
    return table_spline(n, x, f, y2, a);
}

void orc_sensor_bind_image(OrcSensor* s, int xmin, int ymin, int nx, int ny, int dtype_bytes, void* pixels) {
    free_image_state(s);
    s->xmin = xmin; s->ymin = ymin; s->nx = nx; s->ny = ny; s->dtype_bytes = dtype_bytes;
    s->target = pixels;
    s->initialized = 0;
}

static inline double tget(const OrcSensor* s, int x, int y) {
    size_t k = (size_t)y * s->nx + x;
    return s->dtype_bytes == 4 ? (double)((float*)s->target)[k] : ((double*)s->target)[k];
}

static inline f2* Hp(const OrcSensor* s, int x, int y) { return s->H + ((size_t)y * s->nx + x) * (s->nv + 2); }
static inline f2* Vp(const OrcSensor* s, int x, int y) { return s->V + ((size_t)y * (s->nx + 1) + x) * s->nv; }

/* polygon of pixel (x,y) (array indices) in polygon order, pixel-local coords */
static void pixel_poly(const OrcSensor* s, int x, int y, double* p) {
    int nv = s->nv, n = 0;
    const f2* vl = Vp(s, x, y);
    const f2* vr = Vp(s, x + 1, y);
    const f2* hb = Hp(s, x, y);
    const f2* ht = Hp(s, x, y + 1);
    for (int k = nv / 2 - 1; k >= 0; --k) { p[2 * n] = vl[k].x; p[2 * n + 1] = vl[k].y; n++; }
    for (int k = 0; k <= nv + 1; ++k) { p[2 * n] = hb[k].x; p[2 * n + 1] = hb[k].y; n++; }
    for (int k = 0; k < nv; ++k) { p[2 * n] = (double)vr[k].x + 1.0; p[2 * n + 1] = vr[k].y; n++; }
    for (int k = nv + 1; k >= 0; --k) { p[2 * n] = ht[k].x; p[2 * n + 1] = (double)ht[k].y + 1.0; n++; }
    for (int k = nv - 1; k >= nv / 2; --k) { p[2 * n] = vl[k].x; p[2 * n + 1] = vl[k].y; n++; }
}

/* Silicon::updatePixelBounds */
static void update_bounds(OrcSensor* s, int x, int y) {
    double p[2 * 132];
    pixel_poly(s, x, y, p);
    double oxmin = INFINITY, oxmax = -INFINITY, oymin = INFINITY, oymax = -INFINITY;
    for (int n = 0; n < s->npoly; ++n) {
        double px = p[2 * n], py = p[2 * n + 1];
        if (px < oxmin) oxmin = px;
        if (px > oxmax) oxmax = px;
        if (py < oymin) oymin = py;
        if (py > oymax) oymax = py;
    }
    /* The "trivially inside" box, inscribed by construction: bounded by the innermost vertex of each edge, a
       corner counting for both edges it ends (a box built from the 45-degree wedges around the
       centre can keep the corner triangles of a pixel whose corners have moved in more than its edges; the
       reference's regression moments across 4 / 8 / 32-vertex models favour the inscribed one). */
    double ixmin = -INFINITY, ixmax = INFINITY, iymin = -INFINITY, iymax = INFINITY;
    for (int n = 0; n < s->npoly; ++n) {
        double px = p[2 * n], py = p[2 * n + 1];
        double ex = s->emptypoly[2 * n], ey = s->emptypoly[2 * n + 1];
        if (ex == 0.0 && px > ixmin) ixmin = px;
        if (ex == 1.0 && px < ixmax) ixmax = px;
        if (ey == 0.0 && py > iymin) iymin = py;
        if (ey == 1.0 && py < iymax) iymax = py;
    }
    size_t k = ((size_t)y * s->nx + x) * 4;
    s->outer[k] = oxmin; s->outer[k + 1] = oxmax; s->outer[k + 2] = oymin; s->outer[k + 3] = oymax;
    s->inner[k] = ixmin; s->inner[k + 1] = ixmax; s->inner[k + 2] = iymin; s->inner[k + 3] = iymax;
}

/* Silicon::calculateTreeRingDistortion on one stored boundary point of owner pixel (i,j) (image coords) */
static void treering_point(const OrcSensor* s, f2* pt, int i, int j, int ocx, int ocy) {
    double tx = (double)i + pt->x - s->cfg.treering_center[0] + (double)ocx;
    double ty = (double)j + pt->y - s->cfg.treering_center[1] + (double)ocy;
    double r = sqrt(tx * tx + ty * ty);
    if (r > 0 && r < s->tr_r[s->ntr - 1]) {
        double shift = s->tr_y2 ? table_spline(s->ntr, s->tr_r, s->tr_f, s->tr_y2, r)
                                : table_linear(s->ntr, s->tr_r, s->tr_f, r);
        double dx = shift * tx / r;
        double dy = shift * ty / r;
        pt->x = (float)((double)pt->x + dx);
        pt->y = (float)((double)pt->y + dy);
    }
}

static void init_boundaries(OrcSensor* s, int ocx, int ocy) {
    int nx = s->nx, ny = s->ny, nv = s->nv;
    size_t nH = (size_t)(ny + 1) * nx * (nv + 2), nV = (size_t)ny * (nx + 1) * nv;
    if (!s->H) {
        s->H = (f2*)malloc(nH * sizeof(f2));
        s->V = (f2*)malloc((nV ? nV : 1) * sizeof(f2));
        s->inner = (double*)malloc((size_t)nx * ny * 4 * sizeof(double));
        s->outer = (double*)malloc((size_t)nx * ny * 4 * sizeof(double));
        s->delta = (double*)calloc((size_t)nx * ny, sizeof(double));
    }
    for (int y = 0; y <= ny; ++y)
        for (int x = 0; x <= nx; ++x) {
            if (x < nx) {
                f2* h = Hp(s, x, y);
                h[0].x = 0.f; h[0].y = 0.f;
                for (int k = 0; k < nv; ++k) { h[k + 1].x = (float)edge_frac(nv, k); h[k + 1].y = 0.f; }
                h[nv + 1].x = 1.f; h[nv + 1].y = 0.f;
            }
            if (y < ny) {
                f2* v = Vp(s, x, y);
                for (int k = 0; k < nv; ++k) { v[k].x = 0.f; v[k].y = (float)edge_frac(nv, k); }
            }
        }
    if (s->ntr > 2) {
        for (int y = 0; y <= ny; ++y)
            for (int x = 0; x <= nx; ++x) {
                if (x < nx) {
                    f2* h = Hp(s, x, y);
                    for (int k = 0; k <= nv + 1; ++k) treering_point(s, &h[k], s->xmin + x, s->ymin + y, ocx, ocy);
                }
                if (y < ny) {
                    f2* v = Vp(s, x, y);
                    for (int k = 0; k < nv; ++k) treering_point(s, &v[k], s->xmin + x, s->ymin + y, ocx, ocy);
                }
            }
    }
}

/* Silicon::updatePixelDistortions: add charge * per-electron kernel to every
   boundary point within qdist pixels; charge(x,y) given by q() */
static void update_distortions(OrcSensor* s, const double* qd /* delta or NULL -> target */) {
    int nx = s->nx, ny = s->ny, nv = s->nv, q = s->cfg.qdist;
    int nx9 = s->cfg.nx, ny9 = s->cfg.ny;
    int cxk = (nx9 - 1) / 2, cyk = (ny9 - 1) / 2;
    uint8_t* changed = (uint8_t*)calloc((size_t)nx * ny, 1);
    /* horizontal rows */
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y <= ny; ++y)
        for (int x = 0; x < nx; ++x) {
            int i1 = x - q < 0 ? 0 : x - q, i2 = x + q > nx - 1 ? nx - 1 : x + q;
            int j1 = y - (q + 1) < 0 ? 0 : y - (q + 1), j2 = y + q > ny - 1 ? ny - 1 : y + q;
            int kmax = nv + 1;
            f2* h = Hp(s, x, y);
            int change = 0;
            for (int j = j1; j <= j2; ++j)
                for (int i = i1; i <= i2; ++i) {
                    double charge = qd ? qd[(size_t)j * nx + i] : tget(s, i, j);
                    if (charge == 0.0) continue;
                    change = 1;
                    const f2* kh = s->KH + ((size_t)(y - j + cyk) * nx9 + (x - i + cxk)) * (nv + 2);
                    for (int k = 0; k <= kmax; ++k) {
                        h[k].x = (float)((double)h[k].x + (double)kh[k].x * charge);
                        h[k].y = (float)((double)h[k].y + (double)kh[k].y * charge);
                    }
                }
            if (change) {
                for (int dy = -1; dy <= 0; ++dy) {
                    int py = y + dy;
                    if (py >= 0 && py < ny) changed[(size_t)py * nx + x] = 1;
                }
            }
        }
    /* vertical columns */
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < ny; ++y)
        for (int x = 0; x <= nx; ++x) {
            int i1 = x - (q + 1) < 0 ? 0 : x - (q + 1), i2 = x + q > nx - 1 ? nx - 1 : x + q;
            int j1 = y - q < 0 ? 0 : y - q, j2 = y + q > ny - 1 ? ny - 1 : y + q;
            f2* v = Vp(s, x, y);
            int change = 0;
            for (int j = j1; j <= j2; ++j)
                for (int i = i1; i <= i2; ++i) {
                    double charge = qd ? qd[(size_t)j * nx + i] : tget(s, i, j);
                    if (charge == 0.0) continue;
                    change = 1;
                    const f2* kv = s->KV + ((size_t)(y - j + cyk) * nx9 + (x - i + cxk)) * nv;
                    for (int k = 0; k < nv; ++k) {
                        v[k].x = (float)((double)v[k].x + (double)kv[k].x * charge);
                        v[k].y = (float)((double)v[k].y + (double)kv[k].y * charge);
                    }
                }
            if (change) {
                for (int dx = -1; dx <= 0; ++dx) {
                    int px = x + dx;
                    if (px >= 0 && px < nx) changed[(size_t)y * nx + px] = 1;
                }
            }
        }
#pragma omp parallel for schedule(static)
    for (int y = 0; y < ny; ++y)
        for (int x = 0; x < nx; ++x)
            if (changed[(size_t)y * nx + x]) update_bounds(s, x, y);
    free(changed);
}

static void add_delta(OrcSensor* s, double sign) {
    size_t n = (size_t)s->nx * s->ny;
    if (s->dtype_bytes == 4) {
        float* t = (float*)s->target;
        for (size_t k = 0; k < n; ++k) t[k] = (float)((double)t[k] + sign * s->delta[k]);
    } else {
        double* t = (double*)s->target;
        for (size_t k = 0; k < n; ++k) t[k] = t[k] + sign * s->delta[k];
    }
}

/* Silicon::update */
static void sensor_update(OrcSensor* s) {
    update_distortions(s, s->delta);
    add_delta(s, 1.0);
    memset(s->delta, 0, sizeof(double) * (size_t)s->nx * s->ny);
}

/* Silicon::initialize */
static void sensor_initialize(OrcSensor* s, int ocx, int ocy) {
    init_boundaries(s, ocx, ocy);
    update_distortions(s, NULL);
    for (int y = 0; y < s->ny; ++y)
        for (int x = 0; x < s->nx; ++x) update_bounds(s, x, y);
    memset(s->delta, 0, sizeof(double) * (size_t)s->nx * s->ny);
    s->accum_flux = 0.0;
    s->initialized = 1;
}

/* Polygon::contains (crossing test) */
static int poly_contains(int n, const double* p, double x, double y) {
    int inside = 0;
    double x1 = p[0], y1 = p[1];
    double xinters = 0.0;
    for (int i = 1; i <= n; ++i) {
        double x2 = p[2 * (i % n)], y2 = p[2 * (i % n) + 1];
        if (y > fmin(y1, y2)) {
            if (y <= fmax(y1, y2)) {
                if (x <= fmax(x1, x2)) {
                    if (y1 != y2) xinters = (y - y1) * (x2 - x1) / (y2 - y1) + x1;
                    if ((x1 == x2) || (x <= xinters)) inside = !inside;
                }
            }
        }
        x1 = x2;
        y1 = y2;
    }
    return inside;
}

/* Silicon::insidePixel; ix,iy image coords */
static int inside_pixel(const OrcSensor* s, int ix, int iy, double x, double y, double zconv, int* off_edge,
                        uint64_t* npoly_tests) {
    int ax = ix - s->xmin, ay = iy - s->ymin;
    if (ax < 0 || ax >= s->nx || ay < 0 || ay >= s->ny) {
        if (off_edge) *off_edge = 1;
        return 0;
    }
    size_t k = ((size_t)ay * s->nx + ax) * 4;
    const double* in = s->inner + k;
    const double* out = s->outer + k;
    int inside;
    if (x >= in[0] && x <= in[1] && y >= in[2] && y <= in[3]) {
        inside = 1;
    } else if (!(x >= out[0] && x <= out[1] && y >= out[2] && y <= out[3])) {
        inside = 0;
    } else {
        const double zfit = 12.0;
        const double zfactor = tanh(zconv / zfit);
        double p[2 * 132];
        pixel_poly(s, ax, ay, p);
        for (int n = 0; n < s->npoly; ++n) {
            double ex = s->emptypoly[2 * n], ey = s->emptypoly[2 * n + 1];
            p[2 * n] = ex + (p[2 * n] - ex) * zfactor;
            p[2 * n + 1] = ey + (p[2 * n + 1] - ey) * zfactor;
        }
        inside = poly_contains(s->npoly, p, x, y);
        if (npoly_tests) (*npoly_tests)++;
    }
    if (!inside && off_edge) {
        *off_edge = 0;
        if (ax == 0 && x < in[0]) *off_edge = 1;
        if (ax == s->nx - 1 && x > in[1]) *off_edge = 1;
        if (ay == 0 && y < in[2]) *off_edge = 1;
        if (ay == s->ny - 1 && y > in[3]) *off_edge = 1;
    }
    return inside;
}

static const int xoff[9] = {0, 1, 1, 0, -1, -1, -1, 0, 1};
static const int yoff[9] = {0, 0, 1, 1, 1, 0, -1, -1, -1};

/* Silicon::accumulate on photons [i1,i2) into delta */
static double accumulate_chunk(OrcSensor* s, int64_t i1, int64_t i2, const double* px, const double* py,
                               const double* pdxdz, const double* pdydz, const double* pwl, const double* pflux,
                               const double* rand4, int64_t n, B2AccumStats* st) {
    const double T = s->cfg.sensor_thickness, P = s->cfg.pixel_size;
    const double invPixelSize = 1. / P;
    const double diffStep_pixel_z = s->cfg.diff_step / (T * P);
    double added = 0.0;
    uint64_t npt = 0, nns = 0, nnf = 0, nb9 = 0, ndrop = 0;
    const double* g1 = rand4;
    const double* g2 = rand4 + n;
    const double* unf = rand4 + 2 * n;
    const double* udep = rand4 + 3 * n;
    /* photons of one chunk are independent (GalSim's loop is "#pragma omp parallel for" too) */
#pragma omp parallel for schedule(static) reduction(+ : added, npt, nns, nnf, nb9, ndrop)
    for (int64_t i = i1; i < i2; ++i) {
        double x0 = px[i], y0 = py[i];
        /* calculateConversionDepth */
        double dz;
        if (pwl) {
            double abs_length = table_linear(s->nabs, s->abs_w, s->abs_l, pwl[i]);
            double si_length = -abs_length * log(1.0 - udep[i]);
            if (pdxdz) {
                double a = pdxdz[i], b = pdydz[i];
                double d = si_length / sqrt(1.0 + a * a + b * b);
                dz = fmin(T - 1.0, d);
            } else {
                dz = si_length;
            }
        } else {
            dz = 1.0;
        }
        if (pdxdz) {
            double dz_pixel = dz * invPixelSize;
            x0 += pdxdz[i] * dz_pixel;
            y0 += pdydz[i] * dz_pixel;
        }
        double zconv = T - dz;
        if (zconv < 0.0) { ndrop++; continue; }
        if (s->cfg.diff_step != 0.) {
            double diffStep = fmax(0.0, diffStep_pixel_z * sqrt(zconv * T));
            x0 += diffStep * g1[i];
            y0 += diffStep * g2[i];
        }
        int ix = (int)floor(x0 + 0.5);
        int iy = (int)floor(y0 + 0.5);
        double x = x0 - ix + 0.5;
        double y = y0 - iy + 0.5;
        if (fabs(x) < 1e-9 || fabs(x - 1.0) < 1e-9 || fabs(y) < 1e-9 || fabs(y - 1.0) < 1e-9) nb9++;
        int off_edge = 0;
        int found = inside_pixel(s, ix, iy, x, y, zconv, &off_edge, &npt);
        if (!found && off_edge) continue;
        int step = 0;
        if (!found) {
            nns++;
            if ((x > y) && (x > 1.0 - y)) step = 1;
            else if ((x > y) && (x < 1.0 - y)) step = 7;
            else if ((x < y) && (x > 1.0 - y)) step = 3;
            else step = 5;
            int nn = step;
            for (int m = 1; m < 9; ++m) {
                int ix_off = ix + xoff[nn], iy_off = iy + yoff[nn];
                double x_off = x - xoff[nn], y_off = y - yoff[nn];
                if (inside_pixel(s, ix_off, iy_off, x_off, y_off, zconv, NULL, &npt)) {
                    ix = ix_off;
                    iy = iy_off;
                    found = 1;
                    break;
                }
                nn = ((nn - 1) + step) % 8 + 1;
            }
        }
        if (!found) {
            nnf++;
            int nn = (unf[i] > 0.5) ? 0 : step;
            ix += xoff[nn];
            iy += yoff[nn];
        }
        int ax = ix - s->xmin, ay = iy - s->ymin;
        if (ax >= 0 && ax < s->nx && ay >= 0 && ay < s->ny) {
            double flux = pflux[i];
#pragma omp atomic
            s->delta[(size_t)ay * s->nx + ax] += flux;
            added += flux;
        }
    }
    if (st) {
        st->n_polygon_tests += npt;
        st->n_neighbor_search += nns;
        st->n_not_found += nnf;
        st->n_boundary_1e9 += nb9;
        st->n_dropped_bottom += ndrop;
    }
    return added;
}

/* galsim.SiliconSensor.accumulate(photons, image, orig_center, resume, recalc) */
double orc_sensor_accumulate(OrcSensor* s, int64_t n, const double* x, const double* y, const double* dxdz,
                             const double* dydz, const double* wl_nm, const double* flux, const double* rand4,
                             int ocx, int ocy, int resume, int recalc, B2AccumStats* st) {
    if (st) memset(st, 0, sizeof(*st));
    if (!resume || !s->initialized) {
        sensor_initialize(s, ocx, ocy);
    } else {
        add_delta(s, -1.0); /* subtractDelta: image seen by the caller included the pending delta */
        if (recalc) {
            sensor_update(s);
            s->accum_flux = 0.0;
            if (st) st->n_updates++;
        }
    }
    double added = 0.0;
    double nrecalc = s->cfg.nrecalc;
    int64_t i1 = 0;
    while (i1 < n) {
        int64_t i2 = n;
        int hit = 0;
        if (nrecalc > 0) {
            double acc = s->accum_flux;
            for (int64_t i = i1; i < n; ++i) {
                acc += flux[i];
                if (acc >= nrecalc) {
                    i2 = i + 1;
                    hit = 1;
                    break;
                }
            }
            if (!hit) s->accum_flux = acc;
        }
        added += accumulate_chunk(s, i1, i2, x, y, dxdz, dydz, wl_nm, flux, rand4, n, st);
        if (hit) {
            sensor_update(s);
            s->accum_flux = 0.0;
            if (st) st->n_updates++;
        }
        i1 = i2;
    }
    add_delta(s, 1.0); /* addDelta: show the pending charge in the image, keep it in delta */
    if (st) st->added_flux = added;
    return added;
}

/* Silicon::fillWithPixelAreas */
void orc_sensor_pixel_areas(OrcSensor* s, int ocx, int ocy, int use_flux, double* areas) {
    init_boundaries(s, ocx, ocy);
    if (use_flux) update_distortions(s, NULL);
    s->initialized = 0;
    double p[2 * 132];
    for (int y = 0; y < s->ny; ++y)
        for (int x = 0; x < s->nx; ++x) {
            pixel_poly(s, x, y, p);
            double area = 0.0;
            for (int n = 0; n < s->npoly; ++n) {
                int n2 = (n + 1) % s->npoly;
                area += p[2 * n] * p[2 * n2 + 1];
                area -= p[2 * n2] * p[2 * n + 1];
            }
            areas[(size_t)y * s->nx + x] = fabs(area) / 2.0;
        }
}

/* galsim.Sensor.accumulate = PhotonArray.addTo */
double orc_plain_accumulate(int xmin, int ymin, int nx, int ny, int dtype_bytes, void* pixels, int64_t n,
                            const double* x, const double* y, const double* flux) {
    double added = 0.0;
    for (int64_t i = 0; i < n; ++i) {
        int ix = (int)floor(x[i] + 0.5) - xmin;
        int iy = (int)floor(y[i] + 0.5) - ymin;
        if (ix >= 0 && ix < nx && iy >= 0 && iy < ny) {
            size_t k = (size_t)iy * nx + ix;
            if (dtype_bytes == 4) ((float*)pixels)[k] = (float)((double)((float*)pixels)[k] + flux[i]);
            else ((double*)pixels)[k] += flux[i];
            added += flux[i];
        }
    }
    return added;
}

void orc_sensor_get_pixel(OrcSensor* s, int ix, int iy, double* poly, double* bounds) {
    int ax = ix - s->xmin, ay = iy - s->ymin;
    pixel_poly(s, ax, ay, poly);
    size_t k = ((size_t)ay * s->nx + ax) * 4;
    memcpy(bounds, s->inner + k, 4 * sizeof(double));
    memcpy(bounds + 4, s->outer + k, 4 * sizeof(double));
}

/* thread control for the CPU-baseline timings */
void orc_set_threads(int n) {
#ifdef _OPENMP
    omp_set_num_threads(n);
#endif
}
int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
