// stamps.cu -- the nrecalc cadence of per-object stamps, entirely on the device (sm_100a).
//
// The classic pipeline (imsim/lsst_image.py:342-389, imsim/stamp.py:562-572) draws every object on its own stamp
// image with SiliconSensor.accumulate: brighter-fatter sees only that object's charge and the pixel boundaries of
// the stamp are recomputed every nrecalc (10^4) electrons (config/imsim-config.yaml:230-235).  A chunk of 10^4
// photons is a microsecond of device work, so driving that loop from the host -- chunk finder, deposit, slow list,
// three update kernels, a synchronisation for the chunk bounds -- costs ~100 launches per bright star.
//
// Here one launch handles a whole list of stamps.  A thread block takes a stamp (jobs are handed out through an
// atomic counter, heaviest first), builds the stamp's boundary state in its slice of a device arena and runs
// Silicon::accumulate's loop itself: a block-wide prefix sum of the photon fluxes finds the photon at which the
// charge since the last update reaches nrecalc, the photons up to there are deposited (fast path, then the listed
// slow ones), the boundary points within reach of the new charge are moved, bounding boxes refreshed, the charge
// folded into the stamp image -- all between __syncthreads(), no host in the loop.  Stamps never share state, so
// blocks never wait for each other.  At the end the stamp is added to the full image
// (``full_image[bounds] += stamp[bounds]``, lsst_image.py:359-368).
//
// Arithmetic per photon and per boundary point is the code of sensor.cu (sensor_fast_path, slow_photon, the
// update of k_update_distortions / k_update_bounds), so a stamp comes out bit-identical to
// b2_sensor_bind_image + b2_sensor_accumulate on the same photons (tests/test_gpu_stamps.py).
#include "sensor_device.cuh"

#include <algorithm>
#include <numeric>
#include <vector>

#define ST_THREADS 512
#define ST_PER 4
#define ST_TILE (ST_THREADS * ST_PER)  // photons per pass

struct StampSlot {  // byte offsets of one stamp's state inside the arena
    size_t H, V, inner, outer, delta, target, changed;
};

struct StampPhotons {
    const double *x, *y, *dxdz, *dydz, *wl, *flux, *rand4;
    int64_t ntot;
    uint64_t seed, offset;
};

struct FullImage {
    void* pix;
    int xmin, ymin, nx, ny, dtype_bytes;
};

template <typename T>
__device__ __forceinline__ void full_add(const FullImage& f, int ix, int iy, double v) {
    int ax = ix - f.xmin, ay = iy - f.ymin;
    if (ax >= 0 && ax < f.nx && ay >= 0 && ay < f.ny) atomicAdd(reinterpret_cast<T*>(f.pix) + (size_t)ay * f.nx + ax, (T)v);
}

// Silicon::updatePixelBounds of pixel (x, y) -- the body of k_update_bounds
template <int NV>
__device__ __forceinline__ void stamp_bounds_pixel(const DevSensor& s, int x, int y) {
    double oxmin = INFINITY, oxmax = -INFINITY, oymin = INFINITY, oymax = -INFINITY;
    walk_polygon<NV>(s, x, y, [&](double px, double py, double, double) {
        oxmin = fmin(oxmin, px);
        oxmax = fmax(oxmax, px);
        oymin = fmin(oymin, py);
        oymax = fmax(oymax, py);
    });
    double cx = (oxmin + oxmax) / 2.0, cy = (oymin + oymax) / 2.0;
    double ixmin = oxmin, ixmax = oxmax, iymin = oymin, iymax = oymax;
    walk_polygon<NV>(s, x, y, [&](double px, double py, double, double) {
        if (px - cx >= fabs(py - cy) && px < ixmax) ixmax = px;
        if (px - cx <= -fabs(py - cy) && px > ixmin) ixmin = px;
        if (py - cy >= fabs(px - cx) && py < iymax) iymax = py;
        if (py - cy <= -fabs(px - cx) && py > iymin) iymin = py;
    });
    size_t pix = (size_t)y * s.nx + x;
    *reinterpret_cast<double4*>(s.outer + pix * 4) = make_double4(oxmin, oxmax, oymin, oymax);
    *reinterpret_cast<double4*>(s.inner + pix * 4) = make_double4(ixmin, ixmax, iymin, iymax);
}

// Silicon::updatePixelDistortions for the boundary slot (x, y): the body of k_update_distortions, charge = delta
template <int NV>
__device__ __forceinline__ void stamp_update_slot(const DevSensor& s, const float2* __restrict__ KH,
                                                  const float2* __restrict__ KV, uint8_t* __restrict__ changed, int x,
                                                  int y) {
    const int q = s.qdist, nx = s.nx, ny = s.ny;
    const int cxk = (s.nx9 - 1) / 2, cyk = (s.ny9 - 1) / 2;
    const double* __restrict__ charge = s.delta;
    if (x < nx) {
        int i1 = max(x - q, 0), i2 = min(x + q, nx - 1);
        int j1 = max(y - (q + 1), 0), j2 = min(y + q, ny - 1);
        float2* h = s.H + Hidx(s, x, y);
        bool change = false;
        for (int j = j1; j <= j2; ++j)
            for (int i = i1; i <= i2; ++i) {
                double c = charge[(size_t)j * nx + i];
                if (c == 0.0) continue;
                change = true;
                const float2* kh = KH + ((y - j + cyk) * s.nx9 + (x - i + cxk)) * (NV + 2);
#pragma unroll
                for (int k = 0; k <= NV + 1; ++k) {
                    float2 p = h[k];
                    float2 d = kh[k];
                    p.x = (float)__dadd_rn((double)p.x, __dmul_rn((double)d.x, c));
                    p.y = (float)__dadd_rn((double)p.y, __dmul_rn((double)d.y, c));
                    h[k] = p;
                }
            }
        if (change) {
            if (y < ny) changed[(size_t)y * nx + x] = 1;
            if (y > 0) changed[(size_t)(y - 1) * nx + x] = 1;
        }
    }
    if (y < ny) {
        int i1 = max(x - (q + 1), 0), i2 = min(x + q, nx - 1);
        int j1 = max(y - q, 0), j2 = min(y + q, ny - 1);
        float2* v = s.V + Vidx(s, x, y);
        bool change = false;
        for (int j = j1; j <= j2; ++j)
            for (int i = i1; i <= i2; ++i) {
                double c = charge[(size_t)j * nx + i];
                if (c == 0.0) continue;
                change = true;
                const float2* kv = KV + ((y - j + cyk) * s.nx9 + (x - i + cxk)) * NV;
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    float2 p = v[k];
                    float2 d = kv[k];
                    p.x = (float)__dadd_rn((double)p.x, __dmul_rn((double)d.x, c));
                    p.y = (float)__dadd_rn((double)p.y, __dmul_rn((double)d.y, c));
                    v[k] = p;
                }
            }
        if (change) {
            if (x < nx) changed[(size_t)y * nx + x] = 1;
            if (x > 0) changed[(size_t)y * nx + x - 1] = 1;
        }
    }
}

template <typename T>
__device__ __forceinline__ void stamp_fold_delta(const DevSensor& s, int x, int y) {
    size_t i = (size_t)y * s.nx + x;
    double d = s.delta[i];
    if (d != 0.0) {
        T* t = reinterpret_cast<T*>(s.target);
        t[i] = (T)__dadd_rn((double)t[i], d);
        s.delta[i] = 0.0;
    }
}

template <int NV>
__global__ void __launch_bounds__(ST_THREADS, 2)
k_stamp_jobs(const __grid_constant__ DevSensor base, const B2StampJob* __restrict__ jobs, const StampSlot* __restrict__ slots,
             const int* __restrict__ order, int njobs, int* __restrict__ next, unsigned char* __restrict__ arena,
             const __grid_constant__ StampPhotons ph, double nrecalc, int ocx, int ocy, const __grid_constant__ FullImage full,
             unsigned long long* __restrict__ stats, double* __restrict__ added_total, double* __restrict__ added_job,
             SlowRec* __restrict__ slow_scratch) {
    extern __shared__ float2 sK[];  // KH then KV
    __shared__ DevSensor s;
    __shared__ int sh_job, sh_cut, pend[4];  // pend: box of the pixels holding charge since the last update
    __shared__ unsigned sh_nslow, sh_nupd;
    __shared__ double sh_warp[ST_THREADS / 32], sh_tile_sum;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int nKH = base.nx9 * base.ny9 * (NV + 2), nKV = base.nx9 * base.ny9 * NV;
    for (int k = tid; k < nKH; k += ST_THREADS) sK[k] = base.KH[k];
    for (int k = tid; k < nKV; k += ST_THREADS) sK[nKH + k] = base.KV[k];
    const float2* KH = sK;
    const float2* KV = sK + nKH;
    SlowRec* slow = slow_scratch + (size_t)blockIdx.x * ST_TILE;
    unsigned npoly = 0, nneigh = 0, nnf = 0, nb9 = 0, ndrop = 0;
    const bool f32 = full.dtype_bytes == 4;
    if (tid == 0) sh_nupd = 0;

    for (;;) {
        __syncthreads();
        if (tid == 0) sh_job = atomicAdd(next, 1);
        __syncthreads();
        if (sh_job >= njobs) break;
        const int jid = order[sh_job];
        const B2StampJob job = jobs[jid];
        const int64_t p0 = job.p0, p1 = job.p0 + job.n;
        double my_added = 0.0;
        if (job.plain) {
            // galsim.Sensor (faint objects, imsim/stamp.py:534-537): photons binned on the stamp, no silicon
            for (int64_t i = p0 + tid; i < p1; i += ST_THREADS) {
                int ix = (int)floor(ph.x[i] + 0.5), iy = (int)floor(ph.y[i] + 0.5);
                if (ix >= job.xmin && ix < job.xmin + job.nx && iy >= job.ymin && iy < job.ymin + job.ny) {
                    double f = ph.flux[i];
                    my_added += f;
                    if (f32) full_add<float>(full, ix, iy, f);
                    else full_add<double>(full, ix, iy, f);
                }
            }
        } else {
            // ---- bind a zero stamp and build its undistorted + tree-ring boundaries (Silicon::initialize)
            if (tid == 0) {
                s = base;
                const StampSlot sl = slots[jid];
                s.xmin = job.xmin; s.ymin = job.ymin; s.nx = job.nx; s.ny = job.ny;
                s.H = reinterpret_cast<float2*>(arena + sl.H);
                s.V = reinterpret_cast<float2*>(arena + sl.V);
                s.inner = reinterpret_cast<double*>(arena + sl.inner);
                s.outer = reinterpret_cast<double*>(arena + sl.outer);
                s.delta = reinterpret_cast<double*>(arena + sl.delta);
                s.target = arena + sl.target;
                s.dtype_bytes = full.dtype_bytes;
                pend[0] = pend[2] = 1 << 30;
                pend[1] = pend[3] = -1;
            }
            __syncthreads();
            uint8_t* changed = arena + slots[jid].changed;
            const int nx = s.nx, ny = s.ny;
            const bool tr = s.ntr > 2;
            for (int idx = tid; idx < (nx + 1) * (ny + 1); idx += ST_THREADS) {
                const int x = idx % (nx + 1), y = idx / (nx + 1);
                if (x < nx) {
                    float2* h = s.H + Hidx(s, x, y);
                    for (int k = 0; k <= NV + 1; ++k) {
                        float2 p;
                        p.x = (k == 0) ? 0.f : (k == NV + 1 ? 1.f : (float)s.frac[k - 1]);
                        p.y = 0.f;
                        if (tr) treering_point(s, p, s.xmin + x, s.ymin + y, ocx, ocy);
                        h[k] = p;
                    }
                }
                if (y < ny) {
                    float2* v = s.V + Vidx(s, x, y);
                    for (int k = 0; k < NV; ++k) {
                        float2 p;
                        p.x = 0.f;
                        p.y = (float)s.frac[k];
                        if (tr) treering_point(s, p, s.xmin + x, s.ymin + y, ocx, ocy);
                        v[k] = p;
                    }
                }
                if (x < nx && y < ny) {
                    size_t i = (size_t)y * nx + x;
                    s.delta[i] = 0.0;
                    changed[i] = 0;
                    if (f32) reinterpret_cast<float*>(s.target)[i] = 0.f;
                    else reinterpret_cast<double*>(s.target)[i] = 0.0;
                }
            }
            __syncthreads();
            for (int idx = tid; idx < nx * ny; idx += ST_THREADS) stamp_bounds_pixel<NV>(s, idx % nx, idx / nx);

            // ---- Silicon::accumulate with the boundary update every nrecalc electrons
            double accum = 0.0;  // flux since the last update (block-uniform)
            int64_t i0 = p0;
            while (i0 < p1) {
                const int64_t tend = (p1 - i0 < ST_TILE) ? p1 : i0 + ST_TILE;
                __syncthreads();  // the previous pass (or the set-up) is complete
                if (tid == 0) {
                    sh_cut = ST_TILE + 1;
                    sh_nslow = 0;
                }
                double f[ST_PER], run = 0.0, incl = 0.0;
                if (nrecalc > 0.0) {
                    // inclusive prefix sums of this pass's fluxes in photon order: thread t holds photons 4t .. 4t+3
#pragma unroll
                    for (int k = 0; k < ST_PER; ++k) {
                        int64_t i = i0 + (int64_t)tid * ST_PER + k;
                        f[k] = (i < tend) ? ph.flux[i] : 0.0;
                        run += f[k];
                    }
                    incl = run;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        double t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (lane >= o) incl += t;
                    }
                    if (lane == 31) sh_warp[wid] = incl;
                }
                __syncthreads();
                if (nrecalc > 0.0) {
                    double before = 0.0;
                    for (int w = 0; w < wid; ++w) before += sh_warp[w];
                    if (tid == ST_THREADS - 1) sh_tile_sum = before + incl;
                    double cum = accum + before + (incl - run);
#pragma unroll
                    for (int k = 0; k < ST_PER; ++k) {
                        cum += f[k];
                        int64_t i = i0 + (int64_t)tid * ST_PER + k;
                        if (i < tend && cum >= nrecalc) {
                            atomicMin(&sh_cut, tid * ST_PER + k);
                            break;
                        }
                    }
                }
                __syncthreads();
                const bool hit = sh_cut <= ST_TILE;
                const int64_t cut = hit ? i0 + sh_cut + 1 : tend;
                // ---- deposit photons [i0, cut): fast path, the rest to the block's list
                int bx0 = 1 << 30, bx1 = -1, by0 = 1 << 30, by1 = -1;
#pragma unroll 1
                for (int k = 0; k < ST_PER; ++k) {
                    int64_t i = i0 + tid + (int64_t)k * ST_THREADS;
                    bool to_slow = false;
                    SlowRec rec;
                    if (i < cut) {
                        double g1, g2, unf, udep;
                        if (ph.rand4) {
                            g1 = ph.rand4[i];
                            g2 = ph.rand4[ph.ntot + i];
                            unf = ph.rand4[2 * ph.ntot + i];
                            udep = ph.rand4[3 * ph.ntot + i];
                        } else {
                            sensor_draws(ph.seed, ph.offset + (uint64_t)i, g1, g2, unf, udep);
                        }
                        double a = 0.0, b = 0.0;
                        if (ph.dxdz) {
                            a = ph.dxdz[i];
                            b = ph.dydz[i];
                        }
                        double add = 0.0;
                        unsigned b9 = 0, dr = 0;
                        int dax, day;
                        to_slow = sensor_fast_path_ex(s, ph.x[i], ph.y[i], ph.dxdz != nullptr, a, b, ph.wl != nullptr,
                                                      ph.wl ? ph.wl[i] : 0.0, ph.flux[i], g1, g2, unf, udep, rec, add, b9,
                                                      dr, dax, day);
                        my_added += add;
                        nb9 += b9;
                        ndrop += dr;
                        if (dax >= 0) {
                            bx0 = min(bx0, dax); bx1 = max(bx1, dax);
                            by0 = min(by0, day); by1 = max(by1, day);
                        }
                    }
                    unsigned m = __ballot_sync(0xffffffffu, to_slow);
                    if (m) {
                        unsigned bs = 0;
                        if (lane == 0) bs = atomicAdd(&sh_nslow, (unsigned)__popc(m));
                        bs = __shfl_sync(0xffffffffu, bs, 0);
                        if (to_slow) slow[bs + __popc(m & ((1u << lane) - 1u))] = rec;
                    }
                }
                __syncthreads();
                for (unsigned j = tid; j < sh_nslow; j += ST_THREADS) {
                    int dax, day;
                    my_added += slow_photon<NV>(s, slow[j], npoly, nneigh, nnf, dax, day);
                    if (dax >= 0) {
                        bx0 = min(bx0, dax); bx1 = max(bx1, dax);
                        by0 = min(by0, day); by1 = max(by1, day);
                    }
                }
                if (bx1 >= 0) {
                    atomicMin(&pend[0], bx0); atomicMax(&pend[1], bx1);
                    atomicMin(&pend[2], by0); atomicMax(&pend[3], by1);
                }
                __syncthreads();
                if (hit) {
                    // ---- Silicon::update, restricted to the reach of the charge deposited since the last one
                    const int q = s.qdist;
                    if (pend[1] >= 0) {
                        const int sx0 = max(pend[0] - q, 0), sx1 = min(pend[1] + q + 1, nx);      // boundary slots
                        const int sy0 = max(pend[2] - q, 0), sy1 = min(pend[3] + q + 1, ny);
                        const int sw = sx1 - sx0 + 1, shh = sy1 - sy0 + 1;
                        for (int idx = tid; idx < sw * shh; idx += ST_THREADS)
                            stamp_update_slot<NV>(s, KH, KV, changed, sx0 + idx % sw, sy0 + idx / sw);
                        __syncthreads();
                        const int cx0 = max(sx0 - 1, 0), cx1 = min(sx1, nx - 1), cy0 = max(sy0 - 1, 0), cy1 = min(sy1, ny - 1);
                        const int cw = cx1 - cx0 + 1, chh = cy1 - cy0 + 1;
                        for (int idx = tid; idx < cw * chh; idx += ST_THREADS) {
                            const int x = cx0 + idx % cw, y = cy0 + idx / cw;
                            const size_t pix = (size_t)y * nx + x;
                            if (changed[pix]) {
                                changed[pix] = 0;
                                stamp_bounds_pixel<NV>(s, x, y);
                            }
                            if (x >= pend[0] && x <= pend[1] && y >= pend[2] && y <= pend[3]) {
                                if (f32) stamp_fold_delta<float>(s, x, y);
                                else stamp_fold_delta<double>(s, x, y);
                            }
                        }
                        __syncthreads();
                    }
                    if (tid == 0) {
                        pend[0] = pend[2] = 1 << 30;
                        pend[1] = pend[3] = -1;
                        sh_nupd++;
                    }
                    accum = 0.0;
                } else if (nrecalc > 0.0) {
                    accum += sh_tile_sum;
                }
                i0 = cut;
            }
            __syncthreads();
            // ---- Silicon::addDelta, then full_image[bounds] += stamp[bounds]
            for (int idx = tid; idx < nx * ny; idx += ST_THREADS) {
                const int x = idx % nx, y = idx / nx;
                double v;
                if (f32) {
                    stamp_fold_delta<float>(s, x, y);
                    v = (double)reinterpret_cast<float*>(s.target)[idx];
                    if (v != 0.0) full_add<float>(full, s.xmin + x, s.ymin + y, v);
                } else {
                    stamp_fold_delta<double>(s, x, y);
                    v = reinterpret_cast<double*>(s.target)[idx];
                    if (v != 0.0) full_add<double>(full, s.xmin + x, s.ymin + y, v);
                }
            }
        }
        // flux that landed on this stamp
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) my_added += __shfl_xor_sync(0xffffffffu, my_added, o);
        __syncthreads();
        if (lane == 0) sh_warp[wid] = my_added;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < ST_THREADS / 32; ++w) t += sh_warp[w];
            if (added_job) added_job[jid] = t;
            if (t != 0.0) atomicAdd(added_total, t);
        }
    }
    unsigned long long w0 = warp_sum(npoly), w1 = warp_sum(nneigh), w2 = warp_sum(nnf), w3 = warp_sum(nb9), w4 = warp_sum(ndrop);
    if (lane == 0) {
        if (w0) atomicAdd(&stats[ST_POLY], w0);
        if (w1) atomicAdd(&stats[ST_NEIGH], w1);
        if (w2) atomicAdd(&stats[ST_NOTFOUND], w2);
        if (w3) atomicAdd(&stats[ST_B9], w3);
        if (w4) atomicAdd(&stats[ST_DROP], w4);
    }
    __syncthreads();
    if (tid == 0 && sh_nupd) atomicAdd(&stats[ST_N - 1], (unsigned long long)sh_nupd);
}

// ------------------------------------------------------------------ host side
static inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

static size_t stamp_state_bytes(int nx, int ny, int nv, int dtype_bytes, StampSlot* sl) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += up256(bytes);
        return o;
    };
    const size_t npix = (size_t)nx * ny;
    StampSlot t;
    t.H = take((size_t)(ny + 1) * nx * (nv + 2) * sizeof(float2));
    t.V = take(((size_t)ny * (nx + 1) * nv + nv) * sizeof(float2));
    t.inner = take(npix * 4 * sizeof(double));
    t.outer = take(npix * 4 * sizeof(double));
    t.delta = take(npix * sizeof(double));
    t.target = take(npix * dtype_bytes);
    t.changed = take(npix);
    if (sl) *sl = t;
    return off;
}

// galsim.SiliconSensor.accumulate for a list of objects, each on its own zero stamp (fresh boundaries, the
// sensor's nrecalc cadence inside the stamp), followed by full_image[bounds] += stamp[bounds]:
// the object loop of imsim/lsst_image.py:342-389 with imsim/stamp.py:562-572 inside, one launch per arena load.
extern "C" int b2_sensor_accumulate_stamps(b2_sensor* s, int32_t njobs, const B2StampJob* jobs, int64_t n,
                                           const double* x, const double* y, const double* dxdz, const double* dydz,
                                           const double* wl, const double* flux, const double* rand4, uint64_t seed,
                                           uint64_t offset, int32_t ocx, int32_t ocy, void* full_pixels,
                                           int32_t full_xmin, int32_t full_ymin, int32_t full_nx, int32_t full_ny,
                                           int32_t dtype_bytes, B2AccumStats* stats, double* added_per_job) {
    B2_REQUIRE(s && jobs && njobs >= 0, "b2_sensor_accumulate_stamps: null argument");
    B2_REQUIRE(n == 0 || (x && y && flux), "b2_sensor_accumulate_stamps: null photon array (device pointers expected)");
    B2_REQUIRE((dxdz == nullptr) == (dydz == nullptr), "b2_sensor_accumulate_stamps: dxdz and dydz go together");
    B2_REQUIRE(!wl || s->d.nabs > 0, "b2_sensor_accumulate_stamps: wavelengths given but the sensor has no absorption table");
    B2_REQUIRE(full_pixels && full_nx > 0 && full_ny > 0 && (dtype_bytes == 4 || dtype_bytes == 8),
               "b2_sensor_accumulate_stamps: the full image must be a float32 / float64 device array");
    B2_REQUIRE(s->d.nv == 4 || s->d.nv == 8, "b2_sensor_accumulate_stamps: sensor models with 4 or 8 vertices per edge");
    b2_ctx* ctx = s->ctx;
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    if (stats) memset(stats, 0, sizeof(*stats));
    if (njobs == 0) return 0;
    const int nv = s->d.nv;
    std::vector<StampSlot> slots(njobs);
    std::vector<size_t> need(njobs, 0);
    for (int j = 0; j < njobs; ++j) {
        const B2StampJob& jb = jobs[j];
        B2_REQUIRE(jb.n >= 0 && jb.p0 >= 0 && jb.p0 + jb.n <= n, "b2_sensor_accumulate_stamps: photon range outside the arrays");
        B2_REQUIRE(jb.nx > 0 && jb.ny > 0, "b2_sensor_accumulate_stamps: empty stamp");
        if (!jb.plain) need[j] = stamp_state_bytes(jb.nx, jb.ny, nv, dtype_bytes, &slots[j]);
    }
    // heaviest first (photons, plus the set-up of the stamp's boundary state)
    std::vector<int> order(njobs);
    std::iota(order.begin(), order.end(), 0);
    auto cost = [&](int j) { return (double)jobs[j].n + (jobs[j].plain ? 0.0 : 0.25 * jobs[j].nx * (double)jobs[j].ny); };
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return cost(a) > cost(b); });
    size_t budget = (size_t)6144 << 20;
    if (const char* e = getenv("B2_STAMP_ARENA_MB")) budget = (size_t)std::max(64, atoi(e)) << 20;
    size_t biggest = 0;
    for (size_t v : need) biggest = std::max(biggest, v);
    budget = std::max(budget, biggest);
    // device copies of the job table, counters, per-block slow lists
    const int grid_max = 2 * s->sm_count;
    const size_t jobs_b = up256((size_t)njobs * sizeof(B2StampJob)), slots_b = up256((size_t)njobs * sizeof(StampSlot));
    const size_t order_b = up256((size_t)njobs * sizeof(int)), added_b = up256((size_t)njobs * sizeof(double));
    const size_t slow_b = up256((size_t)grid_max * ST_TILE * sizeof(SlowRec));
    if (b2_scratch_reserve(ctx, s->stamp_meta, jobs_b + slots_b + order_b + added_b + slow_b + 512)) return 1;
    unsigned char* m = (unsigned char*)s->stamp_meta.ptr;
    B2StampJob* djobs = (B2StampJob*)m;
    StampSlot* dslots = (StampSlot*)(m + jobs_b);
    int* dorder = (int*)(m + jobs_b + slots_b);
    double* dadded_job = (double*)(m + jobs_b + slots_b + order_b);
    SlowRec* dslow = (SlowRec*)(m + jobs_b + slots_b + order_b + added_b);
    int* dnext = (int*)(m + jobs_b + slots_b + order_b + added_b + slow_b);
    B2_CUDA(cudaMemsetAsync(s->dstats, 0, ST_N * sizeof(unsigned long long) + 64, st));
    B2_CUDA(cudaMemcpyAsync(djobs, jobs, (size_t)njobs * sizeof(B2StampJob), cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(dorder, order.data(), (size_t)njobs * sizeof(int), cudaMemcpyHostToDevice, st));
    const StampPhotons ph{x, y, dxdz, dydz, wl, flux, rand4, n, seed, offset};
    const FullImage full{full_pixels, full_xmin, full_ymin, full_nx, full_ny, dtype_bytes};
    const size_t smem = (size_t)s->d.nx9 * s->d.ny9 * (2 * nv + 2) * sizeof(float2);
    // waves: consecutive jobs of the sorted list whose stamp states fit the arena together
    int w0 = 0;
    bool slots_sent = false;
    std::vector<StampSlot> abs_slots(njobs);
    while (w0 < njobs) {
        size_t used = 0;
        int w1 = w0;
        while (w1 < njobs && (w1 == w0 || used + need[order[w1]] <= budget)) {
            const int j = order[w1];
            StampSlot t = slots[j];
            t.H += used; t.V += used; t.inner += used; t.outer += used; t.delta += used; t.target += used; t.changed += used;
            abs_slots[j] = t;
            used += need[j];
            ++w1;
        }
        if (b2_scratch_reserve(ctx, s->stamp_arena, std::max(used, (size_t)256))) return 1;
        // slot offsets of this wave (the table is rewritten per wave; the stream orders it after the previous launch)
        B2_CUDA(cudaMemcpyAsync(dslots, abs_slots.data(), (size_t)njobs * sizeof(StampSlot), cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaStreamSynchronize(st));  // abs_slots / pageable staging is reused by the next wave
        slots_sent = true;
        B2_CUDA(cudaMemsetAsync(dnext, 0, sizeof(int), st));
        const int nw = w1 - w0;
        const unsigned grid = (unsigned)std::min(nw, grid_max);
        {
            B2_TIMED("k_stamp_jobs", st);
            if (nv == 4) {
                if (smem > 48 * 1024) B2_CUDA(cudaFuncSetAttribute(k_stamp_jobs<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_stamp_jobs<4><<<grid, ST_THREADS, smem, st>>>(s->d, djobs, dslots, dorder + w0, nw, dnext,
                                                               (unsigned char*)s->stamp_arena.ptr, ph, s->cfg.nrecalc, ocx, ocy,
                                                               full, s->dstats, s->dadded, dadded_job, dslow);
            } else {
                if (smem > 48 * 1024) B2_CUDA(cudaFuncSetAttribute(k_stamp_jobs<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                k_stamp_jobs<8><<<grid, ST_THREADS, smem, st>>>(s->d, djobs, dslots, dorder + w0, nw, dnext,
                                                               (unsigned char*)s->stamp_arena.ptr, ph, s->cfg.nrecalc, ocx, ocy,
                                                               full, s->dstats, s->dadded, dadded_job, dslow);
            }
            B2_CHECK_LAUNCH();
        }
        w0 = w1;
    }
    (void)slots_sent;
    // the sensor's own bound image (if any) is untouched; a later accumulate(resume=True) on it stays valid
    if (stats || added_per_job) {
        unsigned long long h[ST_N];
        double added = 0.0;
        B2_CUDA(cudaMemcpyAsync(h, s->dstats, sizeof(h), cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaMemcpyAsync(&added, s->dadded, sizeof(double), cudaMemcpyDeviceToHost, st));
        if (added_per_job)
            B2_CUDA(cudaMemcpyAsync(added_per_job, dadded_job, (size_t)njobs * sizeof(double), cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
        if (stats) {
            stats->added_flux = added;
            stats->n_polygon_tests = h[ST_POLY];
            stats->n_neighbor_search = h[ST_NEIGH];
            stats->n_not_found = h[ST_NOTFOUND];
            stats->n_boundary_1e9 = h[ST_B9];
            stats->n_dropped_bottom = h[ST_DROP];
            stats->n_updates = h[ST_N - 1];
        }
    }
    return 0;
}
