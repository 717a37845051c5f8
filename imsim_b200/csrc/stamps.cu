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

#include <cooperative_groups.h>

#include <algorithm>
#include <numeric>
#include <vector>

#define ST_THREADS 512
#define ST_PER 4
#define ST_TILE (ST_THREADS * ST_PER)  // photons per pass
#define ST_QCAP 4096                  // owned slots / pixels queued per round of the boundary update
#define ST_PIXCAP 32768                // charged pixels listed per boundary update (beyond: box scan)

struct StampSlot {  // byte offsets of one stamp's state inside the arena
    size_t H, V, inner, outer, delta, target, changed, cbits;
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
    double ixmin = -INFINITY, ixmax = INFINITY, iymin = -INFINITY, iymax = INFINITY;
    walk_polygon<NV>(s, x, y, [&](double px, double py, double ex, double ey) {
        if (ex == 0.0 && px > ixmin) ixmin = px;
        if (ex == 1.0 && px < ixmax) ixmax = px;
        if (ey == 0.0 && py > iymin) iymin = py;
        if (ey == 1.0 && py < iymax) iymax = py;
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
                double c = __ldcg(charge + (size_t)j * nx + i);
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
                double c = __ldcg(charge + (size_t)j * nx + i);
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

// ---- the update driven by the list of charged pixels ----------------------------------------------------------
// Every listed pixel proposes the boundary slots (and, in a second pass, the pixels) within its reach.  A proposal
// is taken up by exactly one proposer without any atomics: all charged pixels that reach a slot see the same
// (2q+2) x (2q+2) window of the stamp's occupancy bitmap around it, and the first set bit of that window, in
// (row, column) order, is the owner.  The owner then gathers the window's charges in that same order -- the order
// of k_update_distortions, so the sums are bit-identical -- reading only the bitmap rows it already holds and the
// delta of pixels that do carry charge.  Bitmap and delta are written by atomics: they are read with __ldcg.
struct StampBits {  // occupancy bitmap of the stamp in flight: in shared memory when it fits, else in the arena
    unsigned* w;
    int wpr;      // words per pixel row
    bool global;  // arena copy: written by atomics, so read with __ldcg (L2), never through L1
    __device__ __forceinline__ unsigned word(int k) const { return global ? __ldcg(w + k) : w[k]; }
};

__device__ __forceinline__ unsigned row_bits(const StampBits& cb, int j, int i1, int i2) {
    // bits of row j for columns i1 .. i2 (i2 - i1 < 32), bit 0 = column i1
    const int w1 = i1 >> 5, w2 = i2 >> 5;
    unsigned long long v = cb.word(j * cb.wpr + w1);
    if (w2 != w1) v |= (unsigned long long)cb.word(j * cb.wpr + w2) << 32;
    v >>= (i1 & 31);
    const int n = i2 - i1 + 1;
    return (unsigned)v & (n >= 32 ? 0xffffffffu : ((1u << n) - 1u));
}

#define ST_SH(nv) ((nv) + 3)  // padded lengths [double2] of a table entry in shared memory: horizontal, vertical
#define ST_SV(nv) ((nv) + 1)
#define ST_QMAX 3  // the list path is written for qdist = 3 (GalSim's default, what imSim's models use): 8 x 8 windows

// q == 3: the window of slot (x, y) is 8 rows x 8 columns, one 64-bit word, bit 8 r + c = pixel (x - 4 + c, y - 4 + r)
__device__ __forceinline__ unsigned long long stamp_slot_window(const DevSensor& s, const StampBits& cb, int x, int y) {
    const int c0 = x - 4, r0 = y - 4;
    const int cl = max(c0, 0), ch = min(x + 3, s.nx - 1);
    unsigned long long W = 0ull;
    if (cl > ch) return W;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int j = r0 + r;
        if (j >= 0 && j < s.ny) W |= (unsigned long long)(row_bits(cb, j, cl, ch) << (cl - c0)) << (8 * r);
    }
    return W;
}

// is (pi, pj) the first charged pixel of the window (row, column order)?
__device__ __forceinline__ bool stamp_slot_owner(unsigned long long W, int x, int y, int pi, int pj) {
    if (!W) return false;
    const int ob = __ffsll((long long)W) - 1;
    return (y - 4 + (ob >> 3)) == pj && (x - 4 + (ob & 7)) == pi;
}

// the gather of slot (x, y): boundary points in registers, one delta load per charged pixel, one store at the end
template <int NV>
__device__ __forceinline__ void stamp_slot_apply(const DevSensor& s, const double2* __restrict__ KH,
                                                 const double2* __restrict__ KV, unsigned long long W, int x, int y) {
    const int nx = s.nx, ny = s.ny;
    const int c0 = x - 4, r0 = y - 4;
    const int cxk = (s.nx9 - 1) / 2, cyk = (s.ny9 - 1) / 2;
    const double* __restrict__ charge = s.delta;
    if (x < nx) {
        // horizontal boundary: all rows of the window, columns x-3 .. x+3 (not column c0)
        unsigned long long Wh = W & 0xfefefefefefefefeull;
        // points as float-valued doubles (bf_term, sensor_device.cuh): widened once, narrowed once
        float2* h = s.H + Hidx(s, x, y);
        double hx[NV + 2], hy[NV + 2];
#pragma unroll
        for (int k = 0; k <= NV + 1; ++k) {
            const float2 t = h[k];
            hx[k] = (double)t.x;
            hy[k] = (double)t.y;
        }
        bool any = false;
        while (Wh) {
            const int b = __ffsll((long long)Wh) - 1;
            Wh &= Wh - 1;
            const int j = r0 + (b >> 3), i = c0 + (b & 7);
            const double c = __ldcg(charge + (size_t)j * nx + i);
            if (c == 0.0) continue;
            any = true;
            const double2* kh = KH + ((y - j + cyk) * s.nx9 + (x - i + cxk)) * ST_SH(NV);
#pragma unroll
            for (int k = 0; k <= NV + 1; ++k) {
                const double2 d = kh[k];
                bf_term(hx[k], d.x, c);
                bf_term(hy[k], d.y, c);
            }
        }
        if (any) {
#pragma unroll
            for (int k = 0; k <= NV + 1; ++k) h[k] = make_float2((float)hx[k], (float)hy[k]);
        }
    }
    if (y < ny) {
        // vertical boundary: rows y-3 .. y+3 (not row r0), all columns of the window
        unsigned long long Wv = W & ~0xffull;
        float2* v = s.V + Vidx(s, x, y);
        double vx[NV], vy[NV];
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const float2 t = v[k];
            vx[k] = (double)t.x;
            vy[k] = (double)t.y;
        }
        bool any = false;
        while (Wv) {
            const int b = __ffsll((long long)Wv) - 1;
            Wv &= Wv - 1;
            const int j = r0 + (b >> 3), i = c0 + (b & 7);
            const double c = __ldcg(charge + (size_t)j * nx + i);
            if (c == 0.0) continue;
            any = true;
            const double2* kv = KV + ((y - j + cyk) * s.nx9 + (x - i + cxk)) * ST_SV(NV);
#pragma unroll
            for (int k = 0; k < NV; ++k) {
                const double2 d = kv[k];
                bf_term(vx[k], d.x, c);
                bf_term(vy[k], d.y, c);
            }
        }
        if (any) {
#pragma unroll
            for (int k = 0; k < NV; ++k) v[k] = make_float2((float)vx[k], (float)vy[k]);
        }
    }
}

// second pass: is (pi, pj) the first charged pixel within q+1 of pixel (x, y)?  (then it refreshes that pixel's boxes)
__device__ __forceinline__ bool stamp_owns_pixel(const DevSensor& s, const StampBits& cb, int x, int y, int pi, int pj) {
    const int q = s.qdist;
    const int cl = max(x - q - 1, 0), ch = min(x + q + 1, s.nx - 1);
    const int jl = max(y - q - 1, 0), jh = min(y + q + 1, s.ny - 1);
    for (int j = jl; j <= jh; ++j) {
        const unsigned mm = row_bits(cb, j, cl, ch);
        if (mm) return j == pj && cl + __ffs(mm) - 1 == pi;
    }
    return false;
}

template <typename T>
__device__ __forceinline__ void stamp_fold_delta(const DevSensor& s, int x, int y) {
    size_t i = (size_t)y * s.nx + x;
    double d = __ldcg(s.delta + i);  // written by atomics: read where they live (L2), not through L1
    if (d != 0.0) {
        T* t = reinterpret_cast<T*>(s.target);
        t[i] = (T)__dadd_rn((double)t[i], d);
        s.delta[i] = 0.0;
    }
}

// A stamp is worked on by a team: one thread block, or -- for the few stamps that hold most of the photons, whose
// update loop would otherwise keep one SM busy long after the others have finished -- a thread-block cluster of CS
// blocks on CS SMs.  The boundary state lives in the arena (global memory) either way; the members of a cluster
// split photons, charged pixels and pixels by index, keep their own lists and queues, read each other's few control
// words through distributed shared memory and meet at cluster barriers (release / acquire, so plain loads of
// boundary points another member has moved are safe afterwards).
namespace cg = cooperative_groups;

template <int CS>
struct Team {
    __device__ static __forceinline__ unsigned rank() { return cg::this_cluster().block_rank(); }
    __device__ static __forceinline__ void sync() { cg::this_cluster().sync(); }
    template <typename T>
    __device__ static __forceinline__ T* peer(T* p, unsigned r) { return cg::this_cluster().map_shared_rank(p, r); }
};
template <>
struct Team<1> {
    __device__ static __forceinline__ unsigned rank() { return 0u; }
    __device__ static __forceinline__ void sync() { __syncthreads(); }
    template <typename T>
    __device__ static __forceinline__ T* peer(T* p, unsigned) { return p; }
};

template <int NV, int CS>
__global__ void __launch_bounds__(ST_THREADS, 2)
k_stamp_jobs(const __grid_constant__ DevSensor base, const B2StampJob* __restrict__ jobs, const StampSlot* __restrict__ slots,
             const int* __restrict__ order, int njobs, int* __restrict__ next, unsigned char* __restrict__ arena,
             const __grid_constant__ StampPhotons ph, double nrecalc, int ocx, int ocy, const __grid_constant__ FullImage full,
             unsigned long long* __restrict__ stats, double* __restrict__ added_total, double* __restrict__ added_job,
             SlowRec* __restrict__ slow_scratch, int* __restrict__ pix_scratch, int smem_bit_words,
             unsigned long long* __restrict__ prof) {
    extern __shared__ double2 sK[];  // KH then KV, widened to double (bf_term)
    __shared__ DevSensor s;
    __shared__ int sh_job, sh_cut, pend[4];  // pend: box of the pixels holding charge since the last update
    __shared__ unsigned sh_nslow, sh_nupd, sh_npix, sh_nq;
    __shared__ unsigned queue[ST_QCAP];  // sh_npix: length of the charged-pixel list (may exceed its capacity)
    __shared__ double sh_warp[ST_THREADS / 32], sh_tile_sum;  // sh_tile_sum: flux of this block's share of the pass
    using T = Team<CS>;
    constexpr int TT = ST_THREADS * CS;  // threads and photons per pass of the team
    constexpr int TTILE = ST_TILE * CS;
    const unsigned rank = T::rank();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int gtid = (int)rank * ST_THREADS + tid;
    // entries padded by one point: the lanes of a warp read different entries, which then spread over the banks
    const int nent = base.nx9 * base.ny9;
    const int nKH = nent * ST_SH(NV), nKV = nent * ST_SV(NV);
    for (int k = tid; k < nent * (NV + 2); k += ST_THREADS) sK[(k / (NV + 2)) * ST_SH(NV) + k % (NV + 2)] = base.KHd[k];
    for (int k = tid; k < nent * NV; k += ST_THREADS) sK[nKH + (k / NV) * ST_SV(NV) + k % NV] = base.KVd[k];
    const double2* KH = sK;
    const double2* KV = sK + nKH;
    SlowRec* slow = slow_scratch + (size_t)blockIdx.x * ST_TILE;
    int* pixlist = pix_scratch + (size_t)blockIdx.x * ST_PIXCAP;
    unsigned npoly = 0, nneigh = 0, nnf = 0, nb9 = 0, ndrop = 0;
    const bool f32 = full.dtype_bytes == 4;
    if (tid == 0) sh_nupd = 0;
    // B2_STAMP_PROFILE: cycles of thread 0 per phase, summed over the blocks (0 set-up, 1 cut search, 2 deposit, 3 slow
    // list, 4 boundary slots, 5 boxes, 6 fold, 7 final add)
    long long tlast = prof ? clock64() : 0;
#define ST_PROF(k)                                                   \
    if (prof && tid == 0) {                                          \
        const long long tnow = clock64();                            \
        atomicAdd(&prof[k], (unsigned long long)(tnow - tlast));     \
        tlast = tnow;                                                \
    }

    for (;;) {
        T::sync();
        if (gtid == 0) sh_job = atomicAdd(next, 1);
        T::sync();
        const int job_k = *T::peer(&sh_job, 0);
        if (job_k >= njobs) break;
        const int jid = order[job_k];
        const B2StampJob job = jobs[jid];
        const int64_t p0 = job.p0, p1 = job.p0 + job.n;
        double my_added = 0.0;
        if (job.plain) {
            // galsim.Sensor (faint objects, imsim/stamp.py:534-537): photons binned on the stamp, no silicon
            for (int64_t i = p0 + gtid; i < p1; i += TT) {
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
                sh_npix = 0;
            }
            __syncthreads();
            uint8_t* changed = arena + slots[jid].changed;
            StampBits cb;
            cb.wpr = (s.nx + 31) / 32;
            cb.global = CS > 1 || cb.wpr * s.ny > smem_bit_words;
            cb.w = cb.global ? reinterpret_cast<unsigned*>(arena + slots[jid].cbits)
                             : reinterpret_cast<unsigned*>(sK + nKH + nKV);
            unsigned* const cbits = cb.w;
            const int wpr = cb.wpr;
            const int nx = s.nx, ny = s.ny;
            const bool tr = s.ntr > 2;
            for (int idx = gtid; idx < (nx + 1) * (ny + 1); idx += TT) {
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
                    if ((x & 31) == 0) cbits[y * wpr + (x >> 5)] = 0u;
                    if (f32) reinterpret_cast<float*>(s.target)[i] = 0.f;
                    else reinterpret_cast<double*>(s.target)[i] = 0.0;
                }
            }
            T::sync();
            for (int idx = gtid; idx < nx * ny; idx += TT) stamp_bounds_pixel<NV>(s, idx % nx, idx / nx);
            T::sync();
            ST_PROF(0)

            // ---- Silicon::accumulate with the boundary update every nrecalc electrons
            double accum = 0.0;  // flux since the last update (block-uniform)
            int64_t i0 = p0;
            while (i0 < p1) {
                const int64_t tend = (p1 - i0 < TTILE) ? p1 : i0 + TTILE;
                T::sync();  // the previous pass (or the set-up) is complete, its control words have been read
                if (tid == 0) {
                    sh_cut = TTILE + 1;
                    sh_nslow = 0;
                }
                double f[ST_PER], run = 0.0, incl = 0.0;
                if (nrecalc > 0.0) {
                    // inclusive prefix sums of this pass's fluxes in photon order: thread t holds photons 4t .. 4t+3
#pragma unroll
                    for (int k = 0; k < ST_PER; ++k) {
                        int64_t i = i0 + (int64_t)gtid * ST_PER + k;
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
                double tile_sum = 0.0;
                if (nrecalc > 0.0) {
                    double before = 0.0;
                    for (int w = 0; w < wid; ++w) before += sh_warp[w];
                    if (tid == ST_THREADS - 1) sh_tile_sum = before + incl;
                    if (CS > 1) {
                        // the blocks ahead of this one in the pass
                        T::sync();
                        for (unsigned r = 0; r < (unsigned)CS; ++r) {
                            const double t = *T::peer(&sh_tile_sum, r);
                            if (r < rank) before += t;
                            tile_sum += t;
                        }
                    }
                    double cum = accum + before + (incl - run);
#pragma unroll
                    for (int k = 0; k < ST_PER; ++k) {
                        cum += f[k];
                        int64_t i = i0 + (int64_t)gtid * ST_PER + k;
                        if (i < tend && cum >= nrecalc) {
                            atomicMin(&sh_cut, gtid * ST_PER + k);
                            break;
                        }
                    }
                }
                T::sync();
                int cut_at = sh_cut;
                if (CS > 1) {
                    for (unsigned r = 0; r < (unsigned)CS; ++r) cut_at = min(cut_at, *T::peer(&sh_cut, r));
                } else {
                    tile_sum = sh_tile_sum;
                }
                const bool hit = cut_at <= TTILE;
                ST_PROF(1)
                const int64_t cut = hit ? i0 + cut_at + 1 : tend;
                // ---- deposit photons [i0, cut): fast path, the rest to the block's list
                int bx0 = 1 << 30, bx1 = -1, by0 = 1 << 30, by1 = -1;
#pragma unroll 1
                for (int k = 0; k < ST_PER; ++k) {
                    int64_t i = i0 + gtid + (int64_t)k * TT;
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
                            const int pix = day * nx + dax;
                            if (nrecalc > 0.0) {
                                const unsigned bit = 1u << (dax & 31);
                                if (!(atomicOr(&cbits[day * wpr + (dax >> 5)], bit) & bit)) {  // first charge since the update
                                    unsigned at = atomicAdd(&sh_npix, 1u);
                                    if (at < ST_PIXCAP) pixlist[at] = pix;
                                }
                            }
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
                ST_PROF(2)
                for (unsigned j = tid; j < sh_nslow; j += ST_THREADS) {
                    int dax, day;
                    my_added += slow_photon<NV>(s, slow[j], npoly, nneigh, nnf, dax, day);
                    if (dax >= 0) {
                        bx0 = min(bx0, dax); bx1 = max(bx1, dax);
                        by0 = min(by0, day); by1 = max(by1, day);
                        const int pix = day * nx + dax;
                        if (nrecalc > 0.0) {
                            const unsigned bit = 1u << (dax & 31);
                            if (!(atomicOr(&cbits[day * wpr + (dax >> 5)], bit) & bit)) {
                                unsigned at = atomicAdd(&sh_npix, 1u);
                                if (at < ST_PIXCAP) pixlist[at] = pix;
                            }
                        }
                    }
                }
                if (bx1 >= 0) {
                    atomicMin(&pend[0], bx0); atomicMax(&pend[1], bx1);
                    atomicMin(&pend[2], by0); atomicMax(&pend[3], by1);
                }
                T::sync();
                ST_PROF(3)
                if (hit) {
                    // ---- Silicon::update, restricted to the reach of the charge deposited since the last one
                    const int q = s.qdist;
                    const unsigned npix = sh_npix;  // this block's list; the team's lists together hold every charged pixel
                    unsigned npix_team = npix;
                    bool listed = npix <= ST_PIXCAP;
                    int pb[4] = {pend[0], pend[1], pend[2], pend[3]};
                    if (CS > 1) {
                        npix_team = 0;
                        for (unsigned r = 0; r < (unsigned)CS; ++r) {
                            const unsigned m = *T::peer(&sh_npix, r);
                            npix_team += m;
                            listed = listed && m <= ST_PIXCAP;
                            const int* pr = T::peer(&pend[0], r);
                            pb[0] = min(pb[0], pr[0]); pb[1] = max(pb[1], pr[1]);
                            pb[2] = min(pb[2], pr[2]); pb[3] = max(pb[3], pr[3]);
                        }
                    }
                    if (npix_team > 0 && listed && q == ST_QMAX && nx < 65536 && ny < 65536) {
                        // the usual case: work proportional to the number of charged pixels (see
                        // stamp_update_slot_owned): boundary slots within reach first, then the boxes of the pixels
                        // Owners are few and scattered among the proposals (one in ~50 in a star's core), so the
                        // proposals of a round are only judged; the owned slots / pixels go to a queue in shared memory
                        // and are then worked off with every lane busy.
                        // A charged pixel with no other charge within 8 pixels (most of a star's wings) owns every
                        // slot and pixel it proposes: marked once here (top bit of its list entry), it skips the windows.
                        for (unsigned k = tid; k < npix; k += ST_THREADS) {
                            const int pix = pixlist[k];
                            const int pi = pix % nx, pj = pix / nx;
                            const int cl = max(pi - 8, 0), ch = min(pi + 8, nx - 1);
                            bool alone = true;
                            for (int j = max(pj - 8, 0); j <= min(pj + 8, ny - 1) && alone; ++j) {
                                unsigned mrow = row_bits(cb, j, cl, ch);
                                if (j == pj) mrow &= ~(1u << (pi - cl));
                                alone = (mrow == 0u);
                            }
                            if (alone) pixlist[k] = pix | (int)0x80000000;
                        }
                        __syncthreads();
                        for (int pass = 0; pass < 2; ++pass) {
                            const int B = pass == 0 ? 8 : 9, BB = B * B;     // slots within reach; pixels next to those
                            const unsigned total = npix * (unsigned)BB;
                            unsigned idx0 = 0;
                            while (idx0 < total) {
                                if (tid == 0) sh_nq = 0;
                                __syncthreads();
                                // fill: judge one proposal per thread and round until the queue could overflow
                                unsigned nq = 0;
                                while (idx0 < total && nq <= ST_QCAP - ST_THREADS) {
                                    const unsigned idx = idx0 + tid;
                                    if (idx < total) {
                                        const int raw = pixlist[idx / BB], c = (int)(idx % BB);
                                        const int pix = raw & 0x7fffffff;
                                        const bool alone = raw < 0;
                                        const int pi = pix % nx, pj = pix / nx;
                                        const int x = pi - (pass == 0 ? 3 : 4) + c % B, y = pj - (pass == 0 ? 3 : 4) + c / B;
                                        bool own = false;
                                        if (pass == 0) {
                                            if (x >= 0 && y >= 0 && x <= nx && y <= ny)
                                                own = alone || stamp_slot_owner(stamp_slot_window(s, cb, x, y), x, y, pi, pj);
                                        } else if (x >= 0 && y >= 0 && x < nx && y < ny) {
                                            own = alone || stamp_owns_pixel(s, cb, x, y, pi, pj);
                                        }
                                        if (own) queue[atomicAdd(&sh_nq, 1u)] = ((unsigned)y << 16) | (unsigned)x;
                                    }
                                    idx0 += ST_THREADS;
                                    __syncthreads();
                                    nq = sh_nq;
                                }
                                // drain: every lane takes owned slots / pixels
                                for (unsigned k = tid; k < nq; k += ST_THREADS) {
                                    const int x = (int)(queue[k] & 0xffffu), y = (int)(queue[k] >> 16);
                                    if (pass == 0) stamp_slot_apply<NV>(s, KH, KV, stamp_slot_window(s, cb, x, y), x, y);
                                    else stamp_bounds_pixel<NV>(s, x, y);
                                }
                                __syncthreads();
                            }
                            T::sync();  // every block's slots are in place before the boxes, the boxes before the fold
                            ST_PROF(4 + pass)
                        }
                        for (unsigned k = tid; k < npix; k += ST_THREADS) {
                            const int pix = pixlist[k] & 0x7fffffff;
                            const int x = pix % nx, y = pix / nx;
                            if (f32) stamp_fold_delta<float>(s, x, y);
                            else stamp_fold_delta<double>(s, x, y);
                            atomicAnd(&cbits[y * wpr + (x >> 5)], ~(1u << (x & 31)));
                        }
                        T::sync();
                        ST_PROF(6)
                    } else if (pb[1] >= 0) {
                        // more charged pixels than the list holds: scan the box of the pending charge
                        const int sx0 = max(pb[0] - q, 0), sx1 = min(pb[1] + q + 1, nx);      // boundary slots
                        const int sy0 = max(pb[2] - q, 0), sy1 = min(pb[3] + q + 1, ny);
                        const int sw = sx1 - sx0 + 1, shh = sy1 - sy0 + 1;
                        for (int idx = gtid; idx < sw * shh; idx += TT)
                            stamp_update_slot<NV>(s, s.KH, s.KV, changed, sx0 + idx % sw, sy0 + idx / sw);
                        T::sync();
                        const int cx0 = max(sx0 - 1, 0), cx1 = min(sx1, nx - 1), cy0 = max(sy0 - 1, 0), cy1 = min(sy1, ny - 1);
                        const int cw = cx1 - cx0 + 1, chh = cy1 - cy0 + 1;
                        for (int idx = gtid; idx < cw * chh; idx += TT) {
                            const int x = cx0 + idx % cw, y = cy0 + idx / cw;
                            const size_t pix = (size_t)y * nx + x;
                            if (CS > 1 ? __ldcg(changed + pix) : changed[pix]) {
                                changed[pix] = 0;
                                stamp_bounds_pixel<NV>(s, x, y);
                            }
                            if (x >= pb[0] && x <= pb[1] && y >= pb[2] && y <= pb[3]) {
                                if (f32) stamp_fold_delta<float>(s, x, y);
                                else stamp_fold_delta<double>(s, x, y);
                            }
                        }
                        for (int idx = gtid; idx < ny * wpr; idx += TT) cbits[idx] = 0u;
                        T::sync();
                    }
                    if (tid == 0) {
                        pend[0] = pend[2] = 1 << 30;
                        pend[1] = pend[3] = -1;
                        sh_npix = 0;
                        if (rank == 0) sh_nupd++;
                    }
                    accum = 0.0;
                } else if (nrecalc > 0.0) {
                    accum += tile_sum;
                }
                i0 = cut;
            }
            T::sync();
            // ---- Silicon::addDelta, then full_image[bounds] += stamp[bounds]
            for (int idx = gtid; idx < nx * ny; idx += TT) {
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
        __syncthreads();
        ST_PROF(7)
        // flux that landed on this stamp
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) my_added += __shfl_xor_sync(0xffffffffu, my_added, o);
        __syncthreads();
        if (lane == 0) sh_warp[wid] = my_added;
        __syncthreads();
        if (tid == 0) {
            double t = 0.0;
            for (int w = 0; w < ST_THREADS / 32; ++w) t += sh_warp[w];
            if (added_job) {
                if (CS > 1) atomicAdd(&added_job[jid], t);  // zeroed by the host
                else added_job[jid] = t;
            }
            if (t != 0.0) atomicAdd(added_total, t);
        }
    }
    if (CS > 1) T::sync();  // nobody leaves while a team mate may still read its shared memory
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
    t.cbits = take((size_t)ny * ((nx + 31) / 32) * sizeof(unsigned));  // one bit per pixel: charge since the last update
    if (sl) *sl = t;
    return off;
}

struct StampLaunch {
    DevSensor d;
    const B2StampJob* jobs;
    const StampSlot* slots;
    unsigned char* arena;
    StampPhotons ph;
    double nrecalc;
    int ocx, ocy;
    FullImage full;
    unsigned long long* stats;
    double *added_total, *added_job;
    int bit_words;
    unsigned long long* prof;
};

// one launch of k_stamp_jobs<NV, CS> over ``nw`` entries of the order table; CS > 1: thread-block clusters
template <int NV, int CS>
static int stamp_launch(const StampLaunch& a, const int* order, int nw, int* next, SlowRec* slow, int* pix, int teams,
                        size_t smem, cudaStream_t st) {
    auto kern = k_stamp_jobs<NV, CS>;
    B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(teams * CS));
    cfg.blockDim = dim3(ST_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CS > 1 ? 1 : 0;
    B2_CUDA(cudaLaunchKernelEx(&cfg, kern, a.d, a.jobs, a.slots, order, nw, next, a.arena, a.ph, a.nrecalc, a.ocx, a.ocy,
                               a.full, a.stats, a.added_total, a.added_job, slow, pix, a.bit_words, a.prof));
    return 0;
}

template <int CS>
static int stamp_launch_nv(int nv, const StampLaunch& a, const int* order, int nw, int* next, SlowRec* slow, int* pix,
                           int teams, size_t smem, cudaStream_t st) {
    return nv == 4 ? stamp_launch<4, CS>(a, order, nw, next, slow, pix, teams, smem, st)
                   : stamp_launch<8, CS>(a, order, nw, next, slow, pix, teams, smem, st);
}

// how many clusters of CS blocks the device runs at once
template <int CS>
static int stamp_max_clusters(int nv, size_t smem) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CS * 64);
    cfg.blockDim = dim3(ST_THREADS);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e;
    if (nv == 4) {
        cudaFuncSetAttribute(k_stamp_jobs<4, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        e = cudaOccupancyMaxActiveClusters(&n, k_stamp_jobs<4, CS>, &cfg);
    } else {
        cudaFuncSetAttribute(k_stamp_jobs<8, CS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        e = cudaOccupancyMaxActiveClusters(&n, k_stamp_jobs<8, CS>, &cfg);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
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
    B2_REQUIRE(s->ctx, "b2_sensor_accumulate_stamps: the context this sensor was created on has been destroyed");
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
    // per-block lists: the blocks of the one-block-per-stamp launch, then those of the cluster launch
    const size_t slow_b = up256((size_t)2 * grid_max * ST_TILE * sizeof(SlowRec));
    const size_t pix_b = up256((size_t)2 * grid_max * ST_PIXCAP * sizeof(int));
    if (b2_scratch_reserve(ctx, s->stamp_meta, jobs_b + slots_b + order_b + added_b + slow_b + pix_b + 512)) return 1;
    unsigned char* m = (unsigned char*)s->stamp_meta.ptr;
    B2StampJob* djobs = (B2StampJob*)m;
    StampSlot* dslots = (StampSlot*)(m + jobs_b);
    int* dorder = (int*)(m + jobs_b + slots_b);
    double* dadded_job = (double*)(m + jobs_b + slots_b + order_b);
    SlowRec* dslow = (SlowRec*)(m + jobs_b + slots_b + order_b + added_b);
    int* dpix = (int*)(m + jobs_b + slots_b + order_b + added_b + slow_b);
    int* dnext = (int*)(m + jobs_b + slots_b + order_b + added_b + slow_b + pix_b);
    unsigned long long* dprof = getenv("B2_STAMP_PROFILE") ? (unsigned long long*)(dnext + 16) : nullptr;
    if (dprof) B2_CUDA(cudaMemsetAsync(dprof, 0, 8 * sizeof(unsigned long long), st));
    B2_CUDA(cudaMemsetAsync(dadded_job, 0, (size_t)njobs * sizeof(double), st));
    // stamps that hold a large share of the photons go to clusters of ``cs`` blocks (B2_STAMP_CLUSTER = 1 turns that off)
    int cs = 8;
    if (const char* e = getenv("B2_STAMP_CLUSTER")) cs = atoi(e);
    cs = cs >= 8 ? 8 : (cs >= 4 ? 4 : 1);
    // "heavy": a stamp that alone would take more than half of what a block's fair share of the whole list takes
    // (1000 equal stars: none; a catalogue whose photons sit in a few stars: those), never below 5e4
    double total_cost = 0.0;
    for (int j = 0; j < njobs; ++j) total_cost += cost(j);
    double heavy_cost = std::max(5.0e4, 0.5 * total_cost / (2.0 * s->sm_count));
    if (const char* e = getenv("B2_STAMP_HEAVY")) heavy_cost = atof(e);
    B2_CUDA(cudaMemsetAsync(s->dstats, 0, ST_N * sizeof(unsigned long long) + 64, st));
    B2_CUDA(cudaMemcpyAsync(djobs, jobs, (size_t)njobs * sizeof(B2StampJob), cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(dorder, order.data(), (size_t)njobs * sizeof(int), cudaMemcpyHostToDevice, st));
    const StampPhotons ph{x, y, dxdz, dydz, wl, flux, rand4, n, seed, offset};
    const FullImage full{full_pixels, full_xmin, full_ymin, full_nx, full_ny, dtype_bytes};
    const size_t smem_k = (size_t)s->d.nx9 * s->d.ny9 * (ST_SH(nv) + ST_SV(nv)) * sizeof(double2);
    // occupancy bitmap of the stamp in flight in shared memory: two blocks per SM share ~220 KB
    size_t bit_words = 0;
    for (int j = 0; j < njobs; ++j)
        if (!jobs[j].plain) bit_words = std::max(bit_words, (size_t)((jobs[j].nx + 31) / 32) * jobs[j].ny);
    // (two blocks per SM: ~110 KB each, minus the tables and ~18 KB of static lists)
    bit_words = std::min(bit_words, ((size_t)92 * 1024 - smem_k) / sizeof(unsigned));
    const size_t smem = smem_k + bit_words * sizeof(unsigned);
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
            t.H += used; t.V += used; t.inner += used; t.outer += used; t.delta += used; t.target += used; t.changed += used; t.cbits += used;
            abs_slots[j] = t;
            used += need[j];
            ++w1;
        }
        if (b2_scratch_reserve(ctx, s->stamp_arena, std::max(used, (size_t)256))) return 1;
        // slot offsets of this wave (the table is rewritten per wave; the stream orders it after the previous launch)
        B2_CUDA(cudaMemcpyAsync(dslots, abs_slots.data(), (size_t)njobs * sizeof(StampSlot), cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaStreamSynchronize(st));  // abs_slots / pageable staging is reused by the next wave
        slots_sent = true;
        B2_CUDA(cudaMemsetAsync(dnext, 0, 8 * sizeof(int), st));
        // the sorted list starts with the heaviest stamps: those above the threshold are the cluster launch's
        int wh = w0;
        int max_clusters = 0;
        if (cs > 1) {
            while (wh < w1 && !jobs[order[wh]].plain && cost(order[wh]) >= heavy_cost) ++wh;
            if (wh > w0) {
                max_clusters = cs == 8 ? stamp_max_clusters<8>(nv, smem_k) : stamp_max_clusters<4>(nv, smem_k);
                max_clusters = std::min(max_clusters, grid_max / cs);
                if (max_clusters < 1) wh = w0;
            }
        }
        StampLaunch a{s->d, djobs, dslots, (unsigned char*)s->stamp_arena.ptr, ph, s->cfg.nrecalc, ocx, ocy, full,
                      s->dstats, s->dadded, dadded_job, (int)bit_words, dprof};
        {
            B2_TIMED("k_stamp_jobs", st);
            if (wh > w0) {
                if (!s->stamp_aux) {
                    int lo = 0, hi = 0;
                    B2_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                    B2_CUDA(cudaStreamCreateWithPriority(&s->stamp_aux, cudaStreamNonBlocking, hi));
                    B2_CUDA(cudaEventCreateWithFlags(&s->stamp_ev[0], cudaEventDisableTiming));
                    B2_CUDA(cudaEventCreateWithFlags(&s->stamp_ev[1], cudaEventDisableTiming));
                }
                B2_CUDA(cudaEventRecord(s->stamp_ev[0], st));
                B2_CUDA(cudaStreamWaitEvent(s->stamp_aux, s->stamp_ev[0], 0));
                const int teams = std::min(wh - w0, max_clusters);
                SlowRec* hslow = dslow + (size_t)grid_max * ST_TILE;
                int* hpix = dpix + (size_t)grid_max * ST_PIXCAP;
                int rc = cs == 8 ? stamp_launch_nv<8>(nv, a, dorder + w0, wh - w0, dnext + 4, hslow, hpix, teams, smem_k, s->stamp_aux)
                                 : stamp_launch_nv<4>(nv, a, dorder + w0, wh - w0, dnext + 4, hslow, hpix, teams, smem_k, s->stamp_aux);
                if (rc) return rc;
                B2_CUDA(cudaEventRecord(s->stamp_ev[1], s->stamp_aux));
            }
            if (w1 > wh) {
                const int nw = w1 - wh;
                if (stamp_launch_nv<1>(nv, a, dorder + wh, nw, dnext, dslow, dpix, std::min(nw, grid_max), smem, st)) return 1;
            }
            if (wh > w0) B2_CUDA(cudaStreamWaitEvent(st, s->stamp_ev[1], 0));
            B2_CHECK_LAUNCH();
        }
        w0 = w1;
    }
    (void)slots_sent;
    if (dprof) {
        unsigned long long hp[8];
        B2_CUDA(cudaMemcpyAsync(hp, dprof, sizeof(hp), cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
        fprintf(stderr, "k_stamp_jobs cycles of thread 0 summed over blocks: setup %llu cut %llu deposit %llu slow %llu slots %llu boxes %llu fold %llu final %llu\n",
                hp[0], hp[1], hp[2], hp[3], hp[4], hp[5], hp[6], hp[7]);
    }
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
