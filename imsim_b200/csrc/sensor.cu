// sensor.cu -- silicon sensor kernels (sm_100a): conversion depth, diffusion,
// tree-ring + brighter-fatter distorted pixel polygons, charge deposition,
// boundary updates at the nrecalc cadence, pixel areas.
//
// Replaces galsim.SiliconSensor.accumulate / calculate_pixel_areas as called from
// imsim/photon_pooling.py:195-225, imsim/stamp.py:562-572, imsim/flat.py:220-264.
//
// Device layout (per bound image of nx x ny pixels, nv vertices per pixel edge):
//   H  float2 [(ny+1)][nx][nv+2]      bottom edge of pixel (x,y): BL corner, nv points, BR corner,
//                                     in the owning pixel's frame (y ~ 0); corners live here only
//   V  float2 [ny][(nx+1)][nv]        left edge of pixel (x,y), bottom -> top (x ~ 0)
//   inner/outer double4 [ny][nx]      bounding boxes used for the fast accept / reject
//   delta double [ny][nx]             charge deposited since the last boundary update
//   target T [ny][nx]                 the image (float32 or float64)
// A pixel's polygon is 3 contiguous segments: H[y][x][0..nv+2), H[y+1][x][0..nv+2)
// and V[y][x][0..2nv) (its own left edge followed by the right neighbour's).
#include <algorithm>
#include <cstdlib>

#include "sensor_device.cuh"

// ------------------------------------------------------------------ accumulate
// Silicon::accumulate over photons [i1, i2), in two phases so that warps stay converged:
//   k_accumulate       every photon: conversion depth, diffusion, nominal pixel, inner-box test.
//                      ~97 % of the photons are decided here and deposited.  The rest (outside the
//                      inner bounding box of their nominal pixel) are appended to a compact list.
//   k_accumulate_slow  the listed photons, all lanes busy: outer box, polygon test, neighbour
//                      search, coin flip -- exactly the reference's sequence for those photons.
// Deposits are atomic adds into `delta`, so the split does not change any result.
__global__ void __launch_bounds__(256)
k_accumulate(const __grid_constant__ DevSensor s, int64_t i1, int64_t i2, int64_t ntot,
             const double* __restrict__ px, const double* __restrict__ py, const double* __restrict__ pdxdz,
             const double* __restrict__ pdydz, const double* __restrict__ pwl, const double* __restrict__ pflux,
             const double* __restrict__ rand4, uint64_t seed, uint64_t offset, unsigned long long* __restrict__ stats,
             double* __restrict__ added, SlowRec* __restrict__ slow, unsigned long long* __restrict__ nslow) {
    int64_t i = i1 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool active = i < i2;
    unsigned nb9 = 0, ndrop = 0;
    double my_added = 0.0;
    bool to_slow = false;
    SlowRec rec;
    if (active) {
        double g1, g2, unf, udep;
        if (rand4) {
            g1 = rand4[i];
            g2 = rand4[ntot + i];
            unf = rand4[2 * ntot + i];
            udep = rand4[3 * ntot + i];
        } else {
            sensor_draws(seed, offset + (uint64_t)i, g1, g2, unf, udep);
        }
        double a = 0.0, b = 0.0;
        if (pdxdz) {
            a = pdxdz[i];
            b = pdydz[i];
        }
        to_slow = sensor_fast_path(s, px[i], py[i], pdxdz != nullptr, a, b, pwl != nullptr, pwl ? pwl[i] : 0.0,
                                   pflux[i], g1, g2, unf, udep, rec, my_added, nb9, ndrop);
    }
    slow_append(to_slow, rec, slow, nslow);
    // warp-aggregated statistics
    unsigned long long w3 = warp_sum(nb9), w4 = warp_sum(ndrop);
    double wa = my_added;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wa += __shfl_xor_sync(0xffffffffu, wa, o);
    if ((threadIdx.x & 31) == 0) {
        if (w3) atomicAdd(&stats[ST_B9], w3);
        if (w4) atomicAdd(&stats[ST_DROP], w4);
        if (wa != 0.0) atomicAdd(added, wa);
    }
}

template <int NVT>
__global__ void __launch_bounds__(256)
k_accumulate_slow(const __grid_constant__ DevSensor s, const SlowRec* __restrict__ slow,
                  const unsigned long long* __restrict__ nslow, unsigned long long* __restrict__ stats,
                  double* __restrict__ added) {
    const unsigned long long total = *nslow;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    unsigned npoly = 0, nneigh = 0, nnf = 0;
    double my_added = 0.0;
    // whole warps iterate together so the shuffles below stay converged
    const unsigned long long first = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (unsigned long long base = first - (threadIdx.x & 31); base < total; base += stride) {
        unsigned long long j = base + (threadIdx.x & 31);
        if (j < total) {
            int dax, day;
            my_added += slow_photon<NVT>(s, slow[j], npoly, nneigh, nnf, dax, day);
        }
    }
    unsigned long long w0 = warp_sum(npoly), w1 = warp_sum(nneigh), w2 = warp_sum(nnf);
    double wa = my_added;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) wa += __shfl_xor_sync(0xffffffffu, wa, o);
    if ((threadIdx.x & 31) == 0) {
        if (w0) atomicAdd(&stats[ST_POLY], w0);
        if (w1) atomicAdd(&stats[ST_NEIGH], w1);
        if (w2) atomicAdd(&stats[ST_NOTFOUND], w2);
        if (wa != 0.0) atomicAdd(added, wa);
    }
}

// galsim.Sensor.accumulate = PhotonArray.addTo
template <typename T>
__global__ void __launch_bounds__(256)
k_plain_accumulate(const __grid_constant__ DevSensor s, int64_t n, const double* __restrict__ px,
                   const double* __restrict__ py, const double* __restrict__ pflux, double* __restrict__ added) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    double my = 0.0;
    if (i < n) {
        int ax = (int)floor(px[i] + 0.5) - s.xmin;
        int ay = (int)floor(py[i] + 0.5) - s.ymin;
        if (ax >= 0 && ax < s.nx && ay >= 0 && ay < s.ny) {
            double f = pflux[i];
            atomicAdd(&s.delta[(size_t)ay * s.nx + ax], f);
            my = f;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) my += __shfl_xor_sync(0xffffffffu, my, o);
    if ((threadIdx.x & 31) == 0 && my != 0.0) atomicAdd(added, my);
}

// ------------------------------------------------------------------ boundaries
// undistorted boundaries + tree rings: one thread per (x, y) slot, x in [0,nx], y in [0,ny]
__global__ void __launch_bounds__(256)
k_init_boundaries(const __grid_constant__ DevSensor s, int ocx, int ocy) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x > s.nx) return;
    const int nv = s.nv;
    const bool tr = s.ntr > 2;
    if (x < s.nx) {
        float2* h = s.H + Hidx(s, x, y);
        for (int k = 0; k <= nv + 1; ++k) {
            float2 p;
            p.x = (k == 0) ? 0.f : (k == nv + 1 ? 1.f : (float)s.frac[k - 1]);
            p.y = 0.f;
            if (tr) treering_point(s, p, s.xmin + x, s.ymin + y, ocx, ocy);
            h[k] = p;
        }
    }
    if (y < s.ny) {
        float2* v = s.V + Vidx(s, x, y);
        for (int k = 0; k < nv; ++k) {
            float2 p;
            p.x = 0.f;
            p.y = (float)s.frac[k];
            if (tr) treering_point(s, p, s.xmin + x, s.ymin + y, ocx, ocy);
            v[k] = p;
        }
    }
}

// Charge occupancy per 32x32-pixel tile: lets the boundary update skip the (usually large) part
// of the CCD that received no charge since the last update.
#define B2_TILE 32
template <typename CT>
__global__ void __launch_bounds__(256)
k_charge_tiles(const CT* __restrict__ charge, int nx, int ny, int tnx, uint8_t* __restrict__ tiles) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    bool nz = (x < nx) && (charge[(size_t)y * nx + x] != (CT)0);
    // one store per warp that saw charge (a warp spans one tile row segment)
    unsigned any = __ballot_sync(0xffffffffu, nz);
    if (any && (threadIdx.x & 31) == 0) tiles[(y / B2_TILE) * tnx + (x / B2_TILE)] = 1;
}

// true if any tile overlapping pixels [xa, xb] x [ya, yb] holds charge (uniform per block)
__device__ __forceinline__ bool tiles_any(const uint8_t* __restrict__ tiles, int tnx, int tny, int xa, int xb, int ya,
                                          int yb, int nx, int ny) {
    xa = max(xa, 0); ya = max(ya, 0);
    xb = min(xb, nx - 1); yb = min(yb, ny - 1);
    if (xa > xb || ya > yb) return false;
    for (int ty = ya / B2_TILE; ty <= yb / B2_TILE; ++ty)
        for (int tx = xa / B2_TILE; tx <= xb / B2_TILE; ++tx)
            if (tiles[ty * tnx + tx]) return true;
    return false;
}

// Silicon::updatePixelDistortions: charge-weighted sum of the per-electron kernels.
// One thread per (x, y) slot; CHARGE_T: the delta image (double) or the target image.
template <typename CT>
__global__ void __launch_bounds__(128)
k_update_distortions(const __grid_constant__ DevSensor s, const CT* __restrict__ charge, uint8_t* __restrict__ changed,
                     const uint8_t* __restrict__ tiles, int tnx, int tny) {
    extern __shared__ float2 sK[];  // KH then KV
    const int nv = s.nv;
    {
        // block-uniform early exit: no charge within reach of this strip of slots
        int xs = blockIdx.x * blockDim.x, ys = blockIdx.y;
        if (!tiles_any(tiles, tnx, tny, xs - s.qdist - 1, xs + (int)blockDim.x - 1 + s.qdist, ys - s.qdist - 1,
                       ys + s.qdist, s.nx, s.ny))
            return;
    }
    const int nKH = s.nx9 * s.ny9 * (nv + 2), nKV = s.nx9 * s.ny9 * nv;
    for (int k = threadIdx.x; k < nKH; k += blockDim.x) sK[k] = s.KH[k];
    for (int k = threadIdx.x; k < nKV; k += blockDim.x) sK[nKH + k] = s.KV[k];
    __syncthreads();
    const float2* KH = sK;
    const float2* KV = sK + nKH;
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x > s.nx) return;
    const int q = s.qdist, nx = s.nx, ny = s.ny;
    const int cxk = (s.nx9 - 1) / 2, cyk = (s.ny9 - 1) / 2;
    // horizontal slot
    if (x < nx) {
        int i1 = max(x - q, 0), i2 = min(x + q, nx - 1);
        int j1 = max(y - (q + 1), 0), j2 = min(y + q, ny - 1);
        int kmax = nv + 1;
        float2* h = s.H + Hidx(s, x, y);
        bool change = false;
        for (int j = j1; j <= j2; ++j)
            for (int i = i1; i <= i2; ++i) {
                double c = (double)charge[(size_t)j * nx + i];
                if (c == 0.0) continue;
                change = true;
                const float2* kh = KH + ((y - j + cyk) * s.nx9 + (x - i + cxk)) * (nv + 2);
                for (int k = 0; k <= kmax; ++k) {
                    float2 p = h[k];
                    float2 d = kh[k];
                    p.x = (float)__dadd_rn((double)p.x, __dmul_rn((double)d.x, c));
                    p.y = (float)__dadd_rn((double)p.y, __dmul_rn((double)d.y, c));
                    h[k] = p;
                }
            }
        if (change) {
            for (int dy = -1; dy <= 0; ++dy) {
                int pyy = y + dy;
                if (pyy >= 0 && pyy < ny) changed[(size_t)pyy * nx + x] = 1;
            }
        }
    }
    if (y < ny) {
        int i1 = max(x - (q + 1), 0), i2 = min(x + q, nx - 1);
        int j1 = max(y - q, 0), j2 = min(y + q, ny - 1);
        float2* v = s.V + Vidx(s, x, y);
        bool change = false;
        for (int j = j1; j <= j2; ++j)
            for (int i = i1; i <= i2; ++i) {
                double c = (double)charge[(size_t)j * nx + i];
                if (c == 0.0) continue;
                change = true;
                const float2* kv = KV + ((y - j + cyk) * s.nx9 + (x - i + cxk)) * nv;
                for (int k = 0; k < nv; ++k) {
                    float2 p = v[k];
                    float2 d = kv[k];
                    p.x = (float)__dadd_rn((double)p.x, __dmul_rn((double)d.x, c));
                    p.y = (float)__dadd_rn((double)p.y, __dmul_rn((double)d.y, c));
                    v[k] = p;
                }
            }
        if (change) {
            for (int dx = -1; dx <= 0; ++dx) {
                int pxx = x + dx;
                if (pxx >= 0 && pxx < nx) changed[(size_t)y * nx + pxx] = 1;
            }
        }
    }
}

// Tiled version for NV <= 8 (the models imSim ships: 4 and 8 vertices per edge).
// One block = 32 x 8 boundary slots.  The charge halo of the tile (39 x 15 pixels for qdist = 3) is
// staged in shared memory together with one occupancy bit word per halo row, so a slot only visits
// the pixels that actually hold charge (photon pools are sparse outside star cores), in the
// reference's (row, column) order; its boundary points live in registers and are written back once.
// Arithmetic per visited pixel is identical to k_update_distortions (bit-identical results).
// The running points are float-valued doubles.  MODE 0: float tables, GalSim's conversions (3 conversions + 2 FP64
// operations per term); 1: double tables, rounding on the FP64 adder (bf_term: 4 FP64 + 2 integer); 2: double tables,
// rounding by the narrowing / widening conversion pair (2 + 2: both pipes share the term); 3: as 2 with the tables
// staged in shared memory, entries padded by one point so that the per-lane reads of a warp (every lane another
// entry) spread over the banks -- the same reads through L1 cost a wavefront per distinct line.  Same bits in all.
template <typename CT, int NV, int MODE>
__global__ void __launch_bounds__(256)
k_update_distortions_tiled(const __grid_constant__ DevSensor s, const CT* __restrict__ charge,
                           uint8_t* __restrict__ changed) {
    constexpr int TX = 32, TY = 8, HW = 64, MAXHR = 24;
    extern __shared__ unsigned char smem_raw[];
    const int q = s.qdist;
    const int hrows = TY + 2 * q + 1;  // <= MAXHR
    const int hcols = TX + 2 * q + 1;  // <= HW
    double* sc = reinterpret_cast<double*>(smem_raw);                       // [hrows][HW]
    unsigned long long* rowbits = reinterpret_cast<unsigned long long*>(sc + MAXHR * HW);  // [MAXHR]
    // the 9 x 9 distortion tables (6.5 / 11.7 KB) are read through the L1 / read-only path: every block touches
    // all of them, so they stay cached, and not staging them saves a pass and a barrier per tile
    const double2* __restrict__ KH = s.KHd;
    const double2* __restrict__ KV = s.KVd;
    const float2* __restrict__ KHf = s.KH;
    const float2* __restrict__ KVf = s.KV;
    const int tid = threadIdx.y * TX + threadIdx.x;
    const int x0 = blockIdx.x * TX, y0 = blockIdx.y * TY;
    const int nx = s.nx, ny = s.ny;
    // stage the halo: warp w takes halo rows w, w + 8, w + 16
    const int lane = tid & 31, warp = tid >> 5;
    bool any_local = false;
    for (int r = warp; r < hrows; r += 8) {
        int gy = y0 - q - 1 + r;
        unsigned long long bits = 0ull;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            int c = lane + 32 * half;
            int gx = x0 - q - 1 + c;
            double v = 0.0;
            if (c < hcols && gx >= 0 && gx < nx && gy >= 0 && gy < ny) v = (double)charge[(size_t)gy * nx + gx];
            sc[r * HW + c] = v;
            unsigned b = __ballot_sync(0xffffffffu, v != 0.0);
            bits |= (unsigned long long)b << (32 * half);
        }
        if (lane == 0) rowbits[r] = bits;
        any_local |= (bits != 0ull);
    }
    if (!__syncthreads_or(any_local)) return;  // no charge within reach of this tile
    const int cxk = (s.nx9 - 1) / 2, cyk = (s.ny9 - 1) / 2;
    constexpr int SH = NV + 3, SV = NV + 1;  // padded entry lengths of the shared-memory tables [double2]
    double2* sKH = reinterpret_cast<double2*>(rowbits + MAXHR);
    double2* sKV = sKH + (MODE == 3 ? s.nx9 * s.ny9 * SH : 0);
    if (MODE == 3) {
        const int nent = s.nx9 * s.ny9;
        for (int k = tid; k < nent * (NV + 2); k += TX * TY) sKH[(k / (NV + 2)) * SH + k % (NV + 2)] = KH[k];
        for (int k = tid; k < nent * NV; k += TX * TY) sKV[(k / NV) * SV + k % NV] = KV[k];
        // (visible to every thread after the barriers of the slot sort below)
    }
    // Each slot's window of charged pixels is packed once into one 64-bit word, 8 bits per halo row (qdist <= 3:
    // at most 8 rows of at most 8 columns), lowest bit = first pixel in the reference's (row, column) order, so
    // the loop below visits exactly the charged pixels with one find-first-set each and no per-row scanning.
    // Slots without charge in reach (about half of them under star fields) are squeezed out first: the active
    // slots of the tile are compacted into a list in shared memory and the threads take them in order, so the
    // lanes of a warp all have work and only differ in how many pixels their slot sees.
    __shared__ unsigned long long s_bits[TX * TY];
    __shared__ unsigned short s_slot[TX * TY];
    constexpr int NBIN = 32;
    __shared__ int s_bin[NBIN];
    __shared__ int s_total;
    const int tx0 = threadIdx.x, ty0 = threadIdx.y;
#pragma unroll 1
    for (int phase = 0; phase < 2; ++phase) {  // 0: horizontal slots, 1: vertical slots
        unsigned long long bits = 0ull;
        {
            const int x = x0 + tx0, y = y0 + ty0;
            if (phase == 0) {
                // rows j = y-q-1 .. y+q (halo rows ty .. ty+2q+1), cols i = x-q .. x+q
                if (x < nx && y <= ny) {
                    const unsigned long long wmask = (1ull << (2 * q + 1)) - 1ull;
                    for (int dj = 0; dj < 2 * q + 2; ++dj) bits |= ((rowbits[ty0 + dj] >> (tx0 + 1)) & wmask) << (8 * dj);
                }
            } else {
                // rows j = y-q .. y+q (halo rows ty+1 .. ty+2q+1), cols i = x-q-1 .. x+q
                if (x <= nx && y < ny) {
                    const unsigned long long wmask = (1ull << (2 * q + 2)) - 1ull;
                    for (int dj = 1; dj < 2 * q + 2; ++dj) bits |= ((rowbits[ty0 + dj] >> tx0) & wmask) << (8 * (dj - 1));
                }
            }
        }
        // counting sort of the active slots by the number of charged pixels they see (descending), so that the
        // 32 slots a warp takes have similar trip counts: the loop below then runs with nearly full warps
        const int cnt = __popcll(bits);
        const int bin = cnt == 0 ? -1 : (NBIN - 1 - min(cnt - 1, NBIN - 1));  // bin 0 = most pixels
        if (tid < NBIN) s_bin[tid] = 0;
        __syncthreads();
        int rank = 0;
        if (bin >= 0) rank = atomicAdd(&s_bin[bin], 1);
        __syncthreads();
        if (tid < NBIN) {  // exclusive prefix over the bins: one warp, five shuffles (NBIN == 32)
            const int c = s_bin[tid];
            int incl = c;
#pragma unroll
            for (int o = 1; o < NBIN; o <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, o);
                if (tid >= o) incl += t;
            }
            s_bin[tid] = incl - c;
            if (tid == NBIN - 1) s_total = incl;
        }
        __syncthreads();
        if (bin >= 0) {
            const int k = s_bin[bin] + rank;
            s_bits[k] = bits;
            s_slot[k] = (unsigned short)tid;
        }
        __syncthreads();
        const int total = s_total;
        if (tid < total) {
            bits = s_bits[tid];
            const int sid = s_slot[tid];
            const int tx = sid & 31, ty = sid >> 5;
            const int x = x0 + tx, y = y0 + ty;
            if (phase == 0) {
                // the points live in registers as float-valued doubles (bf_term): widened once, narrowed once
                float2* hp = s.H + Hidx(s, x, y);
                double hx[NV + 2], hy[NV + 2];
#pragma unroll
                for (int k = 0; k < NV + 2; ++k) {
                    const float2 t = hp[k];
                    hx[k] = (double)t.x;
                    hy[k] = (double)t.y;
                }
                while (bits) {
                    const int pos = __ffsll((long long)bits) - 1;
                    bits &= bits - 1;
                    const int dj = pos >> 3, di = pos & 7;
                    const double c = sc[(ty + dj) * HW + tx + 1 + di];
                    const int kk = ((q + 1 - dj + cyk) * s.nx9 + (q - di + cxk)) * (NV + 2);
#pragma unroll
                    for (int k = 0; k < NV + 2; ++k) {
                        if (MODE == 1) {
                            const double2 d = __ldg(KH + kk + k);
                            bf_term(hx[k], d.x, c);
                            bf_term(hy[k], d.y, c);
                        } else if (MODE == 3) {
                            const double2 d = sKH[(kk / (NV + 2)) * SH + k];
                            hx[k] = (double)(float)__dadd_rn(hx[k], __dmul_rn(d.x, c));
                            hy[k] = (double)(float)__dadd_rn(hy[k], __dmul_rn(d.y, c));
                        } else if (MODE == 2) {
                            const double2 d = __ldg(KH + kk + k);
                            hx[k] = (double)(float)__dadd_rn(hx[k], __dmul_rn(d.x, c));
                            hy[k] = (double)(float)__dadd_rn(hy[k], __dmul_rn(d.y, c));
                        } else {
                            const float2 d = __ldg(KHf + kk + k);
                            hx[k] = (double)(float)__dadd_rn(hx[k], __dmul_rn((double)d.x, c));
                            hy[k] = (double)(float)__dadd_rn(hy[k], __dmul_rn((double)d.y, c));
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < NV + 2; ++k) hp[k] = make_float2((float)hx[k], (float)hy[k]);
                if (y < ny) changed[(size_t)y * nx + x] = 1;
                if (y > 0) changed[(size_t)(y - 1) * nx + x] = 1;
            } else {
                float2* vp = s.V + Vidx(s, x, y);
                double vx[NV], vy[NV];
#pragma unroll
                for (int k = 0; k < NV; ++k) {
                    const float2 t = vp[k];
                    vx[k] = (double)t.x;
                    vy[k] = (double)t.y;
                }
                while (bits) {
                    const int pos = __ffsll((long long)bits) - 1;
                    bits &= bits - 1;
                    const int dj = (pos >> 3) + 1, di = pos & 7;
                    const double c = sc[(ty + dj) * HW + tx + di];
                    const int kk = ((q + 1 - dj + cyk) * s.nx9 + (q + 1 - di + cxk)) * NV;
#pragma unroll
                    for (int k = 0; k < NV; ++k) {
                        if (MODE == 1) {
                            const double2 d = __ldg(KV + kk + k);
                            bf_term(vx[k], d.x, c);
                            bf_term(vy[k], d.y, c);
                        } else if (MODE == 3) {
                            const double2 d = sKV[(kk / NV) * SV + k];
                            vx[k] = (double)(float)__dadd_rn(vx[k], __dmul_rn(d.x, c));
                            vy[k] = (double)(float)__dadd_rn(vy[k], __dmul_rn(d.y, c));
                        } else if (MODE == 2) {
                            const double2 d = __ldg(KV + kk + k);
                            vx[k] = (double)(float)__dadd_rn(vx[k], __dmul_rn(d.x, c));
                            vy[k] = (double)(float)__dadd_rn(vy[k], __dmul_rn(d.y, c));
                        } else {
                            const float2 d = __ldg(KVf + kk + k);
                            vx[k] = (double)(float)__dadd_rn(vx[k], __dmul_rn((double)d.x, c));
                            vy[k] = (double)(float)__dadd_rn(vy[k], __dmul_rn((double)d.y, c));
                        }
                    }
                }
#pragma unroll
                for (int k = 0; k < NV; ++k) vp[k] = make_float2((float)vx[k], (float)vy[k]);
                if (x < nx) changed[(size_t)y * nx + x] = 1;
                if (x > 0) changed[(size_t)y * nx + x - 1] = 1;
            }
        }
        __syncthreads();  // the lists are reused by the next phase
    }
}

// Silicon::updatePixelBounds for every pixel (all = 1) or the flagged ones
__global__ void __launch_bounds__(256)
k_update_bounds(const __grid_constant__ DevSensor s, const uint8_t* __restrict__ changed, int all,
                const uint8_t* __restrict__ tiles, int tnx, int tny) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (!all && tiles != nullptr) {
        int xs = blockIdx.x * blockDim.x;
        if (!tiles_any(tiles, tnx, tny, xs - s.qdist - 2, xs + (int)blockDim.x + s.qdist + 1, y - s.qdist - 2,
                       y + s.qdist + 1, s.nx, s.ny))
            return;
    }
    if (x >= s.nx) return;
    size_t pix = (size_t)y * s.nx + x;
    if (!all && !changed[pix]) return;
    double oxmin = INFINITY, oxmax = -INFINITY, oymin = INFINITY, oymax = -INFINITY;
    walk_polygon(s, x, y, [&](double px, double py, double, double) {
        oxmin = fmin(oxmin, px);
        oxmax = fmax(oxmax, px);
        oymin = fmin(oymin, py);
        oymax = fmax(oymax, py);
    });
    // the "trivially inside" box, inscribed by construction: the innermost vertex of each edge, corners counting for
    // both edges they end (oracle_sensor.c: update_bounds)
    double ixmin = -INFINITY, ixmax = INFINITY, iymin = -INFINITY, iymax = INFINITY;
    walk_polygon(s, x, y, [&](double px, double py, double ex, double ey) {
        if (ex == 0.0 && px > ixmin) ixmin = px;
        if (ex == 1.0 && px < ixmax) ixmax = px;
        if (ey == 0.0 && py > iymin) iymin = py;
        if (ey == 1.0 && py < iymax) iymax = py;
    });
    *reinterpret_cast<double4*>(s.outer + pix * 4) = make_double4(oxmin, oxmax, oymin, oymax);
    *reinterpret_cast<double4*>(s.inner + pix * 4) = make_double4(ixmin, ixmax, iymin, iymax);
}

// The same computation for NV = 4 / 8 with the pixel's polygon held in registers: all four edge segments are loaded
// up front (independent loads in flight together), the outer box is a float min / max per edge (the stored points
// are floats, conversion and the +1 offsets of the right / top edges are monotone, so the result is the same
// double), and the inner box walks the registers instead of global memory a second time.
template <int NV>
__global__ void __launch_bounds__(256)
k_update_bounds_t(const __grid_constant__ DevSensor s, const uint8_t* __restrict__ changed, int all) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= s.nx) return;
    size_t pix = (size_t)y * s.nx + x;
    if (!all && !changed[pix]) return;
    const float2* hb = s.H + Hidx(s, x, y);
    const float2* ht = s.H + Hidx(s, x, y + 1);
    const float2* vl = s.V + Vidx(s, x, y);
    const float2* vr = vl + NV;
    float2 B[NV + 2], T[NV + 2], L[NV], R[NV];
#pragma unroll
    for (int k = 0; k < NV + 2; ++k) B[k] = hb[k];
#pragma unroll
    for (int k = 0; k < NV + 2; ++k) T[k] = ht[k];
#pragma unroll
    for (int k = 0; k < NV; ++k) L[k] = vl[k];
#pragma unroll
    for (int k = 0; k < NV; ++k) R[k] = vr[k];
    // outer box
    float bx0 = B[0].x, bx1 = B[0].x, by0 = B[0].y, by1 = B[0].y;
    float tx0 = T[0].x, tx1 = T[0].x, ty0 = T[0].y, ty1 = T[0].y;
    float lx0 = L[0].x, lx1 = L[0].x, ly0 = L[0].y, ly1 = L[0].y;
    float rx0 = R[0].x, rx1 = R[0].x, ry0 = R[0].y, ry1 = R[0].y;
#pragma unroll
    for (int k = 1; k < NV + 2; ++k) {
        bx0 = fminf(bx0, B[k].x); bx1 = fmaxf(bx1, B[k].x); by0 = fminf(by0, B[k].y); by1 = fmaxf(by1, B[k].y);
        tx0 = fminf(tx0, T[k].x); tx1 = fmaxf(tx1, T[k].x); ty0 = fminf(ty0, T[k].y); ty1 = fmaxf(ty1, T[k].y);
    }
#pragma unroll
    for (int k = 1; k < NV; ++k) {
        lx0 = fminf(lx0, L[k].x); lx1 = fmaxf(lx1, L[k].x); ly0 = fminf(ly0, L[k].y); ly1 = fmaxf(ly1, L[k].y);
        rx0 = fminf(rx0, R[k].x); rx1 = fmaxf(rx1, R[k].x); ry0 = fminf(ry0, R[k].y); ry1 = fmaxf(ry1, R[k].y);
    }
    const double oxmin = fmin((double)fminf(fminf(bx0, tx0), lx0), (double)rx0 + 1.0);
    const double oxmax = fmax((double)fmaxf(fmaxf(bx1, tx1), lx1), (double)rx1 + 1.0);
    const double oymin = fmin((double)fminf(fminf(by0, ly0), ry0), (double)ty0 + 1.0);
    const double oymax = fmax((double)fmaxf(fmaxf(by1, ly1), ry1), (double)ty1 + 1.0);
    // inner box: the innermost vertex of each edge, corners counting for both edges they end -- the per-edge
    // extrema are at hand already (conversion and the +1 offsets are monotone, as for the outer box)
    const double ixmin = (double)fmaxf(fmaxf(lx1, B[0].x), T[0].x);
    const double ixmax = fmin((double)rx0 + 1.0, (double)fminf(B[NV + 1].x, T[NV + 1].x));
    const double iymin = (double)by1;
    const double iymax = (double)ty0 + 1.0;
    *reinterpret_cast<double4*>(s.outer + pix * 4) = make_double4(oxmin, oxmax, oymin, oymax);
    *reinterpret_cast<double4*>(s.inner + pix * 4) = make_double4(ixmin, ixmax, iymin, iymax);
}

// target (+/-)= delta, optionally clearing delta (Silicon::addDelta / subtractDelta / update)
template <typename T>
__global__ void __launch_bounds__(256)
k_add_delta(T* __restrict__ target, double* __restrict__ delta, size_t n, double sign, int clear) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double d = delta[i];
    if (d != 0.0) {
        target[i] = (T)__dadd_rn((double)target[i], __dmul_rn(sign, d));
        if (clear) delta[i] = 0.0;
    }
}

// Silicon::pixelArea (shoelace) for every pixel
__global__ void __launch_bounds__(256)
k_pixel_areas(const __grid_constant__ DevSensor s, double* __restrict__ areas) {
    int x = blockIdx.x * blockDim.x + threadIdx.x;
    int y = blockIdx.y;
    if (x >= s.nx) return;
    // GalSim starts at polygon vertex 0 (mid left edge); the sum is cyclic, but to keep the
    // same rounding sequence start there as well: collect the walk (which starts at BL) and rotate
    const int nv = s.nv;
    const int npoly = 4 * nv + 4;
    const int start = npoly - nv / 2;  // walk index of polygon vertex 0
    double area = 0.0;
    // two passes over the walk: [start, npoly) then [0, start], chaining consecutive vertices
    double x1 = 0, y1 = 0, fx = 0, fy = 0;
    bool have = false;
    for (int pass = 0; pass < 2; ++pass) {
        int idx = 0;
        walk_polygon(s, x, y, [&](double px, double py, double, double) {
            bool use = (pass == 0) ? (idx >= start) : (idx < start);
            if (use) {
                if (!have) {
                    fx = px; fy = py;
                    have = true;
                } else {
                    area = __dadd_rn(area, __dmul_rn(x1, py));
                    area = __dsub_rn(area, __dmul_rn(px, y1));
                }
                x1 = px; y1 = py;
            }
            idx++;
        });
    }
    area = __dadd_rn(area, __dmul_rn(x1, fy));
    area = __dsub_rn(area, __dmul_rn(fx, y1));
    areas[(size_t)y * s.nx + x] = fabs(area) / 2.0;
}

// Chunk boundaries for nrecalc > 0 (Silicon::accumulate updates the pixel boundaries each time the flux
// added since the last update reaches nrecalc).  Two kernels: k_segment_sums reduces the flux array in
// segments of CHUNK_SEG photons; k_find_chunks (one block) walks the segment sums and scans photon by
// photon, in photon order like the reference's loop, only the segments in which a boundary falls.  For
// unit / integer fluxes (stars, flats) the result is exactly the sequential one; for general fluxes the
// segment sums round differently from a photon-by-photon running total, which can move a boundary by
// one photon when the total lands within an ulp of nrecalc.
#define CHUNK_SEG 4096

__global__ void __launch_bounds__(256)
k_segment_sums(const double* __restrict__ flux, int64_t n, double* __restrict__ sums) {
    const int64_t base = (int64_t)blockIdx.x * CHUNK_SEG;
    double a = 0.0;
#pragma unroll 4
    for (int k = threadIdx.x; k < CHUNK_SEG; k += 256) {
        int64_t i = base + k;
        if (i < n) a += flux[i];
    }
    __shared__ double part[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = a;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += part[k];
        sums[blockIdx.x] = t;
    }
}

__global__ void __launch_bounds__(1024)
k_find_chunks(const double* __restrict__ flux, int64_t n, const double* __restrict__ sums, int64_t nseg, double accum0,
              double nrecalc, int64_t* __restrict__ bounds, int max_bounds, int* __restrict__ nb, double* accum_out) {
    __shared__ double ssum[1024];
    __shared__ double tile[CHUNK_SEG];
    __shared__ double carry;
    __shared__ int count, need, kstart;
    if (threadIdx.x == 0) {
        carry = accum0;
        count = 0;
    }
    for (int64_t seg0 = 0; seg0 < nseg; seg0 += 1024) {
        const int lim = (int)((nseg - seg0 < 1024) ? (nseg - seg0) : 1024);
        __syncthreads();
        ssum[threadIdx.x] = (threadIdx.x < lim) ? sums[seg0 + threadIdx.x] : 0.0;
        if (threadIdx.x == 0) kstart = 0;
        __syncthreads();
        for (;;) {
            if (threadIdx.x == 0) {
                int k = kstart;
                double acc = carry;
                while (k < lim && acc + ssum[k] < nrecalc) acc += ssum[k++];
                carry = acc;
                need = (k < lim) ? k : -1;
                kstart = k + 1;
            }
            __syncthreads();
            const int seg = need;
            if (seg < 0) break;
            const int64_t base = (seg0 + seg) * CHUNK_SEG;
            for (int k = threadIdx.x; k < CHUNK_SEG; k += 1024) tile[k] = (base + k < n) ? flux[base + k] : 0.0;
            __syncthreads();
            if (threadIdx.x == 0) {
                // exact sequential semantics inside the segment (order of additions = photon order)
                double acc = carry;
                const int m = (int)((n - base < CHUNK_SEG) ? (n - base) : CHUNK_SEG);
                for (int k = 0; k < m; ++k) {
                    acc += tile[k];
                    if (acc >= nrecalc) {
                        if (count < max_bounds) bounds[count] = base + k + 1;
                        count++;
                        acc = 0.0;
                    }
                }
                carry = acc;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        *nb = count;
        *accum_out = carry;
    }
}

// ------------------------------------------------------------------ host side
static void launch_slow(unsigned blocks, cudaStream_t st, const DevSensor& d, const SlowRec* slow,
                        const unsigned long long* nslow, unsigned long long* stats, double* added) {
    if (d.nv == 4) k_accumulate_slow<4><<<blocks, 256, 0, st>>>(d, slow, nslow, stats, added);
    else if (d.nv == 8) k_accumulate_slow<8><<<blocks, 256, 0, st>>>(d, slow, nslow, stats, added);
    else k_accumulate_slow<0><<<blocks, 256, 0, st>>>(d, slow, nslow, stats, added);
}

static inline dim3 grid2(int nxslots, int ny, int bs) { return dim3((nxslots + bs - 1) / bs, ny, 1); }

template <typename T>
static int dev_upload(b2_ctx* ctx, std::vector<void*>& owned, const T* src, size_t n, const T** out) {
    void* p = nullptr;
    B2_CUDA(cudaMalloc(&p, (n ? n : 1) * sizeof(T)));
    owned.push_back(p);
    if (n) B2_CUDA(cudaMemcpyAsync(p, src, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    *out = (const T*)p;
    return 0;
}

static void spline_y2_host(int n, const double* x, const double* f, std::vector<double>& y2) {
    y2.assign(n, 0.0);
    if (n < 3) return;
    std::vector<double> cp(n, 0.0), dp(n, 0.0);
    for (int i = 1; i < n - 1; ++i) {
        double h0 = x[i] - x[i - 1], h1 = x[i + 1] - x[i];
        double b = 2.0 * (h0 + h1);
        double rhs = 6.0 * ((f[i + 1] - f[i]) / h1 - (f[i] - f[i - 1]) / h0);
        double m = b - h0 * cp[i - 1];
        cp[i] = h1 / m;
        dp[i] = (rhs - h0 * dp[i - 1]) / m;
    }
    for (int i = n - 2; i >= 1; --i) y2[i] = dp[i] - cp[i] * y2[i + 1];
}

extern "C" int b2_sensor_create(b2_ctx* ctx, const B2SensorConfig* cfg, const double* vertex_data,
                                const double* tr_r, const double* tr_f, const double* tr_y2, const double* abs_w,
                                const double* abs_l, b2_sensor** out) {
    B2_REQUIRE(ctx && cfg && vertex_data && out, "b2_sensor_create: null argument");
    B2_REQUIRE(cfg->num_vertices >= 2 && cfg->num_vertices <= B2_MAX_NV && cfg->num_vertices % 2 == 0,
               "b2_sensor_create: num_vertices must be even and <= 32");
    B2_REQUIRE(cfg->nx >= 2 * cfg->qdist + 3 && cfg->ny >= 2 * cfg->qdist + 3,
               "b2_sensor_create: vertex table smaller than 2*qdist+3");
    B2_REQUIRE(!cfg->transpose, "b2_sensor_create: transpose=True is not supported yet");
    B2_REQUIRE(cfg->n_treering <= 2 || (tr_r && tr_f), "b2_sensor_create: tree-ring table missing");
    B2_CUDA(cudaSetDevice(ctx->device));
    b2_sensor* s = new b2_sensor();
    s->ctx = ctx;
    {
        cudaDeviceProp prop;
        B2_CUDA(cudaGetDeviceProperties(&prop, ctx->device));
        s->sm_count = prop.multiProcessorCount;
    }
    s->cfg = *cfg;
    DevSensor& d = s->d;
    memset(&d, 0, sizeof(d));
    const int nv = d.nv = cfg->num_vertices;
    d.nx9 = cfg->nx;
    d.ny9 = cfg->ny;
    d.qdist = cfg->qdist;
    d.diff_step = cfg->diff_step;
    d.pixel_size = cfg->pixel_size;
    d.thickness = cfg->sensor_thickness;
    d.inv_pixel_size = 1. / d.pixel_size;
    d.diff_step_pixel_z = d.diff_step / (d.thickness * d.pixel_size);
    d.trc[0] = cfg->treering_center[0];
    d.trc[1] = cfg->treering_center[1];
    // undistorted edge points (GalSim buildEmptyPoly)
    {
        double theta0 = -PI_D / 4.0, dtheta = PI_D / (2.0 * (nv + 1.0));
        for (int k = 0; k < nv; ++k) d.frac[k] = (tan(theta0 + (k + 1.0) * dtheta) + 1.0) / 2.0;
    }
    // polygon order of the .dat file -> undistorted positions
    const int npoly = 4 * nv + 4;
    std::vector<double> ex(npoly), ey(npoly);
    {
        int n = 0;
        for (int k = nv / 2 - 1; k >= 0; --k) { ex[n] = 0.0; ey[n] = d.frac[k]; n++; }
        ex[n] = 0.0; ey[n] = 0.0; n++;
        for (int k = 0; k < nv; ++k) { ex[n] = d.frac[k]; ey[n] = 0.0; n++; }
        ex[n] = 1.0; ey[n] = 0.0; n++;
        for (int k = 0; k < nv; ++k) { ex[n] = 1.0; ey[n] = d.frac[k]; n++; }
        ex[n] = 1.0; ey[n] = 1.0; n++;
        for (int k = nv - 1; k >= 0; --k) { ex[n] = d.frac[k]; ey[n] = 1.0; n++; }
        ex[n] = 0.0; ey[n] = 1.0; n++;
        for (int k = nv - 1; k >= nv / 2; --k) { ex[n] = 0.0; ey[n] = d.frac[k]; n++; }
    }
    const int nx9 = cfg->nx, ny9 = cfg->ny;
    std::vector<float2> KH((size_t)nx9 * ny9 * (nv + 2)), KV((size_t)nx9 * ny9 * nv);
    for (int i = 0; i < nx9; ++i)
        for (int j = 0; j < ny9; ++j)
            for (int n = 0; n < npoly; ++n) {
                const double* row = vertex_data + 5 * (((size_t)i * ny9 + j) * npoly + n);
                double pxv = (row[3] - row[0]) / cfg->pixel_size + 0.5;
                double pyv = (row[4] - row[1]) / cfg->pixel_size + 0.5;
                float2 dd;
                dd.x = (float)((pxv - ex[n]) / cfg->num_elec);
                dd.y = (float)((pyv - ey[n]) / cfg->num_elec);
                if (n >= nv / 2 && n <= nv / 2 + nv + 1) KH[((size_t)j * nx9 + i) * (nv + 2) + (n - nv / 2)] = dd;
                else if (n < nv / 2) KV[((size_t)j * nx9 + i) * nv + (nv / 2 - 1 - n)] = dd;
                else if (n >= 7 * nv / 2 + 4) KV[((size_t)j * nx9 + i) * nv + (nv - 1 - (n - (7 * nv / 2 + 4)))] = dd;
            }
    if (dev_upload(ctx, s->owned, KH.data(), KH.size(), &d.KH)) return 1;
    if (dev_upload(ctx, s->owned, KV.data(), KV.size(), &d.KV)) return 1;
    {
        std::vector<double2> KHd(KH.size()), KVd(KV.size());
        for (size_t k = 0; k < KH.size(); ++k) KHd[k] = make_double2((double)KH[k].x, (double)KH[k].y);
        for (size_t k = 0; k < KV.size(); ++k) KVd[k] = make_double2((double)KV[k].x, (double)KV[k].y);
        if (dev_upload(ctx, s->owned, KHd.data(), KHd.size(), &d.KHd)) return 1;
        if (dev_upload(ctx, s->owned, KVd.data(), KVd.size(), &d.KVd)) return 1;
    }
    d.ntr = cfg->n_treering;
    if (d.ntr > 2) {
        if (dev_upload(ctx, s->owned, tr_r, (size_t)d.ntr, &d.tr_r)) return 1;
        if (dev_upload(ctx, s->owned, tr_f, (size_t)d.ntr, &d.tr_f)) return 1;
        d.tr_max = tr_r[d.ntr - 1];
        if (tr_y2) {
            if (dev_upload(ctx, s->owned, tr_y2, (size_t)d.ntr, &d.tr_y2)) return 1;
            d.tr_spline = 1;
        }
    }
    d.nabs = cfg->n_abs;
    if (d.nabs > 0) {
        B2_REQUIRE(abs_w && abs_l, "b2_sensor_create: absorption table missing");
        if (dev_upload(ctx, s->owned, abs_w, (size_t)d.nabs, &d.abs_w)) return 1;
        if (dev_upload(ctx, s->owned, abs_l, (size_t)d.nabs, &d.abs_l)) return 1;
        // uniformly spaced wavelengths (GalSim's absorption.dat is, 5 nm): index by arithmetic instead of a search
        d.abs_x0 = abs_w[0];
        d.abs_inv_dx = 0.0;
        if (d.nabs >= 3) {
            const double dx = (abs_w[d.nabs - 1] - abs_w[0]) / (d.nabs - 1);
            bool uniform = dx > 0.0;
            for (int k = 1; k < d.nabs && uniform; ++k) uniform = fabs(abs_w[k] - (abs_w[0] + k * dx)) <= 1e-9 * dx;
            if (uniform) d.abs_inv_dx = 1.0 / dx;
        }
    }
    void* p = nullptr;
    B2_CUDA(cudaMalloc(&p, ST_N * sizeof(unsigned long long) + 64));
    s->owned.push_back(p);
    s->dstats = (unsigned long long*)p;
    s->dadded = (double*)((char*)p + ST_N * sizeof(unsigned long long));
    s->dnslow = (unsigned long long*)((char*)p + ST_N * sizeof(unsigned long long) + 8);
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->sensors.push_back(s);
    *out = s;
    return 0;
}

// replace the tree-ring table / centre of an existing sensor (next detector of the same vendor): the
// per-image boundary arrays, the big allocation, are kept
extern "C" int b2_sensor_set_treerings(b2_sensor* s, double cx, double cy, const double* tr_r, const double* tr_f,
                                       const double* tr_y2, int32_t n) {
    B2_REQUIRE(s && s->ctx, "b2_sensor_set_treerings: null sensor, or its context has been destroyed");
    B2_REQUIRE(s, "b2_sensor_set_treerings: null sensor");
    B2_REQUIRE(n <= 2 || (tr_r && tr_f), "b2_sensor_set_treerings: table missing");
    b2_ctx* ctx = s->ctx;
    B2_CUDA(cudaSetDevice(ctx->device));
    // No stream synchronisation: kernels take DevSensor by value at launch, the table copies below are ordered on
    // the stream behind whatever is already queued, and a small pageable upload is staged before cudaMemcpyAsync
    // returns.  A visit can therefore queue the next detector behind the running one.
    DevSensor& d = s->d;
    d.trc[0] = cx;
    d.trc[1] = cy;
    s->cfg.treering_center[0] = cx;
    s->cfg.treering_center[1] = cy;
    d.ntr = n;
    d.tr_spline = 0;
    d.tr_r = d.tr_f = d.tr_y2 = nullptr;
    if (n > 2) {
        // one set of buffers, reused from detector to detector: the copies are ordered on the stream behind the
        // kernels that still read the previous detector's tables
        if ((size_t)n > s->tr_cap) {
            for (int k = 0; k < 3; ++k) {
                void* p = nullptr;
                B2_CUDA(cudaMalloc(&p, (size_t)n * sizeof(double)));
                s->owned.push_back(p);
                s->tr_buf[k] = (double*)p;
            }
            s->tr_cap = (size_t)n;
        }
        const double* src[3] = {tr_r, tr_f, tr_y2};
        for (int k = 0; k < 3; ++k)
            if (src[k])
                B2_CUDA(cudaMemcpyAsync(s->tr_buf[k], src[k], (size_t)n * sizeof(double), cudaMemcpyHostToDevice,
                                        ctx->stream));
        d.tr_r = s->tr_buf[0];
        d.tr_f = s->tr_buf[1];
        d.tr_max = tr_r[n - 1];
        if (tr_y2) {
            d.tr_y2 = s->tr_buf[2];
            d.tr_spline = 1;
        }
    }
    s->initialized = false;
    return 0;
}

static void free_list(std::vector<void*>& v) {
    for (void* p : v) cudaFree(p);
    v.clear();
}

// device resources of a sensor; the struct itself stays until b2_sensor_destroy
static void sensor_release_device(b2_sensor* s) {
    if (!s->ctx) return;
    cudaSetDevice(s->ctx->device);
    cudaStreamSynchronize(s->ctx->stream);
    free_list(s->image_owned);
    free_list(s->owned);
    if (s->cum.ptr) cudaFree(s->cum.ptr);
    if (s->slow.ptr) cudaFree(s->slow.ptr);
    if (s->stamp_meta.ptr) cudaFree(s->stamp_meta.ptr);
    if (s->stamp_arena.ptr) cudaFree(s->stamp_arena.ptr);
    s->cum = s->slow = s->stamp_meta = s->stamp_arena = Scratch{};
    if (s->stamp_aux) {
        cudaStreamSynchronize(s->stamp_aux);
        cudaStreamDestroy(s->stamp_aux);
        cudaEventDestroy(s->stamp_ev[0]);
        cudaEventDestroy(s->stamp_ev[1]);
        s->stamp_aux = nullptr;
    }
    s->bound = s->initialized = false;
}

// called by b2_ctx_destroy for the sensors still alive on it: host languages with garbage collection may
// finalise the context before its sensors; their handles stay valid to destroy, every other call fails loudly
void b2_sensor_orphan(b2_sensor* s) {
    sensor_release_device(s);
    s->ctx = nullptr;
}

extern "C" int b2_sensor_destroy(b2_sensor* s) {
    if (!s) return 0;
    if (s->ctx) {
        auto& v = s->ctx->sensors;
        v.erase(std::remove(v.begin(), v.end(), s), v.end());
        sensor_release_device(s);
    }
    delete s;
    return 0;
}

extern "C" int b2_sensor_bind_image(b2_sensor* s, int32_t xmin, int32_t ymin, int32_t nx, int32_t ny,
                                    int32_t dtype_bytes, const void* pixels, int where) {
    B2_REQUIRE(s, "b2_sensor_bind_image: null sensor");
    B2_REQUIRE(s->ctx, "b2_sensor_bind_image: the context this sensor was created on has been destroyed");
    B2_REQUIRE(nx > 0 && ny > 0, "b2_sensor_bind_image: empty image");
    B2_REQUIRE(dtype_bytes == 4 || dtype_bytes == 8, "b2_sensor_bind_image: image must be float32 or float64");
    b2_ctx* ctx = s->ctx;
    B2_CUDA(cudaSetDevice(ctx->device));
    DevSensor& d = s->d;
    // The per-image arrays are kept while the new image fits their capacity (per-object stamps of the
    // classic pipeline rebind thousands of times with varying sizes); they only ever grow.
    const int nv = d.nv;
    const size_t nH = (size_t)(ny + 1) * nx * (nv + 2), nV = (size_t)ny * (nx + 1) * nv + nv;
    const size_t npix = (size_t)nx * ny;
    const size_t ntile = (size_t)((nx + B2_TILE - 1) / B2_TILE) * ((ny + B2_TILE - 1) / B2_TILE);
    bool fits = s->bound && nH <= s->cap_H && nV <= s->cap_V && npix * 8 <= s->cap_pix_bytes && ntile <= s->cap_tiles;
    if (!fits) {
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
        free_list(s->image_owned);
        s->cap_H = nH + nH / 4;
        s->cap_V = nV + nV / 4;
        const size_t cpix = npix + npix / 4;
        s->cap_pix_bytes = cpix * 8;
        s->cap_tiles = ntile + ntile / 4 + 1;
        void* p;
        B2_CUDA(cudaMalloc(&p, s->cap_H * sizeof(float2))); s->image_owned.push_back(p); d.H = (float2*)p;
        B2_CUDA(cudaMalloc(&p, s->cap_V * sizeof(float2))); s->image_owned.push_back(p); d.V = (float2*)p;
        B2_CUDA(cudaMalloc(&p, cpix * 4 * sizeof(double))); s->image_owned.push_back(p); d.inner = (double*)p;
        B2_CUDA(cudaMalloc(&p, cpix * 4 * sizeof(double))); s->image_owned.push_back(p); d.outer = (double*)p;
        B2_CUDA(cudaMalloc(&p, cpix * sizeof(double))); s->image_owned.push_back(p); d.delta = (double*)p;
        B2_CUDA(cudaMalloc(&p, cpix * 8)); s->image_owned.push_back(p); d.target = p;  // float32 or float64 pixels
        B2_CUDA(cudaMalloc(&p, cpix)); s->image_owned.push_back(p); s->changed = (uint8_t*)p;
        B2_CUDA(cudaMalloc(&p, s->cap_tiles)); s->image_owned.push_back(p); s->tiles = (uint8_t*)p;
    }
    s->tnx = (nx + B2_TILE - 1) / B2_TILE;
    s->tny = (ny + B2_TILE - 1) / B2_TILE;
    d.xmin = xmin; d.ymin = ymin; d.nx = nx; d.ny = ny; d.dtype_bytes = dtype_bytes;
    size_t bytes = (size_t)nx * ny * dtype_bytes;
    if (pixels && where == B2_HOST && bytes >= ((size_t)8 << 20) && bytes % 8 == 0 && getenv("B2_IMAGE_PLAIN_COPY") == nullptr) {
        const double* hin[1] = {(const double*)pixels};
        double* din[1] = {(double*)d.target};
        if (b2_pipe_run(ctx, (int64_t)(bytes / 8), 1, hin, din, 0, nullptr, nullptr, nullptr)) return 1;
    } else if (pixels)
        B2_CUDA(cudaMemcpyAsync(d.target, pixels, bytes, where == B2_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, ctx->stream));
    else
        B2_CUDA(cudaMemsetAsync(d.target, 0, bytes, ctx->stream));
    B2_CUDA(cudaMemsetAsync(d.delta, 0, (size_t)nx * ny * sizeof(double), ctx->stream));
    if (where == B2_HOST) B2_CUDA(cudaStreamSynchronize(ctx->stream));
    s->bound = true;
    s->initialized = false;
    s->accum_flux = 0.0;
    return 0;
}

extern "C" int b2_sensor_read_image(b2_sensor* s, void* pixels, int where) {
    B2_REQUIRE(s && s->bound && pixels, "b2_sensor_read_image: no image bound");
    b2_ctx* ctx = s->ctx;
    B2_CUDA(cudaSetDevice(ctx->device));
    size_t bytes = (size_t)s->d.nx * s->d.ny * s->d.dtype_bytes;
    if (where == B2_HOST && bytes >= ((size_t)8 << 20) && bytes % 8 == 0 && getenv("B2_IMAGE_PLAIN_COPY") == nullptr) {
        // a full CCD into a pageable array: through the pinned ring (the driver's own staging runs at ~6 GB/s)
        double* hout[1] = {(double*)pixels};
        const double* dout[1] = {(const double*)s->d.target};
        return b2_pipe_run(ctx, (int64_t)(bytes / 8), 0, nullptr, nullptr, 1, hout, dout, nullptr);
    }
    B2_CUDA(cudaMemcpyAsync(pixels, s->d.target, bytes, where == B2_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, ctx->stream));
    if (where == B2_HOST) B2_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

static int launch_add_delta(b2_sensor* s, double sign, int clear) {
    B2_TIMED("k_add_delta", s->ctx->stream);
    DevSensor& d = s->d;
    size_t n = (size_t)d.nx * d.ny;
    int nb = (int)((n + 255) / 256);
    if (d.dtype_bytes == 4) k_add_delta<float><<<nb, 256, 0, s->ctx->stream>>>((float*)d.target, d.delta, n, sign, clear);
    else k_add_delta<double><<<nb, 256, 0, s->ctx->stream>>>((double*)d.target, d.delta, n, sign, clear);
    B2_CHECK_LAUNCH();
    return 0;
}

template <typename CT>
static int launch_update_tiled(b2_sensor* s, const CT* charge) {
    DevSensor& d = s->d;
    cudaStream_t st = s->ctx->stream;
    size_t smem = (size_t)24 * 64 * sizeof(double) + 24 * sizeof(unsigned long long);
    dim3 block(32, 8, 1);
    dim3 grid((d.nx + 1 + 31) / 32, (d.ny + 1 + 7) / 8, 1);
    static const int mode = getenv("B2_UPDATE_MODE") ? atoi(getenv("B2_UPDATE_MODE")) : 3;
    const size_t smem3 = smem + (size_t)d.nx9 * d.ny9 * ((d.nv + 3) + (d.nv + 1)) * sizeof(double2);
#define B2_TILED(NVV, MM, SM)                                                                                          \
    {                                                                                                                 \
        if ((SM) > 48 * 1024)                                                                                         \
            B2_CUDA(cudaFuncSetAttribute(k_update_distortions_tiled<CT, NVV, MM>,                                      \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SM)));                    \
        k_update_distortions_tiled<CT, NVV, MM><<<grid, block, (SM), st>>>(d, charge, s->changed);                    \
    }
    if (d.nv == 4) {
        if (mode == 1) B2_TILED(4, 1, smem) else if (mode == 2) B2_TILED(4, 2, smem) else if (mode == 3) B2_TILED(4, 3, smem3)
        else B2_TILED(4, 0, smem)
    } else {
        if (mode == 1) B2_TILED(8, 1, smem) else if (mode == 2) B2_TILED(8, 2, smem) else if (mode == 3) B2_TILED(8, 3, smem3)
        else B2_TILED(8, 0, smem)
    }
#undef B2_TILED
    B2_CHECK_LAUNCH();
    return 0;
}

// charge_from_target: distortions from the bound image (initialize) or from delta (update)
static int launch_update_distortions(b2_sensor* s, bool from_target) {
    DevSensor& d = s->d;
    b2_ctx* ctx = s->ctx;
    cudaStream_t st = ctx->stream;
    B2_TIMED("update_distortions(total)", st);
    B2_CUDA(cudaMemsetAsync(s->changed, 0, (size_t)d.nx * d.ny, st));
    const bool tiled = (d.nv == 4 || d.nv == 8) && d.qdist <= 3 && getenv("B2_UPDATE_GENERIC") == nullptr;
    if (tiled) {
        if (!from_target) return launch_update_tiled<double>(s, d.delta);
        if (d.dtype_bytes == 4) return launch_update_tiled<float>(s, (const float*)d.target);
        return launch_update_tiled<double>(s, (const double*)d.target);
    }
    B2_CUDA(cudaMemsetAsync(s->tiles, 0, (size_t)s->tnx * s->tny, st));
    dim3 gt = grid2(d.nx, d.ny, 256);
    size_t smem = ((size_t)d.nx9 * d.ny9 * (2 * d.nv + 2)) * sizeof(float2);
    dim3 g = grid2(d.nx + 1, d.ny + 1, 128);
    if (!from_target) {
        k_charge_tiles<double><<<gt, 256, 0, st>>>(d.delta, d.nx, d.ny, s->tnx, s->tiles);
        B2_CHECK_LAUNCH();
        if (smem > 48 * 1024) B2_CUDA(cudaFuncSetAttribute(k_update_distortions<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_update_distortions<double><<<g, 128, smem, st>>>(d, d.delta, s->changed, s->tiles, s->tnx, s->tny);
    } else if (d.dtype_bytes == 4) {
        k_charge_tiles<float><<<gt, 256, 0, st>>>((const float*)d.target, d.nx, d.ny, s->tnx, s->tiles);
        B2_CHECK_LAUNCH();
        if (smem > 48 * 1024) B2_CUDA(cudaFuncSetAttribute(k_update_distortions<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_update_distortions<float><<<g, 128, smem, st>>>(d, (const float*)d.target, s->changed, s->tiles, s->tnx, s->tny);
    } else {
        k_charge_tiles<double><<<gt, 256, 0, st>>>((const double*)d.target, d.nx, d.ny, s->tnx, s->tiles);
        B2_CHECK_LAUNCH();
        if (smem > 48 * 1024) B2_CUDA(cudaFuncSetAttribute(k_update_distortions<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_update_distortions<double><<<g, 128, smem, st>>>(d, (const double*)d.target, s->changed, s->tiles, s->tnx, s->tny);
    }
    B2_CHECK_LAUNCH();
    return 0;
}

static int launch_bounds_update(b2_sensor* s, int all) {
    B2_TIMED("k_update_bounds", s->ctx->stream);
    DevSensor& d = s->d;
    const bool tiled = (d.nv == 4 || d.nv == 8) && d.qdist <= 3 && getenv("B2_UPDATE_GENERIC") == nullptr;
    if (tiled && d.nv == 4 && getenv("B2_BOUNDS_GENERIC") == nullptr)
        k_update_bounds_t<4><<<grid2(d.nx, d.ny, 256), 256, 0, s->ctx->stream>>>(d, s->changed, all);
    else if (tiled && d.nv == 8 && getenv("B2_BOUNDS_GENERIC") == nullptr)
        k_update_bounds_t<8><<<grid2(d.nx, d.ny, 256), 256, 0, s->ctx->stream>>>(d, s->changed, all);
    else
        k_update_bounds<<<grid2(d.nx, d.ny, 256), 256, 0, s->ctx->stream>>>(d, s->changed, all, tiled ? nullptr : s->tiles,
                                                                           s->tnx, s->tny);
    B2_CHECK_LAUNCH();
    return 0;
}

// Silicon::update
static int sensor_update(b2_sensor* s) {
    if (launch_update_distortions(s, false)) return 1;
    if (launch_bounds_update(s, 0)) return 1;
    if (launch_add_delta(s, 1.0, 1)) return 1;
    return 0;
}

static int init_boundaries(b2_sensor* s, int ocx, int ocy) {
    B2_TIMED("k_init_boundaries", s->ctx->stream);
    DevSensor& d = s->d;
    k_init_boundaries<<<grid2(d.nx + 1, d.ny + 1, 256), 256, 0, s->ctx->stream>>>(d, ocx, ocy);
    B2_CHECK_LAUNCH();
    return 0;
}

// Silicon::initialize
static int sensor_initialize(b2_sensor* s, int ocx, int ocy) {
    DevSensor& d = s->d;
    if (init_boundaries(s, ocx, ocy)) return 1;
    if (launch_update_distortions(s, true)) return 1;
    if (launch_bounds_update(s, 1)) return 1;
    B2_CUDA(cudaMemsetAsync(d.delta, 0, (size_t)d.nx * d.ny * sizeof(double), s->ctx->stream));
    s->accum_flux = 0.0;
    s->initialized = true;
    return 0;
}

extern "C" int b2_sensor_accumulate(b2_sensor* s, int64_t n, const double* x, const double* y, const double* dxdz,
                                    const double* dydz, const double* wl, const double* flux, const double* rand4,
                                    uint64_t seed, uint64_t offset, int32_t ocx, int32_t ocy, int32_t resume,
                                    int32_t recalc, int where, B2AccumStats* stats) {
    B2_REQUIRE(s && s->bound, "b2_sensor_accumulate: no image bound");
    B2_REQUIRE(n == 0 || (x && y && flux), "b2_sensor_accumulate: null photon array");
    B2_REQUIRE((dxdz == nullptr) == (dydz == nullptr), "b2_sensor_accumulate: dxdz and dydz go together");
    B2_REQUIRE(!wl || s->d.nabs > 0, "b2_sensor_accumulate: wavelengths given but the sensor has no absorption table");
    // galsim/sensor.py: resume=True needs the image of the previous call
    B2_REQUIRE(!resume || s->initialized, "b2_sensor_accumulate: resume=True but there was no previous accumulate on this image");
    b2_ctx* ctx = s->ctx;
    B2_CUDA(cudaSetDevice(ctx->device));
    DevSensor& d = s->d;
    cudaStream_t st = ctx->stream;
    if (stats) memset(stats, 0, sizeof(*stats));
    B2_CUDA(cudaMemsetAsync(s->dstats, 0, ST_N * sizeof(unsigned long long) + 64, st));
    uint64_t n_updates = 0;
    if (!resume) {
        if (sensor_initialize(s, ocx, ocy)) return 1;
    } else {
        if (launch_add_delta(s, -1.0, 0)) return 1;  // subtractDelta
        if (recalc) {
            if (sensor_update(s)) return 1;
            s->accum_flux = 0.0;
            n_updates++;
        }
    }
    // stage host photons
    const double *dx = x, *dy = y, *da = dxdz, *db = dydz, *dw = wl, *df = flux, *dr = rand4;
    if (where == B2_HOST && n > 0) {
        Stager sg{ctx};
        size_t narr = 3 + (dxdz ? 2 : 0) + (wl ? 1 : 0) + (rand4 ? 4 : 0);
        if (sg.init(narr * pad256(n * 8))) return 1;
        const double* hin[10];
        double* din[10];
        int nin = 0;
        auto stage = [&](const double* h, const double** d) {
            double* t = sg.take<double>(n);
            hin[nin] = h;
            din[nin++] = t;
            *d = t;
        };
        stage(x, &dx);
        stage(y, &dy);
        stage(flux, &df);
        if (dxdz) {
            stage(dxdz, &da);
            stage(dydz, &db);
        }
        if (wl) stage(wl, &dw);
        if (rand4) {
            // keep the [4][n] layout contiguous
            double* t = (double*)(sg.base + sg.off);
            sg.off += pad256((size_t)4 * n * 8);
            for (int j = 0; j < 4; ++j) {
                hin[nin] = rand4 + (size_t)j * n;
                din[nin++] = t + (size_t)j * n;
            }
            dr = t;
        }
        if (b2_pipe_enabled(n)) {
            // large pageable arrays: uploaded through the pinned ring by the host copy threads (hostpipe.cu)
            if (b2_pipe_run(ctx, n, nin, hin, din, 0, nullptr, nullptr, nullptr)) return 1;
        } else {
            for (int f = 0; f < nin; ++f) H2D(din[f], hin[f], n);
        }
    }
    // chunk boundaries (photon order) at the nrecalc cadence
    std::vector<int64_t> bounds;
    double nrecalc = s->cfg.nrecalc;
    if (nrecalc > 0 && n > 0) {
        const int max_bounds = 1 << 20;
        const int64_t nseg = (n + CHUNK_SEG - 1) / CHUNK_SEG;
        if (b2_scratch_reserve(ctx, s->cum, (size_t)max_bounds * 8 + 64 + (size_t)nseg * 8)) return 1;
        int64_t* dbounds = (int64_t*)s->cum.ptr;
        int* dnb = (int*)((char*)s->cum.ptr + (size_t)max_bounds * 8);
        double* dacc = (double*)((char*)s->cum.ptr + (size_t)max_bounds * 8 + 8);
        double* dsums = (double*)((char*)s->cum.ptr + (size_t)max_bounds * 8 + 64);
        {
            B2_TIMED("k_find_chunks", st);
            k_segment_sums<<<(unsigned)nseg, 256, 0, st>>>(df, n, dsums);
            B2_CHECK_LAUNCH();
            k_find_chunks<<<1, 1024, 0, st>>>(df, n, dsums, nseg, s->accum_flux, nrecalc, dbounds, max_bounds, dnb, dacc);
            B2_CHECK_LAUNCH();
        }
        int nb = 0;
        double acc = 0.0;
        B2_CUDA(cudaMemcpyAsync(&nb, dnb, sizeof(int), cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaMemcpyAsync(&acc, dacc, sizeof(double), cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
        B2_REQUIRE(nb <= max_bounds, "b2_sensor_accumulate: more than 2^20 boundary updates in one call");
        bounds.resize(nb);
        if (nb) {
            B2_CUDA(cudaMemcpyAsync(bounds.data(), dbounds, (size_t)nb * 8, cudaMemcpyDeviceToHost, st));
            B2_CUDA(cudaStreamSynchronize(st));
        }
        s->accum_flux = acc;
    }
    {
        // worst case every photon of the largest chunk needs the slow path
        int64_t maxchunk = n, prev = 0;
        if (!bounds.empty()) {
            maxchunk = 0;
            for (int64_t bnd : bounds) { maxchunk = std::max(maxchunk, bnd - prev); prev = bnd; }
            maxchunk = std::max(maxchunk, n - prev);
        }
        if (maxchunk > 0 && b2_scratch_reserve(ctx, s->slow, (size_t)maxchunk * sizeof(SlowRec))) return 1;
    }
    int64_t i1 = 0;
    size_t ib = 0;
    while (i1 < n) {
        int64_t i2 = (ib < bounds.size()) ? bounds[ib] : n;
        bool hit = ib < bounds.size();
        int64_t cnt = i2 - i1;
        if (cnt > 0) {
            {
                B2_TIMED("k_accumulate", st);
                B2_CUDA(cudaMemsetAsync(s->dnslow, 0, sizeof(unsigned long long), st));
                k_accumulate<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>(d, i1, i2, n, dx, dy, da, db, dw, df, dr,
                                                                             seed, offset, s->dstats, s->dadded,
                                                                             (SlowRec*)s->slow.ptr, s->dnslow);
                B2_CHECK_LAUNCH();
            }
            {
                B2_TIMED("k_accumulate_slow", st);
                int64_t want = (cnt + 255) / 256;
                unsigned blocks = (unsigned)(want < (int64_t)s->sm_count * 8 ? want : (int64_t)s->sm_count * 8);
                launch_slow(blocks, st, d, (const SlowRec*)s->slow.ptr, s->dnslow, s->dstats,
                                                          s->dadded);
                B2_CHECK_LAUNCH();
            }
        }
        if (hit) {
            if (sensor_update(s)) return 1;
            n_updates++;
            ib++;
        }
        i1 = i2;
    }
    if (launch_add_delta(s, 1.0, 0)) return 1;  // addDelta (pending charge stays in delta)
    if (stats || where == B2_HOST) {
        unsigned long long h[ST_N];
        double added = 0.0;
        B2_CUDA(cudaMemcpyAsync(h, s->dstats, sizeof(h), cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaMemcpyAsync(&added, s->dadded, sizeof(double), cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
        if (stats) {
            stats->added_flux = added;
            stats->n_polygon_tests = h[ST_POLY];
            stats->n_neighbor_search = h[ST_NEIGH];
            stats->n_not_found = h[ST_NOTFOUND];
            stats->n_boundary_1e9 = h[ST_B9];
            stats->n_dropped_bottom = h[ST_DROP];
            stats->n_updates = n_updates;
        }
    }
    return 0;
}

// ---- pieces of the accumulate protocol used by the fused pool step (pool.cu) ----
int b2_sensor_begin_accumulate(b2_sensor* s, int32_t ocx, int32_t ocy, int32_t resume, int32_t recalc, int64_t n,
                               uint64_t* n_updates) {
    B2_REQUIRE(s && s->bound, "accumulate: no image bound");
    B2_REQUIRE(!resume || s->initialized, "accumulate: resume=True but there was no previous accumulate on this image");
    b2_ctx* ctx = s->ctx;
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    B2_CUDA(cudaMemsetAsync(s->dstats, 0, ST_N * sizeof(unsigned long long) + 64, st));
    *n_updates = 0;
    if (!resume) {
        if (sensor_initialize(s, ocx, ocy)) return 1;
    } else {
        if (launch_add_delta(s, -1.0, 0)) return 1;  // subtractDelta
        if (recalc) {
            if (sensor_update(s)) return 1;
            s->accum_flux = 0.0;
            (*n_updates)++;
        }
    }
    if (n > 0 && b2_scratch_reserve(ctx, s->slow, (size_t)n * sizeof(SlowRec))) return 1;
    B2_CUDA(cudaMemsetAsync(s->dnslow, 0, sizeof(unsigned long long), st));
    return 0;
}

int b2_sensor_run_slow(b2_sensor* s, int64_t n) {
    cudaStream_t st = s->ctx->stream;
    B2_TIMED("k_accumulate_slow", st);
    int64_t want = (n + 255) / 256;
    unsigned blocks = (unsigned)(want < (int64_t)s->sm_count * 8 ? want : (int64_t)s->sm_count * 8);
    if (blocks == 0) return 0;
    launch_slow(blocks, st, s->d, (const SlowRec*)s->slow.ptr, s->dnslow, s->dstats, s->dadded);
    B2_CHECK_LAUNCH();
    return 0;
}

int b2_sensor_end_accumulate(b2_sensor* s) { return launch_add_delta(s, 1.0, 0); }

// Silicon::update right now (distortions from delta, image += delta, delta = 0)
int b2_sensor_update_now(b2_sensor* s) {
    if (sensor_update(s)) return 1;
    s->accum_flux = 0.0;
    return 0;
}

extern "C" int b2_plain_accumulate(b2_sensor* s, int64_t n, const double* x, const double* y, const double* flux,
                                   int where, double* added_flux) {
    B2_REQUIRE(s && s->bound, "b2_plain_accumulate: no image bound");
    b2_ctx* ctx = s->ctx;
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevSensor& d = s->d;
    B2_CUDA(cudaMemsetAsync(s->dadded, 0, sizeof(double), st));
    B2_CUDA(cudaMemsetAsync(d.delta, 0, (size_t)d.nx * d.ny * sizeof(double), st));
    s->initialized = false;
    if (n > 0) {
        const double *dx = x, *dy = y, *df = flux;
        if (where == B2_HOST) {
            Stager sg{ctx};
            if (sg.init(3 * pad256(n * 8))) return 1;
            double* t;
            t = sg.take<double>(n); H2D(t, x, n); dx = t;
            t = sg.take<double>(n); H2D(t, y, n); dy = t;
            t = sg.take<double>(n); H2D(t, flux, n); df = t;
        }
        if (d.dtype_bytes == 4) k_plain_accumulate<float><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d, n, dx, dy, df, s->dadded);
        else k_plain_accumulate<double><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d, n, dx, dy, df, s->dadded);
        B2_CHECK_LAUNCH();
        if (launch_add_delta(s, 1.0, 1)) return 1;
    }
    if (added_flux) {
        B2_CUDA(cudaMemcpyAsync(added_flux, s->dadded, sizeof(double), cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
    }
    return 0;
}

extern "C" int b2_sensor_pixel_areas(b2_sensor* s, int32_t ocx, int32_t ocy, int32_t use_flux, double* areas, int where) {
    B2_REQUIRE(s && s->bound && areas, "b2_sensor_pixel_areas: no image bound");
    b2_ctx* ctx = s->ctx;
    B2_CUDA(cudaSetDevice(ctx->device));
    DevSensor& d = s->d;
    if (init_boundaries(s, ocx, ocy)) return 1;
    if (use_flux && launch_update_distortions(s, true)) return 1;
    s->initialized = false;
    size_t npix = (size_t)d.nx * d.ny;
    double* dareas = areas;
    if (where == B2_HOST) {
        if (b2_scratch_reserve(ctx, ctx->scratch, npix * 8)) return 1;
        dareas = (double*)ctx->scratch.ptr;
    }
    k_pixel_areas<<<grid2(d.nx, d.ny, 256), 256, 0, ctx->stream>>>(d, dareas);
    B2_CHECK_LAUNCH();
    if (where == B2_HOST) {
        B2_CUDA(cudaMemcpyAsync(areas, dareas, npix * 8, cudaMemcpyDeviceToHost, ctx->stream));
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return 0;
}

__global__ void k_get_pixel(const __grid_constant__ DevSensor s, int ax, int ay, double* poly, double* bounds) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int nv = s.nv, npoly = 4 * nv + 4;
    const int start = npoly - nv / 2;  // walk index of polygon vertex 0
    int idx = 0;
    walk_polygon(s, ax, ay, [&](double px, double py, double, double) {
        int n = (idx - start + npoly) % npoly;
        poly[2 * n] = px;
        poly[2 * n + 1] = py;
        idx++;
    });
    size_t k = ((size_t)ay * s.nx + ax) * 4;
    for (int j = 0; j < 4; ++j) {
        bounds[j] = s.inner[k + j];
        bounds[4 + j] = s.outer[k + j];
    }
}

extern "C" int b2_sensor_get_pixel(b2_sensor* s, int32_t ix, int32_t iy, double* poly, double* bounds) {
    B2_REQUIRE(s && s->bound, "b2_sensor_get_pixel: no image bound");
    b2_ctx* ctx = s->ctx;
    B2_CUDA(cudaSetDevice(ctx->device));
    DevSensor& d = s->d;
    int ax = ix - d.xmin, ay = iy - d.ymin;
    B2_REQUIRE(ax >= 0 && ax < d.nx && ay >= 0 && ay < d.ny, "b2_sensor_get_pixel: pixel outside the image");
    int npoly = 4 * d.nv + 4;
    if (b2_scratch_reserve(ctx, ctx->scratch, (size_t)(2 * npoly + 8) * 8)) return 1;
    double* dp = (double*)ctx->scratch.ptr;
    k_get_pixel<<<1, 32, 0, ctx->stream>>>(d, ax, ay, dp, dp + 2 * npoly);
    B2_CHECK_LAUNCH();
    B2_CUDA(cudaMemcpyAsync(poly, dp, (size_t)2 * npoly * 8, cudaMemcpyDeviceToHost, ctx->stream));
    B2_CUDA(cudaMemcpyAsync(bounds, dp + 2 * npoly, 64, cudaMemcpyDeviceToHost, ctx->stream));
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---- test aid: the rounding primitive of the boundary update
__global__ void k_round_f32(int64_t n, const double* __restrict__ in, double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = round_to_f32(in[i]);
}

extern "C" int b2_test_round_f32(b2_ctx* ctx, int64_t n, const double* in, double* out) {
    B2_REQUIRE(ctx && n >= 0 && (n == 0 || (in && out)), "b2_test_round_f32: null argument");
    if (n == 0) return 0;
    B2_CUDA(cudaSetDevice(ctx->device));
    double* d = nullptr;
    B2_CUDA(cudaMalloc(&d, 2 * (size_t)n * sizeof(double)));
    B2_CUDA(cudaMemcpyAsync(d, in, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    k_round_f32<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, d, d + n);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, d + n, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFree(d);
    B2_CUDA(e);
    return 0;
}

extern "C" int b2_copy_through_ring(b2_ctx* ctx, void* host, void* device, int64_t bytes, int32_t to_device) {
    B2_REQUIRE(ctx && bytes >= 0 && (bytes == 0 || (host && device)), "b2_copy_through_ring: null argument");
    B2_REQUIRE(bytes % 8 == 0, "b2_copy_through_ring: the size must be a multiple of 8 bytes");
    if (bytes == 0) return 0;
    B2_CUDA(cudaSetDevice(ctx->device));
    if (bytes < ((int64_t)1 << 20) || getenv("B2_IMAGE_PLAIN_COPY")) {
        B2_CUDA(cudaMemcpyAsync(to_device ? device : host, to_device ? host : device, (size_t)bytes,
                                to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, ctx->stream));
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
        return 0;
    }
    if (to_device) {
        const double* hin[1] = {(const double*)host};
        double* din[1] = {(double*)device};
        return b2_pipe_run(ctx, bytes / 8, 1, hin, din, 0, nullptr, nullptr, nullptr);
    }
    double* hout[1] = {(double*)host};
    const double* dout[1] = {(const double*)device};
    return b2_pipe_run(ctx, bytes / 8, 0, nullptr, nullptr, 1, hout, dout, nullptr);
}

// dst[start[g] .. start[g+1]) = value[g]: a per-stamp constant field of a pooled upload written on the device
__global__ void k_fill_segments(int64_t n, int64_t nseg, const int64_t* __restrict__ start, const double* __restrict__ value,
                                double* __restrict__ dst) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int64_t lo = 0, hi = nseg;  // start[lo] <= i < start[hi]
    while (hi - lo > 1) {
        const int64_t mid = (lo + hi) >> 1;
        if (__ldg(start + mid) <= i) lo = mid; else hi = mid;
    }
    dst[i] = __ldg(value + lo);
}

extern "C" int b2_fill_segments(b2_ctx* ctx, int64_t nseg, const int64_t* seg_len, const double* value, double* dst) {
    B2_REQUIRE(ctx && nseg >= 0 && (nseg == 0 || (seg_len && value && dst)), "b2_fill_segments: null argument");
    if (nseg == 0) return 0;
    B2_CUDA(cudaSetDevice(ctx->device));
    std::vector<int64_t> start((size_t)nseg + 1, 0);
    for (int64_t g = 0; g < nseg; ++g) {
        B2_REQUIRE(seg_len[g] >= 0, "b2_fill_segments: negative segment length");
        start[g + 1] = start[g] + seg_len[g];
    }
    const int64_t n = start[nseg];
    if (n == 0) return 0;
    // the two small tables go to a scratch buffer of the context (stream-ordered: a later call may reuse it)
    const size_t sb = ((size_t)nseg + 1) * sizeof(int64_t), vb = (size_t)nseg * sizeof(double);
    if (b2_scratch_reserve(ctx, ctx->fill_scratch, sb + vb)) return 1;
    char* d = (char*)ctx->fill_scratch.ptr;
    B2_CUDA(cudaMemcpyAsync(d, start.data(), sb, cudaMemcpyHostToDevice, ctx->stream));
    B2_CUDA(cudaMemcpyAsync(d + sb, value, vb, cudaMemcpyHostToDevice, ctx->stream));
    k_fill_segments<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(n, nseg, (const int64_t*)d, (const double*)(d + sb), dst);
    B2_CHECK_LAUNCH();
    return 0;
}
