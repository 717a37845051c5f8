// pool.cu -- the pooled hot loop as ONE kernel per photon batch (sm_100a).
//
// imsim/photon_pooling.py:141-160 applies, to every pooled sub-batch, TimeSampler,
// PupilAnnulusSampler, PhotonDCR, RubinDiffractionOptics, FocusDepth, Refraction (each a pass over
// the photon arrays) and then SiliconSensor.accumulate.  Here a thread carries one photon from its
// pooled pixel position all the way to the charge deposit: the FP64-bound ray trace hides the
// latency of the sensor's gathers and atomics, and the intermediate PhotonArray fields (time,
// pupil, dxdz, dydz, traced x/y) never touch HBM.  Photons that need the polygon / neighbour
// treatment go to the sensor's compact slow list exactly as in the unfused path, and the Philox
// streams are the same, so fused == unfused bit for bit (tests/test_gpu_pool.py).
#include "optics_device.cuh"
#include "sensor_device.cuh"

#include <algorithm>

struct PoolParams {
    double t0, exptime, r_in, r_out;
    uint64_t sampler_seed, sensor_seed, offset, sensor_offset;
    int write_back;  // also store the traced photons (x, y, dxdz, dydz, flux) like the unfused ops
};

template <int MINB, int PROG>
__global__ void __launch_bounds__(256, MINB)
k_pool_step(const __grid_constant__ DevOptics o, const __grid_constant__ B2OpticsOptions opt,
            const __grid_constant__ DevSensor s, const __grid_constant__ PoolParams pp, int64_t n,
            double* __restrict__ x, double* __restrict__ y, double* __restrict__ dxdz, double* __restrict__ dydz,
            double* __restrict__ flux, const double* __restrict__ wl_nm, unsigned long long* __restrict__ ostats,
            unsigned long long* __restrict__ sstats, double* __restrict__ added, SlowRec* __restrict__ slow,
            unsigned long long* __restrict__ nslow) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool active = i < n;
    bool vig = false, fail = false, offz = false, to_slow = false;
    unsigned nb9 = 0, ndrop = 0;
    double my_added = 0.0;
    SlowRec rec;
    if (active) {
        const uint64_t idx = pp.offset + (uint64_t)i;
        double time, pu, pv;
        sample_time_pupil(pp.sampler_seed, idx, pp.t0, pp.exptime, pp.r_in, pp.r_out, time, pu, pv);
        double g = o.dif.enabled ? philox_normal(opt.seed, opt.photon_offset + (uint64_t)i, 0u) : 0.0;
        const double wl = wl_nm[i];
        OpticsOut r = optics_photon<PROG>(o, opt, x[i], y[i], wl, pu, pv, time, flux[i], g);
        vig = r.vig;
        fail = r.fail;
        offz = r.offz;
        if (pp.write_back) {
            x[i] = r.x;
            y[i] = r.y;
            dxdz[i] = r.dxdz;
            dydz[i] = r.dydz;
            flux[i] = r.flux;
        }
        double g1, g2, unf, udep;
        sensor_draws(pp.sensor_seed, pp.sensor_offset + (uint64_t)i, g1, g2, unf, udep);
        to_slow = sensor_fast_path(s, r.x, r.y, true, r.dxdz, r.dydz, true, wl, r.flux, g1, g2, unf, udep, rec,
                                   my_added, nb9, ndrop);
    }
    slow_append(to_slow, rec, slow, nslow);
    unsigned nv = __popc(__ballot_sync(0xffffffffu, vig));
    unsigned nf = __popc(__ballot_sync(0xffffffffu, fail));
    unsigned nz = __popc(__ballot_sync(0xffffffffu, offz));
    unsigned long long w3 = warp_sum(nb9), w4 = warp_sum(ndrop);
    double wa = my_added;
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) wa += __shfl_xor_sync(0xffffffffu, wa, k);
    if ((threadIdx.x & 31) == 0) {
        if (nv) atomicAdd(&ostats[0], (unsigned long long)nv);
        if (nf) atomicAdd(&ostats[1], (unsigned long long)nf);
        if (nz) atomicAdd(&ostats[2], (unsigned long long)nz);
        if (w3) atomicAdd(&sstats[ST_B9], w3);
        if (w4) atomicAdd(&sstats[ST_DROP], w4);
        if (wa != 0.0) atomicAdd(added, wa);
    }
}

static int pool_occ(int program) {
    static int occ = -1;
    if (occ < 0) {
        const char* e = getenv("B2_POOL_OCC");
        occ = e ? atoi(e) : 0;
        if (occ < 2 || occ > 5) occ = 0;
    }
    if (occ) return (program == B2_PROG_LSST || occ <= 4) ? occ : 4;
    // measured on B200, ms per 2^25 photons at 2 / 3 / 4 (/ 5) blocks per SM: interpreter 8.2 / 7.0 / 6.7;
    // LSST program (no spills down to 64 registers) - / 6.25 / 6.19 / 5.77
    return program == B2_PROG_LSST ? 5 : 4;
}

// One photon batch of the pooled pipeline on device-resident arrays:
//   TimeSampler + PupilAnnulusSampler -> [PhotonDCR] -> RubinDiffractionOptics -> FocusDepth ->
//   Refraction -> SiliconSensor.accumulate(photons, image, resume, recalc)
extern "C" int b2_pool_step(b2_ctx* ctx, b2_sensor* sensor, int64_t n, double* x, double* y, double* dxdz,
                            double* dydz, double* flux, const double* wl_nm, const B2OpticsOptions* opt, double t0,
                            double exptime, double r_inner, double r_outer, uint64_t sampler_seed,
                            uint64_t sensor_seed, uint64_t photon_offset, uint64_t sensor_offset, int32_t resume,
                            int32_t recalc, int32_t write_back, B2OpticsStats* ostats, B2AccumStats* astats) {
    B2_REQUIRE(ctx && sensor && opt, "b2_pool_step: null argument");
    B2_REQUIRE(sensor->ctx == ctx, "b2_pool_step: the sensor belongs to another context");
    B2_REQUIRE(ctx->have_tel && ctx->have_wcs && ctx->have_det, "b2_pool_step: telescope / wcs / detector not uploaded");
    B2_REQUIRE(n == 0 || (x && y && flux && wl_nm), "b2_pool_step: null photon array");
    B2_REQUIRE(!write_back || (dxdz && dydz), "b2_pool_step: write_back needs dxdz / dydz arrays");
    B2_REQUIRE(sensor->cfg.nrecalc == 0.0,
               "b2_pool_step: the fused step is the pooled cadence (nrecalc = 0, recalc at batch starts); use "
               "b2_rubin_optics + b2_sensor_accumulate for nrecalc > 0");
    B2_REQUIRE(sensor->d.nabs > 0, "b2_pool_step: the sensor has no absorption table");
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    uint64_t n_updates = 0;
    if (b2_sensor_begin_accumulate(sensor, 0, 0, resume, recalc, n, &n_updates)) return 1;
    if (b2_scratch_reserve(ctx, ctx->stats, 64)) return 1;
    unsigned long long* dostats = (unsigned long long*)ctx->stats.ptr;
    B2_CUDA(cudaMemsetAsync(dostats, 0, 64, st));
    if (n > 0) {
        PoolParams pp{t0, exptime, r_inner, r_outer, sampler_seed, sensor_seed, photon_offset, sensor_offset, write_back};
        unsigned blocks = (unsigned)((n + 255) / 256);
        {
            B2_TIMED("k_pool_step", st);
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            if (ctx->record_events) {
                B2_CUDA(cudaEventCreate(&e0));
                B2_CUDA(cudaEventCreate(&e1));
                B2_CUDA(cudaEventRecord(e0, st));
            }
#define B2_LAUNCH_POOL(MINB, PROG)                                                                                \
    k_pool_step<MINB, PROG><<<blocks, 256, 0, st>>>(ctx->opt, *opt, sensor->d, pp, n, x, y, dxdz, dydz, flux, wl_nm,   \
                                                    dostats, sensor->dstats, sensor->dadded,                         \
                                                    (SlowRec*)sensor->slow.ptr, sensor->dnslow)
            const int occ = pool_occ(ctx->program);
            if (ctx->program == B2_PROG_LSST) {
                if (occ == 2) B2_LAUNCH_POOL(2, B2_PROG_LSST);
                else if (occ == 3) B2_LAUNCH_POOL(3, B2_PROG_LSST);
                else if (occ == 5) B2_LAUNCH_POOL(5, B2_PROG_LSST);
                else B2_LAUNCH_POOL(4, B2_PROG_LSST);
            } else {
                if (occ == 2) B2_LAUNCH_POOL(2, B2_PROG_GENERIC);
                else if (occ == 3) B2_LAUNCH_POOL(3, B2_PROG_GENERIC);
                else B2_LAUNCH_POOL(4, B2_PROG_GENERIC);
            }
#undef B2_LAUNCH_POOL
            B2_CHECK_LAUNCH();
            if (ctx->record_events) {
                B2_CUDA(cudaEventRecord(e1, st));
                ctx->events.emplace_back(e0, e1);
            }
        }
        if (b2_sensor_run_slow(sensor, n)) return 1;
    }
    if (b2_sensor_end_accumulate(sensor)) return 1;
    if (ostats || astats) {
        unsigned long long ho[3], hs[ST_N];
        double added = 0.0;
        B2_CUDA(cudaMemcpyAsync(ho, dostats, sizeof(ho), cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaMemcpyAsync(hs, sensor->dstats, sizeof(hs), cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaMemcpyAsync(&added, sensor->dadded, sizeof(double), cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
        if (ostats) {
            ostats->n_vignetted = ho[0];
            ostats->n_failed = ho[1];
            ostats->n_offdetector_z = ho[2];
        }
        if (astats) {
            astats->added_flux = added;
            astats->n_polygon_tests = hs[ST_POLY];
            astats->n_neighbor_search = hs[ST_NEIGH];
            astats->n_not_found = hs[ST_NOTFOUND];
            astats->n_boundary_1e9 = hs[ST_B9];
            astats->n_dropped_bottom = hs[ST_DROP];
            astats->n_updates = n_updates;
        }
    }
    return 0;
}

// ------------------------------------------------------------------ photon-shot flat, fused
// imsim/flat.py:239-264 shoots, per iteration, Poisson(counts * area) photons uniform over the bordered
// section, samples their wavelengths and calls sensor.accumulate.  Uniform photons in random order make
// every inner-box gather and every charge deposit a random DRAM access.  Here the photons of one
// iteration are generated tile by tile (32 x 32 pixels) with per-tile Poisson counts drawn on the host --
// the same distribution as N uniform photons in any order -- and go straight from registers to the
// charge deposit: no photon array exists, gathers and atomics stay in one tile's cache lines.
struct FlatParams {
    int tiles_x, tiles_y, tile;      // tile grid over the bound image
    int tile0;                       // first tile of this launch
    double xlo, xhi, ylo, yhi;       // photon rectangle (image bounds +- 0.5)
    uint64_t seed, sensor_seed, offset;
    int ncdf;
};

__global__ void __launch_bounds__(256)
k_flat_step(const __grid_constant__ DevSensor s, const __grid_constant__ FlatParams fp,
            const int64_t* __restrict__ tile_cum, const double* __restrict__ cdf, const double* __restrict__ cdf_wave,
            unsigned long long* __restrict__ sstats, double* __restrict__ added, SlowRec* __restrict__ slow,
            unsigned long long* __restrict__ nslow) {
    const int t = fp.tile0 + blockIdx.x;
    const int tx = t % fp.tiles_x, ty = t / fp.tiles_x;
    // tile rectangle clipped to the photon rectangle
    const double x0 = fmax(fp.xlo, fp.xlo + (double)tx * fp.tile), x1 = fmin(fp.xhi, fp.xlo + (double)(tx + 1) * fp.tile);
    const double y0 = fmax(fp.ylo, fp.ylo + (double)ty * fp.tile), y1 = fmin(fp.yhi, fp.ylo + (double)(ty + 1) * fp.tile);
    // blockIdx.y: slice of the tile's photons (keeps every SM busy when a pass holds few tiles)
    const int64_t tfirst = tile_cum[t], tcount = tile_cum[t + 1] - tfirst;
    const int64_t per = ((tcount + gridDim.y - 1) / gridDim.y + 31) & ~(int64_t)31;
    const int64_t first = tfirst + per * blockIdx.y;
    const int64_t last = (first + per < tfirst + tcount) ? first + per : tfirst + tcount;
    unsigned nb9 = 0, ndrop = 0;
    double my_added = 0.0;
    // whole warps iterate together (ballots in slow_append)
    for (int64_t base = first + (threadIdx.x & ~31); base < last; base += blockDim.x) {
        int64_t i = base + (threadIdx.x & 31);
        bool to_slow = false;
        SlowRec rec;
        if (i < last) {
            const uint64_t idx = fp.offset + (uint64_t)i;
            uint32_t r[4];
            philox4(fp.seed, idx, 5u, r);
            double x = x0 + (x1 - x0) * u01(r[0], r[1]);
            double y = y0 + (y1 - y0) * u01(r[2], r[3]);
            double g1, g2, unf, udep;
            sensor_draws(fp.sensor_seed, idx, g1, g2, unf, udep);
            double wl = 0.0;
            const bool has_wl = fp.ncdf >= 2;
            if (has_wl) {
                uint32_t q[4];
                philox4(fp.seed, idx, 6u, q);
                double u = u01(q[0], q[1]);
                int lo = 0, hi = fp.ncdf - 1;
                while (hi - lo > 1) {
                    int mid = (lo + hi) >> 1;
                    if (__ldg(cdf + mid) <= u) lo = mid; else hi = mid;
                }
                double c0 = __ldg(cdf + lo), c1 = __ldg(cdf + hi);
                double f = (c1 > c0) ? (u - c0) / (c1 - c0) : 0.0;
                wl = __ldg(cdf_wave + lo) + f * (__ldg(cdf_wave + hi) - __ldg(cdf_wave + lo));
            }
            double add1 = 0.0;
            unsigned b9 = 0, dr = 0;
            to_slow = sensor_fast_path(s, x, y, false, 0.0, 0.0, has_wl, wl, 1.0, g1, g2, unf, udep, rec, add1, b9, dr);
            my_added += add1;
            nb9 += b9;
            ndrop += dr;
        }
        slow_append(to_slow, rec, slow, nslow);
    }
    unsigned long long w3 = warp_sum(nb9), w4 = warp_sum(ndrop);
    double wa = my_added;
#pragma unroll
    for (int k = 16; k > 0; k >>= 1) wa += __shfl_xor_sync(0xffffffffu, wa, k);
    if ((threadIdx.x & 31) == 0) {
        if (w3) atomicAdd(&sstats[ST_B9], w3);
        if (w4) atomicAdd(&sstats[ST_DROP], w4);
        if (wa != 0.0) atomicAdd(added, wa);
    }
}

// One iteration of the photon-shot flat on the sensor's bound image: tile_cum (HOST, int64,
// tiles_x*tiles_y + 1 entries, tile_cum[0] = 0) holds the cumulative per-tile photon counts of this
// iteration.  The tiles are processed in passes of at most FLAT_PASS_PHOTONS photons so that the
// compact list of slow-path photons stays bounded however large the section is.
static const int64_t FLAT_PASS_PHOTONS = (int64_t)1 << 27;

extern "C" int b2_flat_step(b2_ctx* ctx, b2_sensor* sensor, const int64_t* tile_cum, int64_t n_total, int32_t tile,
                            const double* cdf, const double* cdf_wave, int32_t ncdf, uint64_t seed,
                            uint64_t sensor_seed, uint64_t photon_offset, int32_t resume, int32_t update_after,
                            B2AccumStats* astats) {
    B2_REQUIRE(ctx && sensor && tile_cum, "b2_flat_step: null argument");
    B2_REQUIRE(sensor->ctx == ctx && sensor->bound, "b2_flat_step: sensor not bound to an image of this context");
    B2_REQUIRE(tile >= 8 && tile <= 256, "b2_flat_step: tile must be 8..256 pixels");
    B2_REQUIRE(ncdf == 0 || (cdf && cdf_wave && ncdf >= 2), "b2_flat_step: bad wavelength CDF");
    B2_REQUIRE(ncdf == 0 || sensor->d.nabs > 0, "b2_flat_step: the sensor has no absorption table");
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    DevSensor& d = sensor->d;
    uint64_t n_updates = 0;
    FlatParams fp;
    fp.tile = tile;
    fp.tiles_x = (d.nx + tile - 1) / tile;
    fp.tiles_y = (d.ny + tile - 1) / tile;
    const int ntiles = fp.tiles_x * fp.tiles_y;
    B2_REQUIRE(tile_cum[0] == 0 && tile_cum[ntiles] == n_total, "b2_flat_step: tile_cum does not sum to n_total");
    // largest pass, for the slow-list reservation
    int64_t max_pass = 0;
    for (int a = 0; a < ntiles;) {
        int b = a + 1;
        while (b < ntiles && tile_cum[b + 1] - tile_cum[a] <= FLAT_PASS_PHOTONS) b++;
        max_pass = std::max(max_pass, tile_cum[b] - tile_cum[a]);
        a = b;
    }
    // chunking at the nrecalc cadence is done by the caller (whole iterations between updates)
    if (b2_sensor_begin_accumulate(sensor, 0, 0, resume, 0, max_pass, &n_updates)) return 1;
    if (b2_scratch_reserve(ctx, ctx->scratch, (size_t)(ntiles + 1) * sizeof(int64_t))) return 1;
    int64_t* dcum = (int64_t*)ctx->scratch.ptr;
    B2_CUDA(cudaMemcpyAsync(dcum, tile_cum, (size_t)(ntiles + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    fp.xlo = d.xmin - 0.5;
    fp.xhi = d.xmin + d.nx - 0.5;
    fp.ylo = d.ymin - 0.5;
    fp.yhi = d.ymin + d.ny - 0.5;
    fp.seed = seed;
    fp.sensor_seed = sensor_seed;
    fp.offset = photon_offset;
    fp.ncdf = ncdf;
    for (int a = 0; a < ntiles && n_total > 0;) {
        int b = a + 1;
        while (b < ntiles && tile_cum[b + 1] - tile_cum[a] <= FLAT_PASS_PHOTONS) b++;
        const int64_t n_pass = tile_cum[b] - tile_cum[a];
        fp.tile0 = a;
        if (n_pass > 0) {
            if (a > 0) B2_CUDA(cudaMemsetAsync(sensor->dnslow, 0, sizeof(unsigned long long), st));
            {
                B2_TIMED("k_flat_step", st);
                cudaEvent_t e0 = nullptr, e1 = nullptr;
                if (ctx->record_events) {
                    B2_CUDA(cudaEventCreate(&e0));
                    B2_CUDA(cudaEventCreate(&e1));
                    B2_CUDA(cudaEventRecord(e0, st));
                }
                int split = (sensor->sm_count * 16 + (b - a) - 1) / (b - a);
                split = std::max(1, std::min(split, 64));
                k_flat_step<<<dim3(b - a, split), 256, 0, st>>>(d, fp, dcum, cdf, cdf_wave, sensor->dstats, sensor->dadded,
                                                   (SlowRec*)sensor->slow.ptr, sensor->dnslow);
                B2_CHECK_LAUNCH();
                if (ctx->record_events) {
                    B2_CUDA(cudaEventRecord(e1, st));
                    ctx->events.emplace_back(e0, e1);
                }
            }
            if (b2_sensor_run_slow(sensor, n_pass)) return 1;
        }
        a = b;
    }
    if (update_after) {
        if (b2_sensor_update_now(sensor)) return 1;
        n_updates++;
    }
    if (b2_sensor_end_accumulate(sensor)) return 1;
    if (astats) {
        unsigned long long hs[ST_N];
        double addedv = 0.0;
        B2_CUDA(cudaMemcpyAsync(hs, sensor->dstats, sizeof(hs), cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaMemcpyAsync(&addedv, sensor->dadded, sizeof(double), cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
        memset(astats, 0, sizeof(*astats));
        astats->added_flux = addedv;
        astats->n_polygon_tests = hs[ST_POLY];
        astats->n_neighbor_search = hs[ST_NEIGH];
        astats->n_not_found = hs[ST_NOTFOUND];
        astats->n_boundary_1e9 = hs[ST_B9];
        astats->n_dropped_bottom = hs[ST_DROP];
        astats->n_updates = n_updates;
    }
    return 0;
}
