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

struct PoolParams {
    double t0, exptime, r_in, r_out;
    uint64_t sampler_seed, sensor_seed, offset;
    int write_back;  // also store the traced photons (x, y, dxdz, dydz, flux) like the unfused ops
};

template <int MINB>
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
        OpticsOut r = optics_photon(o, opt, x[i], y[i], wl, pu, pv, time, flux[i], g);
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
        sensor_draws(pp.sensor_seed, idx, g1, g2, unf, udep);
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

static int pool_occ() {
    static int occ = -1;
    if (occ < 0) {
        const char* e = getenv("B2_POOL_OCC");
        occ = e ? atoi(e) : 4;  // measured on B200: 8.2 / 7.0 / 6.7 ms per 2^25 photons at 2 / 3 / 4 blocks per SM
        if (occ < 2 || occ > 4) occ = 4;
    }
    return occ;
}

// One photon batch of the pooled pipeline on device-resident arrays:
//   TimeSampler + PupilAnnulusSampler -> [PhotonDCR] -> RubinDiffractionOptics -> FocusDepth ->
//   Refraction -> SiliconSensor.accumulate(photons, image, resume, recalc)
extern "C" int b2_pool_step(b2_ctx* ctx, b2_sensor* sensor, int64_t n, double* x, double* y, double* dxdz,
                            double* dydz, double* flux, const double* wl_nm, const B2OpticsOptions* opt, double t0,
                            double exptime, double r_inner, double r_outer, uint64_t sampler_seed,
                            uint64_t sensor_seed, uint64_t photon_offset, int32_t resume, int32_t recalc,
                            int32_t write_back, B2OpticsStats* ostats, B2AccumStats* astats) {
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
        PoolParams pp{t0, exptime, r_inner, r_outer, sampler_seed, sensor_seed, photon_offset, write_back};
        unsigned blocks = (unsigned)((n + 255) / 256);
        {
            B2_TIMED("k_pool_step", st);
            cudaEvent_t e0 = nullptr, e1 = nullptr;
            if (ctx->record_events) {
                B2_CUDA(cudaEventCreate(&e0));
                B2_CUDA(cudaEventCreate(&e1));
                B2_CUDA(cudaEventRecord(e0, st));
            }
            switch (pool_occ()) {
                case 2:
                    k_pool_step<2><<<blocks, 256, 0, st>>>(ctx->opt, *opt, sensor->d, pp, n, x, y, dxdz, dydz, flux, wl_nm, dostats, sensor->dstats, sensor->dadded, (SlowRec*)sensor->slow.ptr, sensor->dnslow);
                    break;
                case 4:
                    k_pool_step<4><<<blocks, 256, 0, st>>>(ctx->opt, *opt, sensor->d, pp, n, x, y, dxdz, dydz, flux, wl_nm, dostats, sensor->dstats, sensor->dadded, (SlowRec*)sensor->slow.ptr, sensor->dnslow);
                    break;
                default:
                    k_pool_step<3><<<blocks, 256, 0, st>>>(ctx->opt, *opt, sensor->d, pp, n, x, y, dxdz, dydz, flux, wl_nm, dostats, sensor->dstats, sensor->dadded, (SlowRec*)sensor->slow.ptr, sensor->dnslow);
            }
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
