// readout.cu -- the post-path electronics on the device (sm_100a), SURVEY.md section 8 f4:
// bleed trails (imsim/bleed_trails.py:26-147), dark current, amplifier split / gain / flips, intra-CCD
// crosstalk, prescan / overscan, charge-transfer inefficiency, bias and read noise, int32 raw segments
// (imsim/readout.py:391-480).  The e-image stays in HBM from the last photon to the amp segments.
//
// Rounding follows the reference's numpy expressions (float32 images, float64 where numpy >= 2 promotes):
// bleed trails, crosstalk and CTE are bit-identical to the reference's own functions (tests/golden/
// readout.npz); dark current and read noise use Philox streams (statistical parity).
#include "b2_common.cuh"

#include <utility>

// ------------------------------------------------------------------ bleed trails
// One thread per (column, half): a channel is a sequential redistribution problem (each saturated run
// spills alternately downwards and upwards until its excess is gone), columns are independent.  Threads
// of a warp hold adjacent columns, so every row access is one coalesced line.
__global__ void __launch_bounds__(128)
k_bleed(float* __restrict__ e, int nx, int ny, double full_well, int midline_stop) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int half = blockIdx.y;
    if (x >= nx) return;
    int ylo = 0, yhi = ny;
    if (midline_stop) {
        const int ymid = ny / 2;
        ylo = half ? ymid : 0;
        yhi = half ? ny : ymid;
    } else if (half) {
        return;
    }
    const int n = yhi - ylo;
    float* c = e + (size_t)ylo * nx + x;  // c[k * nx]
    const float fw32 = (float)full_well;
    // quick reject: the reference only touches channels with a saturated pixel
    bool any = false;
    for (int k = 0; k < n; ++k) any |= (c[(size_t)k * nx] > fw32);
    if (!any) return;
    // The runs are those of the ORIGINAL channel (np.diff(padded > full_well) before any change).  Bleeding
    // never lifts a pixel above full well, and a pixel of a later run that an earlier run reaches is set
    // to full well by "bled = min(full_well - value, excess)" with a negative room -- still inside that
    // later run's original extent.  So scanning for "> full_well" as we go would miss such a pixel; keep
    // the original saturation flags in a bit mask instead (channels are <= 4096 rows).
    unsigned mask[128];
    const int nwords = (n + 31) >> 5;
    if (nwords > 128) return;  // guarded on the host
    for (int w = 0; w < nwords; ++w) {
        unsigned m = 0u;
        for (int b = 0; b < 32; ++b) {
            int k = (w << 5) + b;
            if (k < n && c[(size_t)k * nx] > fw32) m |= 1u << b;
        }
        mask[w] = m;
    }
    int y = 0;
    while (y < n) {
        if (!((mask[y >> 5] >> (y & 31)) & 1u)) {
            ++y;
            continue;
        }
        const int y0 = y;
        float s = 0.0f;  // sum() of float32 scalars
        while (y < n && ((mask[y >> 5] >> (y & 31)) & 1u)) {
            s = __fadd_rn(s, c[(size_t)y * nx]);
            ++y;
        }
        const int y1 = y;
        double excess = __dsub_rn((double)s, __dmul_rn((double)(y1 - y0), full_well));
        for (int k = y0; k < y1; ++k) c[(size_t)k * nx] = fw32;
        const int reach = max(y0, n - y1);
        for (int dy = 0; dy < reach; ++dy) {
#pragma unroll
            for (int side = 0; side < 2; ++side) {
                const int yp = side == 0 ? y0 - dy - 1 : y1 + dy;
                if (yp >= 0 && yp < n) {
                    float v = c[(size_t)yp * nx];
                    float room = __fsub_rn(fw32, v);
                    if ((double)room <= excess) {  // min(room, excess) -> room
                        c[(size_t)yp * nx] = __fadd_rn(v, room);
                        excess = __dsub_rn(excess, (double)room);
                    } else {
                        c[(size_t)yp * nx] = (float)__dadd_rn((double)v, excess);
                        excess = 0.0;
                    }
                } else if (yp < 0) {
                    excess = __dsub_rn(excess, fmin(full_well, excess));
                }
                if (excess == 0.0) break;
            }
            if (excess == 0.0) break;
        }
    }
}

extern "C" int b2_bleed_trails(b2_ctx* ctx, float* eimage, int32_t nx, int32_t ny, double full_well,
                               int32_t midline_stop, int where) {
    B2_REQUIRE(ctx && eimage && nx > 0 && ny > 0, "b2_bleed_trails: bad argument");
    B2_REQUIRE(ny <= 4096, "b2_bleed_trails: channels of up to 4096 rows");
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    float* d = eimage;
    size_t bytes = (size_t)nx * ny * sizeof(float);
    if (where == B2_HOST) {
        if (b2_scratch_reserve(ctx, ctx->scratch, bytes)) return 1;
        d = (float*)ctx->scratch.ptr;
        B2_CUDA(cudaMemcpyAsync(d, eimage, bytes, cudaMemcpyHostToDevice, st));
    }
    {
        B2_TIMED("k_bleed", st);
        k_bleed<<<dim3((nx + 127) / 128, 2), 128, 0, st>>>(d, nx, ny, full_well, midline_stop);
        B2_CHECK_LAUNCH();
    }
    if (where == B2_HOST) {
        B2_CUDA(cudaMemcpyAsync(eimage, d, bytes, cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
    }
    return 0;
}

// ------------------------------------------------------------------ Poisson deviates
// Exact Poisson sampler, counter based: inversion by sequential search below a mean of 30, Hoermann's
// transformed rejection (PTRS, 1993 -- the algorithm behind numpy's and, via boost, GalSim's large-mean
// generators) above.  Uniforms come from Philox(seed, index, stream) with the rejection round in the
// high counter word, so a pixel's deviate depends only on (seed, pixel index).
__device__ __forceinline__ double poisson_exact(double lam, uint64_t seed, uint64_t index, uint32_t stream) {
    if (!(lam > 0.0)) return 0.0;
    uint32_t r[4];
    if (lam < 30.0) {
        philox4(seed, index, stream, r);
        const double u = u01(r[0], r[1]);
        double p = exp(-lam), cdf = p;
        int k = 0;
        while (u > cdf && k < 400) {
            ++k;
            p *= lam / (double)k;
            cdf += p;
        }
        return (double)k;
    }
    const double slam = sqrt(lam), loglam = log(lam);
    const double b = 0.931 + 2.53 * slam, a = -0.059 + 0.02483 * b;
    const double invalpha = 1.1239 + 1.1328 / (b - 3.4), vr = 0.9277 - 3.6224 / (b - 2.0);
    for (uint32_t round = 0; round < 64; ++round) {
        philox4(seed ^ ((uint64_t)round << 40), index, stream + 1u, r);
        const double U = u01(r[0], r[1]) - 0.5, V = u01(r[2], r[3]);
        const double us = 0.5 - fabs(U);
        const double k = floor((2.0 * a / us + b) * U + lam + 0.43);
        if (us >= 0.07 && V <= vr) return k;
        if (k < 0.0 || (us < 0.013 && V > us)) continue;
        if (log(V) + log(invalpha) - log(a / (us * us) + b) <= -lam + k * loglam - lgamma(k + 1.0)) return k;
    }
    return rint(lam);  // 64 rejections in a row: probability ~ 1e-60
}

// image += Poisson(sky_level * area * modulation): LSST_ImageBuilderBase.addNoise (imsim/lsst_image.py:128-199:
// "image += sky", sky optionally multiplied by gradient / vignetting / fringing maps) followed by the config
// CCD-noise builder, which for photon-shot images only has the sky's shot noise left to add; `areas` are the
// tree-ring / brighter-fatter pixel areas of sensor.calculate_pixel_areas when the sky is drawn through the
// sensor model (config/imsim-config.yaml:222-228).
template <typename T>
__global__ void __launch_bounds__(256)
k_add_sky(T* __restrict__ image, size_t n, double sky_level, const double* __restrict__ areas,
          const float* __restrict__ modulation, uint64_t seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double mean = sky_level;
    if (areas) mean *= areas[i];
    if (modulation) mean *= (double)modulation[i];
    image[i] = (T)((double)image[i] + poisson_exact(mean, seed, (uint64_t)i, 22u));
}

extern "C" int b2_add_sky(b2_ctx* ctx, void* image, int32_t dtype_bytes, int64_t npix, double sky_level,
                          const double* areas, const float* modulation, uint64_t seed) {
    B2_REQUIRE(ctx && image && npix > 0, "b2_add_sky: bad argument");
    B2_REQUIRE(dtype_bytes == 4 || dtype_bytes == 8, "b2_add_sky: image must be float32 or float64");
    B2_REQUIRE(sky_level >= 0.0, "b2_add_sky: negative sky level");
    B2_CUDA(cudaSetDevice(ctx->device));
    B2_TIMED("k_add_sky", ctx->stream);
    unsigned nb = (unsigned)((npix + 255) / 256);
    if (dtype_bytes == 4) k_add_sky<float><<<nb, 256, 0, ctx->stream>>>((float*)image, (size_t)npix, sky_level, areas, modulation, seed);
    else k_add_sky<double><<<nb, 256, 0, ctx->stream>>>((double*)image, (size_t)npix, sky_level, areas, modulation, seed);
    B2_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------ cosmic rays
// image[iy][ix] += value for a list of pixels with numpy's indexing rules: CosmicRays.paint_cr
// (imsim/cosmic_rays.py:74-111) adds span values inside "try: image_array[y, x] += value / except IndexError",
// so a negative index wraps around and only an index beyond the array is skipped.
template <typename T>
__global__ void __launch_bounds__(256)
k_scatter_add(T* __restrict__ image, int nx, int ny, int64_t n, const int32_t* __restrict__ iy,
              const int32_t* __restrict__ ix, const float* __restrict__ values) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int y = iy[i], x = ix[i];
    if (y < -ny || y >= ny || x < -nx || x >= nx) return;  // IndexError -> pass
    if (y < 0) y += ny;
    if (x < 0) x += nx;
    atomicAdd(&image[(size_t)y * nx + x], (T)values[i]);
}

extern "C" int b2_scatter_add(b2_ctx* ctx, void* image, int32_t dtype_bytes, int32_t nx, int32_t ny, int64_t n,
                              const int32_t* iy, const int32_t* ix, const float* values) {
    B2_REQUIRE(ctx && image && nx > 0 && ny > 0, "b2_scatter_add: bad argument");
    B2_REQUIRE(dtype_bytes == 4 || dtype_bytes == 8, "b2_scatter_add: image must be float32 or float64");
    if (n <= 0) return 0;
    B2_REQUIRE(iy && ix && values, "b2_scatter_add: null pixel list");
    B2_CUDA(cudaSetDevice(ctx->device));
    B2_TIMED("k_scatter_add", ctx->stream);
    unsigned nb = (unsigned)((n + 255) / 256);
    if (dtype_bytes == 4) k_scatter_add<float><<<nb, 256, 0, ctx->stream>>>((float*)image, nx, ny, n, iy, ix, values);
    else k_scatter_add<double><<<nb, 256, 0, ctx->stream>>>((double*)image, nx, ny, n, iy, ix, values);
    B2_CHECK_LAUNCH();
    return 0;
}

// ------------------------------------------------------------------ readout
// small-mean Poisson by inversion (dark current: 0.02 e-/s x 32 s), Gaussian approximation above 64
__device__ __forceinline__ double poisson_draw(double mean, double u, double g) {
    if (mean <= 0.0) return 0.0;
    if (mean > 64.0) return fmax(0.0, rint(mean + sqrt(mean) * g));
    double p = exp(-mean), cdf = p;
    int k = 0;
    while (u > cdf && k < 1024) {
        ++k;
        p *= mean / (double)k;
        cdf += p;
    }
    return (double)k;
}

// dark current added to the e-image in place (ImageF += float64 array)
__global__ void __launch_bounds__(256)
k_dark_current(float* __restrict__ e, size_t n, double mean, uint64_t seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t r[4];
    philox4(seed, (uint64_t)i, 20u, r);
    double u = u01(r[0], r[1]);
    double g = 0.0;
    if (mean > 64.0) {
        double u2 = u01(r[2], r[3]);
        g = sqrt(-2.0 * log(u)) * cospi(2.0 * u2);
    }
    e[i] = (float)((double)e[i] + poisson_draw(mean, u, g));
}

// amp split + gain + flips (+ crosstalk) into the raw segment's imaging area; prescan / overscan = 0
__global__ void __launch_bounds__(256)
k_amp_segments(const float* __restrict__ e, int nx, const B2Amp* __restrict__ amps, int namp,
               const double* __restrict__ xtalk, float* __restrict__ seg, int raw_nx, int raw_ny) {
    const int a = blockIdx.z;
    const int ix = blockIdx.x * blockDim.x + threadIdx.x;  // position in readout order inside the imaging area
    const int iy = blockIdx.y;
    const B2Amp A = amps[a];
    if (ix >= A.nx || iy >= A.ny) return;
    // value of amp b at the same readout position (all amps share the imaging-area shape)
    auto amp_value = [&](const B2Amp& B) -> float {
        int sx = B.flip_x ? (B.nx - 1 - ix) : ix;
        int sy = B.flip_y ? (B.ny - 1 - iy) : iy;
        return __fdiv_rn(e[(size_t)(B.y0 + sy) * nx + (B.x0 + sx)], (float)B.gain);
    };
    const float own = amp_value(A);
    float out = own;
    if (xtalk) {
        // amp_arrays[a] + sum([x * y for x, y in zip(amp_arrays, xtalk_row)]): float64, left to right from 0
        double acc = 0.0;
        for (int b = 0; b < namp; ++b)
            acc = __dadd_rn(acc, __dmul_rn((double)amp_value(amps[b]), xtalk[a * namp + b]));
        out = (float)__dadd_rn((double)own, acc);
    }
    seg[((size_t)a * raw_ny + (A.data_y0 + iy)) * raw_nx + (A.data_x0 + ix)] = out;
}

// one direction of the charge-transfer inefficiency: out_i = sum_{k = nt .. 0} band[i][k] * in_{i-k} in
// float64, stored as float32 (cte_matrix @ vector assigned into an ImageF).  stride_i / stride_o: element
// strides along / across the transfer direction.
__global__ void __launch_bounds__(256)
k_cte(const float* __restrict__ in, float* __restrict__ out, const double* __restrict__ band, int n_along, int n_across,
      size_t stride_along, size_t stride_across, size_t seg_stride, int nt, int thread_along) {
    // the thread index always runs over the unit-stride dimension so that every tap is a coalesced line
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = thread_along ? t : blockIdx.y;  // along the transfer
    const int j = thread_along ? blockIdx.y : t;  // across
    if (i >= n_along || j >= n_across) return;
    const float* src = in + (size_t)blockIdx.z * seg_stride + (size_t)j * stride_across;
    double acc = 0.0;
    const int kmax = min(nt, i);
    for (int k = kmax; k >= 0; --k)
        acc = __dadd_rn(acc, __dmul_rn(band[(size_t)i * (nt + 1) + k], (double)src[(size_t)(i - k) * stride_along]));
    out[(size_t)blockIdx.z * seg_stride + (size_t)j * stride_across + (size_t)i * stride_along] = (float)acc;
}

// bias + read noise + conversion to int32 (np.array(float32, dtype=np.int32): truncation)
__global__ void __launch_bounds__(256)
k_digitize(const float* __restrict__ seg, int32_t* __restrict__ out, const B2Amp* __restrict__ amps, size_t per_amp,
           size_t n, uint64_t seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const B2Amp A = amps[i / per_amp];
    float v = __fadd_rn(seg[i], (float)A.bias_level);
    if (A.read_noise > 0.0) {
        uint32_t r[4];
        philox4(seed, (uint64_t)i, 21u, r);
        double g = sqrt(-2.0 * log(u01(r[0], r[1]))) * cospi(2.0 * u01(r[2], r[3]));
        v = (float)((double)v + A.read_noise * g);
    }
    out[i] = (int32_t)v;
}

extern "C" int b2_readout(b2_ctx* ctx, float* eimage, int32_t nx, int32_t ny, const B2Amp* amps, int32_t namp,
                          const double* xtalk, const double* pband, const double* sband, int32_t ntransfers,
                          double full_well, int32_t midline_stop, double dark_mean, uint64_t seed,
                          float* segments, int32_t* raw) {
    B2_REQUIRE(ctx && eimage && amps && namp >= 1 && namp <= 64, "b2_readout: bad argument");
    B2_REQUIRE(segments || raw, "b2_readout: nothing to write");
    B2_CUDA(cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const int raw_nx = amps[0].raw_nx, raw_ny = amps[0].raw_ny;
    for (int a = 0; a < namp; ++a) {
        const B2Amp& A = amps[a];
        B2_REQUIRE(A.raw_nx == raw_nx && A.raw_ny == raw_ny && A.nx == amps[0].nx && A.ny == amps[0].ny,
                   "b2_readout: all amplifier segments must have the same shape");
        B2_REQUIRE(A.x0 >= 0 && A.y0 >= 0 && A.x0 + A.nx <= nx && A.y0 + A.ny <= ny, "b2_readout: amp outside the image");
        B2_REQUIRE(A.data_x0 >= 0 && A.data_y0 >= 0 && A.data_x0 + A.nx <= raw_nx && A.data_y0 + A.ny <= raw_ny,
                   "b2_readout: imaging area outside the raw segment");
        B2_REQUIRE(A.gain > 0.0, "b2_readout: gain must be positive");
    }
    const size_t per_amp = (size_t)raw_nx * raw_ny, nseg = per_amp * namp;
    // scratch: amps | xtalk | pband | sband | segment buffers A, B
    const size_t o_amps = 0, o_xt = pad256(sizeof(B2Amp) * namp), o_pb = o_xt + pad256(sizeof(double) * namp * namp),
                 o_sb = o_pb + pad256(sizeof(double) * raw_ny * (ntransfers + 1)),
                 o_a = o_sb + pad256(sizeof(double) * raw_nx * (ntransfers + 1)), o_b = o_a + pad256(nseg * sizeof(float));
    if (b2_scratch_reserve(ctx, ctx->scratch, o_b + pad256(nseg * sizeof(float)))) return 1;
    char* base = (char*)ctx->scratch.ptr;
    B2Amp* damps = (B2Amp*)(base + o_amps);
    double *dxt = (double*)(base + o_xt), *dpb = (double*)(base + o_pb), *dsb = (double*)(base + o_sb);
    float *sa = (float*)(base + o_a), *sb = (float*)(base + o_b);
    B2_CUDA(cudaMemcpyAsync(damps, amps, sizeof(B2Amp) * namp, cudaMemcpyHostToDevice, st));
    if (xtalk) B2_CUDA(cudaMemcpyAsync(dxt, xtalk, sizeof(double) * namp * namp, cudaMemcpyHostToDevice, st));
    if (pband) B2_CUDA(cudaMemcpyAsync(dpb, pband, sizeof(double) * raw_ny * (ntransfers + 1), cudaMemcpyHostToDevice, st));
    if (sband) B2_CUDA(cudaMemcpyAsync(dsb, sband, sizeof(double) * raw_nx * (ntransfers + 1), cudaMemcpyHostToDevice, st));
    if (full_well > 0.0) {
        B2_REQUIRE(ny <= 4096, "b2_readout: bleed trails need channels of up to 4096 rows");
        B2_TIMED("k_bleed", st);
        k_bleed<<<dim3((nx + 127) / 128, 2), 128, 0, st>>>(eimage, nx, ny, full_well, midline_stop);
        B2_CHECK_LAUNCH();
    }
    if (dark_mean > 0.0) {
        B2_TIMED("k_dark_current", st);
        size_t n = (size_t)nx * ny;
        k_dark_current<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(eimage, n, dark_mean, seed);
        B2_CHECK_LAUNCH();
    }
    {
        B2_TIMED("k_amp_segments", st);
        B2_CUDA(cudaMemsetAsync(sa, 0, nseg * sizeof(float), st));
        k_amp_segments<<<dim3((amps[0].nx + 255) / 256, amps[0].ny, namp), 256, 0, st>>>(eimage, nx, damps, namp,
                                                                                         xtalk ? dxt : nullptr, sa, raw_nx,
                                                                                         raw_ny);
        B2_CHECK_LAUNCH();
    }
    float* cur = sa;
    float* oth = sb;
    if (pband) {  // parallel transfers: along rows of the segment (y), columns independent
        B2_TIMED("k_cte", st);
        k_cte<<<dim3((raw_nx + 255) / 256, raw_ny, namp), 256, 0, st>>>(cur, oth, dpb, raw_ny, raw_nx, (size_t)raw_nx, 1,
                                                                      per_amp, ntransfers, 0);
        B2_CHECK_LAUNCH();
        std::swap(cur, oth);
    }
    if (sband) {  // serial transfers: along x, rows independent
        B2_TIMED("k_cte", st);
        k_cte<<<dim3((raw_nx + 255) / 256, raw_ny, namp), 256, 0, st>>>(cur, oth, dsb, raw_nx, raw_ny, 1, (size_t)raw_nx,
                                                                      per_amp, ntransfers, 1);
        B2_CHECK_LAUNCH();
        std::swap(cur, oth);
    }
    if (segments) B2_CUDA(cudaMemcpyAsync(segments, cur, nseg * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (raw) {
        B2_TIMED("k_digitize", st);
        k_digitize<<<(unsigned)((nseg + 255) / 256), 256, 0, st>>>(cur, raw, damps, per_amp, nseg, seed);
        B2_CHECK_LAUNCH();
    }
    return 0;
}
