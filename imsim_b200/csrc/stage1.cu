// stage1.cu -- stage 1 of the photon-shooting path on the device (sm_100a): catalogue objects ->
// pooled photons.
//
// In the reference every object is drawn by GalSim on the host (imsim/stamp.py:727-743,
// drawImage(method='phot') with the PSF photon ops first, stamp.py:550-553) and the stamps' PhotonArrays
// are concatenated (imsim/photon_pooling.py:177-192).  Here one thread per pooled photon finds its
// object in the cumulative count table, samples the object's profile, its wavelength and the PSF kicks
// (imsim/atmPSF.py:298-336: frozen-flow screens + SecondKick, config/imsim-config.yaml:239-256: optics
// Gaussian) and writes the pooled SoA directly: build_stamps + merge_photon_arrays without the host.
// The random streams are Philox, not GalSim's, so parity with the reference is statistical for this
// stage; parity with the numpy restatement used by the tests, on injected uniforms, is to rounding.
#include "b2_common.cuh"

#include <map>
#include <mutex>

#define ARCSEC_PER_RAD 206264.80624709636

struct DevStage1 {
    B2Psf psf;
    const void* screens[B2_MAX_SCREENS];
    const double* kick;   // second-kick radial table [n_kick]
    const double* luts;   // radial profile tables [n_lut][n_ent]
    int n_lut, n_ent;
    double lut_tmax;
};

struct b2_stage1_state {
    DevStage1 d;
    void* kick_mem = nullptr;
    void* lut_mem = nullptr;
};

static std::mutex g_stage1_mu;
static std::map<b2_ctx*, b2_stage1_state> g_stage1;

static b2_stage1_state& stage1_of(b2_ctx* ctx) {
    std::lock_guard<std::mutex> lk(g_stage1_mu);
    auto it = g_stage1.find(ctx);
    if (it == g_stage1.end()) {
        b2_stage1_state st;
        memset(&st.d, 0, sizeof(st.d));
        it = g_stage1.emplace(ctx, st).first;
    }
    return it->second;
}

// linear interpolation in a table of n entries tabulated on t in [0, tmax]; t = -log(1 - u)
__device__ __forceinline__ double table_at(const double* __restrict__ tab, int n, double tmax, double u) {
    double t = -log1p(-u);
    double g = fmin(t, tmax) * ((double)(n - 1) / tmax);
    int i = min((int)g, n - 2);
    double f = g - (double)i;
    double a = __ldg(tab + i), b = __ldg(tab + i + 1);
    return a + f * (b - a);
}

template <typename T>
__device__ __forceinline__ void screen_gradient(const T* __restrict__ tab, int npix, double inv_scale, double X, double Y,
                                                double& gx, double& gy) {
    // galsim.LookupTable2D(..., interpolant='linear', edge_mode='wrap').gradient
    double ax = X * inv_scale, ay = Y * inv_scale;
    double fx0 = floor(ax), fy0 = floor(ay);
    double fx = ax - fx0, fy = ay - fy0;
    // wrap the cell index into [0, npix)
    double np_d = (double)npix;
    int ix0 = (int)(fx0 - np_d * floor(fx0 / np_d));
    int iy0 = (int)(fy0 - np_d * floor(fy0 / np_d));
    ix0 = min(max(ix0, 0), npix - 1);
    iy0 = min(max(iy0, 0), npix - 1);
    int ix1 = ix0 + 1 == npix ? 0 : ix0 + 1;
    int iy1 = iy0 + 1 == npix ? 0 : iy0 + 1;
    const T* r0 = tab + (size_t)iy0 * npix;
    const T* r1 = tab + (size_t)iy1 * npix;
    double f00 = (double)__ldg(r0 + ix0), f10 = (double)__ldg(r0 + ix1);
    double f01 = (double)__ldg(r1 + ix0), f11 = (double)__ldg(r1 + ix1);
    gx = ((f10 - f00) * (1.0 - fy) + (f11 - f01) * fy) * inv_scale;
    gy = ((f01 - f00) * (1.0 - fx) + (f11 - f10) * fx) * inv_scale;
}

// quad-packed float32 screens: one 16-byte load holds the four corners of the cell
__device__ __forceinline__ void screen_gradient_quads(const float4* __restrict__ tab, int npix, double inv_scale, double X,
                                                      double Y, double& gx, double& gy) {
    double ax = X * inv_scale, ay = Y * inv_scale;
    double fx0 = floor(ax), fy0 = floor(ay);
    double fx = ax - fx0, fy = ay - fy0;
    double np_d = (double)npix;
    int ix0 = (int)(fx0 - np_d * floor(fx0 / np_d));
    int iy0 = (int)(fy0 - np_d * floor(fy0 / np_d));
    ix0 = min(max(ix0, 0), npix - 1);
    iy0 = min(max(iy0, 0), npix - 1);
    const float4 q = __ldg(tab + (size_t)iy0 * npix + ix0);
    const double f00 = (double)q.x, f10 = (double)q.y, f01 = (double)q.z, f11 = (double)q.w;
    gx = ((f10 - f00) * (1.0 - fy) + (f11 - f01) * fy) * inv_scale;
    gy = ((f01 - f00) * (1.0 - fx) + (f11 - f10) * fx) * inv_scale;
}

__global__ void __launch_bounds__(256)
k_stage1_photons(const __grid_constant__ DevStage1 s, int64_t n, double* __restrict__ x, double* __restrict__ y,
                 double* __restrict__ flux, double* __restrict__ wl, const B2Object* __restrict__ objects,
                 const int64_t* __restrict__ obj_cum, int nobj, const double* __restrict__ cdf,
                 const double* __restrict__ cdf_wave, int ncdf, const double* __restrict__ rand, uint64_t seed,
                 uint64_t offset) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // object j owns photons [obj_cum[j], obj_cum[j+1])
    int lo = 0, hi = nobj;
    while (hi - lo > 1) {
        int mid = (lo + hi) >> 1;
        if (__ldg(obj_cum + mid) <= i) lo = mid; else hi = mid;
    }
    const B2Object ob = objects[lo];
    double r[B2_STAGE1_NRAND];
    if (rand) {
#pragma unroll
        for (int k = 0; k < B2_STAGE1_NRAND; ++k) r[k] = rand[(size_t)k * n + i];
    } else {
        const uint64_t idx = offset + (uint64_t)i;
#pragma unroll
        for (int q = 0; q < B2_STAGE1_NRAND / 2; ++q) {
            if (q >= 2 && s.psf.n_screens == 0 && s.psf.n_kick == 0 && s.psf.gauss_sigma == 0.0) {
                r[2 * q] = r[2 * q + 1] = 0.5;
                continue;
            }
            uint32_t w[4];
            philox4(seed, idx, 7u + (uint32_t)q, w);
            r[2 * q] = u01(w[0], w[1]);
            r[2 * q + 1] = u01(w[2], w[3]);
        }
    }
    // ---- profile offset in the units of ob.m
    double dx = 0.0, dy = 0.0;
    switch (ob.kind) {
        case B2_PROF_GAUSSIAN: {
            double rad = sqrt(-2.0 * log(r[0]));
            double sn, cs;
            sincospi(2.0 * r[1], &sn, &cs);
            dx = rad * cs;
            dy = rad * sn;
            break;
        }
        case B2_PROF_RADIAL: {
            double rho = (s.luts != nullptr && ob.lut >= 0 && ob.lut < s.n_lut)
                             ? table_at(s.luts + (size_t)ob.lut * s.n_ent, s.n_ent, s.lut_tmax, r[0])
                             : 0.0;
            double sn, cs;
            sincospi(2.0 * r[1], &sn, &cs);
            dx = rho * cs;
            dy = rho * sn;
            break;
        }
        case B2_PROF_KNOTS: {
            // galsim.RandomKnots: the knot positions are a Gaussian random walk frozen per object
            uint64_t k = (uint64_t)(r[2] * (double)ob.n_knots);
            if (k >= (uint64_t)ob.n_knots) k = ob.n_knots - 1;
            uint32_t w[4];
            philox4(ob.knot_seed, k, 11u, w);
            double rad = sqrt(-2.0 * log(u01(w[0], w[1]))) * (1.0 / 1.1774100225154747);  // hlr -> sigma
            double sn, cs;
            sincospi(2.0 * u01(w[2], w[3]), &sn, &cs);
            dx = rad * cs;
            dy = rad * sn;
            break;
        }
        case B2_PROF_BOX:
            dx = (r[0] - 0.5) * ob.p0;
            dy = (r[1] - 0.5) * ob.p1;
            break;
        default:
            break;
    }
    double px = ob.x + (ob.m[0] * dx + ob.m[1] * dy);
    double py = ob.y + (ob.m[2] * dx + ob.m[3] * dy);
    // ---- wavelength: inverse CDF of this object's SED x bandpass
    double wave = 0.0;
    if (ncdf >= 2) {
        const double* c = cdf + (size_t)ob.sed * ncdf;
        const double* cw = cdf_wave + (size_t)ob.sed * ncdf;
        double u = r[3];
        int a = 0, b = ncdf - 1;
        while (b - a > 1) {
            int mid = (a + b) >> 1;
            if (__ldg(c + mid) <= u) a = mid; else b = mid;
        }
        double c0 = __ldg(c + a), c1 = __ldg(c + b);
        double f = (c1 > c0) ? (u - c0) / (c1 - c0) : 0.0;
        wave = __ldg(cw + a) + f * (__ldg(cw + b) - __ldg(cw + a));
    }
    // ---- PSF kicks [arcsec]
    double kx = 0.0, ky = 0.0;
    if (s.psf.n_screens > 0) {
        // the shooter's own pupil position and time (PhaseScreenPSF._shoot)
        double ri2 = s.psf.r_inner * s.psf.r_inner, ro2 = s.psf.r_outer * s.psf.r_outer;
        double rr = sqrt(ri2 + (ro2 - ri2) * r[4]);
        double sn, cs;
        sincospi(2.0 * r[5], &sn, &cs);
        double pu = rr * cs, pv = rr * sn;
        double t = s.psf.t0 + s.psf.exptime * r[6];
        const double tx = ob.tanx, ty = ob.tany;
        double gx = 0.0, gy = 0.0;
        const double inv_scale = 1.0 / s.psf.screen_scale;
        for (int l = 0; l < s.psf.n_screens; ++l) {
            double X = pu - s.psf.vx[l] * t + s.psf.altitude[l] * tx;
            double Y = pv - s.psf.vy[l] * t + s.psf.altitude[l] * ty;
            double ax, ay;
            if (s.psf.screen_f32 == 2) screen_gradient_quads((const float4*)s.screens[l], s.psf.npix, inv_scale, X, Y, ax, ay);
            else if (s.psf.screen_f32 == 1) screen_gradient((const float*)s.screens[l], s.psf.npix, inv_scale, X, Y, ax, ay);
            else screen_gradient((const double*)s.screens[l], s.psf.npix, inv_scale, X, Y, ax, ay);
            gx += ax;
            gy += ay;
        }
        // wavefront gradient [nm / m] -> angle; chromatic dilation of the atmospheric part
        double chrom = (ncdf >= 2 && s.psf.exponent != 0.0) ? exp(s.psf.exponent * log(wave / s.psf.base_wavelength)) : 1.0;
        double sc = 1e-9 * ARCSEC_PER_RAD * chrom;
        kx += gx * sc;
        ky += gy * sc;
    }
    if (s.psf.n_kick > 0 && r[7] >= s.psf.kick_delta_prob) {
        double u = (r[7] - s.psf.kick_delta_prob) / (1.0 - s.psf.kick_delta_prob);
        double th = table_at(s.kick, s.psf.n_kick, s.psf.kick_tmax, u);
        double sn, cs;
        sincospi(2.0 * r[8], &sn, &cs);
        kx += th * cs;
        ky += th * sn;
    }
    if (s.psf.gauss_sigma > 0.0) {
        double rad = s.psf.gauss_sigma * sqrt(-2.0 * log(r[9]));
        double sn, cs;
        sincospi(2.0 * r[10], &sn, &cs);
        kx += rad * cs;
        ky += rad * sn;
    }
    px += s.psf.arcsec_to_pix[0] * kx + s.psf.arcsec_to_pix[1] * ky;
    py += s.psf.arcsec_to_pix[2] * kx + s.psf.arcsec_to_pix[3] * ky;
    x[i] = px;
    y[i] = py;
    flux[i] = 1.0;
    if (wl) wl[i] = wave;
}

// ------------------------------------------------------------------ host side
extern "C" int b2_psf_upload(b2_ctx* ctx, const B2Psf* psf, const void* const* screens, const double* kick_table) {
    B2_REQUIRE(ctx && psf, "b2_psf_upload: null argument");
    B2_REQUIRE(psf->n_screens >= 0 && psf->n_screens <= B2_MAX_SCREENS, "b2_psf_upload: 0..8 screens");
    B2_REQUIRE(psf->n_screens == 0 || (screens && psf->npix >= 2 && psf->screen_scale > 0.0), "b2_psf_upload: bad screens");
    B2_REQUIRE(psf->n_kick == 0 || (kick_table && psf->n_kick >= 2 && psf->kick_tmax > 0.0 && psf->kick_delta_prob < 1.0),
               "b2_psf_upload: bad second-kick table");
    B2_CUDA(cudaSetDevice(ctx->device));
    b2_stage1_state& st = stage1_of(ctx);
    st.d.psf = *psf;
    for (int l = 0; l < B2_MAX_SCREENS; ++l) st.d.screens[l] = (l < psf->n_screens) ? screens[l] : nullptr;
    for (int l = 0; l < psf->n_screens; ++l) B2_REQUIRE(screens[l], "b2_psf_upload: null screen table");
    if (st.kick_mem) {
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(st.kick_mem);
        st.kick_mem = nullptr;
    }
    st.d.kick = nullptr;
    if (psf->n_kick > 0) {
        B2_CUDA(cudaMalloc(&st.kick_mem, (size_t)psf->n_kick * sizeof(double)));
        B2_CUDA(cudaMemcpyAsync(st.kick_mem, kick_table, (size_t)psf->n_kick * sizeof(double), cudaMemcpyHostToDevice,
                                ctx->stream));
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
        st.d.kick = (const double*)st.kick_mem;
    }
    return 0;
}

extern "C" int b2_radial_luts_upload(b2_ctx* ctx, const double* lut, int32_t n_lut, int32_t n_entries, double tmax) {
    B2_REQUIRE(ctx && lut && n_lut >= 1 && n_entries >= 2 && tmax > 0.0, "b2_radial_luts_upload: bad argument");
    B2_CUDA(cudaSetDevice(ctx->device));
    b2_stage1_state& st = stage1_of(ctx);
    if (st.lut_mem) {
        B2_CUDA(cudaStreamSynchronize(ctx->stream));
        cudaFree(st.lut_mem);
        st.lut_mem = nullptr;
    }
    size_t bytes = (size_t)n_lut * n_entries * sizeof(double);
    B2_CUDA(cudaMalloc(&st.lut_mem, bytes));
    B2_CUDA(cudaMemcpyAsync(st.lut_mem, lut, bytes, cudaMemcpyHostToDevice, ctx->stream));
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    st.d.luts = (const double*)st.lut_mem;
    st.d.n_lut = n_lut;
    st.d.n_ent = n_entries;
    st.d.lut_tmax = tmax;
    return 0;
}

void b2_stage1_release(b2_ctx* ctx) {
    std::lock_guard<std::mutex> lk(g_stage1_mu);
    auto it = g_stage1.find(ctx);
    if (it == g_stage1.end()) return;
    if (it->second.kick_mem) cudaFree(it->second.kick_mem);
    if (it->second.lut_mem) cudaFree(it->second.lut_mem);
    g_stage1.erase(it);
}

extern "C" int b2_stage1_photons(b2_ctx* ctx, int64_t n, double* x, double* y, double* flux, double* wl,
                                 const B2Object* objects, const int64_t* obj_cum, int32_t nobj, const double* cdf,
                                 const double* cdf_wave, int32_t n_sed, int32_t ncdf, const double* rand, uint64_t seed,
                                 uint64_t photon_offset) {
    B2_REQUIRE(ctx, "b2_stage1_photons: null context");
    if (n <= 0) return 0;
    B2_REQUIRE(x && y && flux && objects && obj_cum && nobj >= 1, "b2_stage1_photons: null argument");
    B2_REQUIRE(ncdf == 0 || (cdf && cdf_wave && ncdf >= 2 && n_sed >= 1 && wl), "b2_stage1_photons: bad wavelength tables");
    B2_CUDA(cudaSetDevice(ctx->device));
    b2_stage1_state& st = stage1_of(ctx);
    B2_TIMED("k_stage1_photons", ctx->stream);
    k_stage1_photons<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(st.d, n, x, y, flux, wl, objects, obj_cum, nobj,
                                                                         cdf, cdf_wave, ncdf, rand, seed, photon_offset);
    B2_CHECK_LAUNCH();
    return 0;
}
