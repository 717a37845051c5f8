// b2_common.cuh -- shared host/device plumbing of libimsim_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <utility>
#include <vector>

#include "../../include/imsim_b200.h"

// ---------------------------------------------------------------- errors
extern thread_local std::string g_b2_error;
extern std::atomic<uint64_t> g_b2_launches;

static inline int b2_fail(const char* fmt, const char* a = "", const char* b = "") {
    char buf[1024];
    snprintf(buf, sizeof(buf), fmt, a, b);
    g_b2_error = buf;
    return 1;
}

#define B2_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (call);                                                               \
        if (_e != cudaSuccess) {                                                               \
            char _b[1024];                                                                     \
            snprintf(_b, sizeof(_b), "%s:%d: %s failed: %s", __FILE__, __LINE__, #call,        \
                     cudaGetErrorString(_e));                                                  \
            g_b2_error = _b;                                                                   \
            return 1;                                                                          \
        }                                                                                      \
    } while (0)

#define B2_CHECK_LAUNCH()                         \
    do {                                          \
        g_b2_launches.fetch_add(1);               \
        B2_CUDA(cudaGetLastError());              \
    } while (0)

// Optional per-kernel device timing (B2_TIMING=1): cudaEvents around each launch on the launching
// stream, resolved lazily by b2_timing_report().  Off by default: no events, no overhead.
bool b2_timing_enabled();
long b2_timing_begin(const char* name, cudaStream_t st);
void b2_timing_end(long slot, cudaStream_t st);
struct B2TimedScope {  // scopes may nest (a host entry point timing itself around timed launches)
    cudaStream_t st;
    long slot;
    B2TimedScope(const char* name, cudaStream_t s) : st(s), slot(-1) {
        if (b2_timing_enabled()) slot = b2_timing_begin(name, st);
    }
    ~B2TimedScope() {
        if (slot >= 0) b2_timing_end(slot, st);
    }
};
#define B2_TIMED(name, stream) B2TimedScope _b2_timed_scope(name, stream)

#define B2_REQUIRE(cond, msg)             \
    do {                                  \
        if (!(cond)) return b2_fail("%s", msg); \
    } while (0)

#define H2D(dst, src, cnt) B2_CUDA(cudaMemcpyAsync(dst, src, (cnt) * sizeof(double), cudaMemcpyHostToDevice, ctx->stream))
#define D2H(dst, src, cnt) B2_CUDA(cudaMemcpyAsync(dst, src, (cnt) * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream))

// ---------------------------------------------------------------- context
struct DevWcs {
    double crpix[2];
    double cd[4];
    double cdinv[4];
    double ab[2][10];   // packed triangle: index of (i,j), i+j<=3: see sip_index()
    double newton_tol;
    int order;
    int pad;
};

struct DevObsc {
    int kind, negate;
    double p[6];  // radii already squared where the test is on r^2
};

struct DevSurf {
    int kind, interact, med_in, med_out, n_coef, rot_identity, n_obsc, extra_kind;
    double R, k1, invR;
    double coef[B2_MAX_ASPHERE_COEF];
    double dr[3];
    double drot[9];
    DevObsc obsc[B2_MAX_OBSC];
    int poly_n, pad;
    double poly_scale;
    const double* extra;  // device pointer (poly coefficients or bicubic block)
    // fast path for the usual single centred Clear{Circle,Annulus}: vignetted unless in2 <= r^2 < out2
    int simple_clear, pad2;
    double clr_in2, clr_out2;
};

#define B2_DEV_MAX_SURF 16
#define B2_DEV_MAX_MEDIA 4

// pixel -> field-angle tangents compiled into one polynomial per detector (b2_xytov_compile)
#define B2_XYPOLY_N 6
struct DevXyPoly {
    int enabled, pad;
    double box[4];   // xlo, xhi, ylo, yhi [pixels]: where the fit was made and checked
    double c0[2], sc[2];
    double cx[B2_XYPOLY_N][B2_XYPOLY_N], cy[B2_XYPOLY_N][B2_XYPOLY_N];  // c[i][j] X^i Y^j
};

struct DevOptics {
    int n_surf, n_media, medium_stop, pad;
    DevSurf surf[B2_DEV_MAX_SURF];
    B2Medium media[B2_DEV_MAX_MEDIA];
    DevWcs img, field;
    double M_if[9];  // img tangent frame -> field tangent frame (row-major)
    DevXyPoly xyv;
    B2Detector det;
    B2Diffraction dif;
};

struct Scratch {
    void* ptr = nullptr;
    size_t bytes = 0;
};

struct b2_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool have_tel = false, have_wcs = false, have_det = false;
    int program = 0;         // surface program matching the uploaded telescope (optics_device.cuh), 0 = generic
    DevOptics opt;           // host copy, passed to kernels by value (__grid_constant__)
    // extra sag tables (b2_telescope_set_extra): one grow-only device buffer per surface, written in stream
    // order, fed from two alternating pinned staging slots so that re-uploading a telescope per detector
    // neither leaks nor synchronises the stream
    struct ExtraTable {
        Scratch dev;
        void* pin[2] = {nullptr, nullptr};
        size_t pin_bytes[2] = {0, 0};
        cudaEvent_t ev[2] = {nullptr, nullptr};
        int next = 0;
    } extras[B2_DEV_MAX_SURF];
    Scratch scratch;         // staging for B2_HOST calls
    Scratch stats;           // small device buffer for counters
    Scratch fill_scratch;    // segment tables of b2_fill_segments
    B2TanSip img_host, field_host;
    // live timing of the dominant kernel (bench.py roofline): event pairs around each launch
    bool record_events = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> events;
    std::vector<struct b2_sensor*> sensors;  // sensors created on this context and not yet destroyed
    struct b2_hostpipe* pipe = nullptr;  // hostpipe.cu: pinned ring + streams of the pipelined B2_HOST route
};

int b2_scratch_reserve(b2_ctx* ctx, Scratch& s, size_t bytes);

// hostpipe.cu -- pipelined staging of pageable host arrays.  Chunk by chunk: host threads copy `nin` caller
// arrays into pinned slots -> H2D into din[f] + off -> kernel(off, cnt, stream) (optional) -> D2H of `nout`
// arrays -> host threads copy them into hout[f] + off.  Everything has completed when it returns.
bool b2_pipe_enabled(int64_t n);
int b2_pipe_threads();
int b2_pipe_run(b2_ctx* ctx, int64_t n, int nin, const double* const* hin, double* const* din, int nout,
                double* const* hout, const double* const* dout,
                const std::function<int(int64_t, int64_t, cudaStream_t)>* kernel);
void b2_pipe_release(b2_ctx* ctx);
void b2_stage1_release(b2_ctx* ctx);
void b2_sensor_orphan(struct b2_sensor* s);  // sensor.cu  // stage1.cu: per-context PSF / profile tables

// stage helper for B2_HOST calls: carve arrays out of the context scratch
struct Stager {
    b2_ctx* ctx;
    char* base = nullptr;
    size_t off = 0, cap = 0;
    int init(size_t bytes) {
        if (b2_scratch_reserve(ctx, ctx->scratch, bytes)) return 1;
        base = (char*)ctx->scratch.ptr;
        cap = bytes;
        off = 0;
        return 0;
    }
    template <typename T>
    T* take(size_t n) {
        size_t b = (n * sizeof(T) + 255) & ~size_t(255);
        T* p = (T*)(base + off);
        off += b;
        return p;
    }
};
static inline size_t pad256(size_t b) { return (b + 255) & ~size_t(255); }

// ---------------------------------------------------------------- Philox4x32-10
struct Philox {
    uint32_t c[4];
    uint32_t k[2];
};

__host__ __device__ static inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
    uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
#else
    uint64_t p0 = (uint64_t)M0 * c[0], p1 = (uint64_t)M1 * c[2];
    uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
    uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
#endif
    uint32_t n0 = hi1 ^ c[1] ^ k[0];
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c[3] ^ k[1];
    uint32_t n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// 4 x 32 random bits for (seed, index, stream)
__host__ __device__ static inline void philox4(uint64_t seed, uint64_t index, uint32_t stream, uint32_t out[4]) {
    uint32_t c[4] = {(uint32_t)index, (uint32_t)(index >> 32), stream, 0x9E3779B9u};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        philox_round(c, k);
        k[0] += 0x9E3779B9u;
        k[1] += 0xBB67AE85u;
    }
    out[0] = c[0]; out[1] = c[1]; out[2] = c[2]; out[3] = c[3];
}

// uniform in (0,1): 52 random bits, never 0 or 1
__host__ __device__ static inline double u01(uint32_t a, uint32_t b) {
    uint64_t m = ((uint64_t)(a & 0xFFFFFu) << 32) | b;
    return ((double)m + 0.5) * (1.0 / 4503599627370496.0);
}
