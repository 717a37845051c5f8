// abi.cu -- context management and library-level entry points of libimsim_b200.so
#include <cstdlib>
#include <map>
#include <mutex>

#include "b2_common.cuh"

thread_local std::string g_b2_error;
std::atomic<uint64_t> g_b2_launches{0};

extern "C" const char* b2_last_error(void) { return g_b2_error.c_str(); }
extern "C" int b2_abi_version(void) { return B2_ABI_VERSION; }
extern "C" uint64_t b2_launch_count(void) { return g_b2_launches.load(); }

extern "C" int64_t b2_sizeof(int32_t which) {
    switch (which) {
        case 0: return sizeof(B2Telescope);
        case 1: return sizeof(B2Surface);
        case 2: return sizeof(B2TanSip);
        case 3: return sizeof(B2Detector);
        case 4: return sizeof(B2Diffraction);
        case 5: return sizeof(B2OpticsOptions);
        case 6: return sizeof(B2OpticsStats);
        case 7: return sizeof(B2SensorConfig);
        case 8: return sizeof(B2AccumStats);
        case 9: return sizeof(B2Obsc);
        case 10: return sizeof(B2Medium);
        case 11: return sizeof(B2Object);
        case 12: return sizeof(B2Psf);
        case 13: return sizeof(B2Amp);
        case 14: return sizeof(B2StampJob);
    }
    return -1;
}

int b2_scratch_reserve(b2_ctx* ctx, Scratch& s, size_t bytes) {
    if (bytes <= s.bytes) return 0;
    // grow geometrically: staging sizes repeat from call to call
    size_t want = bytes + bytes / 4 + 4096;
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    if (s.ptr) B2_CUDA(cudaFree(s.ptr));
    s.ptr = nullptr;
    s.bytes = 0;
    B2_CUDA(cudaMalloc(&s.ptr, want));
    s.bytes = want;
    return 0;
}

extern "C" int b2_ctx_create(int device, void* cuda_stream, b2_ctx** out) {
    B2_REQUIRE(out, "b2_ctx_create: null out pointer");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return b2_fail("b2_ctx_create: no CUDA device available (%s); there is no CPU fallback", cudaGetErrorString(e));
    B2_REQUIRE(device >= 0 && device < ndev, "b2_ctx_create: bad device index");
    B2_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    B2_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) return b2_fail("b2_ctx_create: built for sm_100a (B200); device is %s", prop.name);
    b2_ctx* ctx = new b2_ctx();
    ctx->device = device;
    ctx->stream = (cudaStream_t)cuda_stream;
    memset(&ctx->opt, 0, sizeof(ctx->opt));
    *out = ctx;
    return 0;
}

extern "C" int b2_ctx_destroy(b2_ctx* ctx) {
    if (!ctx) return 0;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (b2_sensor* s : ctx->sensors) b2_sensor_orphan(s);
    ctx->sensors.clear();
    b2_stage1_release(ctx);
    b2_pipe_release(ctx);
    for (auto& e : ctx->extras) {
        if (e.dev.ptr) cudaFree(e.dev.ptr);
        for (int k = 0; k < 2; ++k) {
            if (e.pin[k]) cudaFreeHost(e.pin[k]);
            if (e.ev[k]) cudaEventDestroy(e.ev[k]);
        }
    }
    if (ctx->scratch.ptr) cudaFree(ctx->scratch.ptr);
    if (ctx->stats.ptr) cudaFree(ctx->stats.ptr);
    if (ctx->fill_scratch.ptr) cudaFree(ctx->fill_scratch.ptr);
    delete ctx;
    return 0;
}

extern "C" int b2_ctx_set_stream(b2_ctx* ctx, void* cuda_stream) {
    B2_REQUIRE(ctx, "null context");
    ctx->stream = (cudaStream_t)cuda_stream;
    return 0;
}

extern "C" int b2_ctx_synchronize(b2_ctx* ctx) {
    B2_REQUIRE(ctx, "null context");
    B2_CUDA(cudaSetDevice(ctx->device));
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
    return 0;
}

// ---------------------------------------------------------------- per-kernel timing (B2_TIMING=1)
namespace {
struct TimedLaunch {
    const char* name;
    cudaEvent_t e0, e1;
};
std::vector<TimedLaunch> g_timed;
std::mutex g_timed_mu;
}  // namespace

bool b2_timing_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("B2_TIMING");
        on = (e && e[0] == '1') ? 1 : 0;
    }
    return on == 1;
}

long b2_timing_begin(const char* name, cudaStream_t st) {
    TimedLaunch t;
    t.name = name;
    cudaEventCreate(&t.e0);
    cudaEventCreate(&t.e1);
    cudaEventRecord(t.e0, st);
    std::lock_guard<std::mutex> lk(g_timed_mu);
    g_timed.push_back(t);
    return (long)g_timed.size() - 1;
}

void b2_timing_end(long slot, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_timed_mu);
    if (slot < (long)g_timed.size()) cudaEventRecord(g_timed[slot].e1, st);
}

// JSON {"kernel": [count, total_ms], ...} of everything timed since the last report; clears the log
extern "C" int b2_timing_report(char* buf, int64_t cap) {
    B2_REQUIRE(buf && cap > 2, "b2_timing_report: bad buffer");
    std::lock_guard<std::mutex> lk(g_timed_mu);
    std::map<std::string, std::pair<long, double>> agg;
    for (auto& t : g_timed) {
        cudaEventSynchronize(t.e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, t.e0, t.e1);
        auto& a = agg[t.name];
        a.first++;
        a.second += ms;
        cudaEventDestroy(t.e0);
        cudaEventDestroy(t.e1);
    }
    g_timed.clear();
    std::string out = "{";
    bool first = true;
    for (auto& kv : agg) {
        char line[256];
        snprintf(line, sizeof(line), "%s\"%s\": [%ld, %.6f]", first ? "" : ", ", kv.first.c_str(), kv.second.first,
                 kv.second.second);
        out += line;
        first = false;
    }
    out += "}";
    B2_REQUIRE((int64_t)out.size() + 1 <= cap, "b2_timing_report: buffer too small");
    memcpy(buf, out.c_str(), out.size() + 1);
    return 0;
}

// ---------------------------------------------------------------- dominant-kernel events
extern "C" int b2_ctx_record_kernel_events(b2_ctx* ctx, int32_t on) {
    B2_REQUIRE(ctx, "null context");
    ctx->record_events = on != 0;
    return 0;
}

// sum of the recorded launches' durations [ms] and their count; clears the list
extern "C" int b2_ctx_kernel_ms(b2_ctx* ctx, double* total_ms, int64_t* count) {
    B2_REQUIRE(ctx && total_ms && count, "b2_ctx_kernel_ms: null argument");
    B2_CUDA(cudaSetDevice(ctx->device));
    double tot = 0.0;
    for (auto& ev : ctx->events) {
        B2_CUDA(cudaEventSynchronize(ev.second));
        float ms = 0.f;
        B2_CUDA(cudaEventElapsedTime(&ms, ev.first, ev.second));
        tot += ms;
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    *total_ms = tot;
    *count = (int64_t)ctx->events.size();
    ctx->events.clear();
    return 0;
}
