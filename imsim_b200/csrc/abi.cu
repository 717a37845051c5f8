// abi.cu -- context management and library-level entry points of libimsim_b200.so
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
    for (void* p : ctx->extras) cudaFree(p);
    if (ctx->scratch.ptr) cudaFree(ctx->scratch.ptr);
    if (ctx->stats.ptr) cudaFree(ctx->stats.ptr);
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
