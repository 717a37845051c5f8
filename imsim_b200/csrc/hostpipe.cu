// hostpipe.cu -- pipelined staging of pageable host photon arrays (sm_100a library, host code only).
//
// The reference hands every photon op and the sensor ordinary numpy arrays (GalSim PhotonArray fields,
// imsim/photon_ops.py:81-127, imsim/photon_pooling.py:195-225).  A cudaMemcpyAsync from pageable memory is
// staged by the driver on one host thread (~11 GB/s measured) and serialises with the kernel, so the plain
// B2_HOST calls were bound by that copy.  Here a call is cut into chunks that move through a ring of pinned
// slots: a small pool of host threads copies chunk k from the caller's arrays into a slot while the DMA
// engines and the kernel work on chunks k-1 and k-2 (two internal streams), and results travel back the
// same way.  Arithmetic is untouched: a chunk is a sub-range of the same per-photon kernel, and the Philox
// counters are offset by the chunk start.
#include <unistd.h>

#include <chrono>

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <thread>

#include <emmintrin.h>

#include "b2_common.cuh"

namespace {

struct Piece {
    char* dst;
    const char* src;
    size_t n;
    bool to_ring = false;  // destination is a pinned ring slot: the CPU never reads it back
};

// Copy into a pinned ring slot with non-temporal stores: the slot is read next by the copy engine, not by the CPU,
// so its lines are neither fetched for ownership nor left to evict the caller's arrays from the cache
// (B2_COPY_NT=0: plain memcpy).
bool g_copy_nt = false;

void copy_piece(const Piece& q) {
    if (!q.to_ring || !g_copy_nt || q.n < 4096) {
        memcpy(q.dst, q.src, q.n);
        return;
    }
    char* d = q.dst;
    const char* s = q.src;
    size_t n = q.n;
    const size_t head = (64 - ((uintptr_t)d & 63)) & 63;
    if (head) {
        memcpy(d, s, head);
        d += head; s += head; n -= head;
    }
    const size_t body = n & ~(size_t)63;
    for (size_t o = 0; o < body; o += 64) {
        const __m128i a = _mm_loadu_si128((const __m128i*)(s + o));
        const __m128i b = _mm_loadu_si128((const __m128i*)(s + o + 16));
        const __m128i c = _mm_loadu_si128((const __m128i*)(s + o + 32));
        const __m128i e = _mm_loadu_si128((const __m128i*)(s + o + 48));
        _mm_stream_si128((__m128i*)(d + o), a);
        _mm_stream_si128((__m128i*)(d + o + 16), b);
        _mm_stream_si128((__m128i*)(d + o + 32), c);
        _mm_stream_si128((__m128i*)(d + o + 48), e);
    }
    _mm_sfence();
    if (n > body) memcpy(d + body, s + body, n - body);
}

// Persistent memcpy workers.  Never destroyed (a joinable std::thread in a static destructor of a forked or
// exiting interpreter is a hang waiting to happen); re-created in a forked child, where the parent's threads
// do not exist.
class CopyPool {
    std::vector<std::thread> th_;
    std::mutex run_mu_;  // one copy job at a time: calls on different contexts may come from different host threads
    std::mutex m_;
    std::condition_variable cv_work_, cv_done_;
    const std::vector<Piece>* pieces_ = nullptr;
    std::atomic<size_t> next_{0};
    size_t busy_ = 0;
    uint64_t gen_ = 0;

    void drain(const std::vector<Piece>& p) {
        for (;;) {
            size_t i = next_.fetch_add(1);
            if (i >= p.size()) break;
            copy_piece(p[i]);
        }
    }
    void worker() {
        uint64_t seen = 0;
        for (;;) {
            const std::vector<Piece>* p;
            {
                std::unique_lock<std::mutex> lk(m_);
                cv_work_.wait(lk, [&] { return gen_ != seen; });
                seen = gen_;
                p = pieces_;
            }
            drain(*p);
            {
                std::lock_guard<std::mutex> lk(m_);
                if (--busy_ == 0) cv_done_.notify_one();
            }
        }
    }

public:
    explicit CopyPool(int nthreads) {
        for (int i = 1; i < nthreads; ++i) {
            th_.emplace_back([this] { worker(); });
            th_.back().detach();
        }
    }
    int threads() const { return (int)th_.size() + 1; }
    // copies every piece; the calling thread takes its share
    void run(const std::vector<Piece>& p) {
        // one job at a time on the workers; a caller that finds them busy (another context, or the checkpoint copy of
        // a builder whose next upload is already running) copies its pieces itself instead of queueing behind it
        std::unique_lock<std::mutex> serial(run_mu_, std::try_to_lock);
        if (!serial.owns_lock() || th_.empty() || p.size() <= 1) {
            for (const Piece& q : p) copy_piece(q);
            return;
        }
        {
            std::lock_guard<std::mutex> lk(m_);
            pieces_ = &p;
            next_.store(0);
            busy_ = th_.size();
            ++gen_;
        }
        cv_work_.notify_all();
        drain(p);
        std::unique_lock<std::mutex> lk(m_);
        cv_done_.wait(lk, [&] { return busy_ == 0; });
    }
};

std::mutex g_pool_mu;
CopyPool* g_pool = nullptr;
pid_t g_pool_pid = 0;

long env_long(const char* name, long dflt) {
    const char* e = getenv(name);
    if (!e || !*e) return dflt;
    char* end = nullptr;
    long v = strtol(e, &end, 10);
    return (end && *end == 0 && v > 0) ? v : dflt;
}

CopyPool& pool() {
    std::lock_guard<std::mutex> lk(g_pool_mu);
    if (!g_pool || g_pool_pid != getpid()) {
        // up to 8 copy threads, fewer when several ranks share the host's cores (torchrun exports LOCAL_WORLD_SIZE)
        long hw = (long)std::thread::hardware_concurrency();
        long share = std::max(1L, hw / env_long("LOCAL_WORLD_SIZE", 1));
        long n = env_long("B2_HOST_THREADS", std::max(1L, std::min(8L, share)));
        g_pool = new CopyPool((int)std::min(n, 64L));  // the previous pool (parent's, after a fork) is abandoned
        g_pool_pid = getpid();
    }
    // measured (profiles/README.md): with eight ranks on one host the copies are bound by the memory system and
    // non-temporal stores give 6 %; a single rank is 2 % faster with plain memcpy
    g_copy_nt = env_long("LOCAL_WORLD_SIZE", 1) > 1;
    if (const char* e = getenv("B2_COPY_NT")) g_copy_nt = atoi(e) != 0;
    return *g_pool;
}

constexpr int NSLOT = 3;
constexpr size_t PIECE = size_t(1) << 20;  // 1 MB per memcpy work item

}  // namespace

struct b2_hostpipe {
    void* pin = nullptr;
    size_t pin_bytes = 0;
    cudaStream_t st[2] = {nullptr, nullptr};
    cudaEvent_t done[NSLOT] = {nullptr, nullptr, nullptr};
    cudaEvent_t entry = nullptr;
};

void b2_pipe_release(b2_ctx* ctx) {
    b2_hostpipe* p = ctx->pipe;
    if (!p) return;
    for (int i = 0; i < 2; ++i)
        if (p->st[i]) cudaStreamDestroy(p->st[i]);
    for (int i = 0; i < NSLOT; ++i)
        if (p->done[i]) cudaEventDestroy(p->done[i]);
    if (p->entry) cudaEventDestroy(p->entry);
    if (p->pin) cudaFreeHost(p->pin);
    delete p;
    ctx->pipe = nullptr;
}

// B2_PIPE_MIN: smallest call [photons] that takes the pipelined route (below it the set-up costs more than the
// single staged copy); B2_PIPE_CHUNK: photons per chunk.  Read on every call so that tests can switch them.
bool b2_pipe_enabled(int64_t n) { return n >= env_long("B2_PIPE_MIN", 1L << 18); }

int b2_pipe_threads() { return pool().threads(); }

int b2_pipe_run(b2_ctx* ctx, int64_t n, int nin, const double* const* hin, double* const* din, int nout,
                double* const* hout, const double* const* dout,
                const std::function<int(int64_t, int64_t, cudaStream_t)>* kernel) {
    if (n <= 0 || (nin == 0 && nout == 0)) return 0;
    B2_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->pipe) ctx->pipe = new b2_hostpipe();
    b2_hostpipe& p = *ctx->pipe;
    if (!p.st[0]) {
        for (int i = 0; i < 2; ++i) B2_CUDA(cudaStreamCreateWithFlags(&p.st[i], cudaStreamNonBlocking));
        for (int i = 0; i < NSLOT; ++i) B2_CUDA(cudaEventCreateWithFlags(&p.done[i], cudaEventDisableTiming));
        B2_CUDA(cudaEventCreateWithFlags(&p.entry, cudaEventDisableTiming));
    }
    const int64_t chunk = std::min<int64_t>(n, env_long("B2_PIPE_CHUNK", 1L << 19));
    const size_t slot_bytes = (size_t)(nin + nout) * (size_t)chunk * sizeof(double);
    if (p.pin_bytes < NSLOT * slot_bytes) {
        if (p.pin) B2_CUDA(cudaFreeHost(p.pin));
        p.pin = nullptr;
        p.pin_bytes = 0;
        B2_CUDA(cudaHostAlloc(&p.pin, NSLOT * slot_bytes, cudaHostAllocDefault));
        p.pin_bytes = NSLOT * slot_bytes;
    }
    auto slot = [&](int s, int f) { return (double*)((char*)p.pin + (size_t)s * slot_bytes) + (size_t)f * chunk; };
    // the device arrays may still be in use by earlier work on the context's stream
    B2_CUDA(cudaEventRecord(p.entry, ctx->stream));
    for (int i = 0; i < 2; ++i) B2_CUDA(cudaStreamWaitEvent(p.st[i], p.entry, 0));
    CopyPool& cp = pool();
    std::vector<Piece> pieces;
    auto host_copy = [&](int64_t k, int s, bool in) {
        const int64_t off = k * chunk, cnt = std::min(chunk, n - off);
        const size_t bytes = (size_t)cnt * sizeof(double);
        pieces.clear();
        const int nf = in ? nin : nout;
        for (int f = 0; f < nf; ++f) {
            char* pinned = (char*)slot(s, in ? f : nin + f);
            char* user = in ? (char*)const_cast<double*>(hin[f] + off) : (char*)(hout[f] + off);
            for (size_t o = 0; o < bytes; o += PIECE) {
                size_t m = std::min(PIECE, bytes - o);
                pieces.push_back(in ? Piece{pinned + o, user + o, m, true} : Piece{user + o, pinned + o, m, false});
            }
        }
        cp.run(pieces);
    };
    const int64_t nch = (n + chunk - 1) / chunk;
    for (int64_t k = 0; k < nch + NSLOT; ++k) {
        const int s = (int)(k % NSLOT);
        if (k >= NSLOT && k - NSLOT < nch) {
            // chunk k - NSLOT lived in this slot: wait for it, hand its results back
            B2_CUDA(cudaEventSynchronize(p.done[s]));
            if (nout) host_copy(k - NSLOT, s, false);
        }
        if (k < nch) {
            const int64_t off = k * chunk, cnt = std::min(chunk, n - off);
            if (nin) host_copy(k, s, true);
            cudaStream_t st = p.st[k & 1];
            for (int f = 0; f < nin; ++f)
                B2_CUDA(cudaMemcpyAsync(din[f] + off, slot(s, f), (size_t)cnt * sizeof(double), cudaMemcpyHostToDevice, st));
            if (kernel && (*kernel)(off, cnt, st)) return 1;
            for (int f = 0; f < nout; ++f)
                B2_CUDA(cudaMemcpyAsync(slot(s, nin + f), dout[f] + off, (size_t)cnt * sizeof(double),
                                        cudaMemcpyDeviceToHost, st));
            B2_CUDA(cudaEventRecord(p.done[s], st));
        }
    }
    for (int i = 0; i < 2; ++i) B2_CUDA(cudaStreamSynchronize(p.st[i]));
    return 0;
}


// Pooled upload: the photon arrays of many stamps (imsim/photon_pooling.py:177-192 merges them on the host with one
// memcpy per stamp and field) go straight from the stamps' own pageable arrays into one device array per field --
// the merge happens inside the pinned ring: the copy threads gather the segments that fall into a chunk while the
// copy engine moves the previous chunk.  seg: nfields * nseg host pointers, field-major; seg_len: photons per segment;
// dst: one device array per field, each holding sum(seg_len) doubles.  Complete when it returns.
extern "C" int b2_photons_upload(b2_ctx* ctx, int32_t nfields, int64_t nseg, const double* const* seg,
                                 const int64_t* seg_len, double* const* dst) {
    B2_REQUIRE(ctx && seg && seg_len && dst && nfields > 0 && nfields <= 16 && nseg >= 0, "b2_photons_upload: bad argument");
    std::vector<int64_t> start((size_t)nseg + 1, 0);
    for (int64_t k = 0; k < nseg; ++k) {
        B2_REQUIRE(seg_len[k] >= 0, "b2_photons_upload: negative segment length");
        start[k + 1] = start[k] + seg_len[k];
    }
    const int64_t n = start[nseg];
    if (n == 0) return 0;
    B2_CUDA(cudaSetDevice(ctx->device));
    if (!ctx->pipe) ctx->pipe = new b2_hostpipe();
    b2_hostpipe& p = *ctx->pipe;
    if (!p.st[0]) {
        for (int i = 0; i < 2; ++i) B2_CUDA(cudaStreamCreateWithFlags(&p.st[i], cudaStreamNonBlocking));
        for (int i = 0; i < NSLOT; ++i) B2_CUDA(cudaEventCreateWithFlags(&p.done[i], cudaEventDisableTiming));
        B2_CUDA(cudaEventCreateWithFlags(&p.entry, cudaEventDisableTiming));
    }
    const int64_t chunk = std::min<int64_t>(n, env_long("B2_PIPE_CHUNK", 1L << 19));
    const size_t slot_bytes = (size_t)nfields * (size_t)chunk * sizeof(double);
    for (int i = 0; i < 2; ++i) B2_CUDA(cudaStreamSynchronize(p.st[i]));  // copies of an earlier call still read the ring
    if (p.pin_bytes < NSLOT * slot_bytes) {
        if (p.pin) B2_CUDA(cudaFreeHost(p.pin));
        p.pin = nullptr;
        p.pin_bytes = 0;
        B2_CUDA(cudaHostAlloc(&p.pin, NSLOT * slot_bytes, cudaHostAllocDefault));
        p.pin_bytes = NSLOT * slot_bytes;
    }
    auto slot = [&](int s, int f) { return (char*)p.pin + (size_t)s * slot_bytes + (size_t)f * chunk * sizeof(double); };
    B2_CUDA(cudaEventRecord(p.entry, ctx->stream));
    for (int i = 0; i < 2; ++i) B2_CUDA(cudaStreamWaitEvent(p.st[i], p.entry, 0));
    CopyPool& cp = pool();
    std::vector<Piece> pieces;
    const int64_t nch = (n + chunk - 1) / chunk;
    int64_t sfirst = 0;  // first segment that reaches into the current chunk
    const bool prof = getenv("B2_PIPE_PROFILE") != nullptr;
    double t_wait = 0.0, t_copy = 0.0, t_issue = 0.0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    for (int64_t k = 0; k < nch; ++k) {
        const int s = (int)(k % NSLOT);
        const int64_t off = k * chunk, cnt = std::min(chunk, n - off);
        double t0 = prof ? now() : 0.0;
        if (k >= NSLOT) B2_CUDA(cudaEventSynchronize(p.done[s]));  // the copy that last read this slot
        if (prof) t_wait += now() - t0;
        while (sfirst < nseg && start[sfirst + 1] <= off) ++sfirst;
        pieces.clear();
        for (int64_t g = sfirst; g < nseg && start[g] < off + cnt; ++g) {
            const int64_t a = std::max(start[g], off), b = std::min(start[g + 1], off + cnt);
            if (b <= a) continue;
            for (int f = 0; f < nfields; ++f) {
                const char* src = (const char*)(seg[(size_t)f * nseg + g] + (a - start[g]));
                char* dstp = slot(s, f) + (size_t)(a - off) * sizeof(double);
                const size_t bytes = (size_t)(b - a) * sizeof(double);
                for (size_t o = 0; o < bytes; o += PIECE) pieces.push_back(Piece{dstp + o, src + o, std::min(PIECE, bytes - o), true});
            }
        }
        t0 = prof ? now() : 0.0;
        cp.run(pieces);
        if (prof) t_copy += now() - t0;
        t0 = prof ? now() : 0.0;
        cudaStream_t st = p.st[k & 1];
        for (int f = 0; f < nfields; ++f)
            B2_CUDA(cudaMemcpyAsync(dst[f] + off, slot(s, f), (size_t)cnt * sizeof(double), cudaMemcpyHostToDevice, st));
        B2_CUDA(cudaEventRecord(p.done[s], st));
        if (prof) t_issue += now() - t0;
    }
    if (prof)
        fprintf(stderr, "b2_photons_upload: %lld photons x %d fields, %lld chunks, %d threads: wait %.1f ms, host copy %.1f ms, issue %.1f ms\n",
                (long long)n, nfields, (long long)nch, cp.threads(), 1e3 * t_wait, 1e3 * t_copy, 1e3 * t_issue);
    // later work on the context's stream sees the uploaded arrays; the host does not wait for the copies
    for (int i = 0; i < 2; ++i) {
        B2_CUDA(cudaEventRecord(p.entry, p.st[i]));
        B2_CUDA(cudaStreamWaitEvent(ctx->stream, p.entry, 0));
    }
    return 0;
}


// memcpy between two host buffers on the copy threads (a pinned snapshot of a CCD image into the caller's pageable
// array: 66 MB take ~13 ms on one thread and ~1.5 ms on eight)
extern "C" int b2_host_memcpy(void* dst, const void* src, int64_t bytes) {
    B2_REQUIRE(bytes >= 0 && (bytes == 0 || (dst && src)), "b2_host_memcpy: bad argument");
    std::vector<Piece> pieces;
    for (size_t o = 0; o < (size_t)bytes; o += PIECE)
        pieces.push_back(Piece{(char*)dst + o, (const char*)src + o, std::min(PIECE, (size_t)bytes - o)});
    pool().run(pieces);
    return 0;
}

// Is every segment a constant array?  GalSim's shooters give all photons of an object the same flux, so the flux
// field of a pooled upload is one number per stamp: reading it once (host threads, no writes) is cheaper than copying
// it into the ring and over PCIe.  value[g] = first element of segment g (0 for an empty one); *all_constant = 1 iff
// every element of every segment equals its segment's first.
extern "C" int b2_segments_constant(int64_t nseg, const double* const* seg, const int64_t* seg_len, double* value,
                                    int32_t* all_constant) {
    B2_REQUIRE(nseg >= 0 && (nseg == 0 || (seg && seg_len && value)) && all_constant, "b2_segments_constant: null argument");
    std::atomic<int> ok{1};
    std::atomic<int64_t> next{0};
    auto work = [&] {
        for (;;) {
            const int64_t g = next.fetch_add(1);
            if (g >= nseg) break;
            const double* a = seg[g];
            const int64_t m = seg_len[g];
            const double v = m > 0 ? a[0] : 0.0;
            value[g] = v;
            if (!ok.load(std::memory_order_relaxed)) continue;  // values are still wanted
            // bitwise comparison in blocks: any difference (a NaN included) makes the segment non-constant
            uint64_t vb, acc = 0;
            memcpy(&vb, &v, 8);
            const uint64_t* b = reinterpret_cast<const uint64_t*>(a);
            for (int64_t i = 0; i < m; ++i) acc |= b[i] ^ vb;
            if (acc) ok.store(0);
        }
    };
    long total = 0;
    for (int64_t g = 0; g < nseg; ++g) total += (long)seg_len[g];
    int nth = (int)std::min<long>(pool().threads(), std::max<long>(1, total >> 20));
    std::vector<std::thread> th;
    for (int i = 1; i < nth; ++i) th.emplace_back(work);
    work();
    for (auto& t : th) t.join();
    *all_constant = ok.load();
    return 0;
}
