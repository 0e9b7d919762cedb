// vf_abi.cpp — implementation of include/b200vf.h: contexts, LUT upload, frame
// validation/dispatch, and the pinned H2D → kernel → D2H stream pipeline for
// system-memory frames.  No exception leaves this file; there is no CPU fallback.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200vf.h"
#include "vf_internal.h"

using namespace vf;

namespace {

thread_local std::string g_last_error;  // for calls that have no context

struct Slot {  // one stage buffer set of the host-frame pipeline
    void *d_in = nullptr, *d_out = nullptr;
    size_t d_in_cap = 0, d_out_cap = 0;
    void *h_in = nullptr, *h_out = nullptr;  // pinned bounce buffers for pageable frames
    size_t h_in_cap = 0, h_out_cap = 0;
    cudaEvent_t ev_h2d = nullptr, ev_k = nullptr, ev_d2h = nullptr;
    bool busy = false;
    bool used = false;  // its events have been recorded at least once
    // deferred copy-out of a pageable destination
    uint8_t *user_out = nullptr;
    int64_t user_stride = 0;
    size_t row_bytes = 0, rows = 0, d_pitch = 0;
};

constexpr int kTickets = 16;  // "host.async": host-frame calls in flight at most (minus one)
constexpr int kMaxSlots = 8;  // stage buffer sets of the host-frame pipeline; "host.slots" of them are used

// A few helper threads for the row copies between pageable frames and the pinned bounce
// buffers (one core's memcpy is ~10 GB/s, well below PCIe).  Created on first pageable frame.
class CopyPool {
public:
    explicit CopyPool(int n_threads) {
        for (int i = 0; i < n_threads; i++) workers_.emplace_back([this] { loop(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> g(mu_);
            quit_ = true;
        }
        cv_.notify_all();
        for (std::thread &t : workers_) t.join();
    }
    // Runs fn(i) for i in [0, n) on the pool and the calling thread; returns when all are done.
    void parallel_for(size_t n, const std::function<void(size_t)> &fn) {
        if (n == 0) return;
        {
            std::lock_guard<std::mutex> g(mu_);
            fn_ = &fn;
            next_ = 0;
            total_ = n;
            pending_ = n;
            epoch_++;
        }
        cv_.notify_all();
        run_tasks();
        std::unique_lock<std::mutex> lk(mu_);
        done_cv_.wait(lk, [this] { return pending_ == 0; });
        fn_ = nullptr;
    }

private:
    void run_tasks() {
        for (;;) {
            size_t i;
            const std::function<void(size_t)> *fn;
            {
                std::lock_guard<std::mutex> g(mu_);
                if (!fn_ || next_ >= total_) return;
                i = next_++;
                fn = fn_;
            }
            (*fn)(i);
            std::lock_guard<std::mutex> g(mu_);
            if (--pending_ == 0) done_cv_.notify_all();
        }
    }
    void loop() {
        uint64_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return quit_ || epoch_ != seen; });
                if (quit_) return;
                seen = epoch_;
            }
            run_tasks();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_cv_;
    const std::function<void(size_t)> *fn_ = nullptr;
    size_t next_ = 0, total_ = 0, pending_ = 0;
    uint64_t epoch_ = 0;
    bool quit_ = false;
};

// hsvfilter, hsvdetector and the colorlut ! hsvfilter chain are pure functions of a pixel's three
// colour bytes for as long as their settings stand.  Once the settings have been stable for a
// while, the element's own exact kernel is run once over all 2^24 triples and frames are served
// from that table (one 4-byte gather per pixel, HBM-bound) instead of ~70 instructions per pixel.
// Gathers are content-sensitive (random colours: one L2 sector per pixel), the compute kernel is
// not, so in auto mode both ways are timed on the stream's real frames and the faster one runs.
template <class T>
void key_put(std::vector<uint8_t> &k, const T &v) {  // raw bytes of v appended to a table key
    const uint8_t *p = reinterpret_cast<const uint8_t *>(&v);
    k.insert(k.end(), p, p + sizeof(T));
}

enum FnPath { kFnAuto = 0, kFnCompute = 1, kFnTable = 2 };
constexpr uint64_t kFnStablePixels = 1ull << 25;  // ~4 frames of 4K before a table is worth building
constexpr uint64_t kProbeMinPixels = 1ull << 20;
constexpr uint64_t kReprobeMinNs = 250ull * 1000 * 1000;         // re-time the idle kind after 0.25 s …
constexpr uint64_t kReprobeMaxNs = 8ull * 1000 * 1000 * 1000;    // … backing off to 8 s while it keeps losing
constexpr uint64_t kRefreshComputeNs = 30ull * 1000 * 1000 * 1000;
constexpr uint32_t kReprobeMinLaunches = 8;

// Which of two kernels serves a stream: [1] the table gather (content-sensitive: random colours cost
// one L2 sector per pixel) or [0] the per-pixel compute / interpolating kernel (content-insensitive).
// Launches are timed with CUDA events on the stream's real frames, never blocking the caller (at most
// one measurement is outstanding; results are collected by a later call):
//   * both kinds are timed once, the faster one serves;
//   * the serving kind keeps being sampled (every 8th launch), so a change of content that slows the
//     table down is seen within a few launches and the compute kernel takes over;
//   * the kind that is not serving is re-timed after `interval` of wall-clock time (and at least 8
//     launches) — while the compute kernel serves, that is the only way to notice that the content
//     has become table-friendly again (a scene change); the interval doubles (0.25 s .. 8 s) while a
//     re-timing confirms the choice clearly, so steady content pays well under 1 % for it.  While
//     the table serves, the compute kernel's figure does not age (it does not depend on content)
//     and is only refreshed every 30 s.
struct PathPolicy {
    float ns_per_px[2] = {-1.0f, -1.0f};  // measured device time; < 0 = not known yet
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool ev_failed = false;
    int pending = -1;  // which kind the outstanding timing belongs to
    uint64_t pending_pixels = 0;
    uint32_t since_probe = 0;       // launches since the idle kind was last timed
    uint64_t last_probe_ns = 0;     // steady-clock time of that
    uint64_t interval = kReprobeMinNs;
    int chosen = 1;

    static uint64_t clock_ns() {
        return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
                   std::chrono::steady_clock::now().time_since_epoch()).count();
    }
    void reset() {
        ns_per_px[0] = ns_per_px[1] = -1.0f;
        pending = -1;
        since_probe = 0;
        last_probe_ns = clock_ns();
        interval = kReprobeMinNs;
        chosen = 1;
    }
    // Kind to launch now; *timed = bracket it with begin() / end().
    int next(uint64_t pixels, bool *timed) {
        *timed = false;
        if (pending >= 0 && cudaEventQuery(ev[1]) == cudaSuccess) {
            float ms = 0.0f;
            const int kind = pending;
            pending = -1;
            if (cudaEventElapsedTime(&ms, ev[0], ev[1]) == cudaSuccess && pending_pixels) {
                ns_per_px[kind] = ms * 1e6f / (float)pending_pixels;
                if (ns_per_px[0] >= 0.0f && ns_per_px[1] >= 0.0f) {
                    if (kind != chosen) {  // a re-timing of the kind that is not serving
                        const bool confirmed = ns_per_px[chosen] * 1.25f < ns_per_px[kind];
                        interval = confirmed ? std::min(interval * 2, kReprobeMaxNs) : kReprobeMinNs;
                    } else if (ns_per_px[chosen ^ 1] * 1.05f < ns_per_px[chosen]) {
                        // Switch only on a sample of the SERVING kind: it is re-timed with every
                        // launch, so the figure that loses is never a stale (or one-off) one; 5 %
                        // hysteresis against flapping between two equally fast kinds.
                        chosen ^= 1;
                    }
                }
            }
        }
        cudaGetLastError();  // cudaErrorNotReady from the query is not an error
        since_probe++;
        if (pending >= 0 || pixels < kProbeMinPixels || ev_failed) return chosen;
        if (!ev[0] && (cudaEventCreate(&ev[0]) != cudaSuccess || cudaEventCreate(&ev[1]) != cudaSuccess)) {
            cudaGetLastError();
            ev_failed = true;
            return chosen;
        }
        for (int m = 1; m >= 0; m--)
            if (ns_per_px[m] < 0.0f) {  // never measured
                *timed = true;
                return m;
            }
        // the serving kind is sampled on every 8th launch (two event records and a query cost a
        // few microseconds of host time, which is what bounds single-frame 1080p calls)
        *timed = (since_probe & 7u) == 0;
        if (since_probe >= kReprobeMinLaunches) {
            const uint64_t now = clock_ns();
            if (now - last_probe_ns >= (chosen == 1 ? kRefreshComputeNs : interval)) {
                since_probe = 0;
                last_probe_ns = now;
                *timed = true;
                return chosen ^ 1;
            }
        }
        return chosen;
    }
    void begin(cudaStream_t s) { cudaEventRecord(ev[0], s); }
    void end(cudaStream_t s, int mode, uint64_t pixels) {
        cudaEventRecord(ev[1], s);
        pending = mode;
        pending_pixels = pixels;
    }
    void destroy() {
        for (cudaEvent_t &e : ev)
            if (e) cudaEventDestroy(e), e = nullptr;
    }
};

struct FnTable {
    SharedTable *shared = nullptr;  // from the device-wide cache (vf_tables.cpp); null until due
    bool alloc_failed = false;
    std::vector<uint8_t> key;   // element, settings, layouts: what the table is (to be) for
    uint64_t stable_pixels = 0; // processed with this key before the table exists
    bool built = false;
    bool last_used_table = false;
    PathPolicy policy;
};

}  // namespace

struct b200vf_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;  // compute stream for device frames and kernels
    bool own_stream = false;
    cudaEvent_t ev_order = nullptr;  // b200vf_ctx_wait_for
    cudaStream_t s_in = nullptr, s_out = nullptr;  // copy streams of the host path
    Slot slots[kMaxSlots];
    int n_slots = 4;  // "host.slots": chunks in flight (H2D / kernel / D2H overlap needs >= 3)
    // "host.register": page-lock recurring pageable frames in place instead of bouncing them
    // through pinned staging buffers (a GStreamer pool hands the same few buffers round and round)
    int register_mode = 0;
    struct Registered {
        void *base;
        size_t bytes;
        uint64_t last_use;
    };
    std::vector<Registered> registered;            // ranges this context page-locked
    std::vector<std::pair<const void *, size_t>> seen;  // pageable ranges met once (ring)
    size_t seen_next = 0;
    size_t registered_bytes = 0, register_budget = (size_t)2 << 30;
    uint64_t use_clock = 0;
    // host-path diagnostics ("host.dbg_*", read-only): where a host-frame call spends its time
    uint64_t dbg_chunks = 0, dbg_wait_ns = 0, dbg_call_ns = 0, dbg_copy_ns = 0;
    std::vector<cudaEvent_t> dbg_tl;  // "host.dbg_mode" = 4: timing events around every op of a call
    bool in_host_call = false;  // chunks of a host frame are being launched: no policy timing (PCIe-bound)
    int dbg_mode = 0;  // "host.dbg_mode" (experiments): 1 = no kernel, 2 = no kernel and no cross-stream waits
    int next_slot = 0;
    // "host.async": calls with page-locked host frames return once queued; tickets tell them apart
    int host_async = 0;
    uint64_t host_ticket = 0;       // host-frame calls issued so far (the last call's ticket)
    uint64_t host_ticket_done = 0;  // every call up to this ticket is known to be complete
    cudaEvent_t ticket_ev[kTickets] = {};  // ticket t: recorded on s_out behind the call's last D2H
    uint64_t ticket_of[kTickets] = {};
    DeviceLut lut;
    std::string last_error;
    b200vf_stats stats{};
    int math_mode = kMathFast;
    int lut_path = kLutAuto;
    int lut_interp = kInterpTrilinear;
    int64_t chunk_bytes = 0;  // "host.chunk_bytes"; 0 = auto (a sixth of the call's bytes, within 4..17 MiB)
    // "host.copy_threads": threads for pageable-frame row copies; half the cores, within 2..8
    // (one core's memcpy is ~10 GB/s; a 4K pageable stream goes 410 -> 590 frames/s from 4 to 8)
    int copy_threads = (int)std::min(8u, std::max(2u, std::thread::hardware_concurrency() / 2));
    CopyPool *pool = nullptr;
    int lut_path_active = -1;     // resolved path of the last colorlut launch ("lut.path_active")
    std::vector<uint8_t> lut_key; // identifies the LUT's content (hash, size, domain): key of its tables
    SharedTable *baked = nullptr; // the LUT baked to 8-bit resolution, from the device-wide cache
    bool baked_failed = false;    // no memory for it: the interpolating kernels serve
    PathPolicy lut_policy;        // auto: baked table vs direct interpolation, measured
    PathPolicy lut64_policy;      // auto, 16-bit frames: delta-table op vs direct kernel, measured
    bool share_tables = true;     // "tables.share": 0 = private tables (keys salted with the context)
    int fn_path = kFnAuto;        // "hsv.path"
    FnTable fn;                   // tabulated hsvfilter / hsvdetector / chain function
};

namespace {

int fail(b200vf_ctx *ctx, int code, const std::string &msg) {
    if (ctx)
        ctx->last_error = msg;
    else
        g_last_error = msg;
    return code;
}

}  // namespace
int vf::fail_global(int code, const std::string &msg) { return fail(nullptr, code, msg); }
int vf::ctx_fail(b200vf_ctx *ctx, int code, const std::string &msg) { return fail(ctx, code, msg); }
namespace {

int cuda_fail(b200vf_ctx *ctx, cudaError_t e, const char *what) {
    return fail(ctx, B200VF_ERR_CUDA,
                std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}

#define VF_CUDA(ctx, call)                                      \
    do {                                                        \
        cudaError_t e__ = (call);                               \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
    } while (0)

struct FormatInfo {
    const char *name;
    int bpp, r, g, b, a;  // byte offsets for 8-bit formats
};

// Byte layouts: SURVEY.md Appendix C (hsvfilter/imp.rs:327-371, hsvdetector/imp.rs:428-704)
const FormatInfo kFormats[B200VF_FORMAT_COUNT] = {
    {"RGBA", 4, 0, 1, 2, 3},      {"RGBx", 4, 0, 1, 2, 3}, {"xRGB", 4, 1, 2, 3, 0},
    {"ARGB", 4, 1, 2, 3, 0},      {"BGRx", 4, 2, 1, 0, 3}, {"BGRA", 4, 2, 1, 0, 3},
    {"xBGR", 4, 3, 2, 1, 0},      {"ABGR", 4, 3, 2, 1, 0}, {"RGB", 3, 0, 1, 2, -1},
    {"BGR", 3, 2, 1, 0, -1},      {"RGBA64_LE", 8, 0, 2, 4, 6}, {"RGBA64_BE", 8, 0, 2, 4, 6},
};

PixLayout layout_of(uint32_t fmt) {
    const FormatInfo &f = kFormats[fmt];
    return PixLayout{f.bpp, f.r, f.g, f.b, f.a};
}

bool hsvfilter_accepts(uint32_t f) { return f <= B200VF_FORMAT_BGR; }  // hsvfilter/imp.rs:278-289
bool hsvdetector_accepts_in(uint32_t f) {                               // hsvdetector/imp.rs:78-87
    return f == B200VF_FORMAT_RGBX || f == B200VF_FORMAT_XRGB || f == B200VF_FORMAT_BGRX ||
           f == B200VF_FORMAT_XBGR || f == B200VF_FORMAT_RGB || f == B200VF_FORMAT_BGR;
}
bool hsvdetector_accepts_out(uint32_t f) {  // hsvdetector/imp.rs:89-96
    return f == B200VF_FORMAT_RGBA || f == B200VF_FORMAT_ARGB || f == B200VF_FORMAT_BGRA ||
           f == B200VF_FORMAT_ABGR;
}
bool colorlut_accepts(uint32_t f) {  // colorlut/imp.rs:122-134
    return f == B200VF_FORMAT_RGBA || f == B200VF_FORMAT_RGBA64_LE || f == B200VF_FORMAT_RGBA64_BE;
}

int check_frame(b200vf_ctx *ctx, const b200vf_frame *f, const char *who) {
    if (!f) return fail(ctx, B200VF_ERR_INVALID_ARG, std::string(who) + ": frame is NULL");
    if (f->format >= B200VF_FORMAT_COUNT)
        return fail(ctx, B200VF_ERR_UNSUPPORTED_FORMAT, std::string(who) + ": unknown format");
    if (f->memory > B200VF_MEM_DEVICE)
        return fail(ctx, B200VF_ERR_INVALID_ARG, std::string(who) + ": unknown memory kind");
    if (f->width == 0 || f->height == 0) return B200VF_OK;  // empty frame: nothing to touch
    if (!f->data) return fail(ctx, B200VF_ERR_INVALID_ARG, std::string(who) + ": data is NULL");
    const int64_t row = (int64_t)f->width * kFormats[f->format].bpp;
    if (f->stride < row)
        return fail(ctx, B200VF_ERR_INVALID_ARG,
                    std::string(who) + ": stride smaller than width * bytes_per_pixel");
    if (kFormats[f->format].bpp == 8 && (f->stride & 1))
        return fail(ctx, B200VF_ERR_INVALID_ARG,
                    std::string(who) + ": RGBA64 stride must be a whole number of u16");
    return B200VF_OK;
}

int activate(b200vf_ctx *ctx) {
    if (!ctx) return fail(nullptr, B200VF_ERR_INVALID_ARG, "context is NULL");
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != ctx->device)
        VF_CUDA(ctx, cudaSetDevice(ctx->device));
    return B200VF_OK;
}

bool same_geometry(const b200vf_frame &a, const b200vf_frame &b) {
    return a.stride == b.stride && a.width == b.width && a.height == b.height &&
           a.format == b.format && a.memory == b.memory;
}

bool is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// One element's kernel launch over frames already in device memory.
struct Launcher {
    virtual cudaError_t run(b200vf_ctx *ctx, const FrameSet &fs, int n, const Geom &g) = 0;
    virtual ~Launcher() = default;
};

int ensure_cap(b200vf_ctx *ctx, void **p, size_t *cap, size_t need, bool pinned_host) {
    if (*cap >= need) return B200VF_OK;
    if (*p) {
        if (pinned_host)
            cudaFreeHost(*p);
        else
            cudaFree(*p);
        *p = nullptr;
        *cap = 0;
    }
    size_t want = need + need / 4;
    cudaError_t e = pinned_host ? cudaMallocHost(p, want) : cudaMalloc(p, want);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, B200VF_ERR_NOMEM, "staging buffer allocation failed");
    }
    *cap = want;
    return B200VF_OK;
}

// rows x row_bytes from (src, src_pitch) to (dst, dst_pitch), split over the copy pool
void copy_rows(b200vf_ctx *ctx, uint8_t *dst, int64_t dst_pitch, const uint8_t *src,
               int64_t src_pitch, size_t row_bytes, size_t rows) {
    const size_t total = row_bytes * rows;
    const int nt = ctx->copy_threads;
    if (nt <= 1 || total < (1u << 20)) {
        for (size_t r = 0; r < rows; r++)
            std::memcpy(dst + (int64_t)r * dst_pitch, src + (int64_t)r * src_pitch, row_bytes);
        return;
    }
    if (!ctx->pool) {
        try {
            ctx->pool = new CopyPool(nt - 1);
        } catch (...) {  // no threads to be had: copy on the caller, nothing crosses the ABI
            ctx->pool = nullptr;
            ctx->copy_threads = 1;
            for (size_t r = 0; r < rows; r++)
                std::memcpy(dst + (int64_t)r * dst_pitch, src + (int64_t)r * src_pitch, row_bytes);
            return;
        }
    }
    const size_t parts = (size_t)nt * 2;
    const size_t per = (rows + parts - 1) / parts;
    ctx->pool->parallel_for(parts, [&](size_t p) {
        const size_t r0 = p * per, r1 = std::min(rows, r0 + per);
        for (size_t r = r0; r < r1; r++)
            std::memcpy(dst + (int64_t)r * dst_pitch, src + (int64_t)r * src_pitch, row_bytes);
    });
}

// dbg_mode 4: a timing event on `st`, kept for the dump at the end of the call
void tl_mark(b200vf_ctx *ctx, cudaStream_t st) {
    if (ctx->dbg_mode != 4) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    ctx->dbg_tl.push_back(e);
}

void tl_dump(b200vf_ctx *ctx) {
    if (ctx->dbg_tl.empty()) return;
    std::fprintf(stderr, "timeline (us from the first mark): per chunk h2d[start,end] k[start,end] d2h[start,end]\n");
    for (size_t i = 0; i + 5 < ctx->dbg_tl.size(); i += 6) {
        float t[6];
        for (int j = 0; j < 6; j++) cudaEventElapsedTime(&t[j], ctx->dbg_tl[0], ctx->dbg_tl[i + j]);
        std::fprintf(stderr, "chunk %3zu  h2d %8.1f %8.1f  k %8.1f %8.1f  d2h %8.1f %8.1f\n", i / 6, t[0] * 1e3,
                     t[1] * 1e3, t[2] * 1e3, t[3] * 1e3, t[4] * 1e3, t[5] * 1e3);
    }
    for (cudaEvent_t e : ctx->dbg_tl) cudaEventDestroy(e);
    ctx->dbg_tl.clear();
}

uint64_t now_ns() {
    return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(
               std::chrono::steady_clock::now().time_since_epoch()).count();
}

int drain_slot(b200vf_ctx *ctx, Slot &s) {
    if (!s.busy) return B200VF_OK;
    const uint64_t t0 = now_ns();
    VF_CUDA(ctx, cudaEventSynchronize(s.ev_d2h));
    ctx->dbg_wait_ns += now_ns() - t0;
    if (s.user_out) {  // pageable destination: copy rows out of the pinned bounce buffer
        copy_rows(ctx, s.user_out, s.user_stride, (const uint8_t *)s.h_out, (int64_t)s.d_pitch,
                  s.row_bytes, s.rows);
        s.user_out = nullptr;
    }
    s.busy = false;
    return B200VF_OK;
}

// Blocks until the host-frame call with this ticket (and, the copy-out stream running in order,
// every earlier one) is complete.
int host_wait(b200vf_ctx *ctx, uint64_t ticket) {
    if (ticket <= ctx->host_ticket_done) return B200VF_OK;
    for (uint64_t t = ticket; t <= ctx->host_ticket; t++) {
        if (ctx->ticket_of[t % kTickets] != t) continue;  // a synchronous call, or overwritten by a later one
        const uint64_t t0 = now_ns();
        VF_CUDA(ctx, cudaEventSynchronize(ctx->ticket_ev[t % kTickets]));
        ctx->dbg_wait_ns += now_ns() - t0;
        ctx->host_ticket_done = t;
        return B200VF_OK;
    }
    // no event at or after the ticket: the calls since were synchronous ones
    ctx->host_ticket_done = ctx->host_ticket;
    return B200VF_OK;
}

// Page-locked ranges may only be unregistered when no "host.async" copy can still touch them.
void quiesce_async(b200vf_ctx *ctx) {
    if (ctx->host_ticket_done == ctx->host_ticket) return;
    cudaStreamSynchronize(ctx->s_out);
    ctx->host_ticket_done = ctx->host_ticket;
}

void unregister_all(b200vf_ctx *ctx) {
    if (!ctx->registered.empty()) quiesce_async(ctx);
    for (const b200vf_ctx::Registered &r : ctx->registered) cudaHostUnregister(r.base);
    cudaGetLastError();
    ctx->registered.clear();
    ctx->registered_bytes = 0;
    ctx->seen.clear();
}

// Drops every registration of this context that overlaps [p, p + bytes) (bytes == 0: contains p).
void forget_range(b200vf_ctx *ctx, const void *p, size_t bytes) {
    const uintptr_t lo = (uintptr_t)p, hi = lo + (bytes ? bytes : 1);
    for (size_t i = 0; i < ctx->registered.size();) {
        const uintptr_t b = (uintptr_t)ctx->registered[i].base, e = b + ctx->registered[i].bytes;
        if (b < hi && lo < e) {
            quiesce_async(ctx);
            cudaHostUnregister(ctx->registered[i].base);
            cudaGetLastError();
            ctx->registered_bytes -= ctx->registered[i].bytes;
            ctx->registered.erase(ctx->registered.begin() + (long)i);
        } else {
            i++;
        }
    }
    for (auto &s : ctx->seen)
        if ((uintptr_t)s.first < hi && lo < (uintptr_t)s.first + s.second) s = {nullptr, 0};
}

// "host.register": a pageable range that comes by a second time is page-locked in place (LRU within
// a byte budget) so that the copy engines read / write it directly.  Returns true if [p, p+bytes)
// is page-locked when the function returns.  Any failure just leaves the bounce path in charge.
bool maybe_register(b200vf_ctx *ctx, const void *p, size_t bytes) {
    if (!ctx->register_mode || !p || bytes < (1u << 16)) return false;
    ctx->use_clock++;
    for (b200vf_ctx::Registered &r : ctx->registered)
        if (r.base == p && r.bytes >= bytes) {
            r.last_use = ctx->use_clock;
            return true;
        }
    bool met_before = false;
    for (const auto &s : ctx->seen)
        if (s.first == p && s.second == bytes) met_before = true;
    if (!met_before) {  // first sight: remember, bounce this time
        if (ctx->seen.size() < 64)
            ctx->seen.emplace_back(p, bytes);
        else
            ctx->seen[ctx->seen_next++ % 64] = {p, bytes};
        return false;
    }
    if (bytes > ctx->register_budget) return false;
    forget_range(ctx, p, bytes);  // a stale, differently sized registration of the same memory
    while (ctx->registered_bytes + bytes > ctx->register_budget && !ctx->registered.empty()) {
        size_t lru = 0;
        for (size_t i = 1; i < ctx->registered.size(); i++)
            if (ctx->registered[i].last_use < ctx->registered[lru].last_use) lru = i;
        forget_range(ctx, ctx->registered[lru].base, ctx->registered[lru].bytes);
    }
    if (cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterPortable) != cudaSuccess) {
        cudaGetLastError();  // already registered by someone else, or not registrable: bounce
        return false;
    }
    ctx->registered.push_back({const_cast<void *>(p), bytes, ctx->use_clock});
    ctx->registered_bytes += bytes;
    return true;
}

// System-memory frames: split into row chunks and run them through a ring of stage slots so
// that H2D of chunk i+1, the kernel of chunk i and D2H of chunk i-1 overlap.  Pinned
// (page-locked) frames are copied directly; pageable ones bounce through pinned buffers.
// Only width*bpp bytes of each row are read or written.
int run_host_chunks(b200vf_ctx *ctx, const b200vf_frame *in, const b200vf_frame *out, size_t n_frames,
                    int in_bpp, int out_bpp, Launcher &L) {
    struct CallScope {
        b200vf_ctx *c;
        uint64_t t0 = now_ns();
        explicit CallScope(b200vf_ctx *ctx) : c(ctx) { c->in_host_call = true; }
        ~CallScope() {
            c->in_host_call = false;
            c->dbg_call_ns += now_ns() - t0;
        }
    } call_scope(ctx);
    size_t call_bytes = 0;
    for (size_t fi = 0; fi < n_frames; fi++)
        call_bytes += (size_t)in[fi].width * in[fi].height * (size_t)std::max(in_bpp, out_bpp);
    bool all_pinned = true;
    if (ctx->host_async && ctx->host_ticket + 1 >= kTickets) {  // bound the calls in flight
        int rc = host_wait(ctx, ctx->host_ticket + 2 - kTickets);
        if (rc) return rc;
    }
    for (size_t fi = 0; fi < n_frames; fi++) {
        const b200vf_frame &fin = in[fi], &fout = out[fi];
        if (fin.width == 0 || fin.height == 0) continue;
        const size_t rb_in = (size_t)fin.width * in_bpp, rb_out = (size_t)fin.width * out_bpp;
        const size_t p_in = (rb_in + 15) & ~(size_t)15, p_out = (rb_out + 15) & ~(size_t)15;
        bool pin_in = is_pinned(fin.data), pin_out = fout.data == fin.data ? pin_in : is_pinned(fout.data);
        if (!pin_in)
            pin_in = maybe_register(ctx, fin.data, (size_t)fin.stride * (fin.height - 1) + rb_in);
        if (!pin_out)
            pin_out = fout.data == fin.data
                          ? pin_in
                          : maybe_register(ctx, fout.data, (size_t)fout.stride * (fout.height - 1) + rb_out);
        all_pinned = all_pinned && pin_in && pin_out;
        // Chunk size.  Every chunk costs ~30 us of copy-engine idle time (the hand-over between the
        // engines through events, measured with "host.dbg_mode" = 4), the first H2D and the last D2H
        // of a call overlap with nothing: big chunks for big calls, small ones for a single frame.
        size_t chunk_bytes = (size_t)ctx->chunk_bytes;
        if (chunk_bytes == 0)
            chunk_bytes = std::min<size_t>(17u << 20, std::max<size_t>(4u << 20, call_bytes / 6));
        // "host.async": the next call's first H2D overlaps this call's last D2H, so only the
        // per-chunk cost is left: whole 4K frames (tools/host_async_probe.py: 1,233 frames/s with
        // ~5 MiB pieces, 1,436 with whole frames, one 4K frame per call)
        if (ctx->chunk_bytes == 0 && ctx->host_async && pin_in && pin_out) chunk_bytes = 34u << 20;
        size_t rows_per_chunk = std::max<size_t>(1, chunk_bytes / std::max(p_in, p_out));
        // pageable frames: the host's own row copies into / out of the bounce buffers are part of
        // the pipeline, so a frame is cut into at least n_slots pieces for them to overlap the DMA
        if ((!pin_in || !pin_out) && ctx->chunk_bytes == 0 && fin.height >= 64)
            rows_per_chunk = std::min(rows_per_chunk, ((size_t)fin.height + ctx->n_slots - 1) / ctx->n_slots);
        for (size_t r0 = 0; r0 < fin.height; r0 += rows_per_chunk) {
            const size_t rows = std::min(rows_per_chunk, (size_t)fin.height - r0);
            Slot &s = ctx->slots[ctx->next_slot];
            ctx->next_slot = (ctx->next_slot + 1) % ctx->n_slots;
            ctx->dbg_chunks++;
            int rc = B200VF_OK;
            // Re-using a slot.  Page-locked frames: the ring's dependencies are enforced on the
            // device (the H2D waits for the kernel that last read d_in, the kernel for the D2H that
            // last read d_out), so the host runs ahead and the copy engines never wait for a host
            // wake-up.  Pageable frames go through the slot's pinned bounce buffers, which the host
            // itself reads and writes: it has to wait for the slot's previous chunk to finish.
            if (!pin_in || !pin_out || s.user_out) {
                if ((rc = drain_slot(ctx, s))) return rc;
            } else if (s.used) {
                VF_CUDA(ctx, cudaStreamWaitEvent(ctx->s_in, s.ev_k, 0));
                VF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, s.ev_d2h, 0));
            }
            if (s.d_in_cap < rows * p_in || s.d_out_cap < rows * p_out) {
                // growing a staging buffer frees the old one: nothing queued may still use it
                if ((rc = drain_slot(ctx, s))) return rc;
                VF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
            }
            if ((rc = ensure_cap(ctx, &s.d_in, &s.d_in_cap, rows * p_in, false))) return rc;
            if ((rc = ensure_cap(ctx, &s.d_out, &s.d_out_cap, rows * p_out, false))) return rc;
            const uint8_t *src = (const uint8_t *)fin.data + (int64_t)r0 * fin.stride;
            uint8_t *dst = (uint8_t *)fout.data + (int64_t)r0 * fout.stride;
            // An in-place element reads what a previous chunk's D2H may still be writing only
            // if rows overlapped; chunks are disjoint row ranges, so no hazard.
            tl_mark(ctx, ctx->s_in);
            if (pin_in && p_in == rb_in && (size_t)fin.stride == rb_in) {  // contiguous: one 1-D copy
                VF_CUDA(ctx, cudaMemcpyAsync(s.d_in, src, rows * rb_in, cudaMemcpyHostToDevice, ctx->s_in));
            } else if (pin_in) {
                VF_CUDA(ctx, cudaMemcpy2DAsync(s.d_in, p_in, src, (size_t)fin.stride, rb_in, rows,
                                               cudaMemcpyHostToDevice, ctx->s_in));
            } else {
                if ((rc = ensure_cap(ctx, &s.h_in, &s.h_in_cap, rows * p_in, true))) return rc;
                copy_rows(ctx, (uint8_t *)s.h_in, (int64_t)p_in, src, fin.stride, rb_in, rows);
                VF_CUDA(ctx, cudaMemcpyAsync(s.d_in, s.h_in, rows * p_in, cudaMemcpyHostToDevice,
                                             ctx->s_in));
            }
            ctx->stats.h2d_bytes += rows * rb_in;
            tl_mark(ctx, ctx->s_in);
            VF_CUDA(ctx, cudaEventRecord(s.ev_h2d, ctx->s_in));
            if (ctx->dbg_mode != 2) VF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, s.ev_h2d, 0));
            tl_mark(ctx, ctx->stream);
            FrameSet fs;
            fs.in[0] = (const uint8_t *)s.d_in;
            fs.out[0] = (uint8_t *)s.d_out;
            Geom g{(long long)p_in, (long long)p_out, fin.width, (uint32_t)rows};
            if (ctx->dbg_mode == 3) g.height = 1;  // experiment: the launch without the memory traffic
            cudaError_t e = ctx->dbg_mode == 1 || ctx->dbg_mode == 2 ? cudaSuccess : L.run(ctx, fs, 1, g);
            if (e != cudaSuccess) return cuda_fail(ctx, e, "kernel launch");
            tl_mark(ctx, ctx->stream);
            VF_CUDA(ctx, cudaEventRecord(s.ev_k, ctx->stream));
            if (ctx->dbg_mode != 2) VF_CUDA(ctx, cudaStreamWaitEvent(ctx->s_out, s.ev_k, 0));
            tl_mark(ctx, ctx->s_out);
            if (pin_out && p_out == rb_out && (size_t)fout.stride == rb_out) {
                VF_CUDA(ctx, cudaMemcpyAsync(dst, s.d_out, rows * rb_out, cudaMemcpyDeviceToHost, ctx->s_out));
                s.user_out = nullptr;
            } else if (pin_out) {
                VF_CUDA(ctx, cudaMemcpy2DAsync(dst, (size_t)fout.stride, s.d_out, p_out, rb_out,
                                               rows, cudaMemcpyDeviceToHost, ctx->s_out));
                s.user_out = nullptr;
            } else {
                if ((rc = ensure_cap(ctx, &s.h_out, &s.h_out_cap, rows * p_out, true))) return rc;
                VF_CUDA(ctx, cudaMemcpyAsync(s.h_out, s.d_out, rows * p_out,
                                             cudaMemcpyDeviceToHost, ctx->s_out));
                s.user_out = dst;
                s.user_stride = fout.stride;
                s.row_bytes = rb_out;
                s.rows = rows;
                s.d_pitch = p_out;
            }
            ctx->stats.d2h_bytes += rows * rb_out;
            tl_mark(ctx, ctx->s_out);
            VF_CUDA(ctx, cudaEventRecord(s.ev_d2h, ctx->s_out));
            s.busy = s.used = true;
        }
        ctx->stats.frames++;
    }
    const uint64_t ticket = ++ctx->host_ticket;
    if (ctx->host_async && all_pinned) {
        // "host.async": the frames are complete once this event is (b200vf_ctx_host_wait); nothing
        // here waits for the device, so the next call's first H2D overlaps this call's last D2H
        cudaEvent_t &e = ctx->ticket_ev[ticket % kTickets];
        if (!e) VF_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        VF_CUDA(ctx, cudaEventRecord(e, ctx->s_out));
        ctx->ticket_of[ticket % kTickets] = ticket;
        return B200VF_OK;
    }
    for (int i = 0; i < ctx->n_slots; i++) {  // host frames are complete when the call returns
        int rc = drain_slot(ctx, ctx->slots[(ctx->next_slot + i) % ctx->n_slots]);
        if (rc) return rc;
    }
    ctx->host_ticket_done = ticket;  // s_out runs in order: earlier calls are complete too
    tl_dump(ctx);
    return B200VF_OK;
}

int run_host(b200vf_ctx *ctx, const b200vf_frame *in, const b200vf_frame *out, size_t n_frames,
             int in_bpp, int out_bpp, Launcher &L) {
    int rc;
    try {
        rc = run_host_chunks(ctx, in, out, n_frames, in_bpp, out_bpp, L);
    } catch (const std::bad_alloc &) {
        rc = fail(ctx, B200VF_ERR_NOMEM, "host allocation failed");
    }
    if (rc != B200VF_OK) {
        // A failed call must not leave chunks in flight: their copy-out targets belong to the
        // caller's frames, which die when this call returns.  Let the device finish, then forget.
        cudaStreamSynchronize(ctx->s_in);
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->s_out);
        cudaGetLastError();
        for (Slot &s : ctx->slots) {
            s.busy = false;
            s.user_out = nullptr;
        }
        ctx->host_ticket_done = ctx->host_ticket;
    }
    return rc;
}

// Device-memory frames: group runs of identical geometry into batched launches.
int run_device(b200vf_ctx *ctx, const b200vf_frame *in, const b200vf_frame *out, size_t n_frames,
               Launcher &L) {
    size_t i = 0;
    while (i < n_frames) {
        if (in[i].width == 0 || in[i].height == 0) {
            i++;
            continue;
        }
        FrameSet fs;
        int n = 0;
        size_t j = i;
        while (j < n_frames && n < kMaxBatch && same_geometry(in[j], in[i]) &&
               same_geometry(out[j], out[i])) {
            fs.in[n] = (const uint8_t *)in[j].data;
            fs.out[n] = (uint8_t *)out[j].data;
            n++, j++;
        }
        Geom g{(long long)in[i].stride, (long long)out[i].stride, in[i].width, in[i].height};
        cudaError_t e = L.run(ctx, fs, n, g);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "kernel launch");
        ctx->stats.frames += (uint64_t)n;
        i = j;
    }
    return B200VF_OK;
}

int run_frames(b200vf_ctx *ctx, const b200vf_frame *in, const b200vf_frame *out, size_t n_frames,
               int in_bpp, int out_bpp, Launcher &L) {
    if (n_frames == 0) return B200VF_OK;
    const uint32_t mem = in[0].memory;
    for (size_t i = 0; i < n_frames; i++)
        if (in[i].memory != mem || out[i].memory != mem)
            return fail(ctx, B200VF_ERR_INVALID_ARG,
                        "all frames of one call must share the same memory kind");
    try {  // nothing may unwind through the C ABI
        if (mem == B200VF_MEM_HOST) return run_host(ctx, in, out, n_frames, in_bpp, out_bpp, L);
        return run_device(ctx, in, out, n_frames, L);
    } catch (const std::bad_alloc &) {
        return fail(ctx, B200VF_ERR_NOMEM, "host allocation failed");
    } catch (...) {
        return fail(ctx, B200VF_ERR_CUDA, "unexpected internal error");
    }
}

void free_device_lut(DeviceLut &lut) {
    if (lut.lut3d) cudaFree(lut.lut3d);
    if (lut.lut3d_rx) cudaFree(lut.lut3d_rx);
    if (lut.lut3d_rg) cudaFree(lut.lut3d_rg);
    if (lut.lut3d_d) cudaFree(lut.lut3d_d);
    if (lut.lut1d) cudaFree(lut.lut1d);
    lut = DeviceLut();  // lut3d_baked is borrowed from the table cache
}

void drop_baked(b200vf_ctx *ctx) {
    table_release(ctx->baked);
    ctx->baked = nullptr;
    ctx->baked_failed = false;
    ctx->lut.lut3d_baked = nullptr;
    ctx->lut.baked_interp = -1;
    ctx->lut_policy.reset();
    ctx->lut64_policy.reset();
}

void free_lut(b200vf_ctx *ctx) {
    drop_baked(ctx);
    free_device_lut(ctx->lut);
    ctx->lut_key.clear();
}

// 128 bits of FNV-1a-style hashing over the LUT's floats + kind, size, domain: the identity of a LUT
// for the table cache.
void make_lut_key(std::vector<uint8_t> &key, uint32_t kind, uint32_t size, const float *data, size_t n_floats,
                  const float scale[3], const float offset[3]) {
    uint64_t h1 = 0xcbf29ce484222325ull, h2 = 0x84222325cbf29ce4ull;
    const uint8_t *p = reinterpret_cast<const uint8_t *>(data);
    for (size_t i = 0; i < n_floats * sizeof(float); i++) {
        h1 = (h1 ^ p[i]) * 0x100000001b3ull;
        h2 = (h2 ^ p[i]) * 0x100000001b3ull + (h2 >> 29);
    }
    key.clear();
    auto put = [&](const void *v, size_t n) {
        const uint8_t *b = reinterpret_cast<const uint8_t *>(v);
        key.insert(key.end(), b, b + n);
    };
    put("LUT", 3), put(&kind, 4), put(&size, 4), put(&h1, 8), put(&h2, 8);
    put(scale, 12), put(offset, 12);
}

}  // namespace

// ============================================================================
// library
// ============================================================================
extern "C" {

const char *b200vf_version(void) { return "b200vf 0.1 (sm_100a)"; }

const char *b200vf_status_string(int status) {
    switch (status) {
    case B200VF_OK: return "ok";
    case B200VF_ERR_INVALID_ARG: return "invalid argument";
    case B200VF_ERR_UNSUPPORTED_FORMAT: return "unsupported format";
    case B200VF_ERR_CUDA: return "CUDA error";
    case B200VF_ERR_NO_LUT: return "No LUT configured";
    case B200VF_ERR_PARSE: return "invalid LUT file";
    case B200VF_ERR_IO: return "I/O error";
    case B200VF_ERR_NO_DEVICE: return "no CUDA device";
    case B200VF_ERR_NOMEM: return "out of memory";
    case B200VF_ERR_SETTINGS: return "LUT file location is not configured";
    default: return "unknown status";
    }
}

int b200vf_device_count(int *count) {
    if (!count) return fail(nullptr, B200VF_ERR_INVALID_ARG, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *count = 0;
        return fail(nullptr, B200VF_ERR_NO_DEVICE,
                    std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
    }
    *count = n;
    return B200VF_OK;
}

uint32_t b200vf_format_bytes_per_pixel(uint32_t format) {
    return format < B200VF_FORMAT_COUNT ? (uint32_t)kFormats[format].bpp : 0u;
}

const char *b200vf_format_name(uint32_t format) {
    return format < B200VF_FORMAT_COUNT ? kFormats[format].name : nullptr;
}

int b200vf_format_from_name(const char *name) {
    if (!name) return -1;
    for (int i = 0; i < B200VF_FORMAT_COUNT; i++)
        if (std::strcmp(kFormats[i].name, name) == 0) return i;
    return -1;
}

// ============================================================================
// context
// ============================================================================
int b200vf_ctx_create(int device, b200vf_ctx **out) {
    if (!out) return fail(nullptr, B200VF_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    int n = 0;
    int rc = b200vf_device_count(&n);
    if (rc) return rc;
    if (n == 0) return fail(nullptr, B200VF_ERR_NO_DEVICE, "no CUDA device present; no CPU fallback");
    if (device < 0 || device >= n)
        return fail(nullptr, B200VF_ERR_NO_DEVICE, "device index out of range");
    b200vf_ctx *ctx = new (std::nothrow) b200vf_ctx();
    if (!ctx) return fail(nullptr, B200VF_ERR_NOMEM, "context allocation failed");
    ctx->device = device;
    cudaError_t e = cudaSetDevice(device);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    ctx->own_stream = (e == cudaSuccess);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_in, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_out, cudaStreamNonBlocking);
    for (int i = 0; i < kMaxSlots && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&ctx->slots[i].ev_h2d, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->slots[i].ev_k, cudaEventDisableTiming);
        if (e == cudaSuccess)
            e = cudaEventCreateWithFlags(&ctx->slots[i].ev_d2h, cudaEventDisableTiming);
    }
    if (e != cudaSuccess) {
        int code = cuda_fail(nullptr, e, "context creation");
        b200vf_ctx_destroy(ctx);
        return code;
    }
    *out = ctx;
    return B200VF_OK;
}

void b200vf_ctx_destroy(b200vf_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->s_in) cudaStreamSynchronize(ctx->s_in);
    if (ctx->s_out) cudaStreamSynchronize(ctx->s_out);
    free_lut(ctx);
    unregister_all(ctx);
    for (Slot &s : ctx->slots) {
        if (s.d_in) cudaFree(s.d_in);
        if (s.d_out) cudaFree(s.d_out);
        if (s.h_in) cudaFreeHost(s.h_in);
        if (s.h_out) cudaFreeHost(s.h_out);
        if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
        if (s.ev_k) cudaEventDestroy(s.ev_k);
        if (s.ev_d2h) cudaEventDestroy(s.ev_d2h);
    }
    delete ctx->pool;
    table_release(ctx->fn.shared);
    ctx->fn.shared = nullptr;
    ctx->fn.policy.destroy();
    ctx->lut_policy.destroy();
    ctx->lut64_policy.destroy();
    if (ctx->ev_order) cudaEventDestroy(ctx->ev_order);
    for (cudaEvent_t e : ctx->ticket_ev)
        if (e) cudaEventDestroy(e);
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->s_in) cudaStreamDestroy(ctx->s_in);
    if (ctx->s_out) cudaStreamDestroy(ctx->s_out);
    cudaGetLastError();
    delete ctx;
}

const char *b200vf_last_error(const b200vf_ctx *ctx) {
    return ctx ? ctx->last_error.c_str() : g_last_error.c_str();
}

int b200vf_ctx_device(const b200vf_ctx *ctx) { return ctx ? ctx->device : -1; }

int b200vf_ctx_synchronize(b200vf_ctx *ctx) {
    int rc = activate(ctx);
    if (rc) return rc;
    VF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    VF_CUDA(ctx, cudaStreamSynchronize(ctx->s_in));
    VF_CUDA(ctx, cudaStreamSynchronize(ctx->s_out));
    ctx->host_ticket_done = ctx->host_ticket;
    return B200VF_OK;
}

uint64_t b200vf_ctx_host_ticket(const b200vf_ctx *ctx) { return ctx ? ctx->host_ticket : 0; }

int b200vf_ctx_host_wait(b200vf_ctx *ctx, uint64_t ticket) {
    int rc = activate(ctx);
    if (rc) return rc;
    if (ticket > ctx->host_ticket)
        return fail(ctx, B200VF_ERR_INVALID_ARG, "host_wait: no host-frame call has this ticket yet");
    return host_wait(ctx, ticket);
}

void *b200vf_ctx_get_stream(const b200vf_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

int b200vf_ctx_set_stream(b200vf_ctx *ctx, void *cuda_stream) {
    int rc = activate(ctx);
    if (rc) return rc;
    // tables built on the old stream must be complete before work on the new one reads them
    // (a NULL handle is the legacy default stream and synchronises like any other)
    VF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
    ctx->stream = (cudaStream_t)cuda_stream;
    ctx->own_stream = false;
    return B200VF_OK;
}

int b200vf_ctx_wait_for(b200vf_ctx *ctx, b200vf_ctx *upstream) {
    int rc = activate(ctx);
    if (rc) return rc;
    if (!upstream) return fail(ctx, B200VF_ERR_INVALID_ARG, "wait_for: upstream is NULL");
    if (upstream == ctx || upstream->stream == ctx->stream) return B200VF_OK;  // already ordered
    if (upstream->device != ctx->device)
        return fail(ctx, B200VF_ERR_INVALID_ARG, "wait_for: the two contexts are on different devices");
    if (!ctx->ev_order) VF_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_order, cudaEventDisableTiming));
    VF_CUDA(ctx, cudaEventRecord(ctx->ev_order, upstream->stream));
    VF_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_order, 0));
    return B200VF_OK;
}

int b200vf_ctx_set_option(b200vf_ctx *ctx, const char *key, int64_t value) {
    if (!ctx || !key) return fail(ctx, B200VF_ERR_INVALID_ARG, "set_option: NULL argument");
    if (!std::strcmp(key, "hsv.math")) {
        if (value != kMathFast && value != kMathPlain)
            return fail(ctx, B200VF_ERR_INVALID_ARG, "hsv.math must be 0 or 1");
        ctx->math_mode = (int)value;
    } else if (!std::strcmp(key, "lut.path")) {
        if (value < kLutAuto || value > kLutBaked)
            return fail(ctx, B200VF_ERR_INVALID_ARG, "lut.path must be 0..4");
        ctx->lut_path = (int)value;
    } else if (!std::strcmp(key, "hsv.path")) {
        if (value < kFnAuto || value > kFnTable)
            return fail(ctx, B200VF_ERR_INVALID_ARG, "hsv.path must be 0..2");
        ctx->fn_path = (int)value;
    } else if (!std::strcmp(key, "lut.interpolation")) {
        if (value < kInterpTrilinear || value > kInterpNearest)
            return fail(ctx, B200VF_ERR_INVALID_ARG, "lut.interpolation must be 0..2");
        ctx->lut_interp = (int)value;
    } else if (!std::strcmp(key, "tables.share")) {
        if (value != 0 && value != 1) return fail(ctx, B200VF_ERR_INVALID_ARG, "tables.share must be 0 or 1");
        ctx->share_tables = value != 0;
    } else if (!std::strcmp(key, "host.chunk_bytes")) {
        if (value != 0 && value < 4096) return fail(ctx, B200VF_ERR_INVALID_ARG, "host.chunk_bytes too small");
        ctx->chunk_bytes = value;
    } else if (!std::strcmp(key, "host.dbg_mode")) {
        // 4 = CUDA-event timeline of a host-frame call on stderr.  1-3 leave the kernels out (wrong
        // pixels!) to attribute the pipeline's time; only with B200VF_ALLOW_DEBUG_MODES=1 in the
        // environment (tools/host_path_probe.py).
        const bool skips_work = value >= 1 && value <= 3;
        if (value < 0 || value > 4 || (skips_work && !std::getenv("B200VF_ALLOW_DEBUG_MODES")))
            return fail(ctx, B200VF_ERR_INVALID_ARG, "host.dbg_mode must be 0 or 4");
        ctx->dbg_mode = (int)value;
    } else if (!std::strcmp(key, "host.async")) {
        if (value != 0 && value != 1) return fail(ctx, B200VF_ERR_INVALID_ARG, "host.async must be 0 or 1");
        if (!value && ctx->host_async) {  // back to synchronous calls: nothing stays in flight
            int rc = b200vf_ctx_synchronize(ctx);
            if (rc) return rc;
        }
        ctx->host_async = (int)value;
    } else if (!std::strcmp(key, "host.slots")) {
        if (value < 2 || value > kMaxSlots) return fail(ctx, B200VF_ERR_INVALID_ARG, "host.slots must be 2..8");
        if (ctx->host_ticket_done != ctx->host_ticket) {  // "host.async" calls in flight use the ring
            int rc = b200vf_ctx_synchronize(ctx);
            if (rc) return rc;
        }
        ctx->n_slots = (int)value;
        ctx->next_slot = 0;  // every slot is idle now
    } else if (!std::strcmp(key, "host.register")) {
        if (value != 0 && value != 1) return fail(ctx, B200VF_ERR_INVALID_ARG, "host.register must be 0 or 1");
        if (!value) unregister_all(ctx);
        ctx->register_mode = (int)value;
    } else if (!std::strcmp(key, "host.register_budget")) {
        if (value < 0) return fail(ctx, B200VF_ERR_INVALID_ARG, "host.register_budget must be >= 0");
        ctx->register_budget = (size_t)value;
    } else if (!std::strcmp(key, "host.copy_threads")) {
        if (value < 1 || value > 64)
            return fail(ctx, B200VF_ERR_INVALID_ARG, "host.copy_threads must be 1..64");
        if (ctx->pool && value != ctx->copy_threads) {
            delete ctx->pool;
            ctx->pool = nullptr;
        }
        ctx->copy_threads = (int)value;
    } else {
        return fail(ctx, B200VF_ERR_INVALID_ARG, std::string("unknown option ") + key);
    }
    return B200VF_OK;
}

int b200vf_ctx_get_option(const b200vf_ctx *ctx, const char *key, int64_t *value) {
    if (!ctx || !key || !value) return B200VF_ERR_INVALID_ARG;
    if (!std::strcmp(key, "hsv.math"))
        *value = ctx->math_mode;
    else if (!std::strcmp(key, "lut.path"))
        *value = ctx->lut_path;
    else if (!std::strcmp(key, "lut.interpolation"))
        *value = ctx->lut_interp;
    else if (!std::strcmp(key, "lut.path_active"))  // read-only: kernel of the last colorlut launch:
        *value = ctx->lut_path_active;              // 0 direct, 1 R-, 3 RG-resampled, 2 1D, 4 baked, 5/6 extensions, 7 16-bit delta-table op
    else if (!std::strcmp(key, "hsv.path"))
        *value = ctx->fn_path;
    else if (!std::strcmp(key, "hsv.table_active"))  // read-only: did the last launch use the table
        *value = ctx->fn.last_used_table ? 1 : 0;
    else if (!std::strcmp(key, "tables.share"))
        *value = ctx->share_tables ? 1 : 0;
    else if (!std::strcmp(key, "lut.tables_built"))  // read-only bit mask: 1 R-resampled, 2 RG-resampled, 4 baked
        *value = (ctx->lut.lut3d_rx ? 1 : 0) | (ctx->lut.lut3d_rg ? 2 : 0) | (ctx->lut.lut3d_baked ? 4 : 0);
    else if (!std::strcmp(key, "tables.device_bytes")) {  // read-only: function tables cached on this device
        uint64_t bytes = 0;
        table_cache_stats(ctx->device, nullptr, &bytes);
        *value = (int64_t)bytes;
    } else if (!std::strcmp(key, "tables.device_count")) {
        uint64_t cnt = 0;
        table_cache_stats(ctx->device, &cnt, nullptr);
        *value = (int64_t)cnt;
    } else if (!std::strcmp(key, "host.chunk_bytes"))
        *value = ctx->chunk_bytes;
    else if (!std::strcmp(key, "host.copy_threads"))
        *value = ctx->copy_threads;
    else if (!std::strcmp(key, "host.slots"))
        *value = ctx->n_slots;
    else if (!std::strcmp(key, "host.async"))
        *value = ctx->host_async;
    else if (!std::strcmp(key, "host.dbg_chunks"))
        *value = (int64_t)ctx->dbg_chunks;
    else if (!std::strcmp(key, "host.dbg_wait_ns"))   // blocked in cudaEventSynchronize for a slot
        *value = (int64_t)ctx->dbg_wait_ns;
    else if (!std::strcmp(key, "host.dbg_call_ns"))   // inside host-frame calls, total
        *value = (int64_t)ctx->dbg_call_ns;
    else if (!std::strcmp(key, "host.register"))
        *value = ctx->register_mode;
    else if (!std::strcmp(key, "host.register_budget"))
        *value = (int64_t)ctx->register_budget;
    else if (!std::strcmp(key, "host.registered_bytes"))  // read-only: page-locked in place by this context
        *value = (int64_t)ctx->registered_bytes;
    else
        return B200VF_ERR_INVALID_ARG;
    return B200VF_OK;
}

int b200vf_ctx_get_stats(const b200vf_ctx *ctx, b200vf_stats *out) {
    if (!ctx || !out) return B200VF_ERR_INVALID_ARG;
    *out = ctx->stats;
    return B200VF_OK;
}

int b200vf_ctx_reset_stats(b200vf_ctx *ctx) {
    if (!ctx) return B200VF_ERR_INVALID_ARG;
    ctx->stats = b200vf_stats{};
    return B200VF_OK;
}

// ============================================================================
// memory helpers
// ============================================================================
int b200vf_host_alloc(size_t bytes, void **out) {
    if (!out) return fail(nullptr, B200VF_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, B200VF_ERR_NOMEM, std::string("cudaMallocHost: ") + cudaGetErrorString(e));
    }
    return B200VF_OK;
}

int b200vf_host_free(void *p) {
    if (p && cudaFreeHost(p) != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, B200VF_ERR_CUDA, "cudaFreeHost failed");
    }
    return B200VF_OK;
}

int b200vf_host_is_pinned(const void *p) { return p && is_pinned(p) ? 1 : 0; }

int b200vf_ctx_host_memory_released(b200vf_ctx *ctx, const void *p, size_t bytes) {
    int rc = activate(ctx);
    if (rc) return rc;
    if (!p) return B200VF_OK;
    // copies that touch the range were complete when their call returned, or ("host.async") are
    // waited for before the range is unregistered
    forget_range(ctx, p, bytes);
    return B200VF_OK;
}

int b200vf_device_alloc(b200vf_ctx *ctx, size_t bytes, void **out) {
    if (!out) return fail(ctx, B200VF_ERR_INVALID_ARG, "out is NULL");
    *out = nullptr;
    int rc = activate(ctx);
    if (rc) return rc;
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(ctx, B200VF_ERR_NOMEM, std::string("cudaMalloc: ") + cudaGetErrorString(e));
    }
    return B200VF_OK;
}

int b200vf_device_free(b200vf_ctx *ctx, void *p) {
    int rc = activate(ctx);
    if (rc) return rc;
    if (p) VF_CUDA(ctx, cudaFree(p));
    return B200VF_OK;
}

int b200vf_memcpy(b200vf_ctx *ctx, void *dst, const void *src, size_t bytes, int kind) {
    int rc = activate(ctx);
    if (rc) return rc;
    if (bytes == 0) return B200VF_OK;
    if (!dst || !src) return fail(ctx, B200VF_ERR_INVALID_ARG, "memcpy: NULL pointer");
    cudaMemcpyKind k = kind == 0   ? cudaMemcpyHostToDevice
                       : kind == 1 ? cudaMemcpyDeviceToHost
                       : kind == 2 ? cudaMemcpyDeviceToDevice
                                   : cudaMemcpyDefault;
    if (kind < 0 || kind > 2) return fail(ctx, B200VF_ERR_INVALID_ARG, "memcpy: bad kind");
    VF_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, k, ctx->stream));
    if (kind != 2) VF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return B200VF_OK;
}

// ============================================================================
// .cube parser
// ============================================================================
static int export_cube(const CubeData &cd, b200vf_cube *out) {
    out->kind = (uint32_t)cd.kind;
    out->size = cd.size;
    for (int c = 0; c < 3; c++) {
        out->domain_scale[c] = cd.domain_scale[c];
        out->domain_offset[c] = cd.domain_offset[c];
    }
    out->n_floats = cd.data.size();
    out->data = (float *)std::malloc(std::max<size_t>(1, cd.data.size()) * sizeof(float));
    if (!out->data) return B200VF_ERR_NOMEM;
    std::memcpy(out->data, cd.data.data(), cd.data.size() * sizeof(float));
    return B200VF_OK;
}

static void copy_err(const std::string &msg, char *err, size_t errlen) {
    if (err && errlen) std::snprintf(err, errlen, "%s", msg.c_str());
}

int b200vf_cube_parse(const char *text, size_t len, b200vf_cube *out, char *err, size_t errlen) {
    if (!out || (!text && len)) return fail(nullptr, B200VF_ERR_INVALID_ARG, "cube_parse: NULL argument");
    std::memset(out, 0, sizeof *out);
    try {
        CubeData cd;
        std::string msg;
        int rc = parse_cube_text(text ? text : "", len, cd, msg);
        if (rc) {
            copy_err(msg, err, errlen);
            return fail(nullptr, rc == 2 ? B200VF_ERR_IO : B200VF_ERR_PARSE, msg);
        }
        return export_cube(cd, out);
    } catch (...) {
        return fail(nullptr, B200VF_ERR_NOMEM, "cube_parse: out of memory");
    }
}

int b200vf_cube_parse_file(const char *path, b200vf_cube *out, char *err, size_t errlen) {
    if (!out || !path) return fail(nullptr, B200VF_ERR_INVALID_ARG, "cube_parse_file: NULL argument");
    std::memset(out, 0, sizeof *out);
    try {
        CubeData cd;
        std::string msg;
        int rc = parse_cube_file(path, cd, msg);
        if (rc) {
            copy_err(msg, err, errlen);
            return fail(nullptr, rc == 2 ? B200VF_ERR_IO : B200VF_ERR_PARSE, msg);
        }
        return export_cube(cd, out);
    } catch (...) {
        return fail(nullptr, B200VF_ERR_NOMEM, "cube_parse_file: out of memory");
    }
}

void b200vf_cube_free(b200vf_cube *cube) {
    if (!cube) return;
    std::free(cube->data);
    std::memset(cube, 0, sizeof *cube);
}

// ============================================================================
// colorlut
// ============================================================================
int b200vf_colorlut_set_lut(b200vf_ctx *ctx, uint32_t kind, uint32_t size, const float *data,
                            const float domain_scale[3], const float domain_offset[3]) {
    int rc = activate(ctx);
    if (rc) return rc;
    if (!data || !domain_scale || !domain_offset)
        return fail(ctx, B200VF_ERR_INVALID_ARG, "set_lut: NULL argument");
    if (kind == B200VF_LUT_1D) {
        if (size < 2 || size > 65536) return fail(ctx, B200VF_ERR_INVALID_ARG, "set_lut: 1D size out of 2..=65536");
    } else if (kind == B200VF_LUT_3D) {
        if (size < 2 || size > 256) return fail(ctx, B200VF_ERR_INVALID_ARG, "set_lut: 3D size out of 2..=256");
    } else {
        return fail(ctx, B200VF_ERR_INVALID_ARG, "set_lut: kind must be 1 or 3");
    }
    // The new LUT is built completely in `L` first; the context's LUT is replaced only when the
    // upload has succeeded, so a failed call leaves the previous LUT (or none) in place and a
    // later colorlut_process never sees a half-initialised one.
    DeviceLut L;
    auto abandon = [&](int code, const std::string &msg) {
        free_device_lut(L);
        cudaGetLastError();
        return fail(ctx, code, msg);
    };
    try {
        L.kind = (int)kind;
        L.size = size;
        L.identity_domain = true;
        for (int c = 0; c < 3; c++) {
            L.scale[c] = domain_scale[c];
            L.offset[c] = domain_offset[c];
            if (!(domain_scale[c] == 1.0f && domain_offset[c] == 0.0f)) L.identity_domain = false;
        }
        const size_t n = size, np = n + 1;
        const size_t count = kind == B200VF_LUT_1D ? 3 * n : 4 * n * n * n;
        {
            bool unit = true;
            for (size_t i = 0; i < count && unit; i++) unit = data[i] >= 0.0f && data[i] <= 1.0f;
            L.unit_range = unit;  // false for NaN / inf / out-of-range entries
        }
        std::vector<float> host;
        float **slot;
        if (kind == B200VF_LUT_1D) {
            host.resize(3 * np);
            for (int c = 0; c < 3; c++) {
                std::memcpy(&host[c * np], data + (size_t)c * n, n * sizeof(float));
                host[c * np + n] = data[(size_t)c * n + n - 1];
            }
            slot = &L.lut1d;
        } else {
            // pad to (N+1)^3, duplicating the far faces: corner x0+1 of the reference's
            // min(x0+1, N-1) clamp (imp.rs:500-502) becomes a plain +1 offset
            host.resize(np * np * np * 4);
            for (size_t z = 0; z < np; z++)
                for (size_t y = 0; y < np; y++) {
                    const size_t zs = std::min(z, n - 1), ys = std::min(y, n - 1);
                    const float *srow = data + 4 * (ys * n + zs * n * n);
                    float *drow = &host[4 * (y * np + z * np * np)];
                    // entry = {R(x), R(x+1), G(x), B(x)}: the four floats of corner x plus
                    // the red of corner x+1, so a pixel fetches 16 + 8 bytes per row pair
                    // instead of 16 + 16 (lane 3 of the file layout is the constant 1.0)
                    for (size_t x = 0; x < np; x++) {
                        const float *s0 = srow + 4 * std::min(x, n - 1);
                        const float *s1 = srow + 4 * std::min(x + 1, n - 1);
                        float *d = drow + 4 * x;
                        d[0] = s0[0], d[1] = s1[0], d[2] = s0[1], d[3] = s0[2];
                    }
                }
            slot = reinterpret_cast<float **>(&L.lut3d);
        }
        if (cudaMalloc((void **)slot, host.size() * sizeof(float)) != cudaSuccess)
            return abandon(B200VF_ERR_NOMEM, "set_lut: device allocation failed");
        cudaError_t e = cudaMemcpy(*slot, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice);
        if (e != cudaSuccess)
            return abandon(B200VF_ERR_CUDA, std::string("set_lut: upload failed: ") + cudaGetErrorString(e));
        std::vector<uint8_t> key;
        make_lut_key(key, kind, size, data, count, domain_scale, domain_offset);
        // nobody may still read the old LUT when it is freed
        e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess)
            return abandon(B200VF_ERR_CUDA, std::string("set_lut: ") + cudaGetErrorString(e));
        free_lut(ctx);
        ctx->lut = L;
        ctx->lut_key.swap(key);
        // The tables derived from the LUT (baked to 8-bit resolution, R- / RG-resampled) are built
        // on first use by the path that needs them (ensure_baked / ensure_resampled).
        return B200VF_OK;
    } catch (const std::bad_alloc &) {
        return abandon(B200VF_ERR_NOMEM, "set_lut: host allocation failed");
    }
}

int b200vf_colorlut_set_lut_file(b200vf_ctx *ctx, const char *location) {
    int rc = activate(ctx);
    if (rc) return rc;
    if (!location)  // colorlut/imp.rs:175-180
        return fail(ctx, B200VF_ERR_SETTINGS, "LUT file location is not configured");
    try {
        CubeData cd;
        std::string msg;
        int prc = parse_cube_file(location, cd, msg);
        if (prc)  // colorlut/imp.rs:182-187
            return fail(ctx, prc == 2 ? B200VF_ERR_IO : B200VF_ERR_PARSE,
                        std::string("Failed to parse LUT file ") + location + ": " + msg);
        return b200vf_colorlut_set_lut(ctx, (uint32_t)cd.kind, cd.size, cd.data.data(),
                                       cd.domain_scale, cd.domain_offset);
    } catch (...) {
        return fail(ctx, B200VF_ERR_NOMEM, "set_lut_file: out of memory");
    }
}

int b200vf_colorlut_clear_lut(b200vf_ctx *ctx) {
    int rc = activate(ctx);
    if (rc) return rc;
    VF_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    free_lut(ctx);
    return B200VF_OK;
}

namespace {
// The table baked to native 8-bit resolution (the default for 8-bit frames) is built once per
// LUT content and interpolation mode by the direct kernel — by whichever context on the device asks
// first (device-wide cache, vf_tables.cpp) — stream-ordered before its first use.
cudaError_t ensure_baked(b200vf_ctx *ctx, int bits, bool force = false) {
    const bool want = force || ctx->lut_path == kLutBaked || ctx->lut_path == kLutAuto;
    if (!want || bits != 8 || (ctx->lut.kind != 3 && !force)) return cudaSuccess;
    if (ctx->baked && ctx->lut.baked_interp == ctx->lut_interp) return cudaSuccess;
    if (ctx->baked) drop_baked(ctx);  // interpolation changed
    if (ctx->baked_failed) return cudaSuccess;
    std::vector<uint8_t> key = ctx->lut_key;
    key.push_back((uint8_t)'B');
    key.push_back((uint8_t)ctx->lut_interp);
    if (!ctx->share_tables) key_put(key, ctx);  // a private table: nobody else has this key
    SharedTable *t = table_acquire(ctx->device, key);
    if (!t) {
        ctx->baked_failed = true;  // not enough memory: the interpolating kernels serve instead
        return cudaSuccess;
    }
    cudaError_t e = table_ensure_built(t, ctx->stream, [&](uint32_t *dst) {
        return launch_build_baked(ctx->stream, ctx->lut, dst, ctx->lut_interp, &ctx->stats.kernel_launches);
    });
    if (e != cudaSuccess) {
        table_release(t);
        return e;
    }
    ctx->baked = t;
    ctx->lut.lut3d_baked = t->data;
    ctx->lut.baked_interp = ctx->lut_interp;
    return cudaSuccess;
}

// R- (and RG-) resampled tables of the interpolating 8-bit kernels: built when a launch is about to
// use them ("lut.path" = 2 / 3, or the baked table could not be had), not at every set_lut.
cudaError_t ensure_resampled(b200vf_ctx *ctx, bool want_rg) {
    DeviceLut &L = ctx->lut;
    if (L.kind != 3) return cudaSuccess;
    const size_t np = (size_t)L.size + 1;
    bool build_rx = false, build_rg = false;
    if (!L.lut3d_rx) {
        if (cudaMalloc((void **)&L.lut3d_rx, np * np * 256 * sizeof(float4)) != cudaSuccess) {
            cudaGetLastError();
            L.lut3d_rx = nullptr;  // optional tables: the direct path serves
            return cudaSuccess;
        }
        build_rx = true;
    }
    if (want_rg && !L.lut3d_rg && L.size <= 71) {  // (N+1) MiB, keeps the table inside the 126 MB L2
        if (cudaMalloc((void **)&L.lut3d_rg, np * 65536 * sizeof(float4)) != cudaSuccess) {
            cudaGetLastError();
            L.lut3d_rg = nullptr;
        } else {
            build_rg = true;
        }
    }
    if (!build_rx && !build_rg) return cudaSuccess;
    return launch_build_resampled(ctx->stream, L, build_rx, build_rg, &ctx->stats.kernel_launches);
}

// The 16-bit fast op's coordinate arithmetic (vf_ops.cuh ColorLut64Op::coord) against the reference
// formula `(c as f32 / 65535.0) * (N as f32 - 1.0)`, floor, subtract — for every 16-bit code.
bool coords16_match(uint32_t n, bool pow2, float khi, float klo) {
    const float sm1 = (float)n - 1.0f;
    for (uint32_t code = 0; code < 65536; code++) {
        const float c = (float)code;
        volatile float v = c / 65535.0f;
        volatile float pr = v * sm1;  // volatile: two separately rounded operations, no contraction
        const float p_ref = pr;
        float p;
        if (pow2) {
            volatile float lo = c * klo;
            p = std::fmaf(c, khi, lo);
        } else {
            volatile float lo = c * 0x1.0001p-48f;  // VF_K65535_LO
            volatile float q = std::fmaf(c, 0x1.0001p-16f, lo);
            volatile float pp = q * sm1;
            p = pp;
        }
        if (std::memcmp(&p, &p_ref, 4) != 0) return false;
        // floor and the subtraction are the same operations on both sides once p is equal
    }
    return true;
}

// Table of the RGBA64 fast op (3D LUT, identity domain, N <= 128), built on the first 16-bit frame.
cudaError_t ensure_lut64(b200vf_ctx *ctx, int bits) {
    DeviceLut &L = ctx->lut;
    if (bits != 16 || L.kind != 3 || !L.identity_domain || L.size > 128 || L.lut3d_d || L.lut64_failed)
        return cudaSuccess;
    if (ctx->lut_interp != kInterpTrilinear || ctx->math_mode == kMathPlain) return cudaSuccess;
    const uint32_t m = L.size - 1;
    L.sm1_pow2 = (m & (m - 1)) == 0;
    const double k = (double)m / 65535.0;
    L.k16_hi = (float)k;
    L.k16_lo = (float)(k - (double)L.k16_hi);
    L.coords16_ok = coords16_match(L.size, L.sm1_pow2, L.k16_hi, L.k16_lo);
    if (!L.coords16_ok) {
        L.lut64_failed = true;  // keep the direct kernel
        return cudaSuccess;
    }
    L.lut3d_d_stride = L.size + 1 <= 65 ? 65 : 129;
    const size_t entries = (size_t)L.lut3d_d_stride * L.lut3d_d_stride * (L.size + 1);
    if (cudaMalloc((void **)&L.lut3d_d, entries * 32) != cudaSuccess) {
        cudaGetLastError();
        L.lut3d_d = nullptr;
        L.lut64_failed = true;
        return cudaSuccess;
    }
    return launch_build_lut64(ctx->stream, L, &ctx->stats.kernel_launches);
}

// Tables for the 3D paths of this launch, per "lut.path" / "lut.interpolation".
cudaError_t ensure_lut_tables(b200vf_ctx *ctx, int bits) {
    cudaError_t e = ensure_lut64(ctx, bits);
    if (e != cudaSuccess) return e;
    e = ensure_baked(ctx, bits);
    if (e != cudaSuccess || bits != 8 || ctx->lut.kind != 3) return e;
    const bool baked_serves = ctx->lut.lut3d_baked && (ctx->lut_path == kLutAuto || ctx->lut_path == kLutBaked);
    if (baked_serves || ctx->lut_interp != kInterpTrilinear || ctx->lut_path == kLutDirect) return cudaSuccess;
    return ensure_resampled(ctx, ctx->lut_path != kLutResampledR && ctx->math_mode != kMathPlain);
}

struct ColorLutLauncher : Launcher {
    int bits;
    bool be;
    cudaError_t run(b200vf_ctx *ctx, const FrameSet &fs, int n, const Geom &g) override {
        cudaError_t e = ensure_lut_tables(ctx, bits);
        if (e != cudaSuccess) return e;
        int path = ctx->lut_path;
        const int resolved = resolved_lut_path(ctx->lut, bits, ctx->math_mode, path, ctx->lut_interp);
        // auto: the baked table is a gather (content-sensitive, and it competes for L2 with every
        // other table on the device); the direct kernel interpolates from the small LUT itself.
        // Both are timed on the stream's own frames and the faster one serves.
        bool timed = false;
        int mode = 1;
        const uint64_t pixels = (uint64_t)n * g.width * g.height;
        // (8-bit: baked table vs direct kernel; 16-bit: the delta-table op, bound by the L1 return
        // path, vs the direct kernel, bound by instruction issue — noisy content favours the latter)
        if (path == kLutAuto && (resolved == 4 || resolved == 7) && ctx->math_mode == kMathFast) {
            PathPolicy &pol = resolved == 4 ? ctx->lut_policy : ctx->lut64_policy;
            mode = ctx->in_host_call ? pol.chosen : pol.next(pixels, &timed);
            if (mode == 0) path = kLutDirect;
        }
        ctx->lut_path_active = resolved_lut_path(ctx->lut, bits, ctx->math_mode, path, ctx->lut_interp);
        PathPolicy &pol = bits == 8 ? ctx->lut_policy : ctx->lut64_policy;
        if (timed) pol.begin(ctx->stream);
        e = launch_colorlut(ctx->stream, fs, n, g, bits, be, ctx->lut, ctx->math_mode, path,
                            ctx->lut_interp, &ctx->stats.kernel_launches);
        if (timed) pol.end(ctx->stream, mode, pixels);
        return e;
    }
};
// colorlut with the videoconvert either side folded in: needs the baked table, whatever "lut.path" says
struct ColorLutConvertLauncher : Launcher {
    PixLayout in_lay, out_lay;
    cudaError_t run(b200vf_ctx *ctx, const FrameSet &fs, int n, const Geom &g) override {
        cudaError_t e = ensure_baked(ctx, 8, /*force=*/true);
        if (e != cudaSuccess) return e;
        if (!ctx->lut.lut3d_baked) return cudaErrorMemoryAllocation;
        ctx->lut_path_active = 4;
        return launch_colorlut_convert(ctx->stream, fs, n, g, in_lay, out_lay, ctx->lut.lut3d_baked,
                                       &ctx->stats.kernel_launches);
    }
};

// ---- tabulated element functions ------------------------------------------------------------

// compute(fs, n, g, table_build): the element's exact kernel(s).  With table_build the frame is
// the 4096x4096 table itself, 4 bytes per pixel in place, whatever the stream's real format is.
using ComputeFn = std::function<cudaError_t(const FrameSet &, int, const Geom &, bool)>;

cudaError_t fn_build(b200vf_ctx *ctx, bool colour_at_1, const ComputeFn &compute) {
    FnTable &t = ctx->fn;
    if (!t.shared) {
        if (t.alloc_failed) return cudaErrorMemoryAllocation;
        std::vector<uint8_t> key = t.key;
        if (!ctx->share_tables) key_put(key, ctx);
        t.shared = table_acquire(ctx->device, key);  // another context may already hold this function
        if (!t.shared) {
            t.alloc_failed = true;  // stay on the compute kernels
            return cudaErrorMemoryAllocation;
        }
    }
    cudaError_t e = table_ensure_built(t.shared, ctx->stream, [&](uint32_t *table) {
        cudaError_t r = launch_table_fill(ctx->stream, table, colour_at_1, &ctx->stats.kernel_launches);
        if (r != cudaSuccess) return r;
        FrameSet fs;
        fs.in[0] = reinterpret_cast<const uint8_t *>(table);
        fs.out[0] = reinterpret_cast<uint8_t *>(table);
        const Geom g{4096 * 4, 4096 * 4, 4096, 4096};
        return compute(fs, 1, g, true);
    });
    if (e == cudaSuccess) t.built = true;
    return e;
}

// One launch of an element: table or compute kernel, per "hsv.path" and (auto) the measured times.
cudaError_t fn_dispatch(b200vf_ctx *ctx, const std::vector<uint8_t> &key, bool colour_at_1,
                        bool keep_other, int in_bpp, int out_bpp, const FrameSet &fs, int n,
                        const Geom &g, const ComputeFn &compute) {
    FnTable &t = ctx->fn;
    const uint64_t pixels = (uint64_t)n * g.width * g.height;
    if (ctx->math_mode == kMathPlain || ctx->fn_path == kFnCompute) {
        t.last_used_table = false;
        return compute(fs, n, g, false);
    }
    if (key != t.key) {  // settings changed: the table (if any) describes another function
        table_release(t.shared);
        t.shared = nullptr;
        t.alloc_failed = false;
        t.key = key;
        t.built = false;
        t.stable_pixels = 0;
        t.policy.reset();
    }
    if (!t.built) {
        const bool due = ctx->fn_path == kFnTable || t.stable_pixels >= kFnStablePixels;
        if (due && !t.alloc_failed) {
            cudaError_t e = fn_build(ctx, colour_at_1, compute);
            if (e != cudaSuccess && e != cudaErrorMemoryAllocation) return e;
        }
        if (!t.built) {
            t.stable_pixels += pixels;
            t.last_used_table = false;
            return compute(fs, n, g, false);
        }
    }
    int mode = 1;
    bool timed = false;
    if (ctx->fn_path == kFnAuto) mode = ctx->in_host_call ? t.policy.chosen : t.policy.next(pixels, &timed);
    if (timed) t.policy.begin(ctx->stream);
    cudaError_t e = mode == 1 ? launch_table_map(ctx->stream, fs, n, g, in_bpp, out_bpp, t.shared->data,
                                                 colour_at_1, keep_other, &ctx->stats.kernel_launches)
                              : compute(fs, n, g, false);
    if (timed) t.policy.end(ctx->stream, mode, pixels);
    t.last_used_table = mode == 1;
    return e;
}

struct HsvFilterLauncher : Launcher {
    PixLayout lay;
    HsvFilterArgs a;
    cudaError_t run(b200vf_ctx *ctx, const FrameSet &fs, int n, const Geom &g) override {
        std::vector<uint8_t> key;
        key_put(key, (uint8_t)1);
        key_put(key, lay.r), key_put(key, lay.g), key_put(key, lay.b);
        key_put(key, a);
        const PixLayout lay4{4, lay.r, lay.g, lay.b, lay.r == 0 || lay.b == 0 ? 3 : 0};
        return fn_dispatch(ctx, key, /*colour_at_1=*/lay.r == 1 || lay.b == 1, /*keep_other=*/true,
                           lay.bpp, lay.bpp, fs, n, g,
                           [&](const FrameSet &f, int m, const Geom &gg, bool build) {
                               return launch_hsvfilter(ctx->stream, f, m, gg, build ? lay4 : lay, a,
                                                       ctx->math_mode, &ctx->stats.kernel_launches);
                           });
    }
};
struct HsvDetectLauncher : Launcher {
    PixLayout in_lay, out_lay;
    HsvDetectArgs a;
    cudaError_t run(b200vf_ctx *ctx, const FrameSet &fs, int n, const Geom &g) override {
        std::vector<uint8_t> key;
        key_put(key, (uint8_t)2);
        key_put(key, in_lay.r), key_put(key, in_lay.g), key_put(key, in_lay.b);
        key_put(key, out_lay.r), key_put(key, out_lay.g), key_put(key, out_lay.b);
        key_put(key, a);
        const PixLayout in4{4, in_lay.r, in_lay.g, in_lay.b, in_lay.r == 0 || in_lay.b == 0 ? 3 : 0};
        return fn_dispatch(ctx, key, in_lay.r == 1 || in_lay.b == 1, /*keep_other=*/false,
                           in_lay.bpp, out_lay.bpp, fs, n, g,
                           [&](const FrameSet &f, int m, const Geom &gg, bool build) {
                               return launch_hsvdetector(ctx->stream, f, m, gg, build ? in4 : in_lay,
                                                         out_lay, a, ctx->math_mode,
                                                         &ctx->stats.kernel_launches);
                           });
    }
};
struct ChainLauncher : Launcher {
    HsvFilterArgs a;
    cudaError_t compute(b200vf_ctx *ctx, const FrameSet &fs, int n, const Geom &g) {
        cudaError_t e = ensure_lut_tables(ctx, 8);
        if (e != cudaSuccess) return e;
        e = launch_chain_lut_hsv(ctx->stream, fs, n, g, ctx->lut, a, ctx->lut_path, ctx->lut_interp,
                                 &ctx->stats.kernel_launches);
        if (e != cudaErrorNotSupported) return e;
        // Rows that are not 16-byte aligned (the fused kernel only exists for the vector path) or
        // an extension interpolation without its baked table: run the two elements back to back
        // (the very chain the fused kernel equals).
        cudaGetLastError();
        e = launch_colorlut(ctx->stream, fs, n, g, 8, false, ctx->lut, kMathFast,
                            ctx->lut_path, ctx->lut_interp,
                            &ctx->stats.kernel_launches);
        if (e != cudaSuccess) return e;
        FrameSet inplace = fs;
        for (int i = 0; i < n; i++) inplace.in[i] = fs.out[i];
        Geom g2 = g;
        g2.in_stride = g.out_stride;
        return launch_hsvfilter(ctx->stream, inplace, n, g2, PixLayout{4, 0, 1, 2, 3}, a, kMathFast,
                                &ctx->stats.kernel_launches);
    }
    cudaError_t run(b200vf_ctx *ctx, const FrameSet &fs, int n, const Geom &g) override {
        std::vector<uint8_t> key;  // the chain's function also depends on the LUT and how it is sampled
        key_put(key, (uint8_t)3);
        key_put(key, a);
        key.insert(key.end(), ctx->lut_key.begin(), ctx->lut_key.end());
        key_put(key, ctx->lut_interp);
        return fn_dispatch(ctx, key, /*colour_at_1=*/false, /*keep_other=*/true, 4, 4, fs, n, g,
                           [&](const FrameSet &f, int m, const Geom &gg, bool) {
                               return compute(ctx, f, m, gg);
                           });
    }
};

int check_pairs(b200vf_ctx *ctx, const b200vf_frame *in, const b200vf_frame *out, size_t n,
                const char *who) {
    if (n && (!in || !out)) return fail(ctx, B200VF_ERR_INVALID_ARG, std::string(who) + ": NULL frame array");
    for (size_t i = 0; i < n; i++) {
        int rc = check_frame(ctx, &in[i], who);
        if (rc) return rc;
        if ((rc = check_frame(ctx, &out[i], who))) return rc;
        if (in[i].width != out[i].width || in[i].height != out[i].height)
            return fail(ctx, B200VF_ERR_INVALID_ARG, std::string(who) + ": in/out size mismatch");
    }
    return B200VF_OK;
}
}  // namespace

int b200vf_colorlut_process_batch(b200vf_ctx *ctx, const b200vf_frame *in, const b200vf_frame *out,
                                  size_t n_frames) {
    int rc = activate(ctx);
    if (rc) return rc;
    if ((rc = check_pairs(ctx, in, out, n_frames, "colorlut"))) return rc;
    if (ctx->lut.kind == 0) return fail(ctx, B200VF_ERR_NO_LUT, "No LUT configured");  // imp.rs:210-213
    if (n_frames == 0) return B200VF_OK;
    const uint32_t fmt = in[0].format;
    for (size_t i = 0; i < n_frames; i++) {
        if (!colorlut_accepts(in[i].format) || in[i].format != out[i].format)
            return fail(ctx, B200VF_ERR_UNSUPPORTED_FORMAT,
                        "colorlut: format must be RGBA64_LE, RGBA64_BE or RGBA on both pads");
        if (in[i].format != fmt)
            return fail(ctx, B200VF_ERR_INVALID_ARG, "colorlut: one batch, one format");
    }
    ColorLutLauncher L;
    L.bits = fmt == B200VF_FORMAT_RGBA ? 8 : 16;
    L.be = fmt == B200VF_FORMAT_RGBA64_BE;
    const int bpp = kFormats[fmt].bpp;
    return run_frames(ctx, in, out, n_frames, bpp, bpp, L);
}

int b200vf_colorlut_process(b200vf_ctx *ctx, const b200vf_frame *in, const b200vf_frame *out) {
    return b200vf_colorlut_process_batch(ctx, in, out, 1);
}

int b200vf_colorlut_convert_process_batch(b200vf_ctx *ctx, const b200vf_frame *in, const b200vf_frame *out,
                                          size_t n_frames) {
    int rc = activate(ctx);
    if (rc) return rc;
    if ((rc = check_pairs(ctx, in, out, n_frames, "colorlut_convert"))) return rc;
    if (ctx->lut.kind == 0) return fail(ctx, B200VF_ERR_NO_LUT, "No LUT configured");
    if (n_frames == 0) return B200VF_OK;
    const uint32_t fi = in[0].format, fo = out[0].format;
    for (size_t i = 0; i < n_frames; i++) {
        if (in[i].format > B200VF_FORMAT_BGR || out[i].format > B200VF_FORMAT_BGR)
            return fail(ctx, B200VF_ERR_UNSUPPORTED_FORMAT,
                        "colorlut_convert: 8-bit packed formats only (RGBA RGBx xRGB ARGB BGRx BGRA xBGR ABGR RGB BGR)");
        if (in[i].format != fi || out[i].format != fo)
            return fail(ctx, B200VF_ERR_INVALID_ARG, "colorlut_convert: one batch, one format pair");
    }
    if (fi == B200VF_FORMAT_RGBA && fo == B200VF_FORMAT_RGBA)  // nothing to convert: the element's own path
        return b200vf_colorlut_process_batch(ctx, in, out, n_frames);
    ColorLutConvertLauncher L;
    L.in_lay = layout_of(fi);
    L.out_lay = layout_of(fo);
    // the padding byte of an x format carries no alpha: the result gets the constant 255
    if (fi == B200VF_FORMAT_RGBX || fi == B200VF_FORMAT_XRGB || fi == B200VF_FORMAT_BGRX ||
        fi == B200VF_FORMAT_XBGR)
        L.in_lay.a = -1;
    rc = run_frames(ctx, in, out, n_frames, L.in_lay.bpp, L.out_lay.bpp, L);
    if (rc == B200VF_ERR_CUDA && !ctx->lut.lut3d_baked)
        return fail(ctx, B200VF_ERR_NOMEM, "colorlut_convert: no memory for the baked LUT table (64 MiB)");
    return rc;
}

// ============================================================================
// hsvfilter
// ============================================================================
int b200vf_hsvfilter_process_batch(b200vf_ctx *ctx, const b200vf_frame *frames, size_t n_frames,
                                   const b200vf_hsvfilter_params *params) {
    int rc = activate(ctx);
    if (rc) return rc;
    if (!params) return fail(ctx, B200VF_ERR_INVALID_ARG, "hsvfilter: params is NULL");
    if ((rc = check_pairs(ctx, frames, frames, n_frames, "hsvfilter"))) return rc;
    if (n_frames == 0) return B200VF_OK;
    const uint32_t fmt = frames[0].format;
    for (size_t i = 0; i < n_frames; i++) {
        if (!hsvfilter_accepts(frames[i].format))
            return fail(ctx, B200VF_ERR_UNSUPPORTED_FORMAT, "hsvfilter: format not in caps");
        if (frames[i].format != fmt)
            return fail(ctx, B200VF_ERR_INVALID_ARG, "hsvfilter: one batch, one format");
    }
    HsvFilterLauncher L;
    L.lay = layout_of(fmt);
    L.a = HsvFilterArgs{params->hue_shift, params->saturation_mul, params->saturation_off,
                        params->value_mul, params->value_off};
    return run_frames(ctx, frames, frames, n_frames, L.lay.bpp, L.lay.bpp, L);
}

int b200vf_hsvfilter_process(b200vf_ctx *ctx, const b200vf_frame *frame,
                             const b200vf_hsvfilter_params *params) {
    return b200vf_hsvfilter_process_batch(ctx, frame, 1, params);
}

// ============================================================================
// hsvdetector
// ============================================================================
int b200vf_hsvdetector_process_batch(b200vf_ctx *ctx, const b200vf_frame *in,
                                     const b200vf_frame *out, size_t n_frames,
                                     const b200vf_hsvdetector_params *params) {
    int rc = activate(ctx);
    if (rc) return rc;
    if (!params) return fail(ctx, B200VF_ERR_INVALID_ARG, "hsvdetector: params is NULL");
    if ((rc = check_pairs(ctx, in, out, n_frames, "hsvdetector"))) return rc;
    if (n_frames == 0) return B200VF_OK;
    const uint32_t fi = in[0].format, fo = out[0].format;
    for (size_t i = 0; i < n_frames; i++) {
        if (!hsvdetector_accepts_in(in[i].format) || !hsvdetector_accepts_out(out[i].format))
            return fail(ctx, B200VF_ERR_UNSUPPORTED_FORMAT, "hsvdetector: format not in caps");
        if (in[i].format != fi || out[i].format != fo)
            return fail(ctx, B200VF_ERR_INVALID_ARG, "hsvdetector: one batch, one format pair");
    }
    HsvDetectLauncher L;
    L.in_lay = layout_of(fi);
    L.out_lay = layout_of(fo);
    L.a = HsvDetectArgs{params->hue_ref,        params->hue_var,   params->saturation_ref,
                        params->saturation_var, params->value_ref, params->value_var};
    return run_frames(ctx, in, out, n_frames, L.in_lay.bpp, L.out_lay.bpp, L);
}

int b200vf_hsvdetector_process(b200vf_ctx *ctx, const b200vf_frame *in, const b200vf_frame *out,
                               const b200vf_hsvdetector_params *params) {
    return b200vf_hsvdetector_process_batch(ctx, in, out, 1, params);
}

// ============================================================================
// diagnostics
// ============================================================================
int b200vf_debug_table_indices(const uint32_t *colours, size_t n, uint32_t *out) {
    if (n && (!colours || !out)) return fail(nullptr, B200VF_ERR_INVALID_ARG, "debug_table_indices: NULL pointer");
    table_indices(colours, n, out);
    return B200VF_OK;
}

int b200vf_debug_hsv_from_rgb(b200vf_ctx *ctx, const void *rgba_device, size_t n_pixels,
                              float *hsv_device) {
    int rc = activate(ctx);
    if (rc) return rc;
    if (n_pixels && (!rgba_device || !hsv_device))
        return fail(ctx, B200VF_ERR_INVALID_ARG, "debug_hsv_from_rgb: NULL pointer");
    cudaError_t e = launch_debug_from_rgb(ctx->stream, (const uint32_t *)rgba_device, hsv_device,
                                          n_pixels, ctx->math_mode == kMathPlain,
                                          &ctx->stats.kernel_launches);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "debug_hsv_from_rgb");
    return B200VF_OK;
}

// ============================================================================
// colorlut ! hsvfilter
// ============================================================================
int b200vf_chain_lut_hsv_process_batch(b200vf_ctx *ctx, const b200vf_frame *in,
                                       const b200vf_frame *out, size_t n_frames,
                                       const b200vf_hsvfilter_params *params) {
    int rc = activate(ctx);
    if (rc) return rc;
    if (!params) return fail(ctx, B200VF_ERR_INVALID_ARG, "chain: params is NULL");
    if ((rc = check_pairs(ctx, in, out, n_frames, "chain"))) return rc;
    if (ctx->lut.kind == 0) return fail(ctx, B200VF_ERR_NO_LUT, "No LUT configured");
    if (ctx->lut.kind != 3)
        return fail(ctx, B200VF_ERR_UNSUPPORTED_FORMAT, "chain: fused path needs a 3D LUT");
    for (size_t i = 0; i < n_frames; i++)
        if (in[i].format != B200VF_FORMAT_RGBA || out[i].format != B200VF_FORMAT_RGBA)
            return fail(ctx, B200VF_ERR_UNSUPPORTED_FORMAT, "chain: RGBA only");
    ChainLauncher L;
    L.a = HsvFilterArgs{params->hue_shift, params->saturation_mul, params->saturation_off,
                        params->value_mul, params->value_off};
    return run_frames(ctx, in, out, n_frames, 4, 4, L);
}

}  // extern "C"
