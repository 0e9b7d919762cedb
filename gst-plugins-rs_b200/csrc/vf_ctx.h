// vf_ctx.h — the context behind the C ABI (struct b200vf_ctx) and what its two implementation files
// share: vf_abi.cpp (entry points, LUT and table management, launchers) and vf_host.cpp (the
// system-memory frame pipeline).  Internal; nothing here is exported.
#pragma once
#include <algorithm>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/b200vf.h"
#include "vf_internal.h"

namespace vf {

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

}  // namespace vf

struct b200vf_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;  // compute stream for device frames and kernels
    bool own_stream = false;
    cudaEvent_t ev_order = nullptr;  // b200vf_ctx_wait_for
    cudaStream_t s_in = nullptr, s_out = nullptr;  // copy streams of the host path
    vf::Slot slots[vf::kMaxSlots];
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
    cudaEvent_t ticket_ev[vf::kTickets] = {};  // ticket t: recorded on s_out behind the call's last D2H
    uint64_t ticket_of[vf::kTickets] = {};
    vf::DeviceLut lut;
    std::string last_error;
    b200vf_stats stats{};
    int math_mode = vf::kMathFast;
    int lut_path = vf::kLutAuto;
    int lut_interp = vf::kInterpTrilinear;
    int64_t chunk_bytes = 0;  // "host.chunk_bytes"; 0 = auto (a sixth of the call's bytes, within 4..17 MiB)
    // "host.copy_threads": threads for pageable-frame row copies; half the cores, within 2..8
    // (one core's memcpy is ~10 GB/s; a 4K pageable stream goes 410 -> 590 frames/s from 4 to 8)
    int copy_threads = (int)std::min(8u, std::max(2u, std::thread::hardware_concurrency() / 2));
    vf::CopyPool *pool = nullptr;
    int lut_path_active = -1;     // resolved path of the last colorlut launch ("lut.path_active")
    std::vector<uint8_t> lut_key; // identifies the LUT's content (hash, size, domain): key of its tables
    vf::SharedTable *baked = nullptr; // the LUT baked to 8-bit resolution, from the device-wide cache
    bool baked_failed = false;    // no memory for it: the interpolating kernels serve
    vf::PathPolicy lut_policy;        // auto: baked table vs direct interpolation, measured
    vf::PathPolicy lut64_policy;      // auto, 16-bit frames: delta-table op vs direct kernel, measured
    bool share_tables = true;     // "tables.share": 0 = private tables (keys salted with the context)
    int fn_path = vf::kFnAuto;        // "hsv.path"
    vf::FnTable fn;                   // tabulated hsvfilter / hsvdetector / chain function
};

namespace vf {

int cuda_fail(b200vf_ctx *ctx, cudaError_t e, const char *what);  // records the message, returns B200VF_ERR_CUDA

#define VF_CUDA(ctx, call)                                          \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return vf::cuda_fail(ctx, e__, #call); \
    } while (0)

int activate(b200vf_ctx *ctx);   // makes the context's device current
bool is_pinned(const void *p);   // page-locked host memory (cudaMallocHost'ed or registered)?

// One element's kernel launch over frames already in device memory.
struct Launcher {
    virtual cudaError_t run(b200vf_ctx *ctx, const FrameSet &fs, int n, const Geom &g) = 0;
    virtual ~Launcher() = default;
};

// ---- vf_host.cpp: system-memory frames ----------------------------------------------------------
// Cuts the frames into row chunks and runs them through the ring of stage slots (H2D || kernel ||
// D2H); complete on return unless "host.async" applies.  A failed call leaves nothing in flight.
int run_host(b200vf_ctx *ctx, const b200vf_frame *in, const b200vf_frame *out, size_t n_frames, int in_bpp,
             int out_bpp, Launcher &L);
int host_wait(b200vf_ctx *ctx, uint64_t ticket);  // "host.async": completes the call with this ticket
void quiesce_async(b200vf_ctx *ctx);              // nothing of an earlier call is in flight afterwards
void unregister_all(b200vf_ctx *ctx);             // "host.register": drops every page-locked range
void forget_range(b200vf_ctx *ctx, const void *p, size_t bytes);  // drops the ranges overlapping [p, p + bytes)

}  // namespace vf
