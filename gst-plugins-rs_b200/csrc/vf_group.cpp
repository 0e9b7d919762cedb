// vf_group.cpp — in-process frame-parallel dispatcher: one context, one host
// thread and three streams per device; frame i of a batch goes to member i mod G; no collective
// and no data exchanged between devices.  Built on the per-context entry points of vf_abi.cpp —
// a group is N ordinary contexts plus the plumbing to feed them from one caller.
#include <condition_variable>
#include <deque>
#include <functional>
#include <mutex>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/b200vf.h"
#include "vf_internal.h"

namespace {

struct Member {
    b200vf_ctx *ctx = nullptr;
    int device = 0;
    std::thread thread;
    std::mutex mu;
    std::condition_variable cv;
    std::deque<std::function<void()>> tasks;
    bool quit = false;

    void loop() {
        for (;;) {
            std::function<void()> task;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return quit || !tasks.empty(); });
                if (tasks.empty()) return;  // quit and drained
                task = std::move(tasks.front());
                tasks.pop_front();
            }
            task();
        }
    }
    void post(std::function<void()> f) {
        {
            std::lock_guard<std::mutex> g(mu);
            tasks.push_back(std::move(f));
        }
        cv.notify_one();
    }
};

// completion latch of one fan-out
struct Latch {
    std::mutex mu;
    std::condition_variable cv;
    size_t left;
    explicit Latch(size_t n) : left(n) {}
    void done() {
        std::lock_guard<std::mutex> g(mu);
        if (--left == 0) cv.notify_all();
    }
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&] { return left == 0; });
    }
};

}  // namespace

struct b200vf_group {
    std::vector<Member *> members;
    std::string last_error;
};

namespace {

int gfail(b200vf_group *g, int code, const std::string &msg) {
    if (g)
        g->last_error = msg;
    else
        vf::fail_global(code, msg);
    return code;
}

// Runs fn(member index, ctx) on every member's thread, waits for all, returns the first failure.
int fan_out(b200vf_group *g, const std::function<int(size_t, b200vf_ctx *)> &fn) {
    const size_t n = g->members.size();
    std::vector<int> rc(n, B200VF_OK);
    Latch latch(n);
    size_t posted = 0;
    bool post_failed = false;
    for (size_t m = 0; m < n && !post_failed; m++) {
        Member *mem = g->members[m];
        try {
            mem->post([&, m, mem] {
                try {
                    rc[m] = fn(m, mem->ctx);
                } catch (...) {  // nothing may unwind out of a worker or through the C ABI
                    rc[m] = B200VF_ERR_NOMEM;
                }
                latch.done();
            });
            posted++;
        } catch (...) {  // could not queue the task: the ones already queued still reference this frame
            post_failed = true;
        }
    }
    for (size_t m = posted; m < n; m++) latch.done();
    latch.wait();
    if (post_failed) return gfail(g, B200VF_ERR_NOMEM, "group: could not queue the members' work");
    for (size_t m = 0; m < n; m++)
        if (rc[m] != B200VF_OK)
            return gfail(g, rc[m],
                         "member " + std::to_string(m) + " (device " + std::to_string(g->members[m]->device) +
                             "): " + b200vf_last_error(g->members[m]->ctx));
    return B200VF_OK;
}

// Frames i with i mod G == m, in order.
std::vector<b200vf_frame> share_of(const b200vf_frame *f, size_t n, size_t m, size_t G) {
    std::vector<b200vf_frame> v;
    v.reserve(n / G + 1);
    for (size_t i = m; i < n; i += G) v.push_back(f[i]);
    return v;
}

// Device-memory frames must live on the device of the member that will process them.
int check_residency(b200vf_ctx *ctx, const std::vector<b200vf_frame> &v) {
    for (const b200vf_frame &f : v) {
        if (f.memory != B200VF_MEM_DEVICE || !f.data) continue;
        uint32_t mem = 0;
        int dev = -1;
        if (b200vf_pointer_info(f.data, &mem, &dev) != B200VF_OK) continue;
        if (mem == B200VF_MEM_DEVICE && dev != b200vf_ctx_device(ctx)) return dev;
    }
    return -1;
}

using ProcessFn = std::function<int(b200vf_ctx *, const b200vf_frame *, const b200vf_frame *, size_t)>;

int process(b200vf_group *g, const b200vf_frame *in, const b200vf_frame *out, size_t n, const char *who,
            const ProcessFn &fn) {
    if (!g) return gfail(nullptr, B200VF_ERR_INVALID_ARG, std::string(who) + ": group is NULL");
    if (n && (!in || !out)) return gfail(g, B200VF_ERR_INVALID_ARG, std::string(who) + ": NULL frame array");
    const size_t G = g->members.size();
    try {
        return fan_out(g, [&](size_t m, b200vf_ctx *ctx) -> int {
            const std::vector<b200vf_frame> a = share_of(in, n, m, G);
            if (a.empty()) return B200VF_OK;
            const std::vector<b200vf_frame> b = in == out ? std::vector<b200vf_frame>() : share_of(out, n, m, G);
            int wrong = check_residency(ctx, a);
            if (wrong < 0 && !b.empty()) wrong = check_residency(ctx, b);
            if (wrong >= 0) {
                // recorded on the member's context: fan_out adds member and device to the message
                return vf::ctx_fail(ctx, B200VF_ERR_INVALID_ARG,
                                    std::string(who) + ": a frame of this member's share lives on device " +
                                        std::to_string(wrong) + " (frame i is processed on member i mod G)");
            }
            return fn(ctx, a.data(), b.empty() ? a.data() : b.data(), a.size());
        });
    } catch (const std::bad_alloc &) {
        return gfail(g, B200VF_ERR_NOMEM, std::string(who) + ": host allocation failed");
    }
}

}  // namespace

extern "C" {

int b200vf_group_create(const int *devices, size_t n_devices, b200vf_group **out) {
    if (!out) return gfail(nullptr, B200VF_ERR_INVALID_ARG, "group_create: out is NULL");
    *out = nullptr;
    if (!devices || n_devices == 0 || n_devices > 64)
        return gfail(nullptr, B200VF_ERR_INVALID_ARG, "group_create: need 1..64 devices");
    b200vf_group *g = new (std::nothrow) b200vf_group();
    if (!g) return gfail(nullptr, B200VF_ERR_NOMEM, "group_create: allocation failed");
    try {
        for (size_t i = 0; i < n_devices; i++) {
            Member *m = new Member();
            g->members.push_back(m);
            m->device = devices[i];
            int rc = b200vf_ctx_create(devices[i], &m->ctx);
            if (rc != B200VF_OK) {
                const std::string msg = b200vf_last_error(nullptr);
                b200vf_group_destroy(g);
                return gfail(nullptr, rc, "group_create: device " + std::to_string(devices[i]) + ": " + msg);
            }
            m->thread = std::thread([m] { m->loop(); });
        }
    } catch (...) {
        b200vf_group_destroy(g);
        return gfail(nullptr, B200VF_ERR_NOMEM, "group_create: could not start the member threads");
    }
    *out = g;
    return B200VF_OK;
}

void b200vf_group_destroy(b200vf_group *g) {
    if (!g) return;
    for (Member *m : g->members) {
        if (m->thread.joinable()) {
            {
                std::lock_guard<std::mutex> lk(m->mu);
                m->quit = true;
            }
            m->cv.notify_all();
            m->thread.join();
        }
        if (m->ctx) b200vf_ctx_destroy(m->ctx);
        delete m;
    }
    delete g;
}

size_t b200vf_group_size(const b200vf_group *g) { return g ? g->members.size() : 0; }

b200vf_ctx *b200vf_group_ctx(b200vf_group *g, size_t member) {
    return g && member < g->members.size() ? g->members[member]->ctx : nullptr;
}

const char *b200vf_group_last_error(const b200vf_group *g) {
    return g ? g->last_error.c_str() : b200vf_last_error(nullptr);
}

int b200vf_group_set_option(b200vf_group *g, const char *key, int64_t value) {
    if (!g) return gfail(nullptr, B200VF_ERR_INVALID_ARG, "group_set_option: group is NULL");
    return fan_out(g, [&](size_t, b200vf_ctx *ctx) { return b200vf_ctx_set_option(ctx, key, value); });
}

int b200vf_group_synchronize(b200vf_group *g) {
    if (!g) return gfail(nullptr, B200VF_ERR_INVALID_ARG, "group_synchronize: group is NULL");
    return fan_out(g, [](size_t, b200vf_ctx *ctx) { return b200vf_ctx_synchronize(ctx); });
}

int b200vf_group_colorlut_set_lut(b200vf_group *g, uint32_t kind, uint32_t size, const float *data,
                                  const float domain_scale[3], const float domain_offset[3]) {
    if (!g) return gfail(nullptr, B200VF_ERR_INVALID_ARG, "group_set_lut: group is NULL");
    return fan_out(g, [&](size_t, b200vf_ctx *ctx) {
        return b200vf_colorlut_set_lut(ctx, kind, size, data, domain_scale, domain_offset);
    });
}

int b200vf_group_colorlut_set_lut_file(b200vf_group *g, const char *location) {
    if (!g) return gfail(nullptr, B200VF_ERR_INVALID_ARG, "group_set_lut_file: group is NULL");
    if (!location) return gfail(g, B200VF_ERR_SETTINGS, "LUT file location is not configured");
    b200vf_cube cube;  // parse once, upload to every device
    char err[512] = {0};
    int rc = b200vf_cube_parse_file(location, &cube, err, sizeof err);
    if (rc != B200VF_OK)
        return gfail(g, rc, std::string("Failed to parse LUT file ") + location + ": " + err);
    rc = b200vf_group_colorlut_set_lut(g, cube.kind, cube.size, cube.data, cube.domain_scale, cube.domain_offset);
    b200vf_cube_free(&cube);
    return rc;
}

int b200vf_group_colorlut_clear_lut(b200vf_group *g) {
    if (!g) return gfail(nullptr, B200VF_ERR_INVALID_ARG, "group_clear_lut: group is NULL");
    return fan_out(g, [](size_t, b200vf_ctx *ctx) { return b200vf_colorlut_clear_lut(ctx); });
}

int b200vf_group_colorlut_process_batch(b200vf_group *g, const b200vf_frame *in, const b200vf_frame *out,
                                        size_t n_frames) {
    return process(g, in, out, n_frames, "group colorlut",
                   [](b200vf_ctx *c, const b200vf_frame *a, const b200vf_frame *b, size_t n) {
                       return b200vf_colorlut_process_batch(c, a, b, n);
                   });
}

int b200vf_group_hsvfilter_process_batch(b200vf_group *g, const b200vf_frame *frames, size_t n_frames,
                                         const b200vf_hsvfilter_params *params) {
    if (g && !params) return gfail(g, B200VF_ERR_INVALID_ARG, "group hsvfilter: params is NULL");
    return process(g, frames, frames, n_frames, "group hsvfilter",
                   [params](b200vf_ctx *c, const b200vf_frame *a, const b200vf_frame *, size_t n) {
                       return b200vf_hsvfilter_process_batch(c, a, n, params);
                   });
}

int b200vf_group_hsvdetector_process_batch(b200vf_group *g, const b200vf_frame *in, const b200vf_frame *out,
                                           size_t n_frames, const b200vf_hsvdetector_params *params) {
    if (g && !params) return gfail(g, B200VF_ERR_INVALID_ARG, "group hsvdetector: params is NULL");
    return process(g, in, out, n_frames, "group hsvdetector",
                   [params](b200vf_ctx *c, const b200vf_frame *a, const b200vf_frame *b, size_t n) {
                       return b200vf_hsvdetector_process_batch(c, a, b, n, params);
                   });
}

int b200vf_group_chain_lut_hsv_process_batch(b200vf_group *g, const b200vf_frame *in, const b200vf_frame *out,
                                             size_t n_frames, const b200vf_hsvfilter_params *params) {
    if (g && !params) return gfail(g, B200VF_ERR_INVALID_ARG, "group chain: params is NULL");
    return process(g, in, out, n_frames, "group chain",
                   [params](b200vf_ctx *c, const b200vf_frame *a, const b200vf_frame *b, size_t n) {
                       return b200vf_chain_lut_hsv_process_batch(c, a, b, n, params);
                   });
}

}  // extern "C"
