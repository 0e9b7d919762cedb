// vf_abi.cpp — implementation of include/b200vf.h: contexts, options, LUT upload and derived
// tables, frame validation and dispatch to the kernel launchers.  System-memory frames go through
// vf_host.cpp; the context itself is in vf_ctx.h.  No exception leaves this file; there is no CPU
// fallback.
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
#include "vf_ctx.h"

using namespace vf;

namespace {

thread_local std::string g_last_error;  // for calls that have no context

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
int vf::cuda_fail(b200vf_ctx *ctx, cudaError_t e, const char *what) {
    return fail(ctx, B200VF_ERR_CUDA,
                std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}

int vf::activate(b200vf_ctx *ctx) {
    if (!ctx) return fail(nullptr, B200VF_ERR_INVALID_ARG, "context is NULL");
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess || cur != ctx->device)
        VF_CUDA(ctx, cudaSetDevice(ctx->device));
    return B200VF_OK;
}

bool vf::is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

namespace {

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

bool same_geometry(const b200vf_frame &a, const b200vf_frame &b) {
    return a.stride == b.stride && a.width == b.width && a.height == b.height &&
           a.format == b.format && a.memory == b.memory;
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
