// vf_host.cpp — system-memory frames: the pinned H2D -> kernel -> D2H stream pipeline over a ring of
// stage slots, the pageable bounce path with its copy threads, in-place page-locking of recurring
// buffers ("host.register") and calls in flight ("host.async").  No exception leaves this file.
#include <cstdio>
#include <new>

#include "vf_ctx.h"

namespace vf {
namespace {

int fail(b200vf_ctx *ctx, int code, const std::string &msg) { return ctx_fail(ctx, code, msg); }

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

}  // namespace

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

namespace {

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

}  // namespace

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

}  // namespace vf
