/*
 * b200vf.h — C ABI of libb200vf.so: the B200-native (sm_100a) per-pixel colour
 * transform path of gst-plugins-rs — `colorlut`, `hsvfilter`, `hsvdetector`.
 *
 * This is the drop-in boundary.  A reference-side element keeps
 * its GObject shell (name, properties, caps, start/stop) and replaces only the
 * body of transform_frame / transform_frame_ip with one call below.  Citations
 * are relative to the reference tree (gst-plugins-rs 0.16.0-alpha):
 *
 *   b200vf_ctx_create / _destroy      ↔ BaseTransformImpl::start / stop
 *                                        video/colorlut/src/colorlut/imp.rs:168-199
 *                                        (context model: d3d12colorlut/imp.rs:299-342)
 *   b200vf_cube_parse[_file]          ↔ CubeLut::parse / parse_file
 *                                        video/colorlut/src/parser.rs:104-281
 *   b200vf_colorlut_set_lut[_file]    ↔ `*self.state.lock() = State { lut: Some(lut) }`
 *                                        video/colorlut/src/colorlut/imp.rs:182-191
 *   b200vf_colorlut_process           ↔ ColorLut::transform_frame
 *                                        video/colorlut/src/colorlut/imp.rs:203-223
 *   b200vf_hsvfilter_process          ↔ HsvFilter::transform_frame_ip
 *                                        video/hsv/src/hsvfilter/imp.rs:323-376 (+ 76-120)
 *   b200vf_hsvdetector_process        ↔ HsvDetector::transform_frame
 *                                        video/hsv/src/hsvdetector/imp.rs:423-707 (+ 100-160)
 *   *_process_batch                   — launch amortisation for many small frames;
 *                                        no reference counterpart (one buffer per call there)
 *   b200vf_pool_*, b200vf_pointer_info ↔ buffer pools / device follow of the reference's GPU
 *                                        sibling, d3d12colorlut/imp.rs:385-542
 *
 * Conventions
 *   - Every entry point returns an int status: 0 = B200VF_OK, negative = error.
 *     No C++ exception or abort crosses this boundary.  `b200vf_last_error(ctx)`
 *     returns a human-readable message for the last failing call on that context.
 *   - A context is single-caller (one GStreamer streaming thread per element
 *     instance); different contexts may be used concurrently from different threads.
 *   - Frames are BORROWED for the duration of the call; pointers are never retained.
 *     Only the first width*bytes_per_pixel bytes of each of `height` rows are read
 *     or written; row padding is never touched (reference: colorlut/imp.rs:281-286,
 *     hsvfilter/imp.rs:94-97, hsvdetector/imp.rs:124-137).
 *   - memory == B200VF_MEM_HOST: the call stages the frame through pinned buffers
 *     (H2D → kernel → D2H stream pipeline) and is COMPLETE when it returns.
 *     memory == B200VF_MEM_DEVICE: the kernel is enqueued on the context's stream
 *     and the call returns immediately (stream-ordered hand-off, the CUDA analogue
 *     of the fence in d3d12colorlut/imp.rs:711-714); use b200vf_ctx_synchronize()
 *     or the stream handle to order later work.
 *   - Parameter structs are passed by value per call, so a property changed while
 *     PLAYING takes effect on the next frame (snapshot semantics of
 *     hsvfilter/imp.rs:85).
 *   - There is NO CPU fallback.  Without a usable CUDA device, ctx_create fails.
 */
#ifndef B200VF_H
#define B200VF_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define B200VF_API __attribute__((visibility("default")))
#else
#define B200VF_API
#endif

#define B200VF_VERSION_MAJOR 0
#define B200VF_VERSION_MINOR 1

/* ---- status codes -------------------------------------------------------- */
typedef enum b200vf_status {
    B200VF_OK = 0,
    B200VF_ERR_INVALID_ARG = -1,        /* NULL pointer, zero size, bad enum, stride < row bytes */
    B200VF_ERR_UNSUPPORTED_FORMAT = -2, /* format not in the element's caps (reference: unreachable!()) */
    B200VF_ERR_CUDA = -3,               /* a CUDA runtime call failed; context must be recreated */
    B200VF_ERR_NO_LUT = -4,             /* colorlut_process before set_lut (imp.rs:210-213) */
    B200VF_ERR_PARSE = -5,              /* CubeParseError::InvalidLut (parser.rs:76-80) */
    B200VF_ERR_IO = -6,                 /* CubeParseError::Io (parser.rs:76-80) */
    B200VF_ERR_NO_DEVICE = -7,          /* no CUDA device / device index out of range */
    B200VF_ERR_NOMEM = -8,              /* host or device allocation failed */
    B200VF_ERR_SETTINGS = -9            /* `location` not configured (imp.rs:175-180) */
} b200vf_status;

/* ---- pixel formats (GstVideoFormat names) -------------------------------- */
typedef enum b200vf_format {
    B200VF_FORMAT_RGBA = 0,
    B200VF_FORMAT_RGBX = 1, /* "RGBx" */
    B200VF_FORMAT_XRGB = 2, /* "xRGB" */
    B200VF_FORMAT_ARGB = 3,
    B200VF_FORMAT_BGRX = 4, /* "BGRx" */
    B200VF_FORMAT_BGRA = 5,
    B200VF_FORMAT_XBGR = 6, /* "xBGR" */
    B200VF_FORMAT_ABGR = 7,
    B200VF_FORMAT_RGB = 8,
    B200VF_FORMAT_BGR = 9,
    B200VF_FORMAT_RGBA64_LE = 10,
    B200VF_FORMAT_RGBA64_BE = 11,
    B200VF_FORMAT_COUNT = 12
} b200vf_format;

typedef enum b200vf_memory {
    B200VF_MEM_HOST = 0,  /* system memory (pageable or pinned); synchronous semantics */
    B200VF_MEM_DEVICE = 1 /* CUDA device memory on the context's device; stream-ordered */
} b200vf_memory;

/* One mapped video plane: what gst_video::VideoFrameRef exposes to the reference
 * loops (plane_data(0), plane_stride()[0], width(), height(), format();
 * colorlut/imp.rs:242-249). */
typedef struct b200vf_frame {
    void *data;      /* first byte of row 0 */
    int64_t stride;  /* bytes between rows; >= width * bytes_per_pixel */
    uint32_t width;  /* pixels */
    uint32_t height; /* rows */
    uint32_t format; /* b200vf_format */
    uint32_t memory; /* b200vf_memory */
} b200vf_frame;

typedef struct b200vf_ctx b200vf_ctx;

/* ---- library ------------------------------------------------------------- */
B200VF_API const char *b200vf_version(void); /* "b200vf 0.1 (sm_100a)" */
B200VF_API const char *b200vf_status_string(int status);
B200VF_API int b200vf_device_count(int *count);
B200VF_API uint32_t b200vf_format_bytes_per_pixel(uint32_t format); /* 0 for unknown */
B200VF_API const char *b200vf_format_name(uint32_t format);         /* GstVideoFormat string */
B200VF_API int b200vf_format_from_name(const char *name);           /* -1 if unknown */

/* ---- context lifecycle (start / stop) ------------------------------------ */
B200VF_API int b200vf_ctx_create(int device, b200vf_ctx **out);
B200VF_API void b200vf_ctx_destroy(b200vf_ctx *ctx);
B200VF_API const char *b200vf_last_error(const b200vf_ctx *ctx);
B200VF_API int b200vf_ctx_device(const b200vf_ctx *ctx);
/* Block until everything enqueued by this context has finished. */
B200VF_API int b200vf_ctx_synchronize(b200vf_ctx *ctx);
/* The cudaStream_t (as void*) device-memory work is enqueued on.  By default a
 * context-owned non-blocking stream; set_stream lets the caller supply its own
 * (e.g. the stream of an upstream CUDA element).  NULL = legacy default stream. */
B200VF_API void *b200vf_ctx_get_stream(const b200vf_ctx *ctx);
/* The supplied stream stays the CALLER's: it must outlive the context (or a later set_stream), and
 * it must not be another context's own stream — that one dies with its context (ctx_destroy, or the
 * stop/start of an element that follows its buffers to another device).  To order two contexts use
 * b200vf_ctx_wait_for instead of sharing a stream. */
B200VF_API int b200vf_ctx_set_stream(b200vf_ctx *ctx, void *cuda_stream);
/* Stream-orders `ctx` after `upstream` (same device): work enqueued on ctx from now on starts
 * only when everything enqueued on upstream so far has finished.  Nothing blocks on the host, each
 * context keeps its own stream, no handle is shared.  This is how chained device-memory elements
 * hand a frame over (upstream produced it, ctx consumes it): call it before ctx's *_process. */
B200VF_API int b200vf_ctx_wait_for(b200vf_ctx *ctx, b200vf_ctx *upstream);

/* Tunables / diagnostics.  Unknown key → INVALID_ARG.
 *   "hsv.math"      0 = fast exact sequences (default), 1 = plain IEEE `/` + fmodf translation
 *   "lut.path"      0 = auto, 1 = direct 8-corner trilinear, 2 = R-resampled table,
 *                   3 = R- and G-resampled table, 4 = table baked to native 8-bit resolution.
 *                   All bit-identical; 2-4 are 8-bit RGBA only.  Auto for 8-bit frames is 4: the
 *                   direct kernel evaluates the LUT once for all 2^24 byte triples (64 MiB,
 *                   L2-resident on B200, stored in blocks of 4x4x2 neighbouring colours per cache
 *                   line, built on the first 8-bit frame after set_lut), frames then need one 4-byte
 *                   gather per pixel; 3 serves if that allocation fails.  RGBA64 runs 1, or (auto /
 *                   4; 3D LUT of size <= 128 with the default domain) a variant of it whose table
 *                   stores each corner with its x-difference ("lut.path_active" = 7).  Under auto
 *                   the table path and the direct kernel are both timed on the stream's own frames
 *                   (calls of >= 2^20 pixels, device frames) and the faster one serves, exactly as
 *                   for "hsv.path".  "lut.path_active" (read-only) = the kernel of the last call.
 *   "hsv.path"      hsvfilter / hsvdetector / chain: 0 = auto, 1 = always the compute kernels (the
 *                   reference's f32 sequence per pixel), 2 = always the function table.  The table
 *                   holds the element's result for all 2^24 colour triples under the current
 *                   settings, filled by those same compute kernels (64 MiB, rebuilt when a setting
 *                   changes), so both ways are bit-identical.  Auto serves frames from the compute
 *                   kernels until the settings have been stable for 2^25 pixels, then builds the
 *                   table and keeps whichever way measures faster on the stream's own frames
 *                   (gathers depend on content, the compute kernels do not): the serving kernel is
 *                   sampled on every 8th launch, the other one after 0.25 s (doubling up to 8 s while it
 *                   keeps losing clearly; a table's rival every 30 s); nothing ever blocks on a
 *                   measurement.
 *                   "hsv.table_active" (read-only) = the last launch used the table.
 *   "lut.interpolation" 3D LUTs: 0 = trilinear (the reference, colorlut/imp.rs:493-526; default),
 *                   1 = tetrahedral, 2 = nearest.  1 and 2 are EXTENSIONS: the reference has no
 *                   such modes (no parity claim against it); they are defined by, and bit-exact
 *                   with, the CPU checker's restatement of the published algorithms.
 *                   For 8-bit RGBA they run from a table baked to native resolution (as "lut.path" = 4,
 *                   built on first use); "lut.path" = 1 forces the direct kernel (4 / 1 fetches per
 *                   pixel), which RGBA64 always uses.
 *   "tables.share"  1 (default) = the 64 MiB function tables (baked LUT, hsvfilter / hsvdetector /
 *                   chain tables) come from a device-wide cache keyed by what they compute (LUT
 *                   content + interpolation, or element + layout + every setting bit): contexts
 *                   that need the same function share one table.  0 = private tables for this
 *                   context.  Read-only: "tables.device_count" / "tables.device_bytes" (cached
 *                   tables on this context's device), "lut.tables_built" (bit mask 1 R-resampled,
 *                   2 RG-resampled, 4 baked — derived LUT tables are built by the first launch that
 *                   needs them, not by set_lut).
 *   "host.chunk_bytes"  chunk size of the host-frame stream pipeline; 0 (default) = a sixth of the
 *                   call's bytes within 4..17 MiB (half 4K frames for batches, ~5 MiB pieces for a
 *                   single frame: every chunk costs ~30 us of engine hand-over, the first H2D and
 *                   the last D2H of a call overlap with nothing; with "host.async" = 1: 34 MiB,
 *                   i.e. whole 4K frames — consecutive calls overlap, only the per-chunk cost is left)
 *   "host.copy_threads" threads used for row copies of pageable frames (default: half the cores,
 *                   within 2..8; 1 = caller only)
 *   "host.slots"    chunks in flight in that pipeline, 2..8 (default 4)
 *   "host.register" 0 (default) / 1: page-lock pageable frames in place (cudaHostRegister) when the
 *                   same (pointer, size) comes by a second time, instead of bouncing every frame
 *                   through pinned staging buffers — upstream buffer pools hand the same few
 *                   buffers round and round.  First sight, failures and ranges beyond
 *                   "host.register_budget" bytes (default 2 GiB, least recently used evicted) take
 *                   the bounce path.  Registrations are dropped by b200vf_ctx_host_memory_released,
 *                   by setting the option to 0 and by ctx_destroy.  ONLY for callers that tell
 *                   the context before such memory is freed (see b200vf_ctx_host_memory_released);
 *                   "host.registered_bytes" (read-only) = bytes currently page-locked this way.
 *   "host.async"    0 (default) / 1: a *_process_batch call whose host frames are all page-locked
 *                   returns as soon as its copies and kernels are queued (see b200vf_ctx_host_wait);
 *                   calls with pageable frames stay synchronous.  For callers that keep a frame of
 *                   latency: the next call's first upload then overlaps this call's last download.
 */
B200VF_API int b200vf_ctx_set_option(b200vf_ctx *ctx, const char *key, int64_t value);
B200VF_API int b200vf_ctx_get_option(const b200vf_ctx *ctx, const char *key, int64_t *value);

/* Host-frame calls in flight ("host.async" = 1).  Every *_process_batch call on system-memory frames
 * gets a ticket (1, 2, 3 … per context); b200vf_ctx_host_ticket returns the most recent call's.
 * The output frames of a call are valid, and its input frames may be reused, once
 * b200vf_ctx_host_wait(ticket) has returned — it blocks until that call and every earlier one is
 * complete (immediately for synchronous calls); b200vf_ctx_synchronize() completes all of them.  At
 * most 15 calls are in flight: the 16th blocks until the oldest is complete.  Frames still in
 * flight belong to the library: do not touch, free or re-submit them (the counterpart in the
 * reference is an element that queues buffers: BaseTransform's submit_input_buffer /
 * generate_output pair instead of transform_frame). */
B200VF_API uint64_t b200vf_ctx_host_ticket(const b200vf_ctx *ctx);
B200VF_API int b200vf_ctx_host_wait(b200vf_ctx *ctx, uint64_t ticket);

typedef struct b200vf_stats {
    uint64_t kernel_launches; /* kernels launched by this context */
    uint64_t frames;          /* frames processed */
    uint64_t h2d_bytes;       /* bytes copied host→device by the host-frame path */
    uint64_t d2h_bytes;       /* bytes copied device→host by the host-frame path */
} b200vf_stats;
B200VF_API int b200vf_ctx_get_stats(const b200vf_ctx *ctx, b200vf_stats *out);
B200VF_API int b200vf_ctx_reset_stats(b200vf_ctx *ctx);

/* ---- pinned host / device frame memory (buffer-pool building blocks) ------ */
B200VF_API int b200vf_host_alloc(size_t bytes, void **out); /* page-locked */
B200VF_API int b200vf_host_free(void *p);
/* 1 if p points into page-locked (cudaMallocHost'ed or registered) host memory, else 0. */
B200VF_API int b200vf_host_is_pinned(const void *p);
/* With "host.register" = 1: tell the context that [p, p + bytes) is about to be freed or reused for
 * something else (bytes = 0: whatever registration contains p), BEFORE that happens — CUDA requires
 * registered memory to be unregistered before it is freed.  The reference-side shim calls this from
 * a destroy notify on the upstream GstMemory (gst_mini_object_weak_ref); see INTEGRATION.md. */
B200VF_API int b200vf_ctx_host_memory_released(b200vf_ctx *ctx, const void *p, size_t bytes);
B200VF_API int b200vf_device_alloc(b200vf_ctx *ctx, size_t bytes, void **out);
B200VF_API int b200vf_device_free(b200vf_ctx *ctx, void *p);
/* Plain copies on the context stream (kind: 0 = H2D, 1 = D2H, 2 = D2D); synchronous for host memory. */
B200VF_API int b200vf_memcpy(b200vf_ctx *ctx, void *dst, const void *src, size_t bytes, int kind);

/* ---- .cube parser (host only; no CUDA needed) ----------------------------- */
typedef enum b200vf_lut_kind { B200VF_LUT_1D = 1, B200VF_LUT_3D = 3 } b200vf_lut_kind;

/* Parsed CubeLut (parser.rs:68-74).  3D: `data` holds size^3 entries of
 * [r,g,b,1.0] (4 floats each) in file order, R fastest (parser.rs:43-53,253-256).
 * 1D: `data` holds three planes r[size], g[size], b[size] (parser.rs:226-236). */
typedef struct b200vf_cube {
    uint32_t kind; /* b200vf_lut_kind */
    uint32_t size;
    float domain_scale[3];
    float domain_offset[3];
    float *data;
    size_t n_floats;
} b200vf_cube;

/* Returns B200VF_OK, B200VF_ERR_PARSE or B200VF_ERR_IO; on error `err` (if not
 * NULL) receives the reference's message text ("Invalid LUT: …" / "IO error: …"). */
B200VF_API int b200vf_cube_parse(const char *text, size_t len, b200vf_cube *out, char *err,
                                 size_t errlen);
B200VF_API int b200vf_cube_parse_file(const char *path, b200vf_cube *out, char *err, size_t errlen);
B200VF_API void b200vf_cube_free(b200vf_cube *cube);

/* ---- colorlut -------------------------------------------------------------- */
/* Upload a parsed LUT (called from `start`).  Layout of `data` as in b200vf_cube. */
B200VF_API int b200vf_colorlut_set_lut(b200vf_ctx *ctx, uint32_t kind, uint32_t size,
                                       const float *data, const float domain_scale[3],
                                       const float domain_offset[3]);
/* parse_file + set_lut in one call: what `start` does with the `location` property.
 * location == NULL → B200VF_ERR_SETTINGS (imp.rs:175-180). */
B200VF_API int b200vf_colorlut_set_lut_file(b200vf_ctx *ctx, const char *location);
B200VF_API int b200vf_colorlut_clear_lut(b200vf_ctx *ctx); /* `stop` */
/* in.format == out.format ∈ {RGBA, RGBA64_LE, RGBA64_BE}; in != out (NeverInPlace,
 * imp.rs:163-164) although in == out also works. */
B200VF_API int b200vf_colorlut_process(b200vf_ctx *ctx, const b200vf_frame *in,
                                       const b200vf_frame *out);
B200VF_API int b200vf_colorlut_process_batch(b200vf_ctx *ctx, const b200vf_frame *in,
                                             const b200vf_frame *out, size_t n_frames);

/* colorlut with the `videoconvert` elements either side of it folded in (the
 * reference's own example pipeline is `videoconvert ! colorlut ! videoconvert`, colorlut/imp.rs:17-19):
 * in.format and out.format are any of the ten 8-bit packed formats, independently; the result is
 * what colorlut gives on the RGBA view of the input, stored in the output layout — one pass, one
 * read and one write of the frame.  Colour bytes follow SURVEY.md Appendix C; the alpha byte of the
 * output is the input's alpha, or 255 when the input has none (x formats, RGB, BGR), and an x byte
 * of the output receives that same value.  EXTENSION: the reference element only takes RGBA /
 * RGBA64; the layout rules are GStreamer core's (not in the reference tree), so the conversion part
 * is parity-unpinned — the colour values are exactly colorlut's.  8 <-> 16-bit widening is not offered.
 * Uses the table baked to 8-bit resolution whatever "lut.path" says (1D LUTs are baked too). */
B200VF_API int b200vf_colorlut_convert_process_batch(b200vf_ctx *ctx, const b200vf_frame *in,
                                                     const b200vf_frame *out, size_t n_frames);

/* ---- hsvfilter ------------------------------------------------------------- */
/* Property snapshot; defaults hsvfilter/imp.rs:25-29 = {0, 1, 0, 1, 0}. */
typedef struct b200vf_hsvfilter_params {
    float hue_shift;      /* "hue-shift"      degrees */
    float saturation_mul; /* "saturation-mul" */
    float saturation_off; /* "saturation-off" */
    float value_mul;      /* "value-mul"      */
    float value_off;      /* "value-off"      */
} b200vf_hsvfilter_params;

/* In place (AlwaysInPlace, imp.rs:316-317).  Formats: the 10 of imp.rs:278-289. */
B200VF_API int b200vf_hsvfilter_process(b200vf_ctx *ctx, const b200vf_frame *frame,
                                        const b200vf_hsvfilter_params *params);
B200VF_API int b200vf_hsvfilter_process_batch(b200vf_ctx *ctx, const b200vf_frame *frames,
                                              size_t n_frames,
                                              const b200vf_hsvfilter_params *params);

/* ---- hsvdetector ----------------------------------------------------------- */
/* Property snapshot; defaults hsvdetector/imp.rs:26-31 = {0, 10, 0, 0.15, 0, 0.3}. */
typedef struct b200vf_hsvdetector_params {
    float hue_ref;        /* "hue-ref"        degrees */
    float hue_var;        /* "hue-var"        [0,180] */
    float saturation_ref; /* "saturation-ref" [0,1] */
    float saturation_var; /* "saturation-var" [0,1] */
    float value_ref;      /* "value-ref"      [0,1] */
    float value_var;      /* "value-var"      [0,1] */
} b200vf_hsvdetector_params;

/* in.format ∈ {RGBx,xRGB,BGRx,xBGR,RGB,BGR} (imp.rs:78-87),
 * out.format ∈ {RGBA,ARGB,BGRA,ABGR} (imp.rs:89-96); same width/height. */
B200VF_API int b200vf_hsvdetector_process(b200vf_ctx *ctx, const b200vf_frame *in,
                                          const b200vf_frame *out,
                                          const b200vf_hsvdetector_params *params);
B200VF_API int b200vf_hsvdetector_process_batch(b200vf_ctx *ctx, const b200vf_frame *in,
                                                const b200vf_frame *out, size_t n_frames,
                                                const b200vf_hsvdetector_params *params);

/* ---- colorlut ! hsvfilter chain ---------------------- */
/* One pass equal to colorlut_process(in → out) followed by hsvfilter_process(out):
 * bit-identical to running the two elements back to back, half the HBM traffic.
 * RGBA only. */
B200VF_API int b200vf_chain_lut_hsv_process_batch(b200vf_ctx *ctx, const b200vf_frame *in,
                                                  const b200vf_frame *out, size_t n_frames,
                                                  const b200vf_hsvfilter_params *params);

/* ---- frame-parallel group: one process feeding several GPUs ------------------ */
/* A group is one ordinary context per listed device, each with its own host thread and its three
 * streams.  Every *_process_batch call hands frame i of the batch to member i mod G; members work
 * concurrently, no data moves between devices and there is no collective (every output pixel
 * depends on one input pixel: colorlut/imp.rs:288-292, hsvfilter/imp.rs:97-118,
 * hsvdetector/imp.rs:134-158).  Results land in the caller's out[i], so order is preserved by
 * construction.  Host frames: the call returns when every member has finished (synchronous, like
 * the per-context calls; with "host.async" = 1 set through b200vf_group_set_option, when every
 * member has queued its share — complete them with b200vf_group_synchronize() or per member with
 * b200vf_ctx_host_wait(b200vf_group_ctx(g, m), b200vf_ctx_host_ticket(...))).  Device frames: frame i must live on the device of member i mod G
 * (checked; B200VF_ERR_INVALID_ARG otherwise) and the call returns once every member has enqueued
 * its share on its own stream — b200vf_group_synchronize() or the members' streams order later
 * work.  LUT and options are replicated on every member.  A device may be listed more than once
 * (several members = several concurrent stream pipelines on that GPU).  A group is single-caller,
 * like a context; the reference-side precedent for an element following a device is
 * d3d12colorlut/imp.rs:494-542. */
typedef struct b200vf_group b200vf_group;
B200VF_API int b200vf_group_create(const int *devices, size_t n_devices, b200vf_group **out);
B200VF_API void b200vf_group_destroy(b200vf_group *group);
B200VF_API size_t b200vf_group_size(const b200vf_group *group);
/* Borrowed member context (stream handle, stats, per-member options); NULL if out of range. */
B200VF_API b200vf_ctx *b200vf_group_ctx(b200vf_group *group, size_t member);
B200VF_API const char *b200vf_group_last_error(const b200vf_group *group);
B200VF_API int b200vf_group_set_option(b200vf_group *group, const char *key, int64_t value);
B200VF_API int b200vf_group_synchronize(b200vf_group *group);
B200VF_API int b200vf_group_colorlut_set_lut(b200vf_group *group, uint32_t kind, uint32_t size,
                                             const float *data, const float domain_scale[3],
                                             const float domain_offset[3]);
B200VF_API int b200vf_group_colorlut_set_lut_file(b200vf_group *group, const char *location);
B200VF_API int b200vf_group_colorlut_clear_lut(b200vf_group *group);
B200VF_API int b200vf_group_colorlut_process_batch(b200vf_group *group, const b200vf_frame *in,
                                                   const b200vf_frame *out, size_t n_frames);
B200VF_API int b200vf_group_hsvfilter_process_batch(b200vf_group *group, const b200vf_frame *frames,
                                                    size_t n_frames,
                                                    const b200vf_hsvfilter_params *params);
B200VF_API int b200vf_group_hsvdetector_process_batch(b200vf_group *group, const b200vf_frame *in,
                                                      const b200vf_frame *out, size_t n_frames,
                                                      const b200vf_hsvdetector_params *params);
B200VF_API int b200vf_group_chain_lut_hsv_process_batch(b200vf_group *group, const b200vf_frame *in,
                                                        const b200vf_frame *out, size_t n_frames,
                                                        const b200vf_hsvfilter_params *params);

/* ---- frame pool (memory:CUDAMemory buffer pool; pinned host pool) ---- */
/* What gst_d3d12::D3D12BufferPool is to d3d12colorlut (d3d12colorlut/imp.rs:385-492): the
 * allocator an element proposes upstream / decides on for its own output so that frames stay
 * in HBM between elements.  A pool belongs to a device, not to a context: buffers may
 * outlive the element that allocated them and be shared by several contexts on that device.
 * All frames of a pool have one geometry; the stride is width*bpp when that is a multiple of
 * 16 (contiguous frames take the kernels' long-row path), else the next multiple of 256.
 * Thread-safe. */
typedef struct b200vf_pool b200vf_pool;
typedef struct b200vf_pool_config {
    uint32_t width, height; /* pixels / rows, both > 0 */
    uint32_t format;        /* b200vf_format */
    uint32_t min_buffers;   /* allocated at create time */
    uint32_t max_buffers;   /* 0 = unlimited; otherwise >= min_buffers */
    uint32_t host_pinned;   /* 0 = device memory; 1 = page-locked system memory (frames report
                             * B200VF_MEM_HOST): what a system-memory element proposes upstream so
                             * that its frames reach the GPU without the pageable bounce copy */
} b200vf_pool_config;
typedef struct b200vf_pool_stats {
    uint32_t allocated;   /* buffers that exist */
    uint32_t outstanding; /* acquired and not yet released */
    uint64_t frame_bytes; /* stride * height */
    int64_t stride;
} b200vf_pool_stats;
enum { B200VF_POOL_DONTWAIT = 1 }; /* GST_BUFFER_POOL_ACQUIRE_FLAG_DONTWAIT */

B200VF_API int b200vf_pool_create(int device, const b200vf_pool_config *config, b200vf_pool **out);
/* Frees every buffer, outstanding ones included, after the device has gone idle. */
B200VF_API void b200vf_pool_destroy(b200vf_pool *pool);
/* Fills *out with a frame of the pool (memory = B200VF_MEM_DEVICE, or _HOST for a pinned pool).  At max_buffers the call blocks
 * until a frame is released, or returns B200VF_ERR_NOMEM with B200VF_POOL_DONTWAIT.  A recycled
 * frame is handed out only after the work recorded at its release has finished. */
B200VF_API int b200vf_pool_acquire(b200vf_pool *pool, uint32_t flags, b200vf_frame *out);
/* Returns a frame to the pool.  `last_use_stream` is the cudaStream_t (as void*) whose already
 * enqueued work still touches the frame — e.g. b200vf_ctx_get_stream(ctx) right after a
 * *_process call — or NULL when the frame is idle; nothing blocks here. */
B200VF_API int b200vf_pool_release(b200vf_pool *pool, const b200vf_frame *frame,
                                   void *last_use_stream);
/* Same, with "whatever stream `last_user` works on" — also when that is the legacy default stream,
 * whose NULL handle b200vf_pool_release would read as "idle". */
B200VF_API int b200vf_pool_release_after(b200vf_pool *pool, const b200vf_frame *frame,
                                         const b200vf_ctx *last_user);
B200VF_API int b200vf_pool_get_stats(b200vf_pool *pool, b200vf_pool_stats *out);
B200VF_API int b200vf_pool_device(const b200vf_pool *pool);

/* Which memory a pointer is (b200vf_memory) and, for device memory, on which device — what
 * d3d12colorlut's before_transform asks of the incoming buffer to follow its device
 * (d3d12colorlut/imp.rs:494-542).  Pageable and pinned host memory both report HOST, device -1. */
B200VF_API int b200vf_pointer_info(const void *p, uint32_t *memory, int *device);

/* ---- diagnostics ------------------------------------------------------------ */
/* RGB→HSV (hsvutils.rs:44-84) of n RGBA pixels in device memory → 3 floats (h,s,v) per pixel
 * in device memory, computed by the same device function the kernels use ("hsv.math" selects
 * the variant).  Lets the tests prove float-level equality with the reference over all 2^24
 * inputs; not used by any element. */
B200VF_API int b200vf_debug_hsv_from_rgb(b200vf_ctx *ctx, const void *rgba_device, size_t n_pixels,
                                         float *hsv_device);
/* Position of colour triples (c0 | c1 << 8 | c2 << 16, bits 24..31 ignored) inside the 2^24-entry
 * function tables — the colour-blocked order of DESIGN.md §13 — evaluated on the host by the very
 * function the kernels use.  No device needed; lets the CPU tests check the layout (a bijection
 * whose 128-byte lines are 4 x 4 x 2 blocks of neighbouring colours). */
B200VF_API int b200vf_debug_table_indices(const uint32_t *colours, size_t n, uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif /* B200VF_H */
