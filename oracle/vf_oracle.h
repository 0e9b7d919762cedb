/*
 * vf_oracle.h — CPU ORACLE for the colorlut / hsvfilter / hsvdetector hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C restatement of the reference's
 * Rust arithmetic (gst-plugins-rs 0.16.0-alpha, video/colorlut + video/hsv).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load it; the product library (libb200vf.so)
 * never links, loads or calls anything in oracle/.
 *
 * Parity pin status: the reference ships NO pixel-level fixtures for this path
 *  and cannot be compiled here (no rustc/cargo/GStreamer).  The
 * oracle is pinned against every known-answer test the reference does hold —
 * parser.rs:377-474 (5 tests) and hsvutils.rs:200-280 (4 tests) — and is
 * otherwise pinned by source semantics only: "parity unpinned" at element
 * level (pixel loops, strides, format mappings).
 *
 * Build: gcc -O3 -std=c11 -ffp-contract=off -fno-fast-math (see oracle/Makefile).
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference/).
 */
#ifndef VF_ORACLE_H
#define VF_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Pixel formats — same numbering as include/b200vf.h so tests can share it. */
enum {
    ORC_FMT_RGBA = 0,
    ORC_FMT_RGBX = 1,
    ORC_FMT_XRGB = 2,
    ORC_FMT_ARGB = 3,
    ORC_FMT_BGRX = 4,
    ORC_FMT_BGRA = 5,
    ORC_FMT_XBGR = 6,
    ORC_FMT_ABGR = 7,
    ORC_FMT_RGB = 8,
    ORC_FMT_BGR = 9,
    ORC_FMT_RGBA64_LE = 10,
    ORC_FMT_RGBA64_BE = 11
};

enum { ORC_LUT_1D = 1, ORC_LUT_3D = 3 };

/* video/colorlut/src/parser.rs:18-74 (Lut3D, CubeLutKind, CubeLut) */
typedef struct orc_cube {
    int kind;               /* ORC_LUT_1D | ORC_LUT_3D */
    uint32_t size;          /* entries per axis */
    float domain_scale[3];  /* parser.rs:264-268 */
    float domain_offset[3]; /* parser.rs:270-274 */
    /* 3D: size^3 entries of [r,g,b,1.0] in file order, index x + y*size + z*size^2
     *     (parser.rs:43-53, 253-256).
     * 1D: three planes r[size], g[size], b[size] back to back (parser.rs:226-236). */
    float *data;
    size_t n_floats;
} orc_cube;

/* parser.rs:110-281.  Returns 0 on success; nonzero = CubeParseError.
 * 1 = InvalidLut (message in err), 2 = Io (invalid UTF-8 / unreadable file). */
int orc_cube_parse(const char *text, size_t len, orc_cube *out, char *err, size_t errlen);
int orc_cube_parse_file(const char *path, orc_cube *out, char *err, size_t errlen);
void orc_cube_free(orc_cube *c);

/* colorlut/imp.rs:203-223 dispatch + 226-397 loops.  format ∈ {RGBA, RGBA64_LE,
 * RGBA64_BE}; strides in bytes.  Returns 0, or -1 on bad format. */
int orc_colorlut_frame(const orc_cube *lut, const uint8_t *src, size_t src_stride, uint8_t *dst,
                       size_t dst_stride, uint32_t width, uint32_t height, int format);

/* EXTENSION without a reference counterpart: the same frame loops with the 3D
 * sample replaced by tetrahedral or nearest interpolation (definitions in vf_oracle.c).
 * ORC_INTERP_TRILINEAR is orc_colorlut_frame itself; 1D LUTs ignore the mode. */
enum { ORC_INTERP_TRILINEAR = 0, ORC_INTERP_TETRAHEDRAL = 1, ORC_INTERP_NEAREST = 2 };
int orc_colorlut_frame_ex(const orc_cube *lut, const uint8_t *src, size_t src_stride, uint8_t *dst,
                          size_t dst_stride, uint32_t width, uint32_t height, int format,
                          int interpolation);

/* Single-pixel helpers (colorlut/imp.rs:399-469) for known-answer tests. */
void orc_colorlut_apply_u8(const orc_cube *lut, const uint8_t in[3], uint8_t out[3]);
void orc_colorlut_apply_u16(const orc_cube *lut, const uint16_t in[3], uint16_t out[3]);

/* hsvutils.rs:44-84 / 88-128 / 132-163 / 167-198 */
void orc_hsv_from_rgb(const uint8_t in_p[3], float hsv[3]);
void orc_hsv_from_bgr(const uint8_t in_p[3], float hsv[3]);
/* n RGBA pixels → 3 floats each, orc_hsv_from_rgb on bytes 0..2 (for exhaustive float tests) */
void orc_hsv_from_rgba_batch(const uint8_t *rgba, size_t n, float *hsv);
void orc_hsv_to_rgb(const float hsv[3], uint8_t out[3]);
void orc_hsv_to_bgr(const float hsv[3], uint8_t out[3]);

/* hsvfilter/imp.rs:33-39 */
typedef struct orc_hsvfilter_params {
    float hue_shift, saturation_mul, saturation_off, value_mul, value_off;
} orc_hsvfilter_params;

/* hsvfilter/imp.rs:76-120 + 323-376; in place.  Returns 0, -1 on bad format. */
int orc_hsvfilter_frame(uint8_t *data, size_t stride, uint32_t width, uint32_t height, int format,
                        const orc_hsvfilter_params *p);

/* hsvdetector/imp.rs:34-41 */
typedef struct orc_hsvdetector_params {
    float hue_ref, hue_var, saturation_ref, saturation_var, value_ref, value_var;
} orc_hsvdetector_params;

/* hsvdetector/imp.rs:100-160 + 423-707.  Returns 0, -1 on bad format pair. */
int orc_hsvdetector_frame(const uint8_t *in, size_t in_stride, int in_format, uint8_t *out,
                          size_t out_stride, int out_format, uint32_t width, uint32_t height,
                          const orc_hsvdetector_params *p);

/* Frame-parallel multi-thread drivers used ONLY by bench.py (`cpu_baseline`,
 * `--impl reference`): N independent single-threaded element instances, one
 * frame each at a time — the generous bound of BASELINE.md §3.  `frames` are
 * equally laid out; returns 0 on success. */
int orc_colorlut_frames_mt(const orc_cube *lut, const uint8_t *const *src, uint8_t *const *dst,
                           size_t n_frames, size_t stride, uint32_t width, uint32_t height,
                           int format, int n_threads);
int orc_hsvfilter_frames_mt(uint8_t *const *frames, size_t n_frames, size_t stride, uint32_t width,
                            uint32_t height, int format, const orc_hsvfilter_params *p,
                            int n_threads);
int orc_hsvdetector_frames_mt(const uint8_t *const *in, uint8_t *const *out, size_t n_frames,
                              size_t in_stride, int in_format, size_t out_stride, int out_format,
                              uint32_t width, uint32_t height, const orc_hsvdetector_params *p,
                              int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* VF_ORACLE_H */
