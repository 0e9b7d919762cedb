/* rgb3_paths.c — 3-byte pixel formats through the function tables, end to end through the C ABI,
 * against the oracle and against the per-pixel compute kernels; no Python needed, so it runs in a
 * second on a GPU box.  Covers what routes RGB / BGR frames to vf_map_tile3_staged_kernel (width a
 * multiple of 128 pixels, 16-byte aligned rows) and, right next to it, what does not (other widths,
 * odd strides), in place and out of place, device and system memory, batches.
 *
 *   gcc -O2 -std=c11 -Iinclude -Ioracle tests/cpp/rgb3_paths.c -Lgst-plugins-rs_b200 -lb200vf \
 *       -Loracle -loracle -lm -Wl,-rpath,$PWD/gst-plugins-rs_b200 -Wl,-rpath,$PWD/oracle -o rgb3_paths
 *
 * exit code 0 = every comparison equal.  (tests/test_gpu_rgb3_paths.py builds and runs it.) */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "b200vf.h"
#include "vf_oracle.h"

static uint64_t g_state = 0x5EED0000u;
static uint64_t splitmix(void) {
    uint64_t z = (g_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

/* smooth content with a little noise, or random bytes: both table-friendly and adversarial */
static void fill(uint8_t *p, uint32_t w, uint32_t h, size_t stride, int bpp, int random) {
    for (uint32_t y = 0; y < h; y++)
        for (uint32_t x = 0; x < w; x++) {
            uint8_t *q = p + (size_t)y * stride + (size_t)x * bpp;
            const uint64_t r = splitmix();
            for (int c = 0; c < bpp; c++) {
                const int base = c == 0 ? (int)(x * 255 / (w > 1 ? w - 1 : 1)) : c == 1 ? (int)(y * 255 / (h > 1 ? h - 1 : 1)) : (int)((x + y) & 255);
                int v = random ? (int)((r >> (8 * c)) & 255) : base + (int)((r >> (8 * c)) % 5) - 2;
                q[c] = (uint8_t)(v < 0 ? 0 : v > 255 ? 255 : v);
            }
        }
}

static int g_fail = 0;
#define CHECK(call)                                                                       \
    do {                                                                                  \
        int rc_ = (call);                                                                 \
        if (rc_ != B200VF_OK) {                                                           \
            printf("FAIL %s -> %d (%s) at line %d\n", #call, rc_, b200vf_last_error(ctx), __LINE__); \
            exit(2);                                                                      \
        }                                                                                 \
    } while (0)

static void expect_equal(const char *what, const uint8_t *a, const uint8_t *b, size_t n) {
    if (memcmp(a, b, n) != 0) {
        size_t i = 0, cnt = 0;
        for (size_t k = 0; k < n; k++)
            if (a[k] != b[k]) {
                if (!cnt) i = k;
                cnt++;
            }
        printf("MISMATCH %s: %zu of %zu bytes, first at %zu (%u vs %u)\n", what, cnt, n, i, a[i], b[i]);
        g_fail++;
    }
}

static const b200vf_hsvfilter_params kFilter = {37.5f, 1.2f, 0.05f, 0.9f, 0.02f};
static const b200vf_hsvdetector_params kDetect = {120.0f, 30.0f, 0.6f, 0.4f, 0.6f, 0.4f};

/* One geometry: hsvfilter in place (3 -> 3) and hsvdetector (3 -> 4), n frames, device memory, table
 * path and compute path, both against the oracle; then one host-memory frame through the table. */
static void run_case(b200vf_ctx *ctx, uint32_t w, uint32_t h, size_t pad3, size_t pad4, int n, uint32_t fmt3, uint32_t fmt4,
                     int random) {
    const size_t s3 = (size_t)w * 3 + pad3, s4 = (size_t)w * 4 + pad4;
    const size_t b3 = s3 * h, b4 = s4 * h;
    uint8_t *src = malloc(b3 * n), *want3 = malloc(b3 * n), *want4 = malloc(b4 * n), *got = malloc((b3 > b4 ? b3 : b4) * n);
    memset(src, 0xA5, b3 * n);
    for (int f = 0; f < n; f++) fill(src + f * b3, w, h, s3, 3, random);
    memcpy(want3, src, b3 * n);
    memset(want4, 0x5A, b4 * n);
    const orc_hsvfilter_params of = {kFilter.hue_shift, kFilter.saturation_mul, kFilter.saturation_off, kFilter.value_mul, kFilter.value_off};
    const orc_hsvdetector_params od = {kDetect.hue_ref, kDetect.hue_var, kDetect.saturation_ref, kDetect.saturation_var, kDetect.value_ref, kDetect.value_var};
    for (int f = 0; f < n; f++) {
        orc_hsvfilter_frame(want3 + f * b3, s3, w, h, (int)fmt3, &of);
        orc_hsvdetector_frame(src + f * b3, s3, (int)fmt3, want4 + f * b4, s4, (int)fmt4, w, h, &od);
    }
    void *d3 = NULL, *d4 = NULL;
    CHECK(b200vf_device_alloc(ctx, b3 * n, &d3));
    CHECK(b200vf_device_alloc(ctx, b4 * n, &d4));
    b200vf_frame f3[8], f4[8];
    for (int f = 0; f < n; f++) {
        f3[f] = (b200vf_frame){(uint8_t *)d3 + f * b3, (int64_t)s3, w, h, fmt3, B200VF_MEM_DEVICE};
        f4[f] = (b200vf_frame){(uint8_t *)d4 + f * b4, (int64_t)s4, w, h, fmt4, B200VF_MEM_DEVICE};
    }
    char what[160];
    for (int path = 2; path >= 1; path--) {  /* 2 = function table, 1 = per-pixel compute kernels */
        CHECK(b200vf_ctx_set_option(ctx, "hsv.path", path));
        /* hsvdetector first (reads d3), then hsvfilter in place on d3 */
        CHECK(b200vf_memcpy(ctx, d3, src, b3 * n, 0));
        {   /* the oracle wrote only width * 4 bytes per row of want4: start the device copy from the same padding pattern */
            uint8_t *init4 = malloc(b4 * n);
            memset(init4, 0x5A, b4 * n);
            CHECK(b200vf_memcpy(ctx, d4, init4, b4 * n, 0));
            free(init4);
        }
        CHECK(b200vf_hsvdetector_process_batch(ctx, f3, f4, (size_t)n, &kDetect));
        CHECK(b200vf_memcpy(ctx, got, d4, b4 * n, 1));
        snprintf(what, sizeof what, "hsvdetector %ux%u pad %zu/%zu n=%d fmt %u->%u hsv.path=%d %s", w, h, pad3, pad4, n, fmt3, fmt4, path,
                 random ? "rand" : "smooth");
        expect_equal(what, got, want4, b4 * n);
        CHECK(b200vf_hsvfilter_process_batch(ctx, f3, (size_t)n, &kFilter));
        CHECK(b200vf_memcpy(ctx, got, d3, b3 * n, 1));
        snprintf(what, sizeof what, "hsvfilter %ux%u pad %zu n=%d fmt %u hsv.path=%d %s", w, h, pad3, n, fmt3, path, random ? "rand" : "smooth");
        expect_equal(what, got, want3, b3 * n);
        int64_t active = -1;
        CHECK(b200vf_ctx_get_option(ctx, "hsv.table_active", &active));
        if (active != (path == 2)) printf("NOTE %s: hsv.table_active = %lld\n", what, (long long)active), g_fail++;
    }
    /* a system-memory frame through the table path (chunked rows, pitched staging buffers) */
    CHECK(b200vf_ctx_set_option(ctx, "hsv.path", 2));
    memcpy(got, src, b3);
    b200vf_frame hf = {got, (int64_t)s3, w, h, fmt3, B200VF_MEM_HOST};
    CHECK(b200vf_hsvfilter_process(ctx, &hf, &kFilter));
    snprintf(what, sizeof what, "hsvfilter host frame %ux%u pad %zu fmt %u", w, h, pad3, fmt3);
    expect_equal(what, got, want3, b3);
    CHECK(b200vf_device_free(ctx, d3));
    CHECK(b200vf_device_free(ctx, d4));
    free(src), free(want3), free(want4), free(got);
}

/* colorlut with the conversions folded in, 3-byte input: RGB -> RGBA, BGR -> RGBA and RGB -> RGB must
 * carry exactly the colours colorlut gives on the RGBA view of the input (alpha 255). */
static void run_convert(b200vf_ctx *ctx, uint32_t w, uint32_t h, int n) {
    const size_t b3 = (size_t)w * h * 3, b4 = (size_t)w * h * 4;
    uint8_t *rgb = malloc(b3 * n), *bgr = malloc(b3 * n), *rgba = malloc(b4 * n), *ref = malloc(b4 * n), *got = malloc(b4 * n),
            *ref3 = malloc(b3 * n);
    for (int f = 0; f < n; f++) fill(rgb + f * b3, w, h, (size_t)w * 3, 3, f & 1);
    for (size_t i = 0; i < (size_t)w * h * n; i++) {
        bgr[3 * i] = rgb[3 * i + 2], bgr[3 * i + 1] = rgb[3 * i + 1], bgr[3 * i + 2] = rgb[3 * i];
        rgba[4 * i] = rgb[3 * i], rgba[4 * i + 1] = rgb[3 * i + 1], rgba[4 * i + 2] = rgb[3 * i + 2], rgba[4 * i + 3] = 255;
    }
    void *d_rgb, *d_bgr, *d_rgba, *d_out;
    CHECK(b200vf_device_alloc(ctx, b3 * n, &d_rgb));
    CHECK(b200vf_device_alloc(ctx, b3 * n, &d_bgr));
    CHECK(b200vf_device_alloc(ctx, b4 * n, &d_rgba));
    CHECK(b200vf_device_alloc(ctx, b4 * n, &d_out));
    CHECK(b200vf_memcpy(ctx, d_rgb, rgb, b3 * n, 0));
    CHECK(b200vf_memcpy(ctx, d_bgr, bgr, b3 * n, 0));
    CHECK(b200vf_memcpy(ctx, d_rgba, rgba, b4 * n, 0));
    b200vf_frame f_rgb[8], f_bgr[8], f_rgba[8], f_out4[8], f_out3[8];
    for (int f = 0; f < n; f++) {
        f_rgb[f] = (b200vf_frame){(uint8_t *)d_rgb + f * b3, (int64_t)w * 3, w, h, B200VF_FORMAT_RGB, B200VF_MEM_DEVICE};
        f_bgr[f] = (b200vf_frame){(uint8_t *)d_bgr + f * b3, (int64_t)w * 3, w, h, B200VF_FORMAT_BGR, B200VF_MEM_DEVICE};
        f_rgba[f] = (b200vf_frame){(uint8_t *)d_rgba + f * b4, (int64_t)w * 4, w, h, B200VF_FORMAT_RGBA, B200VF_MEM_DEVICE};
        f_out4[f] = (b200vf_frame){(uint8_t *)d_out + f * b4, (int64_t)w * 4, w, h, B200VF_FORMAT_RGBA, B200VF_MEM_DEVICE};
        f_out3[f] = (b200vf_frame){(uint8_t *)d_out + f * b3, (int64_t)w * 3, w, h, B200VF_FORMAT_RGB, B200VF_MEM_DEVICE};
    }
    CHECK(b200vf_colorlut_process_batch(ctx, f_rgba, f_out4, (size_t)n));
    CHECK(b200vf_memcpy(ctx, ref, d_out, b4 * n, 1));
    for (size_t i = 0; i < (size_t)w * h * n; i++) ref3[3 * i] = ref[4 * i], ref3[3 * i + 1] = ref[4 * i + 1], ref3[3 * i + 2] = ref[4 * i + 2];
    char what[120];
    CHECK(b200vf_colorlut_convert_process_batch(ctx, f_rgb, f_out4, (size_t)n));
    CHECK(b200vf_memcpy(ctx, got, d_out, b4 * n, 1));
    snprintf(what, sizeof what, "colorlut_convert RGB->RGBA %ux%u n=%d", w, h, n);
    expect_equal(what, got, ref, b4 * n);
    CHECK(b200vf_colorlut_convert_process_batch(ctx, f_bgr, f_out4, (size_t)n));
    CHECK(b200vf_memcpy(ctx, got, d_out, b4 * n, 1));
    snprintf(what, sizeof what, "colorlut_convert BGR->RGBA %ux%u n=%d", w, h, n);
    expect_equal(what, got, ref, b4 * n);
    CHECK(b200vf_colorlut_convert_process_batch(ctx, f_rgb, f_out3, (size_t)n));
    CHECK(b200vf_memcpy(ctx, got, d_out, b3 * n, 1));
    snprintf(what, sizeof what, "colorlut_convert RGB->RGB %ux%u n=%d", w, h, n);
    expect_equal(what, got, ref3, b3 * n);
    CHECK(b200vf_device_free(ctx, d_rgb));
    CHECK(b200vf_device_free(ctx, d_bgr));
    CHECK(b200vf_device_free(ctx, d_rgba));
    CHECK(b200vf_device_free(ctx, d_out));
    free(rgb), free(bgr), free(rgba), free(ref), free(got), free(ref3);
}

int main(void) {
    b200vf_ctx *ctx = NULL;
    if (b200vf_ctx_create(0, &ctx) != B200VF_OK) {
        printf("no device: %s\n", b200vf_last_error(NULL));
        return 3;
    }
    /* staged kernel: width % 128 == 0, aligned rows */
    run_case(ctx, 3840, 2160, 0, 0, 1, B200VF_FORMAT_RGB, B200VF_FORMAT_RGBA, 0);
    run_case(ctx, 1920, 1080, 0, 0, 3, B200VF_FORMAT_BGR, B200VF_FORMAT_ABGR, 1);
    run_case(ctx, 256, 37, 0, 0, 2, B200VF_FORMAT_RGB, B200VF_FORMAT_BGRA, 1);   /* rows not a multiple of 32 */
    run_case(ctx, 128, 1, 0, 0, 1, B200VF_FORMAT_BGR, B200VF_FORMAT_ARGB, 0);    /* one segment */
    run_case(ctx, 384, 65, 16, 32, 2, B200VF_FORMAT_RGB, B200VF_FORMAT_RGBA, 0); /* padded but aligned strides */
    /* right next to it: the other kernels */
    run_case(ctx, 384, 40, 4, 4, 2, B200VF_FORMAT_RGB, B200VF_FORMAT_RGBA, 1);   /* 4-byte aligned strides only */
    run_case(ctx, 200, 33, 0, 0, 2, B200VF_FORMAT_BGR, B200VF_FORMAT_ABGR, 0);   /* width % 128 != 0 */
    run_case(ctx, 130, 9, 1, 3, 1, B200VF_FORMAT_RGB, B200VF_FORMAT_ARGB, 1);    /* odd strides */
    {   /* a 9^3 LUT for the conversion cases; the RGBA -> RGBA colorlut it is compared with is itself oracle-tested */
        static char text[64 * 1024];
        size_t len = (size_t)snprintf(text, sizeof text, "LUT_3D_SIZE 9\n");
        for (int b = 0; b < 9; b++)
            for (int g = 0; g < 9; g++)
                for (int r = 0; r < 9; r++)
                    len += (size_t)snprintf(text + len, sizeof text - len, "%.6f %.6f %.6f\n", (r * r) / 64.0, (g + b) / 16.0,
                                            1.0 - b / 8.0 * (1.0 - r / 16.0));
        b200vf_cube cube;
        char err[256] = {0};
        if (b200vf_cube_parse(text, len, &cube, err, sizeof err) != B200VF_OK) {
            printf("FAIL cube parse: %s\n", err);
            return 2;
        }
        CHECK(b200vf_colorlut_set_lut(ctx, cube.kind, cube.size, cube.data, cube.domain_scale, cube.domain_offset));
        b200vf_cube_free(&cube);
        run_convert(ctx, 3840, 2160, 1);
        run_convert(ctx, 256, 37, 2);
        run_convert(ctx, 200, 33, 2);
    }
    b200vf_ctx_destroy(ctx);
    if (g_fail)
        printf("FAILED: %d mismatches\n", g_fail);
    else
        printf("ALL EQUAL\n");
    return g_fail ? 1 : 0;
}
