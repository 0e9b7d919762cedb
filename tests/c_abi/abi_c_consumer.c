/* abi_c_consumer.c — include/b200vf.h consumed from plain C11 (the way a C or Rust `-sys`
 * binding sees it): the header must compile as C, the structs must have the documented layout,
 * and the device-free entry points must work without a GPU.  Exit code 0 = all checks passed. */
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include "b200vf.h"

#define CHECK(c)                                                   \
    do {                                                           \
        if (!(c)) {                                                \
            fprintf(stderr, "CHECK failed: %s (line %d)\n", #c, __LINE__); \
            return 1;                                              \
        }                                                          \
    } while (0)

_Static_assert(sizeof(b200vf_frame) == 32, "b200vf_frame layout");
_Static_assert(offsetof(b200vf_frame, stride) == 8, "b200vf_frame.stride");
_Static_assert(offsetof(b200vf_frame, format) == 24, "b200vf_frame.format");
_Static_assert(sizeof(b200vf_hsvfilter_params) == 20, "hsvfilter params");
_Static_assert(sizeof(b200vf_hsvdetector_params) == 24, "hsvdetector params");
_Static_assert(sizeof(b200vf_pool_config) == 24, "pool config");
_Static_assert(offsetof(b200vf_pool_config, host_pinned) == 20, "pool config.host_pinned");
_Static_assert(sizeof(b200vf_pool_stats) == 24, "pool stats");

int main(void) {
    CHECK(strstr(b200vf_version(), "sm_100a") != NULL);
    CHECK(b200vf_format_from_name("RGBA") == B200VF_FORMAT_RGBA);
    CHECK(b200vf_format_from_name("BGRx") == B200VF_FORMAT_BGRX);
    CHECK(b200vf_format_bytes_per_pixel(B200VF_FORMAT_RGB) == 3);
    CHECK(b200vf_format_bytes_per_pixel(B200VF_FORMAT_RGBA64_BE) == 8);
    CHECK(strcmp(b200vf_status_string(B200VF_ERR_NO_LUT), "No LUT configured") == 0);

    /* the reference's own parser test (parser.rs:381-408) through the C ABI */
    const char *text = "LUT_3D_SIZE 2\n0 0 0\n1 0 0\n0 1 0\n1 1 0\n0 0 1\n1 0 1\n0 1 1\n1 1 1\n";
    b200vf_cube cube;
    char err[256];
    CHECK(b200vf_cube_parse(text, strlen(text), &cube, err, sizeof err) == B200VF_OK);
    CHECK(cube.kind == B200VF_LUT_3D && cube.size == 2 && cube.n_floats == 32);
    CHECK(cube.data[0] == 0.0f && cube.data[3] == 1.0f);             /* at(0,0,0) = [0,0,0,1] */
    CHECK(cube.data[28] == 1.0f && cube.data[29] == 1.0f && cube.data[30] == 1.0f); /* at(1,1,1) */
    b200vf_cube_free(&cube);
    CHECK(cube.data == NULL);

    const char *bad = "LUT_1D_SIZE 2\n0 0 0\nTITLE \"invalid\"\n1 0 0\n"; /* parser.rs:449-460 */
    CHECK(b200vf_cube_parse(bad, strlen(bad), &cube, err, sizeof err) == B200VF_ERR_PARSE);
    CHECK(strstr(err, "Header found after LUT data") != NULL);

    /* no device → no context, and no CPU fallback */
    int n = -1;
    int rc = b200vf_device_count(&n);
    if (rc != B200VF_OK || n == 0) {
        b200vf_ctx *ctx = (b200vf_ctx *)1;
        CHECK(b200vf_ctx_create(0, &ctx) == B200VF_ERR_NO_DEVICE);
        CHECK(ctx == NULL);
        CHECK(strlen(b200vf_last_error(NULL)) > 0);
    }
    /* pools: argument validation needs no device either */
    {
        b200vf_pool *pool = (b200vf_pool *)1;
        b200vf_pool_config cfg = {64, 32, B200VF_FORMAT_RGBA, 0, 0, 0};
        CHECK(b200vf_pool_create(0, NULL, &pool) == B200VF_ERR_INVALID_ARG && pool == NULL);
        cfg.format = 99;
        CHECK(b200vf_pool_create(0, &cfg, &pool) == B200VF_ERR_UNSUPPORTED_FORMAT);
        cfg.format = B200VF_FORMAT_RGBA;
        cfg.min_buffers = 3, cfg.max_buffers = 2;
        CHECK(b200vf_pool_create(0, &cfg, &pool) == B200VF_ERR_INVALID_ARG);
        CHECK(b200vf_pool_device(NULL) == -1);
        b200vf_pool_destroy(NULL);
        uint32_t memory = 7;
        int device = 7;
        CHECK(b200vf_pointer_info(NULL, &memory, &device) == B200VF_ERR_INVALID_ARG);
    }
    printf("abi_c_consumer: ok\n");
    return 0;
}
