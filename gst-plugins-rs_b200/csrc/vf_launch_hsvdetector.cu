// vf_launch_hsvdetector.cu — hsvdetector launcher + RGB→HSV diagnostics (kernels and ops: vf_ops.cuh).
#include "vf_ops.cuh"

namespace vf {

static HsvDetectParams make_detect_params(const HsvDetectArgs &a) {
    HsvDetectParams p;
    p.hue_off = 180.0f - a.hue_ref;  // hsvdetector/imp.rs:141
    p.hue_var = a.hue_var;
    p.sat_ref = a.sat_ref;
    p.sat_var = a.sat_var;
    p.val_ref = a.val_ref;
    p.val_var = a.val_var;
    return p;
}

static uint32_t detect_selector(const PixLayout &in_lay, const PixLayout &out_lay) {
    uint32_t sel = 0;
    for (int j = 0; j < 4; j++) {
        uint32_t nib = 4u;  // alpha byte
        if (j == out_lay.r) nib = (uint32_t)in_lay.r;
        if (j == out_lay.g) nib = (uint32_t)in_lay.g;
        if (j == out_lay.b) nib = (uint32_t)in_lay.b;
        sel |= nib << (4 * j);
    }
    return sel;
}

cudaError_t launch_hsvdetector(cudaStream_t stream, const FrameSet &fs, int n, const Geom &g,
                               const PixLayout &in_lay, const PixLayout &out_lay,
                               const HsvDetectArgs &a, int math_mode, uint64_t *launches) {
    const uint32_t sel = detect_selector(in_lay, out_lay);
    if (math_mode == kMathPlain) {
        HsvDetectPlainOp op;
        op.p = make_detect_params(a);
        op.ri = (uint32_t)in_lay.r, op.gi = (uint32_t)in_lay.g, op.bi = (uint32_t)in_lay.b;
        op.sel = sel;
        return launch_map(stream, fs, n, g, in_lay.bpp, out_lay.bpp, op, launches);
    }
    const int kind = angle_kind(180.0f - a.hue_ref);
#define VF_RUN(K, R, G, B)                                                            \
    {                                                                                 \
        HsvDetectFastOp<K, R, G, B> op;                                               \
        op.p = make_detect_params(a);                                                 \
        op.sel = sel;                                                                 \
        return launch_map(stream, fs, n, g, in_lay.bpp, out_lay.bpp, op, launches);   \
    }
#define VF_CALL(R, G, B)                                  \
    switch (kind) {                                       \
    case kAngleZero: VF_RUN(kAngleZero, R, G, B)          \
    case kAngleNonNeg: VF_RUN(kAngleNonNeg, R, G, B)      \
    case kAngleNeg: VF_RUN(kAngleNeg, R, G, B)            \
    default: VF_RUN(kAngleGeneric, R, G, B)               \
    }
    VF_FOR_LAYOUT(in_lay, VF_CALL)
#undef VF_CALL
#undef VF_RUN
    return cudaErrorInvalidValue;
}

// ---------------------------------------------------------------------------
// diagnostics: RGB → HSV floats of the fast path, for the exhaustive float-level proof
// ---------------------------------------------------------------------------
__global__ void vf_debug_from_rgb_kernel(const uint32_t *px, float *hsv, size_t n, int plain) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t p = px[i];
    Hsv o = plain ? from_rgb_plain((float)(p & 0xFFu), (float)((p >> 8) & 0xFFu),
                                   (float)((p >> 16) & 0xFFu))
                  : from_rgb_fast2(byte_to_float(p, 0), byte_to_float(p, 1), byte_to_float(p, 2));
    hsv[3 * i + 0] = o.h;
    hsv[3 * i + 1] = o.s;
    hsv[3 * i + 2] = o.v;
}

cudaError_t launch_debug_from_rgb(cudaStream_t stream, const uint32_t *px, float *hsv, size_t n,
                                  int plain, uint64_t *launches) {
    if (n == 0) return cudaSuccess;
    vf_debug_from_rgb_kernel<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(px, hsv, n, plain);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace vf
