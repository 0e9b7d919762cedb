// vf_launch_hsvfilter.cu — hsvfilter launcher: instantiates the hsvfilter kernels (kernels and ops: vf_ops.cuh).
#include "vf_ops.cuh"

namespace vf {

cudaError_t launch_hsvfilter(cudaStream_t stream, const FrameSet &fs, int n, const Geom &g,
                             const PixLayout &lay, const HsvFilterArgs &a, int math_mode,
                             uint64_t *launches) {
    if (math_mode == kMathPlain) {
        HsvFilterPlainOp op;
        op.p = make_filter_params(a);
        op.ri = (uint32_t)lay.r, op.gi = (uint32_t)lay.g, op.bi = (uint32_t)lay.b;
        return launch_map(stream, fs, n, g, lay.bpp, lay.bpp, op, launches);
    }
    const int kind = angle_kind(a.hue_shift);
#define VF_RUN(K, R, G, B)                                                      \
    {                                                                           \
        HsvFilterFastOp<K, R, G, B> op;                                         \
        op.p = make_filter_params(a);                                           \
        return launch_map(stream, fs, n, g, lay.bpp, lay.bpp, op, launches);    \
    }
#define VF_CALL(R, G, B)                                  \
    switch (kind) {                                       \
    case kAngleZero: VF_RUN(kAngleZero, R, G, B)          \
    case kAngleNonNeg: VF_RUN(kAngleNonNeg, R, G, B)      \
    case kAngleNeg: VF_RUN(kAngleNeg, R, G, B)            \
    default: VF_RUN(kAngleGeneric, R, G, B)               \
    }
    VF_FOR_LAYOUT(lay, VF_CALL)
#undef VF_CALL
#undef VF_RUN
    return cudaErrorInvalidValue;
}

}  // namespace vf
