// vf_launch_chain.cu — fused colorlut ! hsvfilter launcher (kernels and ops: vf_ops.cuh).
#include "vf_ops.cuh"

namespace vf {

cudaError_t launch_chain_lut_hsv(cudaStream_t stream, const FrameSet &fs, int n, const Geom &g,
                                 const DeviceLut &lut, const HsvFilterArgs &a, int lut_path,
                                 int interp, uint64_t *launches) {
    if (lut.kind != 3) return cudaErrorInvalidValue;
    const int path = resolve_lut_path(lut, 8, kMathFast, lut_path, interp);
    const int kind = angle_kind(a.hue_shift);
    const bool ident = lut.identity_domain;
    if (path == 4) {  // LUT stage = one gather from the baked table (any interpolation mode)
#define VF_CHAIN_BAKED(S)                                                             \
    if (kind == S) {                                                                  \
        ChainOp<ColorLutBakedOp, HsvFilterFastOp<S, 0, 1, 2>> op;                     \
        op.lut.table = lut.lut3d_baked;                                               \
        op.hsv.p = make_filter_params(a);                                             \
        return launch_map<decltype(op), true>(stream, fs, n, g, 4, 4, op, launches);  \
    }
        VF_CHAIN_BAKED(kAngleZero)
        VF_CHAIN_BAKED(kAngleNonNeg)
        VF_CHAIN_BAKED(kAngleNeg)
        VF_CHAIN_BAKED(kAngleGeneric)
#undef VF_CHAIN_BAKED
        return cudaErrorInvalidValue;
    }
    if (interp != kInterpTrilinear) return cudaErrorNotSupported;  // interpolating stages are trilinear
#define VF_CHAIN_RUN(LUTOP, S)                                      \
    {                                                               \
        ChainOp<LUTOP, HsvFilterFastOp<S, 0, 1, 2>> op;             \
        op.lut.L = make_lut_args(lut);                              \
        op.hsv.p = make_filter_params(a);                           \
        return launch_map<decltype(op), true>(stream, fs, n, g, 4, 4, op, launches);    \
    }
#define VF_CHAIN_CASE(I, S)                                                                  \
    if (ident == I && kind == S) {                                                          \
        if (path == 3) {                                                                     \
            if (lut.unit_range) VF_CHAIN_RUN(ColorLutRgOp<I VF_COMMA true>, S)               \
            VF_CHAIN_RUN(ColorLutRgOp<I VF_COMMA false>, S)                                  \
        }                                                                                    \
        VF_CHAIN_RUN(ColorLutOp<8 VF_COMMA false VF_COMMA I VF_COMMA true VF_COMMA 0>, S)    \
    }
#define VF_COMMA ,
    VF_CHAIN_CASE(true, kAngleZero)
    VF_CHAIN_CASE(true, kAngleNonNeg)
    VF_CHAIN_CASE(true, kAngleNeg)
    VF_CHAIN_CASE(true, kAngleGeneric)
    VF_CHAIN_CASE(false, kAngleZero)
    VF_CHAIN_CASE(false, kAngleNonNeg)
    VF_CHAIN_CASE(false, kAngleNeg)
    VF_CHAIN_CASE(false, kAngleGeneric)
#undef VF_COMMA
#undef VF_CHAIN_CASE
#undef VF_CHAIN_RUN
    return cudaErrorInvalidValue;
}

}  // namespace vf
