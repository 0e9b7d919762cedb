// vf_launch_colorlut.cu — colorlut launcher + LUT preparation kernels (kernels and ops: vf_ops.cuh).
#include "vf_ops.cuh"

namespace vf {

template <int BITS, bool BE, bool IDENT, bool FAST>
static cudaError_t launch_colorlut_path(cudaStream_t stream, const FrameSet &fs, int n,
                                        const Geom &g, const DeviceLut &lut, int path,
                                        uint64_t *launches) {
    const int bpp = BITS == 8 ? 4 : 8;
    if constexpr (BITS == 8) {
        if (path == 4) {
            ColorLutBakedOp op;
            op.table = lut.lut3d_baked;
            return launch_map(stream, fs, n, g, bpp, bpp, op, launches);
        }
        if (path == 2) {
            ColorLut1dByteOp<IDENT, FAST> op;
            op.L = make_lut_args(lut);
            return launch_map(stream, fs, n, g, bpp, bpp, op, launches);
        }
    }
    if (path == 2) {
        ColorLutOp<BITS, BE, IDENT, FAST, 2> op;
        op.L = make_lut_args(lut);
        return launch_map(stream, fs, n, g, bpp, bpp, op, launches);
    }
    if constexpr (BITS == 8 && FAST) {
        if (path == 3) {
            if (lut.unit_range) {
                ColorLutRgOp<IDENT, true> op;
                op.L = make_lut_args(lut);
                return launch_map(stream, fs, n, g, bpp, bpp, op, launches);
            }
            ColorLutRgOp<IDENT, false> op;
            op.L = make_lut_args(lut);
            return launch_map(stream, fs, n, g, bpp, bpp, op, launches);
        }
    }
    if constexpr (BITS == 8) {
        if (path == 1) {
            ColorLutOp<8, false, IDENT, FAST, 1> op;
            op.L = make_lut_args(lut);
            return launch_map(stream, fs, n, g, bpp, bpp, op, launches);
        }
    }
    if constexpr (FAST) {  // extension modes: one arithmetic variant
        if (path == 5) {
            ColorLutOp<BITS, BE, IDENT, true, 5> op;
            op.L = make_lut_args(lut);
            return launch_map(stream, fs, n, g, bpp, bpp, op, launches);
        }
        if (path == 6) {
            ColorLutOp<BITS, BE, IDENT, true, 6> op;
            op.L = make_lut_args(lut);
            return launch_map(stream, fs, n, g, bpp, bpp, op, launches);
        }
    }
    if constexpr (BITS == 16 && IDENT && FAST) {  // the 16-bit fast op (vf_ops.cuh ColorLut64Op)
        if (lut.lut3d_d && lut.coords16_ok && path == 7) {
#define VF_LUT64(P2, U, S)                                                       \
    if (lut.sm1_pow2 == P2 && lut.unit_range == U && lut.lut3d_d_stride == S) {      \
        ColorLut64Op<BE, P2, U, S> op;                                           \
        op.L = make_lut_args(lut);                                               \
        return launch_map(stream, fs, n, g, bpp, bpp, op, launches);             \
    }
            VF_LUT64(true, true, 65)
            VF_LUT64(true, false, 65)
            VF_LUT64(false, true, 65)
            VF_LUT64(false, false, 65)
            VF_LUT64(true, true, 129)
            VF_LUT64(true, false, 129)
            VF_LUT64(false, true, 129)
            VF_LUT64(false, false, 129)
#undef VF_LUT64
        }
    }
    ColorLutOp<BITS, BE, IDENT, FAST, 0> op;
    op.L = make_lut_args(lut);
    return launch_map(stream, fs, n, g, bpp, bpp, op, launches);
}

int resolved_lut_path(const DeviceLut &lut, int bits, int math_mode, int lut_path, int interp) {
    return resolve_lut_path(lut, bits, math_mode, lut_path, interp);
}

cudaError_t launch_colorlut(cudaStream_t stream, const FrameSet &fs, int n, const Geom &g,
                            int bits, bool big_endian, const DeviceLut &lut, int math_mode,
                            int lut_path, int interp, uint64_t *launches) {
    const int path = resolve_lut_path(lut, bits, math_mode, lut_path, interp);
    const bool ident = lut.identity_domain;
    const bool fast = math_mode != kMathPlain || path == 5 || path == 6;
#define VF_LUT_CASE(B, E, I, F)            \
    if (bits == B && big_endian == E && ident == I && fast == F) \
        return launch_colorlut_path<B, E, I, F>(stream, fs, n, g, lut, path, launches);
    VF_LUT_CASE(8, false, true, true)
    VF_LUT_CASE(8, false, false, true)
    VF_LUT_CASE(8, false, true, false)
    VF_LUT_CASE(8, false, false, false)
    VF_LUT_CASE(16, false, true, true)
    VF_LUT_CASE(16, false, false, true)
    VF_LUT_CASE(16, true, true, true)
    VF_LUT_CASE(16, true, false, true)
    VF_LUT_CASE(16, false, true, false)
    VF_LUT_CASE(16, false, false, false)
    VF_LUT_CASE(16, true, true, false)
    VF_LUT_CASE(16, true, false, false)
#undef VF_LUT_CASE
    return cudaErrorInvalidValue;
}

// colorlut with the videoconvert steps either side folded in: any 8-bit
// packed layout in, any out, through the table baked to 8-bit resolution.  The table is indexed by
// R | G << 8 | B << 16 and its entries are R' | G' << 8 | B' << 16 | 0xFF << 24, so both conversions are
// PRMT selectors: one gathers the colour bytes into index order, the other scatters the entry's
// bytes into the output layout and puts the source's alpha — or the entry's 0xFF when the source
// has none — into the alpha / padding byte.
cudaError_t launch_colorlut_convert(cudaStream_t stream, const FrameSet &fs, int n, const Geom &g,
                                    const PixLayout &in_lay, const PixLayout &out_lay,
                                    const uint32_t *baked, uint64_t *launches) {
    TableMapOp op;
    op.table = baked;
    op.idx_sel = (uint32_t)in_lay.r | (uint32_t)in_lay.g << 4 | (uint32_t)in_lay.b << 8 | 4u << 12;
    uint32_t sel = 0;
    for (int p = 0; p < 4; p++) {
        uint32_t nib = in_lay.a >= 0 ? 4u + (uint32_t)in_lay.a : 3u;  // alpha / padding byte
        if (p == out_lay.r) nib = 0;
        if (p == out_lay.g) nib = 1;
        if (p == out_lay.b) nib = 2;
        sel |= nib << (4 * p);
    }
    op.out_sel = sel;
    return launch_map(stream, fs, n, g, in_lay.bpp, out_lay.bpp, op, launches);
}

// ---------------------------------------------------------------------------
// LUT preparation kernels (run once per set_lut, i.e. per `start`)
// ---------------------------------------------------------------------------

// lut_rx[z][y][r] = lerp(c(x0,y,z), c(x0+1,y,z), tx) for the 8-bit code r, with the
// reference's coordinate arithmetic (imp.rs:471-474, 438, 496-517).
template <bool IDENT>
__global__ void vf_build_rx_kernel(LutArgs L, float4 *dst, uint32_t total) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    uint32_t r = i & 255u;
    uint32_t line = i >> 8;  // y + z*(N+1)
    float x = lut_coord<8, IDENT, true>((float)r, L.scale[0], L.offset[0], L.sm1);
    uint32_t x0;
    float tx;
    lut_split<IDENT>(x, L.n - 1, x0, tx);
    const float4 *b = L.lut3d + ((size_t)line * L.sy + x0);
    dst[i] = lerp_x_pair(b, tx);
}

// lut_rg[z][g][r] = lerp(lut_rx[z][y0][r], lut_rx[z][y0+1][r], ty) for the 8-bit code g
// (imp.rs:519-520); z runs over the N+1 padded planes.
template <bool IDENT>
__global__ void vf_build_rg_kernel(LutArgs L, float4 *dst, uint32_t total) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    uint32_t r = i & 255u, gcode = (i >> 8) & 255u, z = i >> 16;
    float y = lut_coord<8, IDENT, true>((float)gcode, L.scale[1], L.offset[1], L.sm1);
    uint32_t y0;
    float ty;
    lut_split<IDENT>(y, L.n - 1, y0, ty);
    const float4 *b = L.lut_rx + ((size_t)(z * L.sy + y0) * 256u + r);
    const float4 v = lerp4_ref(b[0], b[256], ty);
    // R of the next plane rides in lane 1 (the last, padded plane repeats itself; it is only
    // ever read as "plane z0+1", whose lane 1 is unused)
    const uint32_t zn = min(z + 1, L.n);
    const float4 *bn = L.lut_rx + ((size_t)(zn * L.sy + y0) * 256u + r);
    const float xn = lerp_ref(bn[0].x, bn[256].x, ty);
    dst[i] = make_float4(v.x, xn, v.y, v.z);
}

// baked[blk_index(r | g<<8 | b<<16)] = the direct path's output for the pixel (r,g,b): colorlut/imp.rs:431-449 in full.
template <bool IDENT, int PATH>
__global__ void vf_build_baked_kernel(LutArgs L, uint32_t *dst) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;  // = r | g<<8 | b<<16, 2^24 threads
    ColorLutOp<8, false, IDENT, true, PATH> op;
    op.L = L;
    // byte 3 = 0xFF: the constant alpha of colorlut-with-conversion when the source has no alpha
    dst[blk_index(i)] = (op.px(i, nullptr) & 0xFFFFFFu) | 0xFF000000u;
}

template <bool IDENT>
static void build_baked(cudaStream_t stream, const LutArgs &L, uint32_t *dst, int interp) {
    const unsigned blocks = (1u << 24) / 256;
    if (L.lut1d && !L.lut3d)  // a 1D LUT bakes just the same (colorlut-with-conversion uses it)
        vf_build_baked_kernel<IDENT, 2><<<blocks, 256, 0, stream>>>(L, dst);
    else if (interp == kInterpTetrahedral)
        vf_build_baked_kernel<IDENT, 5><<<blocks, 256, 0, stream>>>(L, dst);
    else if (interp == kInterpNearest)
        vf_build_baked_kernel<IDENT, 6><<<blocks, 256, 0, stream>>>(L, dst);
    else
        vf_build_baked_kernel<IDENT, 0><<<blocks, 256, 0, stream>>>(L, dst);
}

cudaError_t launch_build_baked(cudaStream_t stream, const DeviceLut &lut, uint32_t *dst, int interp,
                               uint64_t *launches) {
    if (!dst || (lut.kind == 3 ? !lut.lut3d : !lut.lut1d)) return cudaErrorInvalidValue;
    LutArgs L = make_lut_args(lut);
    if (lut.identity_domain)
        build_baked<true>(stream, L, dst, interp);
    else
        build_baked<false>(stream, L, dst, interp);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

// lut3d_d entry (x, y, z), x < N, y, z <= N: corner values and their x-differences (red and green of
// row y, blue of rows y and y + 1), from the pair-packed table {R(x), R(x+1), G(x), B(x)} (padded,
// so x + 1 and the far faces exist).
__global__ void vf_build_lut64_kernel(LutArgs L, float *dst, uint32_t s) {
    const uint32_t x = threadIdx.x, y = blockIdx.x, z = blockIdx.y;
    if (x >= L.n) return;
    const float4 a = L.lut3d[x + y * L.sy + z * L.sz], b = L.lut3d[x + 1 + y * L.sy + z * L.sz];
    float4 *d = reinterpret_cast<float4 *>(dst + (size_t)(x + y * s + z * s * s) * 8);
    // {R, G, dR, dG | B(y), B(y+1), dB(y), dB(y+1)}: row y = N is only ever read as some cell's upper row
    float b1 = 0.0f, db1 = 0.0f;
    if (y < L.n) {
        const float4 a2 = L.lut3d[x + (y + 1) * L.sy + z * L.sz], b2 = L.lut3d[x + 1 + (y + 1) * L.sy + z * L.sz];
        b1 = a2.w, db1 = __fsub_rn(b2.w, a2.w);
    }
    d[0] = make_float4(a.x, a.z, __fsub_rn(a.y, a.x), __fsub_rn(b.z, a.z));
    d[1] = make_float4(a.w, b1, __fsub_rn(b.w, a.w), db1);
}

cudaError_t launch_build_lut64(cudaStream_t stream, const DeviceLut &lut, uint64_t *launches) {
    if (lut.kind != 3 || !lut.lut3d || !lut.lut3d_d || lut.size > 128) return cudaErrorInvalidValue;
    vf_build_lut64_kernel<<<dim3(lut.size + 1, lut.size + 1), 128, 0, stream>>>(
        make_lut_args(lut), lut.lut3d_d, (uint32_t)lut.lut3d_d_stride);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

cudaError_t launch_build_resampled(cudaStream_t stream, const DeviceLut &lut, bool rx, bool rg,
                                   uint64_t *launches) {
    if (lut.kind != 3 || !lut.lut3d || !lut.lut3d_rx || (rg && !lut.lut3d_rg)) return cudaErrorInvalidValue;
    LutArgs L = make_lut_args(lut);
    uint32_t total = (lut.size + 1) * (lut.size + 1) * 256u;
    uint32_t blocks = (total + 255) / 256;
    cudaError_t e = cudaSuccess;
    if (rx) {
        if (lut.identity_domain)
            vf_build_rx_kernel<true><<<blocks, 256, 0, stream>>>(L, lut.lut3d_rx, total);
        else
            vf_build_rx_kernel<false><<<blocks, 256, 0, stream>>>(L, lut.lut3d_rx, total);
        if (launches) *launches += 1;
        e = cudaGetLastError();
    }
    if (e != cudaSuccess || !rg) return e;
    total = (lut.size + 1) * 65536u;
    blocks = (total + 255) / 256;
    if (lut.identity_domain)
        vf_build_rg_kernel<true><<<blocks, 256, 0, stream>>>(L, lut.lut3d_rg, total);
    else
        vf_build_rg_kernel<false><<<blocks, 256, 0, stream>>>(L, lut.lut3d_rg, total);
    if (launches) *launches += 1;
    return cudaGetLastError();
}

}  // namespace vf
