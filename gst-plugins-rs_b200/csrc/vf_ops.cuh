// vf_ops.cuh — sm_100a kernels for colorlut / hsvfilter / hsvdetector.
//
// All three elements are pure per-pixel maps: streaming packed pixels
// with a few dozen FP32 instructions each.  Measured on B200 these kernels are bound by
// warp-instruction ISSUE (1 per clock per SM sub-partition), not by HBM, so the design
// minimises instructions per pixel and keeps work off the half-rate ALU pipe:
//   * "vec" kernels move 16 bytes per thread per access (uint4 = 4 RGBA pixels or
//     2 RGBA64 pixels), fully coalesced, evict-first so frame data does not evict the
//     LUT from L1/L2.  A contiguous frame (stride == row bytes) is flattened into one
//     long row by the launcher so no lane idles at row ends; loops are nested
//     (segment, row) so no per-pixel division is needed.
//   * "any" kernels are the alignment-free path (odd strides / pointers and the
//     3-byte RGB/BGR formats): one pixel per thread, byte accesses, writes only
//     width*bpp bytes per row.
//   * Channel positions are template parameters on the fast paths (PRMT selectors
//     become immediates); the plain cross-check variants take them at run time.
//   * grid = (row segments, row groups, frames of the batch).
//
// Reference loops replaced: colorlut/imp.rs:237-397, hsvfilter/imp.rs:76-120,
// hsvdetector/imp.rs:100-160.
#pragma once
#include <algorithm>
#include <type_traits>

#include "vf_internal.h"
#include "vf_math.cuh"

namespace vf {

#ifndef VF_UNROLL
#define VF_UNROLL 4
#endif
#ifndef VF_RG_ZTABLE
#define VF_RG_ZTABLE 0
#endif
#ifndef VF_CTAS
#define VF_CTAS 64
#endif
constexpr int kThreads = 256;
constexpr int kUnroll = VF_UNROLL;  // 16-byte units per thread per tile (vec path)
constexpr int kSMs = 148;

struct RowGeom {
    long long in_stride, out_stride;
    uint32_t units_per_row;  // vec: 16-byte units per row; any: pixels per row
    uint32_t tail;           // vec: pixels left over at the end of each row
    uint32_t rows;
    uint32_t tiles_per_row;
};

// Per-CTA lookup table in shared memory: 256 entries of 8 bytes.
//   hsvfilter : entries 0..7 = sector table {center, selector}
//   colorlut  : entry b = {byte offset of LUT plane z0(b), tz(b)} for the blue code b
typedef SectorEntry TabEntry;

// ---------------------------------------------------------------------------
// memory access helpers
// ---------------------------------------------------------------------------
// Frame data is touched once: evict-first loads / stores keep the LUT resident in L1/L2.
// (A/B on B200: .cs vs L1::no_allocate vs plain made no measurable difference — round 1 on the flat
// kernel, round 2 on the tile kernel together with L1::evict_last table gathers: every content class
// within 0.2 % of the default.)
__device__ __forceinline__ uint4 ld_stream16(const void *p) {
    return __ldcs(reinterpret_cast<const uint4 *>(p));
}
__device__ __forceinline__ void st_stream16(void *p, uint4 v) {
    __stcs(reinterpret_cast<uint4 *>(p), v);
}
// LUT entry fetch: read-only path, default caching (entries are reused across pixels).
__device__ __forceinline__ float4 ld_lut16(const float4 *p) { return __ldg(p); }

// Index into a 2^24-entry function table (baked LUT, hsvfilter / hsvdetector / chain tables) for the
// colour triple c0 | c1 << 8 | c2 << 16 (bits 24..31 of `x` are ignored).  The tables are not stored
// in natural order but in blocks of 4 x 4 x 2 neighbouring colours per 128-byte line,
//     index = [c2 7..1][c0 7..2][c1 7..2][c1 1..0][c2 0][c0 1..0],
// because neighbouring pixels of real video differ by a little noise in all three channels: the
// +-2-code neighbourhood of a colour spans ~25 lines of a natural [c2][c1][c0] table, ~12 of this one
// (tools/microbench/tilecache.cu: "noise" content 58 % -> 93 % of the HBM copy peak together with the
// 2-D tile traversal below).  Three field moves, seven ALU operations, masks the alpha byte by the way.
__host__ __device__ __forceinline__ uint32_t blk_index(uint32_t x) {
    uint32_t j = x & 0x00FE0003u;      // c2 high and c0 low stay where they are
    j |= (x >> 5) & 0x000007F8u;       // c1: bits 10-15 -> 5-10, bits 8-9 -> 3-4
    j |= (x & 0x000000FCu) << 9;       // c0 high: bits 2-7 -> 11-16
    j |= (x >> 14) & 0x00000004u;      // c2 low: bit 16 -> 2
    return j;
}

template <int N>
__device__ __forceinline__ void ld_bytes(const uint8_t *p, uint32_t (&w)[2]) {
    w[0] = 0, w[1] = 0;
#pragma unroll
    for (int i = 0; i < N; i++) w[i >> 2] |= (uint32_t)p[i] << (8 * (i & 3));
}
template <int N>
__device__ __forceinline__ void st_bytes(uint8_t *p, const uint32_t (&w)[2]) {
#pragma unroll
    for (int i = 0; i < N; i++) p[i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
}

// ---------------------------------------------------------------------------
// hsvfilter (hsvfilter/imp.rs:76-120 + format arms 327-371)
// ---------------------------------------------------------------------------
// Sector → source of (R,G,B) among {0: c+m, 1: x+m, 2: m}; index 7 = NaN hue.
// (hsvutils.rs:138-154: arms (c,x,0) (x,c,0) (0,c,x) (0,x,c) (x,0,c) (c,0,x), else 0.)
static __constant__ uint8_t kSectorSrcDev[8][3] = {{0, 1, 2}, {0, 1, 2}, {1, 0, 2}, {2, 0, 1},
                                            {2, 1, 0}, {1, 2, 0}, {0, 2, 1}, {2, 2, 2}};

// Fast variant: RI/GI/BI = byte index of R, G, B inside the 32-bit pixel.
template <int KIND, int RI, int GI, int BI>
struct HsvFilterFastOp {
    static constexpr int kPixelBytes = 4;
    HsvFilterParams p;

    __device__ __forceinline__ void init(TabEntry *tab) const {
        if (threadIdx.x < 8) {
            uint32_t k = threadIdx.x, sel = 0;
#pragma unroll
            for (int j = 0; j < 4; j++) {
                uint32_t nib = 4u + j;  // keep the original (alpha / x) byte
                if (j == RI) nib = kSectorSrcDev[k][0];
                if (j == GI) nib = kSectorSrcDev[k][1];
                if (j == BI) nib = kSectorSrcDev[k][2];
                sel |= nib << (4 * j);
            }
            tab[k].center = (k >= 5) ? 5.0f : ((k >= 3) ? 3.0f : 1.0f);
            tab[k].sel = sel;
        }
        __syncthreads();
    }

    __device__ __forceinline__ uint32_t px(uint32_t in, const TabEntry *tab) const {
        Hsv a = from_rgb_fast2(byte_to_float(in, RI), byte_to_float(in, GI), byte_to_float(in, BI));
        Hsv b = hsv_adjust_fast<KIND>(a, p);
        return to_rgb_fast<KIND != kAngleGeneric>(b, tab, in);  // bounded shift ⇒ hue finite
    }
};

// Plain cross-check variant: literal translation, run-time layout.
struct HsvFilterPlainOp {
    static constexpr int kPixelBytes = 4;
    HsvFilterParams p;
    uint32_t ri, gi, bi;

    __device__ __forceinline__ void init(TabEntry *) const {}

    __device__ __forceinline__ uint32_t px(uint32_t in, const TabEntry *) const {
        float r8 = (float)((in >> (8 * ri)) & 0xFFu);
        float g8 = (float)((in >> (8 * gi)) & 0xFFu);
        float b8 = (float)((in >> (8 * bi)) & 0xFFu);
        Hsv b = hsv_adjust_plain(from_rgb_plain(r8, g8, b8), p);
        uint32_t rgb = to_rgb_plain(b);
        uint32_t keep = ~((0xFFu << (8 * ri)) | (0xFFu << (8 * gi)) | (0xFFu << (8 * bi)));
        return (in & keep) | ((rgb & 0xFFu) << (8 * ri)) | (((rgb >> 8) & 0xFFu) << (8 * gi)) |
               (((rgb >> 16) & 0xFFu) << (8 * bi));
    }
};

// ---------------------------------------------------------------------------
// hsvdetector (hsvdetector/imp.rs:100-160 + the 16 closures 428-704)
// ---------------------------------------------------------------------------
template <int KIND, int RI, int GI, int BI>
struct HsvDetectFastOp {
    static constexpr int kPixelBytes = 4;
    HsvDetectParams p;
    uint32_t sel;  // PRMT selector: output bytes from {0..3: input, 4: alpha}

    __device__ __forceinline__ void init(TabEntry *) const {}

    __device__ __forceinline__ uint32_t px(uint32_t in, const TabEntry *) const {
        Hsv a = from_rgb_fast2(byte_to_float(in, RI), byte_to_float(in, GI), byte_to_float(in, BI));
        bool m = hsv_match_fast<KIND>(a, p);
        return prmt(in, m ? 0xFFFFFFFFu : 0u, sel);
    }
};

struct HsvDetectPlainOp {
    static constexpr int kPixelBytes = 4;
    HsvDetectParams p;
    uint32_t ri, gi, bi, sel;

    __device__ __forceinline__ void init(TabEntry *) const {}

    __device__ __forceinline__ uint32_t px(uint32_t in, const TabEntry *) const {
        float r8 = (float)((in >> (8 * ri)) & 0xFFu);
        float g8 = (float)((in >> (8 * gi)) & 0xFFu);
        float b8 = (float)((in >> (8 * bi)) & 0xFFu);
        bool m = hsv_match_plain(from_rgb_plain(r8, g8, b8), p);
        return prmt(in, m ? 0xFFFFFFFFu : 0u, sel);
    }
};

// ---------------------------------------------------------------------------
// colorlut (colorlut/imp.rs:399-543)
// ---------------------------------------------------------------------------
struct LutArgs {
    const float4 *lut3d;   // padded (N+1)^3, entry {R(x), R(x+1), G(x), B(x)}
    const float4 *lut_rx;  // [z][y][r] R-resampled, or null
    const float4 *lut_rg;  // [z][g][r] R- and G-resampled, or null
    const uint32_t *lut_baked;  // [b][g][r] packed output bytes, or null
    const float *lut1d;    // 3 planes of N+1
    const float *lut3d_d;  // 16-bit path: 32-byte entries {R, G, dR, dG, B(y), B(y+1), dB(y), dB(y+1)}, strides S, S^2 (S = 65 / 129)
    float k16_hi, k16_lo;  // RN((N-1)/65535) split hi + lo when N-1 is a power of two (else unused)
    uint32_t bias_bits;    // VF_MAGIC_BITS, passed as data (see ColorLut64Op::px64)
    float neg_zero;        // -0.0f, passed as data (see mulz2)
    uint32_t n;            // N
    uint32_t sy, sz;       // 3D strides in entries: N+1, (N+1)^2
    float sm1;             // (N as f32) - 1.0
    float scale[3], offset[3];
};

// code (integer-valued float) → normalised, domain-mapped coordinate in LUT units
template <int BITS, bool IDENT, bool FAST>
__device__ __forceinline__ float lut_coord(float code, float scale, float offset, float sm1) {
    float v;
    if (FAST)
        v = BITS == 8 ? div255_exact(code) : div65535_exact(code);
    else
        v = code / (BITS == 8 ? 255.0f : 65535.0f);
    // imp.rs:471-479: (v * scale + offset).clamp(0,1).  With scale == 1 and offset == ±0
    // both operations are exact identities on v ∈ [0,1].
    float nrm = IDENT ? v : rs_clamp(__fadd_rn(__fmul_rn(v, scale), offset), 0.0f, 1.0f);
    return __fmul_rn(nrm, sm1);
}

// imp.rs:484-488 / 496-508: i0 = min(floor(p) as usize, N-1), t = p - i0 as f32.
// p ∈ [0, N-1] or NaN (only with a non-identity domain): NaN → i0 = 0, t = NaN.
template <bool IDENT>
__device__ __forceinline__ void lut_split(float p, uint32_t nmax, uint32_t &i0, float &t) {
    if (IDENT) {
        float f = __fadd_rd(p, VF_MAGIC);  // 2^23 + floor(p), p is finite here
        i0 = __float_as_uint(f) & 0xFFFFu;
        t = p - (f - VF_MAGIC);
    } else {
        int i = __float2int_rd(p);  // NaN → 0, saturating
        i0 = (uint32_t)min(max(i, 0), (int)nmax);
        t = p - (float)i0;
    }
}

__device__ __forceinline__ float4 lerp4_ref(float4 a, float4 b, float t) {
    float4 o;
    o.x = lerp_ref(a.x, b.x, t);
    o.y = lerp_ref(a.y, b.y, t);
    o.z = lerp_ref(a.z, b.z, t);
    o.w = 0.0f;  // lane 3 is the constant 1.0 the reference computes and discards
    return o;
}

// One x-lerp of sample_3d: the padded table entry at x0 is {R(x0), R(x0+1), G(x0), B(x0)},
// so corner x0+1 needs only the upper 8 bytes {G, B} of the next entry (LDG.128 + LDG.64
// instead of two LDG.128: 6.6 instead of 8.4 L1 data-pipe cycles per warp).
__device__ __forceinline__ float4 lerp_x_pair(const float4 *e, float tx) {
    const float4 a = __ldg(e);
    const float2 b = __ldg(reinterpret_cast<const float2 *>(e + 1) + 1);
    float4 o;
    o.x = lerp_ref(a.x, a.y, tx);
    o.y = lerp_ref(a.z, b.x, tx);
    o.z = lerp_ref(a.w, b.y, tx);
    o.w = 0.0f;  // lane 3 is the constant 1.0 the reference computes and discards
    return o;
}

// sample_3d (imp.rs:493-526) on the padded table.
template <bool IDENT>
__device__ __forceinline__ float4 sample_3d(const LutArgs &L, float x, float y, float z) {
    uint32_t x0, y0, z0;
    float tx, ty, tz;
    lut_split<IDENT>(x, L.n - 1, x0, tx);
    lut_split<IDENT>(y, L.n - 1, y0, ty);
    lut_split<IDENT>(z, L.n - 1, z0, tz);
    const float4 *b = L.lut3d + (x0 + y0 * L.sy + z0 * L.sz);
    float4 c00 = lerp_x_pair(b, tx);
    float4 c10 = lerp_x_pair(b + L.sy, tx);
    float4 c01 = lerp_x_pair(b + L.sz, tx);
    float4 c11 = lerp_x_pair(b + L.sz + L.sy, tx);
    float4 c0 = lerp4_ref(c00, c10, ty);
    float4 c1 = lerp4_ref(c01, c11, ty);
    return lerp4_ref(c0, c1, tz);
}

// EXTENSION (no reference counterpart; definition: DESIGN.md §11): the unit cell
// split into six tetrahedra by the order of the fractional coordinates; four corners, weights 1-max, max-mid, mid-min, min, accumulated
// left to right, unfused.  Corner RGB = lanes (0, 2, 3) of its own pair-packed entry.
template <bool IDENT>
__device__ __forceinline__ float4 sample_3d_tetrahedral(const LutArgs &L, float x, float y, float z) {
    uint32_t x0, y0, z0;
    float tx, ty, tz;
    lut_split<IDENT>(x, L.n - 1, x0, tx);
    lut_split<IDENT>(y, L.n - 1, y0, ty);
    lut_split<IDENT>(z, L.n - 1, z0, tz);
    const float4 *b = L.lut3d + (x0 + y0 * L.sy + z0 * L.sz);
    // The oracle's six-way branch, as predicates (same tie-breaking: A = tx>ty, B = ty>tz,
    // C = tx>tz, D = tz>ty, E = tz>tx):  max axis = x if A&C, z if D&!C, else y;
    // min axis = x if !A&E, y if A&!B, else z.  The sorted fractions themselves do not depend on
    // how ties are broken, so they come from min/max directly.
    const bool A = tx > ty;
    const bool xmax = A && (tx > tz), zmax = (tz > ty) && !(tx > tz);
    const bool xmin = !A && (tz > tx), ymin = A && !(ty > tz);
    const uint32_t oa = xmax ? 1u : (zmax ? L.sz : L.sy);
    const uint32_t oall = 1u + L.sy + L.sz;
    const uint32_t ob = oall - (xmin ? 1u : (ymin ? L.sy : L.sz));
    const float hi = fmaxf(tx, ty), lo = fminf(tx, ty);
    const float tmax = fmaxf(hi, tz), tmin = fminf(lo, tz), tmid = fmaxf(lo, fminf(hi, tz));
    const float w0 = __fsub_rn(1.0f, tmax), wa = __fsub_rn(tmax, tmid), wb = __fsub_rn(tmid, tmin);
    float w1 = tmin;
    if (!IDENT) {  // a NaN coordinate (non-identity domain) must reach every channel, as in the oracle
        const float any = __fadd_rn(__fadd_rn(tx, ty), tz);
        if (any != any) w1 = any;
    }
    const float4 c0 = __ldg(b), ca = __ldg(b + oa), cb = __ldg(b + ob), c1 = __ldg(b + oall);
    float4 o;
    o.x = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w0, c0.x), __fmul_rn(wa, ca.x)), __fmul_rn(wb, cb.x)),
                    __fmul_rn(w1, c1.x));
    o.y = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w0, c0.z), __fmul_rn(wa, ca.z)), __fmul_rn(wb, cb.z)),
                    __fmul_rn(w1, c1.z));
    o.z = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(w0, c0.w), __fmul_rn(wa, ca.w)), __fmul_rn(wb, cb.w)),
                    __fmul_rn(w1, c1.w));
    o.w = 0.0f;
    return o;
}

// EXTENSION: round half up per axis, one fetch.
// x ∈ [0, N-1] or NaN; NaN + 0.5 → NaN → index 0, as rs_as_usize does.
__device__ __forceinline__ float4 sample_3d_nearest(const LutArgs &L, float x, float y, float z) {
    const int nmax = (int)L.n - 1;
    const uint32_t xi = (uint32_t)min(max(__float2int_rd(__fadd_rn(x, 0.5f)), 0), nmax);
    const uint32_t yi = (uint32_t)min(max(__float2int_rd(__fadd_rn(y, 0.5f)), 0), nmax);
    const uint32_t zi = (uint32_t)min(max(__float2int_rd(__fadd_rn(z, 0.5f)), 0), nmax);
    const float4 c = __ldg(L.lut3d + (xi + yi * L.sy + zi * L.sz));
    return make_float4(c.x, c.z, c.w, 0.0f);
}

// Same value from the R-resampled table: entry [z][y][r] already holds
// lerp(c(x0,y,z), c(x0+1,y,z), tx) computed with the reference's arithmetic for the
// 8-bit code r, so only the y and z lerps remain (4 fetches instead of 8).
template <bool IDENT>
__device__ __forceinline__ float4 sample_3d_rx(const LutArgs &L, uint32_t rcode, float y, float z) {
    uint32_t y0, z0;
    float ty, tz;
    lut_split<IDENT>(y, L.n - 1, y0, ty);
    lut_split<IDENT>(z, L.n - 1, z0, tz);
    const float4 *b = L.lut_rx + ((size_t)(z0 * L.sy + y0) * 256u + rcode);
    float4 c00 = __ldg(b), c10 = __ldg(b + 256);
    float4 c01 = __ldg(b + (size_t)L.sy * 256u), c11 = __ldg(b + (size_t)L.sy * 256u + 256);
    float4 c0 = lerp4_ref(c00, c10, ty);
    float4 c1 = lerp4_ref(c01, c11, ty);
    return lerp4_ref(c0, c1, tz);
}

// sample_1d (imp.rs:482-490) on a plane padded by one entry.
template <bool IDENT>
__device__ __forceinline__ float sample_1d(const float *plane, uint32_t n, float x) {
    uint32_t i0;
    float t;
    lut_split<IDENT>(x, n - 1, i0, t);
    float a = __ldg(plane + i0), b = __ldg(plane + i0 + 1);
    return lerp_ref(a, b, t);
}

// PATH: 0 = 3D direct, 1 = 3D via R-resampled table (8-bit only), 2 = 1D,
//       5 = 3D tetrahedral, 6 = 3D nearest (extensions)
template <int BITS, bool BE, bool IDENT, bool FAST, int PATH>
struct ColorLutOp {
    static constexpr int kPixelBytes = BITS == 8 ? 4 : 8;
    LutArgs L;

    __device__ __forceinline__ void init(TabEntry *) const {}

    __device__ __forceinline__ float3 apply(float c0, float c1, float c2, uint32_t rcode) const {
        float x = lut_coord<BITS, IDENT, FAST>(c0, L.scale[0], L.offset[0], L.sm1);
        float y = lut_coord<BITS, IDENT, FAST>(c1, L.scale[1], L.offset[1], L.sm1);
        float z = lut_coord<BITS, IDENT, FAST>(c2, L.scale[2], L.offset[2], L.sm1);
        float3 o;
        if (PATH == 2) {
            o.x = sample_1d<IDENT>(L.lut1d, L.n, x);
            o.y = sample_1d<IDENT>(L.lut1d + (L.n + 1), L.n, y);
            o.z = sample_1d<IDENT>(L.lut1d + 2 * (L.n + 1), L.n, z);
        } else {
            float4 s;
            if (PATH == 1)
                s = sample_3d_rx<IDENT>(L, rcode, y, z);
            else if (PATH == 5)
                s = sample_3d_tetrahedral<IDENT>(L, x, y, z);
            else if (PATH == 6)
                s = sample_3d_nearest(L, x, y, z);
            else
                s = sample_3d<IDENT>(L, x, y, z);
            o.x = s.x, o.y = s.y, o.z = s.z;
        }
        return o;
    }

    template <int B>
    __device__ __forceinline__ uint32_t code(float v) const {
        return FAST ? unit_to_code_bits<B>(v) : unit_to_code_plain<B>(v);
    }

    // RGBA: bytes R,G,B,A (imp.rs:288-292)
    __device__ __forceinline__ uint32_t px(uint32_t in, const TabEntry *) const {
        float3 o = apply(byte_to_float(in, 0), byte_to_float(in, 1), byte_to_float(in, 2),
                         in & 0xFFu);
        uint32_t r = code<8>(o.x), g = code<8>(o.y), b = code<8>(o.z);
        uint32_t rg = __byte_perm(r, g, 0x0040u);  // low bytes only, so the 2^23 bias is harmless
        return __byte_perm(__byte_perm(rg, b, 0x0410u), in, 0x7210u);
    }

    // RGBA64: four u16 words R,G,B,A in LE or BE byte order; alpha word copied raw
    // (imp.rs:374-395).  in.x = R | G<<16, in.y = B | A<<16 as loaded little-endian.
    __device__ __forceinline__ uint2 px64(uint2 in, const TabEntry *) const {
        // PRMT picks the two bytes of each word in numeric order and sets the 2^23 bias.
        const uint32_t lo = BE ? 0x7401u : 0x7410u, hi = BE ? 0x7423u : 0x7432u;
        float c0 = __uint_as_float(__byte_perm(in.x, VF_MAGIC_BITS, lo)) - VF_MAGIC;
        float c1 = __uint_as_float(__byte_perm(in.x, VF_MAGIC_BITS, hi)) - VF_MAGIC;
        float c2 = __uint_as_float(__byte_perm(in.y, VF_MAGIC_BITS, lo)) - VF_MAGIC;
        float3 o = apply(c0, c1, c2, 0);
        uint32_t r = code<16>(o.x), g = code<16>(o.y), b = code<16>(o.z);
        uint2 out;
        out.x = __byte_perm(r, g, BE ? 0x4501u : 0x5410u);
        out.y = __byte_perm(b, in.y, BE ? 0x7601u : 0x7610u);
        return out;
    }
};

// RGBA64 through a 3D LUT, identity domain (the common case; anything else runs ColorLutOp<16,…,0>).
// Same arithmetic as sample_3d on the reference's values, arranged to issue fewer instructions
// (the direct kernel is bound by warp-instruction issue, 131 per pixel; this op: 81):
//   * table entry (x, y, z) = {R, G, dR, dG | B(y), B(y+1), dB(y), dB(y+1)} (one 32-byte sector) with
//     dC = RN(C(x+1) - C(x)) — the reference's own first operation of each x-lerp, done once at
//     upload: a + dC * tx needs two instructions instead of three;
//   * red and green travel as a pair through add / sub / fma.f32x2 (FADD2 / FFMA2 on sm_100a: the
//     same IEEE operation on both halves of a 64-bit register pair), blue of the two rows of a plane
//     likewise through its x-lerp — the entry layout above delivers exactly these pairs in aligned
//     registers; 73 scalar FP instructions per pixel become 46.  (The packed forms take two dispatch
//     slots, so this buys less time than instructions: + 3.5 points)
//   * the two pixels of a 16-byte unit mostly fall into the same LUT cell: the second one then
//     re-uses the first one's six corner loads from registers;
//   * strides are compile-time constants (S = 65 or 129 entries per row, S^2 per plane — odd on
//     purpose: power-of-two strides put the four corner rows into the same L1 sets, measured 39 %
//     instead of 58 % on noisy content), so the corner rows are one base address + immediates
//     and the index is two multiply-adds;
//   * when N - 1 is a power of two (17, 33, 65, 129 — every common .cube size) the scaling by N - 1
//     is folded into the hi / lo constants of the exact /65535 (scaling by 2^k is exact);
//   * floor() and the integer index come from FRND / F2I on the otherwise idle XU pipe.
// vf_abi.cpp enables this op only after checking, on the host, that its (i0, t) equal the
// reference formula's for all 65536 codes of the LUT's size.
template <bool BE, bool POW2, bool UNIT, int S>
struct ColorLut64Op {
    static constexpr int kPixelBytes = 8;
#ifndef VF_LUT64_MINBLOCKS
#define VF_LUT64_MINBLOCKS 5
#endif
    static constexpr int kMinBlocks = VF_LUT64_MINBLOCKS;
    LutArgs L;

    __device__ __forceinline__ void init(TabEntry *) const {}

    // 16-bit code (integer-valued float) -> cell index and fraction (imp.rs:476-479, 438, 496-508)
    __device__ __forceinline__ void coord(float c, uint32_t &i0, float &t) const {
        const float p = POW2 ? __fmaf_rn(c, L.k16_hi, __fmul_rn(c, L.k16_lo))
                             : __fmul_rn(div65535_exact(c), L.sm1);
        const float fl = floorf(p);   // FRND.FLOOR; p is finite and >= 0
        i0 = (uint32_t)__float2int_rd(p);
        t = __fsub_rn(p, fl);
    }

    static constexpr bool kPair64 = true;  // process_unit hands over both pixels of a unit
    struct Cell {
        uint32_t idx;
        float tx, ty, tz;
    };
    __device__ __forceinline__ f32x2 mz(f32x2 a, f32x2 b) const { return mulz2(a, b, L.neg_zero); }
    __device__ __forceinline__ Cell cell(uint2 in) const {
        const uint32_t lo = BE ? 0x7401u : 0x7410u, hi = BE ? 0x7423u : 0x7432u;
        const uint32_t bias = L.bias_bits;
        Cell c;
        uint32_t x0, y0, z0;
        if constexpr (POW2) {
            const f32x2 c01 = sub2(pk2(__uint_as_float(__byte_perm(in.x, bias, lo)),
                                       __uint_as_float(__byte_perm(in.x, bias, hi))), pk2(VF_MAGIC, VF_MAGIC));
            const f32x2 p01 = fma2(c01, pk2(L.k16_hi, L.k16_hi), mz(c01, pk2(L.k16_lo, L.k16_lo)));
            const float px = lo2(p01), py = hi2(p01);
            const float fx = floorf(px), fy = floorf(py);
            x0 = (uint32_t)__float2int_rd(px), y0 = (uint32_t)__float2int_rd(py);
            const f32x2 t01 = sub2(p01, pk2(fx, fy));
            c.tx = lo2(t01), c.ty = hi2(t01);
        } else {
            coord(__uint_as_float(__byte_perm(in.x, bias, lo)) - VF_MAGIC, x0, c.tx);
            coord(__uint_as_float(__byte_perm(in.x, bias, hi)) - VF_MAGIC, y0, c.ty);
        }
        coord(__uint_as_float(__byte_perm(in.y, bias, lo)) - VF_MAGIC, z0, c.tz);
        c.idx = x0 + y0 * S + z0 * (S * S);
        return c;
    }
    struct Plane {  // rows y0 and y0 + 1 of one z plane
        float4 a0, a1, b;
    };
    template <int OFF>
    __device__ __forceinline__ Plane plane(const float *e) const {
        constexpr int kY = 32 * S;
        Plane p;
        p.a0 = __ldg(reinterpret_cast<const float4 *>(e + OFF / 4));
        p.a1 = __ldg(reinterpret_cast<const float4 *>(e + (OFF + kY) / 4));
        p.b = __ldg(reinterpret_cast<const float4 *>(e + OFF / 4 + 4));
        return p;
    }
    // x- and y-lerp of one plane: (R, G) as a pair, B as a scalar
    __device__ __forceinline__ void plane_xy(const Plane &p, const Cell &c, f32x2 &rg, float &b) const {
        const f32x2 tx = pk2(c.tx, c.tx);
        const f32x2 c0 = add2(pk2(p.a0.x, p.a0.y), mz(pk2(p.a0.z, p.a0.w), tx));
        const f32x2 c1 = add2(pk2(p.a1.x, p.a1.y), mz(pk2(p.a1.z, p.a1.w), tx));
        const f32x2 cb = add2(pk2(p.b.x, p.b.y), mz(pk2(p.b.z, p.b.w), tx));  // (B(y0), B(y0+1))
        rg = add2(c0, mz(sub2(c1, c0), pk2(c.ty, c.ty)));
        b = lerp_ref(lo2(cb), hi2(cb), c.ty);
    }
    __device__ __forceinline__ uint2 eval(const Plane &p0, const Plane &p1, const Cell &c, uint32_t in_y) const {
        f32x2 rg0, rg1;
        float b0, b1;
        plane_xy(p0, c, rg0, b0);
        plane_xy(p1, c, rg1, b1);
        const f32x2 rg = add2(rg0, mz(sub2(rg1, rg0), pk2(c.tz, c.tz)));
        const float bo = lerp_ref(b0, b1, c.tz);
        uint32_t r, g;
        if constexpr (UNIT) {
            const f32x2 y = mz(rg, pk2(65535.0f, 65535.0f));
            r = __float_as_uint(__fadd_rd(__fadd_rz(lo2(y), 0.5f), VF_MAGIC));
            g = __float_as_uint(__fadd_rd(__fadd_rz(hi2(y), 0.5f), VF_MAGIC));
        } else {
            r = unit_to_code_bits<16, UNIT>(lo2(rg)), g = unit_to_code_bits<16, UNIT>(hi2(rg));
        }
        const uint32_t b = unit_to_code_bits<16, UNIT>(bo);
        uint2 out;
        out.x = __byte_perm(r, g, BE ? 0x4501u : 0x5410u);
        out.y = __byte_perm(b, in_y, BE ? 0x7601u : 0x7610u);
        return out;
    }
    // The two pixels of a 16-byte unit: neighbours in a row mostly fall into the same LUT cell, and
    // then the second one re-uses the first one's corners from registers.
    __device__ __forceinline__ uint4 px64_pair(uint4 v, const TabEntry *) const {
        constexpr int kZ = 32 * S * S;
        const Cell a = cell(make_uint2(v.x, v.y)), b = cell(make_uint2(v.z, v.w));
        const float *e = L.lut3d_d + (size_t)a.idx * 8;
        Plane p0 = plane<0>(e), p1 = plane<kZ>(e);
        const uint2 oa = eval(p0, p1, a, v.y);
        if (b.idx != a.idx) {
            e = L.lut3d_d + (size_t)b.idx * 8;
            p0 = plane<0>(e), p1 = plane<kZ>(e);
        }
        const uint2 ob = eval(p0, p1, b, v.w);
        return make_uint4(oa.x, oa.y, ob.x, ob.y);
    }
    __device__ __forceinline__ uint2 px64(uint2 in, const TabEntry *) const {  // the odd pixel at a row's end
        constexpr int kZ = 32 * S * S;
        const Cell a = cell(in);
        const float *e = L.lut3d_d + (size_t)a.idx * 8;
        return eval(plane<0>(e), plane<kZ>(e), a, in.y);
    }
};

// 8-bit RGBA through the R- and G-resampled table: entry [z][g][r] holds the x- and
// y-lerps of the reference already applied (with its own arithmetic) for the byte codes
// r and g, so a pixel needs two fetches and the z-lerp.  g*256 + r is simply the low 16
// bits of the pixel; the blue code indexes a shared-memory table {z0, tz}.
// UNIT: every LUT entry is finite and within [0,1] ⇒ every lerp result is too, and the
// output clamp (imp.rs:538) is the identity.
template <bool IDENT, bool UNIT>
struct ColorLutRgOp {
    static constexpr int kPixelBytes = 4;
#ifndef VF_RG_MINBLOCKS
#define VF_RG_MINBLOCKS 8
#endif
    static constexpr int kMinBlocks = VF_RG_MINBLOCKS;  // two gathers in flight per pixel: occupancy hides them
    LutArgs L;

    __device__ __forceinline__ void init(TabEntry *tab) const {
        if (IDENT && !VF_RG_ZTABLE) return;  // coordinates are computed inline
        uint32_t b = threadIdx.x;  // kThreads == 256 codes
        float z = lut_coord<8, IDENT, true>((float)b, L.scale[2], L.offset[2], L.sm1);
        uint32_t z0;
        float tz;
        lut_split<IDENT>(z, L.n - 1, z0, tz);
        tab[b].center = tz;
        tab[b].sel = z0;
        __syncthreads();
    }

    __device__ __forceinline__ uint32_t px(uint32_t in, const TabEntry *tab) const {
        uint32_t zsel;
        float tz;
        if (IDENT && !VF_RG_ZTABLE) {
            // identity domain: z0 and tz straight from the blue code on the FMA pipe (the kernel
            // is bound by the L1 data pipe, so trading one LDS.64 for five FP ops is a win)
            float p = __fmul_rn(div255_exact(byte_to_float(in, 2)), L.sm1);
            float f = __fadd_rd(p, VF_MAGIC);  // bits = 0x4B000000 + z0
            zsel = __float_as_uint(f);
            tz = p - (f - VF_MAGIC);
        } else {
            TabEntry e = tab[__byte_perm(in, 0, 0x4442u)];
            zsel = e.sel;
            tz = e.center;
        }
        // entry index = z0 << 16 | g << 8 | r: one PRMT glues z0 above the pixel's low 16 bits.
        // Entry layout {R(z), R(z+1), G(z), B(z)}: plane z0 comes as one 16-byte load, and of
        // plane z0+1 only the 8 bytes {G, B} are still needed — an LDG.128 gather costs 4.2 SM
        // cycles per warp, an LDG.64 2.4 (tools/microbench/gather.cu).
        const float4 *p4 = L.lut_rg + __byte_perm(in, zsel, 0x5410u);
        const float4 e0 = ld_lut16(p4);
        const float2 e1 = __ldg(reinterpret_cast<const float2 *>(p4 + 65536) + 1);
        uint32_t r = unit_to_code_bits<8, UNIT>(lerp_ref(e0.x, e0.y, tz));
        uint32_t g = unit_to_code_bits<8, UNIT>(lerp_ref(e0.z, e1.x, tz));
        uint32_t b = unit_to_code_bits<8, UNIT>(lerp_ref(e0.w, e1.y, tz));
        uint32_t rg = __byte_perm(r, g, 0x0040u);
        return __byte_perm(__byte_perm(rg, b, 0x0410u), in, 0x7210u);
    }
};

// 8-bit RGBA through the LUT baked to native resolution: one 4-byte gather per pixel from the
// blocked table (blk_index), traversed in 2-D tiles (vf_map_tile_kernel).
struct ColorLutBakedOp {
    static constexpr int kPixelBytes = 4;
    static constexpr bool kTiled = true;
    const uint32_t *table;
    __device__ __forceinline__ void init(TabEntry *) const {}
    __device__ __forceinline__ uint32_t px(uint32_t in, const TabEntry *) const {
        return __byte_perm(__ldg(table + blk_index(in)), in, 0x7210u);
    }
};

// Any element of this path is a pure function of a pixel's three colour bytes (plus bytes that
// pass through).  TableMapOp applies such a function from a table of all 2^24 triples that was
// filled by running the element's own exact kernel once over every triple (vf_launch_table.cu):
// one PRMT to form the index, one 4-byte gather, one PRMT to merge pass-through bytes.
struct TableMapOp {
    static constexpr int kPixelBytes = 4;
    static constexpr bool kTiled = true;
    const uint32_t *table;
    uint32_t idx_sel;  // PRMT over {pixel, 0}: the colour bytes in memory order, zero-extended
    uint32_t out_sel;  // PRMT over {entry, pixel}: entry bytes, or the pixel's own (alpha / x)
    __device__ __forceinline__ void init(TabEntry *) const {}
    __device__ __forceinline__ uint32_t px(uint32_t in, const TabEntry *) const {
        return __byte_perm(__ldg(table + blk_index(__byte_perm(in, 0u, idx_sel))), in, out_sel);
    }
};

// 1D LUT on 8-bit RGBA: apply_1d (imp.rs:399-413) is a function of one 8-bit code per channel,
// so each CTA evaluates it once for all 256 codes of the three channels — with exactly the
// per-pixel arithmetic of the generic path — into a shared table {R'(c) | G'(c)<<8 | B'(c)<<16},
// and a pixel becomes three shared-memory reads and three PRMTs.
template <bool IDENT, bool FAST>
struct ColorLut1dByteOp {
    static constexpr int kPixelBytes = 4;
    LutArgs L;

    __device__ __forceinline__ void init(TabEntry *tab) const {
        const float code = (float)threadIdx.x;  // kThreads == 256 codes
        uint32_t packed = 0;
#pragma unroll
        for (int c = 0; c < 3; c++) {
            float x = lut_coord<8, IDENT, FAST>(code, L.scale[c], L.offset[c], L.sm1);
            float v = sample_1d<IDENT>(L.lut1d + c * (L.n + 1), L.n, x);
            uint32_t o = FAST ? (unit_to_code_bits<8>(v) & 0xFFu) : unit_to_code_plain<8>(v);
            packed |= o << (8 * c);
        }
        reinterpret_cast<uint32_t *>(tab)[threadIdx.x] = packed;
        __syncthreads();
    }

    __device__ __forceinline__ uint32_t px(uint32_t in, const TabEntry *tab) const {
        const uint32_t *t = reinterpret_cast<const uint32_t *>(tab);
        uint32_t tr = t[in & 0xFFu], tg = t[__byte_perm(in, 0, 0x4441u)],
                 tb = t[__byte_perm(in, 0, 0x4442u)];
        uint32_t rg = __byte_perm(tr, tg, 0x0050u);          // [R'(r), G'(g), ., .]
        return __byte_perm(__byte_perm(rg, tb, 0x0610u), in, 0x7210u);  // + B'(b), alpha
    }
};

// colorlut ! hsvfilter in one pass: the hsvfilter step consumes exactly the bytes
// colorlut would have stored, so the result equals the two-element chain.  The two ops
// use disjoint parts of the shared table (hsv: 0..7 of its own copy).
template <class LutOp, class HsvOp>
struct ChainOp {
    static constexpr int kPixelBytes = 4;
    static constexpr bool kTwoTables = true;
    LutOp lut;
    HsvOp hsv;
    __device__ __forceinline__ void init(TabEntry *tab) const {
        lut.init(tab);
        hsv.init(tab + 256);
    }
    __device__ __forceinline__ uint32_t px(uint32_t in, const TabEntry *tab) const {
        return hsv.px(lut.px(in, tab), tab + 256);
    }
};

template <class Op, class = void>
struct TableEntries {
    static constexpr int value = 256;
};
template <class Op>
struct TableEntries<Op, std::enable_if_t<Op::kTwoTables>> {
    static constexpr int value = 256 + 8;
};

// Latency-bound gather ops ask for full occupancy (8 CTAs x 256 threads = 32 registers/thread).
template <class Op, class = void>
struct MinBlocks {
    static constexpr int value = 0;  // 0 = leave the register budget to the compiler's default
};
template <class Op>
struct MinBlocks<Op, std::enable_if_t<(Op::kMinBlocks > 0)>> {
    static constexpr int value = Op::kMinBlocks;
};

// Table-gather ops ask for the 2-D tile traversal.
template <class Op, class = void>
struct Tiled {
    static constexpr bool value = false;
};
template <class Op>
struct Tiled<Op, std::enable_if_t<Op::kTiled>> {
    static constexpr bool value = true;
};

// ---------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------
template <class Op, class = void>
struct Pair64 {
    static constexpr bool value = false;
};
template <class Op>
struct Pair64<Op, std::enable_if_t<Op::kPair64>> {
    static constexpr bool value = true;
};

template <class Op>
__device__ __forceinline__ uint4 process_unit(const Op &op, uint4 v, const TabEntry *tab) {
    uint4 o;
    if constexpr (Pair64<Op>::value) {
        o = op.px64_pair(v, tab);
    } else if constexpr (Op::kPixelBytes == 4) {
        o.x = op.px(v.x, tab);
        o.y = op.px(v.y, tab);
        o.z = op.px(v.z, tab);
        o.w = op.px(v.w, tab);
    } else {
        uint2 a = op.px64(make_uint2(v.x, v.y), tab);
        uint2 b = op.px64(make_uint2(v.z, v.w), tab);
        o = make_uint4(a.x, a.y, b.x, b.y);
    }
    return o;
}

// 16-byte path: row bases are 16-byte aligned.  grid = (segments, row groups, frames).
template <class Op>
__global__ void __launch_bounds__(kThreads, MinBlocks<Op>::value)
    vf_map_vec_kernel(FrameSet fs, RowGeom g, Op op) {
    __shared__ TabEntry tab[TableEntries<Op>::value];
    op.init(tab);
    const uint8_t *in = fs.in[blockIdx.z];
    uint8_t *out = fs.out[blockIdx.z];
    constexpr uint32_t kPxPerUnit = 16 / Op::kPixelBytes;
    constexpr uint32_t kTile = kThreads * kUnroll;
    for (uint32_t row = blockIdx.y; row < g.rows; row += gridDim.y) {
        const uint8_t *src = in + (size_t)row * g.in_stride;
        uint8_t *dst = out + (size_t)row * g.out_stride;
        for (uint32_t seg = blockIdx.x; seg < g.tiles_per_row; seg += gridDim.x) {
            const uint32_t u0 = seg * kTile + threadIdx.x;
            if (u0 + (kUnroll - 1) * kThreads < g.units_per_row) {  // full tile for this thread
                uint4 v[kUnroll];
#pragma unroll
                for (int j = 0; j < kUnroll; j++)
                    v[j] = ld_stream16(src + (size_t)(u0 + j * kThreads) * 16);
#pragma unroll
                for (int j = 0; j < kUnroll; j++)
                    st_stream16(dst + (size_t)(u0 + j * kThreads) * 16, process_unit(op, v[j], tab));
            } else {
#pragma unroll 1
                for (int j = 0; j < kUnroll; j++) {
                    const uint32_t u = u0 + j * kThreads;
                    if (u < g.units_per_row) {
                        uint4 v = ld_stream16(src + (size_t)u * 16);
                        st_stream16(dst + (size_t)u * 16, process_unit(op, v, tab));
                    } else if (u - g.units_per_row < g.tail) {
                        size_t off = ((size_t)g.units_per_row * kPxPerUnit + (u - g.units_per_row)) *
                                     Op::kPixelBytes;
                        if constexpr (Op::kPixelBytes == 4) {
                            uint32_t v = *reinterpret_cast<const uint32_t *>(src + off);
                            *reinterpret_cast<uint32_t *>(dst + off) = op.px(v, tab);
                        } else {
                            uint2 v = *reinterpret_cast<const uint2 *>(src + off);
                            *reinterpret_cast<uint2 *>(dst + off) = op.px64(v, tab);
                        }
                    }
                }
            }
        }
    }
}

// 2-D tile path for the table-gather ops (4-byte pixels, rows 16-byte aligned): one CTA per tile of
// 64 pixels x 64 rows, a thread taking one 16-byte unit in each of four rows 16 apart (a warp reads
// two 256-byte row segments per access).  What a gather kernel pays for is the number of distinct
// table lines its SM touches at a time; natural video is coherent in two dimensions, so a square
// tile has a far smaller colour footprint than the 4096-pixel scan-line segment of the flattened
// path, and the gathers hit in L1.  One tile per CTA (no grid-stride loop) measured best on every
// content class (tools/microbench/tilecache.cu).  grid = (tile columns, tile rows, frames).
constexpr int kTileUnitsX = 16;                        // 16-byte units per tile row = 64 pixels
constexpr int kTileRowStep = kThreads / kTileUnitsX;   // rows between a thread's units
constexpr int kTileRows = kTileRowStep * kUnroll;      // 64

// Four pixels of a row as four [c0,c1,c2,x] words: one 16-byte access for 4-byte pixels, three 32-bit
// words split / re-joined with PRMT for 3-byte pixels (the fourth byte reads as 0 and is not stored).
template <int BPP>
__device__ __forceinline__ uint4 ld_unit(const uint8_t *p) {
    if constexpr (BPP == 4) {
        return ld_stream16(p);
    } else {
        const uint32_t *s = reinterpret_cast<const uint32_t *>(p);
        const uint32_t a = __ldcs(s), b = __ldcs(s + 1), c = __ldcs(s + 2);
        return make_uint4(a, __byte_perm(a, b, 0x4543u), __byte_perm(b, c, 0x4432u), __byte_perm(c, 0u, 0x4321u));
    }
}
template <int BPP>
__device__ __forceinline__ void st_unit(uint8_t *p, uint4 q) {
    if constexpr (BPP == 4) {
        st_stream16(p, q);
    } else {
        uint32_t *d = reinterpret_cast<uint32_t *>(p);
        __stcs(d, __byte_perm(q.x, q.y, 0x4210u));
        __stcs(d + 1, __byte_perm(q.y, q.z, 0x5421u));
        __stcs(d + 2, __byte_perm(q.z, q.w, 0x6542u));
    }
}

// Tile shape by pixel size.  4-byte pixels: 64 x 64.  With 3-byte pixels on either side a warp takes
// one 128-pixel row segment instead of two 64-pixel ones (tile 128 x 32): its 12-byte-stride
// accesses then touch 3 cache lines per instruction instead of 4 (hsvfilter RGB 77 -> 80 %,
// hsvdetector RGB -> RGBA 82 -> 88 % of the roofline on grad, + 3-4 points on noise).
template <int IN_BPP, int OUT_BPP>
struct TileShape {
    static constexpr int kUnitsX = (IN_BPP == 3 || OUT_BPP == 3) ? 32 : kTileUnitsX;
    static constexpr int kRowStep = kThreads / kUnitsX;
    static constexpr int kRows = kRowStep * kUnroll;
};

// IN_BPP / OUT_BPP = 4 or 3 bytes per pixel in memory; rows 16-byte (4-byte pixels) or 4-byte
// (3-byte pixels) aligned.
template <class Op, int IN_BPP = 4, int OUT_BPP = 4>
__global__ void __launch_bounds__(kThreads, 8) vf_map_tile_kernel(FrameSet fs, RowGeom g, Op op) {
    static_assert(Op::kPixelBytes == 4, "tile path: 8-bit pixels");
    __shared__ TabEntry tab[TableEntries<Op>::value];
    op.init(tab);
    constexpr int kTileUnitsX = TileShape<IN_BPP, OUT_BPP>::kUnitsX, kTileRowStep = TileShape<IN_BPP, OUT_BPP>::kRowStep,
                  kTileRows = TileShape<IN_BPP, OUT_BPP>::kRows;
    const uint32_t x = blockIdx.x * kTileUnitsX + threadIdx.x % kTileUnitsX;
    const uint32_t y0 = blockIdx.y * kTileRows + threadIdx.x / kTileUnitsX;
    const uint8_t *src = fs.in[blockIdx.z] + (size_t)x * (4 * IN_BPP);
    uint8_t *dst = fs.out[blockIdx.z] + (size_t)x * (4 * OUT_BPP);
    if (x < g.units_per_row) {
        uint4 v[kUnroll];
#pragma unroll
        for (int j = 0; j < kUnroll; j++) {
            const uint32_t y = y0 + j * kTileRowStep;
            if (y < g.rows) v[j] = ld_unit<IN_BPP>(src + (size_t)y * g.in_stride);
        }
#pragma unroll
        for (int j = 0; j < kUnroll; j++) {
            const uint32_t y = y0 + j * kTileRowStep;
            if (y < g.rows) st_unit<OUT_BPP>(dst + (size_t)y * g.out_stride, process_unit(op, v[j], tab));
        }
    } else if (x == g.units_per_row) {  // the width % 4 pixels at the end of each row
#pragma unroll 1
        for (int j = 0; j < kUnroll; j++) {
            const uint32_t y = y0 + j * kTileRowStep;
            if (y >= g.rows) break;
            for (uint32_t k = 0; k < g.tail; k++) {
                uint32_t w[2];
                ld_bytes<IN_BPP>(src + (size_t)y * g.in_stride + k * IN_BPP, w);
                w[0] = op.px(w[0], tab);
                st_bytes<OUT_BPP>(dst + (size_t)y * g.out_stride + k * OUT_BPP, w);
            }
        }
    }
}

// 3-byte pixels in (RGB / BGR), 16-byte aligned rows, width a multiple of 128 pixels: the same tile
// (128 pixels x 32 rows, a warp per 384-byte row segment, four segments 8 rows apart per warp), but a
// segment travels as 24 x 16 bytes — lanes 0-23, fully coalesced, full sectors — and is handed
// round through a warp-private piece of shared memory: lane l takes its four pixels from words
// 3l .. 3l+2 (3 is coprime with 32: no bank conflicts) and, for 3-byte output, puts them back the
// same way.  The 12-byte-stride accesses of vf_map_tile_kernel<Op, 3, …> touch all three lines of
// the segment with every instruction and their stores reach L2 as partial sectors (1.55 x the
// bytes); tools/microbench/rgb3_stage.cu, byte-identical outputs: RGB -> RGB 72 -> 82-90 %,
// RGB -> RGBA 84 -> 95-101 % of the roofline on smooth content, + 2-8 points on noise.  Safe in
// place: a warp has read its segments completely before it writes them.
// grid = (width / 128, ceil(rows / 32), frames).
template <class Op, int OUT_BPP>
__global__ void __launch_bounds__(kThreads, 8) vf_map_tile3_staged_kernel(FrameSet fs, RowGeom g, Op op) {
    static_assert(Op::kPixelBytes == 4, "tile path: 8-bit pixels");
    __shared__ TabEntry tab[TableEntries<Op>::value];
    __shared__ uint4 stage[kThreads / 32][kUnroll][24];
    op.init(tab);
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    constexpr int kRowStep = kThreads / 32;  // rows between a warp's segments
    const uint32_t y0 = blockIdx.y * (kRowStep * kUnroll) + warp;
    const uint8_t *src = fs.in[blockIdx.z] + (size_t)blockIdx.x * 384;
    uint8_t *dst = fs.out[blockIdx.z] + (size_t)blockIdx.x * (128 * OUT_BPP);
#pragma unroll
    for (int j = 0; j < kUnroll; j++) {
        const uint32_t y = y0 + j * kRowStep;
        if (y < g.rows && lane < 24) stage[warp][j][lane] = ld_stream16(src + (size_t)y * g.in_stride + lane * 16);
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < kUnroll; j++) {
        const uint32_t y = y0 + j * kRowStep;
        if (y >= g.rows) continue;  // warp-uniform
        uint32_t *s = reinterpret_cast<uint32_t *>(stage[warp][j]) + 3 * lane;
        const uint32_t a = s[0], b = s[1], c = s[2];
        const uint4 v = make_uint4(a, __byte_perm(a, b, 0x4543u), __byte_perm(b, c, 0x4432u), __byte_perm(c, 0u, 0x4321u));
        const uint4 q = process_unit(op, v, tab);
        if constexpr (OUT_BPP == 4) {
            st_stream16(dst + (size_t)y * g.out_stride + lane * 16, q);
        } else {
            s[0] = __byte_perm(q.x, q.y, 0x4210u);
            s[1] = __byte_perm(q.y, q.z, 0x5421u);
            s[2] = __byte_perm(q.z, q.w, 0x6542u);
        }
    }
    if constexpr (OUT_BPP == 3) {
        __syncwarp();
#pragma unroll
        for (int j = 0; j < kUnroll; j++) {
            const uint32_t y = y0 + j * kRowStep;
            if (y < g.rows && lane < 24) st_stream16(dst + (size_t)y * g.out_stride + lane * 16, stage[warp][j][lane]);
        }
    }
}

// 3-byte pixels (RGB / BGR), rows 4-byte aligned: a thread takes 4 pixels = three 32-bit
// words, splits them into four [c0,c1,c2,·] pixels with PRMT, and stores either three words
// (OUT_BPP == 3, hsvfilter) or one uint4 (OUT_BPP == 4, hsvdetector; rows 16-byte aligned).
template <class Op, int OUT_BPP>
__global__ void __launch_bounds__(kThreads) vf_map_vec3_kernel(FrameSet fs, RowGeom g, Op op) {
    __shared__ TabEntry tab[TableEntries<Op>::value];
    op.init(tab);
    const uint8_t *in = fs.in[blockIdx.z];
    uint8_t *out = fs.out[blockIdx.z];
    for (uint32_t row = blockIdx.y; row < g.rows; row += gridDim.y) {
        const uint8_t *src = in + (size_t)row * g.in_stride;
        uint8_t *dst = out + (size_t)row * g.out_stride;
        for (uint32_t seg = blockIdx.x; seg < g.tiles_per_row; seg += gridDim.x) {
            const uint32_t u = seg * kThreads + threadIdx.x;
            if (u < g.units_per_row) {
                const uint32_t *s = reinterpret_cast<const uint32_t *>(src) + (size_t)u * 3;
                const uint32_t a = __ldcs(s), b = __ldcs(s + 1), c = __ldcs(s + 2);
                const uint32_t q0 = op.px(a, tab);
                const uint32_t q1 = op.px(__byte_perm(a, b, 0x4543u), tab);
                const uint32_t q2 = op.px(__byte_perm(b, c, 0x4432u), tab);
                const uint32_t q3 = op.px(__byte_perm(c, 0u, 0x4321u), tab);
                if constexpr (OUT_BPP == 3) {
                    uint32_t *d = reinterpret_cast<uint32_t *>(dst) + (size_t)u * 3;
                    __stcs(d, __byte_perm(q0, q1, 0x4210u));
                    __stcs(d + 1, __byte_perm(q1, q2, 0x5421u));
                    __stcs(d + 2, __byte_perm(q2, q3, 0x6542u));
                } else {
                    st_stream16(dst + (size_t)u * 16, make_uint4(q0, q1, q2, q3));
                }
            } else if (u - g.units_per_row < g.tail) {
                const size_t pxi = (size_t)g.units_per_row * 4 + (u - g.units_per_row);
                uint32_t w[2];
                ld_bytes<3>(src + pxi * 3, w);
                w[0] = op.px(w[0], tab);
                st_bytes<OUT_BPP>(dst + pxi * OUT_BPP, w);
            }
        }
    }
}

// 4-byte pixels in, 3-byte pixels out (colorlut with conversion to RGB / BGR): a thread takes one
// 16-byte unit = 4 pixels and stores three 32-bit words; input rows 16-byte, output rows 4-byte aligned.
template <class Op>
__global__ void __launch_bounds__(kThreads) vf_map_vec43_kernel(FrameSet fs, RowGeom g, Op op) {
    __shared__ TabEntry tab[TableEntries<Op>::value];
    op.init(tab);
    const uint8_t *in = fs.in[blockIdx.z];
    uint8_t *out = fs.out[blockIdx.z];
    for (uint32_t row = blockIdx.y; row < g.rows; row += gridDim.y) {
        const uint8_t *src = in + (size_t)row * g.in_stride;
        uint8_t *dst = out + (size_t)row * g.out_stride;
        for (uint32_t seg = blockIdx.x; seg < g.tiles_per_row; seg += gridDim.x) {
            const uint32_t u = seg * kThreads + threadIdx.x;
            if (u < g.units_per_row) {
                const uint4 v = ld_stream16(src + (size_t)u * 16);
                const uint32_t q0 = op.px(v.x, tab), q1 = op.px(v.y, tab), q2 = op.px(v.z, tab),
                               q3 = op.px(v.w, tab);
                uint32_t *d = reinterpret_cast<uint32_t *>(dst) + (size_t)u * 3;
                __stcs(d, __byte_perm(q0, q1, 0x4210u));
                __stcs(d + 1, __byte_perm(q1, q2, 0x5421u));
                __stcs(d + 2, __byte_perm(q2, q3, 0x6542u));
            } else if (u - g.units_per_row < g.tail) {
                const size_t pxi = (size_t)g.units_per_row * 4 + (u - g.units_per_row);
                uint32_t w[2];
                w[0] = op.px(*reinterpret_cast<const uint32_t *>(src + pxi * 4), tab);
                w[1] = 0;
                st_bytes<3>(dst + pxi * 3, w);
            }
        }
    }
}

// Alignment-free path: one pixel per thread, byte accesses.  IN_BPP/OUT_BPP ∈ {3,4,8}.
// A 3-byte pixel is presented to the op as [b0,b1,b2,0]; only OUT_BPP bytes are stored.
// WORD: rows are 4-byte aligned and pixels are 4 bytes, so a pixel moves as one 32-bit access.
template <class Op, int IN_BPP, int OUT_BPP, bool WORD = false>
__global__ void __launch_bounds__(kThreads) vf_map_any_kernel(FrameSet fs, RowGeom g, Op op) {
    __shared__ TabEntry tab[TableEntries<Op>::value];
    op.init(tab);
    const uint8_t *in = fs.in[blockIdx.z];
    uint8_t *out = fs.out[blockIdx.z];
    for (uint32_t row = blockIdx.y; row < g.rows; row += gridDim.y) {
        for (uint32_t seg = blockIdx.x; seg < g.tiles_per_row; seg += gridDim.x) {
            uint32_t u = seg * kThreads + threadIdx.x;
            if (u >= g.units_per_row) continue;
            const uint8_t *src = in + (size_t)row * g.in_stride + (size_t)u * IN_BPP;
            uint8_t *dst = out + (size_t)row * g.out_stride + (size_t)u * OUT_BPP;
            if constexpr (WORD) {
                *reinterpret_cast<uint32_t *>(dst) =
                    op.px(*reinterpret_cast<const uint32_t *>(src), tab);
                continue;
            }
            uint32_t w[2];
            ld_bytes<IN_BPP>(src, w);
            if constexpr (Op::kPixelBytes == 4) {
                w[0] = op.px(w[0], tab);
            } else {
                uint2 o = op.px64(make_uint2(w[0], w[1]), tab);
                w[0] = o.x, w[1] = o.y;
            }
            st_bytes<OUT_BPP>(dst, w);
        }
    }
}

// ---------------------------------------------------------------------------
// launch plumbing
// ---------------------------------------------------------------------------
// every row of every frame starts on an in_align / out_align byte boundary
inline bool rows_aligned(const FrameSet &fs, int n, const Geom &g, bool flat, uintptr_t in_align,
                         uintptr_t out_align) {
    for (int i = 0; i < n; i++) {
        if (((uintptr_t)fs.in[i] & (in_align - 1)) || ((uintptr_t)fs.out[i] & (out_align - 1)))
            return false;
    }
    if (!flat && (((uintptr_t)g.in_stride & (in_align - 1)) ||
                  ((uintptr_t)g.out_stride & (out_align - 1))))
        return false;
    return true;
}

// grid = (segments, row groups, frames): up to VF_CTAS CTAs per SM in total, every loop
// grid-stride (measured best on B200: 4 x 16 B per thread per tile, 64 CTAs per SM).
inline dim3 grid_for(uint32_t tiles_per_row, uint32_t rows, int n_frames) {
    const uint64_t cap = (uint64_t)kSMs * VF_CTAS;
    uint64_t per_frame = std::max<uint64_t>(kSMs, cap / (uint64_t)std::max(1, n_frames));
    uint32_t gx = (uint32_t)std::min<uint64_t>(tiles_per_row, per_frame);
    uint32_t gy = (uint32_t)std::min<uint64_t>(rows, std::max<uint64_t>(1, per_frame / gx));
    gy = std::min<uint32_t>(gy, 65535u);
    return dim3(std::max(1u, gx), std::max(1u, gy), (unsigned)n_frames);
}

// Launches op over the frames; picks vec / any path from alignment.
// in_bpp / out_bpp: bytes per pixel in memory (3, 4 or 8).
// VEC_ONLY: instantiate just the 16-byte kernel and return cudaErrorNotSupported for frames that
// would need another path (the caller then composes the result from other launches).
template <class Op, bool VEC_ONLY = false>
static cudaError_t launch_map(cudaStream_t stream, const FrameSet &fs, int n, const Geom &g,
                              int in_bpp, int out_bpp, const Op &op, uint64_t *launches) {
    if (n <= 0 || g.width == 0 || g.height == 0) return cudaSuccess;
    RowGeom rg;
    rg.in_stride = g.in_stride;
    rg.out_stride = g.out_stride;
    const bool same_bpp_vec = (in_bpp == Op::kPixelBytes && out_bpp == Op::kPixelBytes);
    const uint64_t row_bytes_in = (uint64_t)g.width * in_bpp;
    const uint64_t row_bytes_out = (uint64_t)g.width * out_bpp;
    // contiguous frames become a single long row
    bool flat = (uint64_t)g.in_stride == row_bytes_in && (uint64_t)g.out_stride == row_bytes_out &&
                (uint64_t)g.width * g.height < (1ull << 31);
    uint64_t width = flat ? (uint64_t)g.width * g.height : g.width;
    uint32_t rows = flat ? 1 : g.height;
    rg.rows = rows;
    if constexpr (Tiled<Op>::value) {
        if (in_bpp == 3 && (out_bpp == 3 || out_bpp == 4) && g.width % 128 == 0 &&
            rows_aligned(fs, n, g, false, 16, 16) && g.height <= 65535u * 32u) {
            constexpr uint32_t tr = (kThreads / 32) * kUnroll;  // rows per CTA: 8 warps x 4 segments
            rg.rows = g.height;
            rg.units_per_row = g.width / 4;
            rg.tail = 0;
            rg.tiles_per_row = g.width / 128;
            const dim3 grid(rg.tiles_per_row, (g.height + tr - 1) / tr, (unsigned)n);
            if (out_bpp == 3)
                vf_map_tile3_staged_kernel<Op, 3><<<grid, kThreads, 0, stream>>>(fs, rg, op);
            else
                vf_map_tile3_staged_kernel<Op, 4><<<grid, kThreads, 0, stream>>>(fs, rg, op);
            if (launches) *launches += 1;
            return cudaGetLastError();
        }
        const bool bpp_ok = (in_bpp == 3 || in_bpp == 4) && (out_bpp == 3 || out_bpp == 4);
        if (bpp_ok && rows_aligned(fs, n, g, false, in_bpp == 4 ? 16 : 4, out_bpp == 4 ? 16 : 4) &&
            g.height <= 65535u * 32u) {
            const bool any3 = in_bpp == 3 || out_bpp == 3;
            const uint32_t ux = any3 ? TileShape<3, 3>::kUnitsX : TileShape<4, 4>::kUnitsX;
            const uint32_t tr = any3 ? TileShape<3, 3>::kRows : TileShape<4, 4>::kRows;
            rg.rows = g.height;
            rg.units_per_row = g.width / 4;
            rg.tail = g.width % 4;
            rg.tiles_per_row = (rg.units_per_row + (rg.tail ? 1 : 0) + ux - 1) / ux;
            const dim3 grid(rg.tiles_per_row, (g.height + tr - 1) / tr, (unsigned)n);
            if (in_bpp == 4 && out_bpp == 4)
                vf_map_tile_kernel<Op, 4, 4><<<grid, kThreads, 0, stream>>>(fs, rg, op);
            else if (in_bpp == 3 && out_bpp == 3)
                vf_map_tile_kernel<Op, 3, 3><<<grid, kThreads, 0, stream>>>(fs, rg, op);
            else if (in_bpp == 3)
                vf_map_tile_kernel<Op, 3, 4><<<grid, kThreads, 0, stream>>>(fs, rg, op);
            else
                vf_map_tile_kernel<Op, 4, 3><<<grid, kThreads, 0, stream>>>(fs, rg, op);
            if (launches) *launches += 1;
            return cudaGetLastError();
        }
    }
    if (same_bpp_vec && rows_aligned(fs, n, g, flat, 16, 16)) {
        const uint32_t ppu = 16 / Op::kPixelBytes;
        const uint32_t tile = kThreads * kUnroll;
        rg.units_per_row = (uint32_t)(width / ppu);
        rg.tail = (uint32_t)(width % ppu);
        rg.tiles_per_row = (rg.units_per_row + rg.tail + tile - 1) / tile;
        vf_map_vec_kernel<Op><<<grid_for(rg.tiles_per_row, rows, n), kThreads, 0, stream>>>(fs, rg, op);
    } else if constexpr (VEC_ONLY) {
        return cudaErrorNotSupported;
    } else if (in_bpp == 4 && out_bpp == 3 && Op::kPixelBytes == 4 && rows_aligned(fs, n, g, flat, 16, 4)) {
        rg.units_per_row = (uint32_t)(width / 4);
        rg.tail = (uint32_t)(width % 4);
        rg.tiles_per_row = (rg.units_per_row + rg.tail + kThreads - 1) / kThreads;
        if constexpr (Op::kPixelBytes == 4)
            vf_map_vec43_kernel<Op><<<grid_for(rg.tiles_per_row, rows, n), kThreads, 0, stream>>>(fs, rg, op);
    } else if (in_bpp == 3 && Op::kPixelBytes == 4 &&
               rows_aligned(fs, n, g, flat, 4, out_bpp == 3 ? 4 : 16)) {
        rg.units_per_row = (uint32_t)(width / 4);
        rg.tail = (uint32_t)(width % 4);
        rg.tiles_per_row = (rg.units_per_row + rg.tail + kThreads - 1) / kThreads;
        dim3 grid = grid_for(rg.tiles_per_row, rows, n);
        if constexpr (Op::kPixelBytes == 4) {
            if (out_bpp == 3)
                vf_map_vec3_kernel<Op, 3><<<grid, kThreads, 0, stream>>>(fs, rg, op);
            else
                vf_map_vec3_kernel<Op, 4><<<grid, kThreads, 0, stream>>>(fs, rg, op);
        }
    } else {
        rg.units_per_row = (uint32_t)width;
        rg.tail = 0;
        rg.tiles_per_row = (rg.units_per_row + kThreads - 1) / kThreads;
        dim3 grid = grid_for(rg.tiles_per_row, rows, n);
        if (in_bpp == 4 && out_bpp == 4) {
            if constexpr (Op::kPixelBytes == 4) {
                if (rows_aligned(fs, n, g, flat, 4, 4))
                    vf_map_any_kernel<Op, 4, 4, true><<<grid, kThreads, 0, stream>>>(fs, rg, op);
                else
                    vf_map_any_kernel<Op, 4, 4><<<grid, kThreads, 0, stream>>>(fs, rg, op);
            }
        } else if (in_bpp == 3 && out_bpp == 3) {
            if constexpr (Op::kPixelBytes == 4)
                vf_map_any_kernel<Op, 3, 3><<<grid, kThreads, 0, stream>>>(fs, rg, op);
        } else if (in_bpp == 3 && out_bpp == 4) {
            if constexpr (Op::kPixelBytes == 4)
                vf_map_any_kernel<Op, 3, 4><<<grid, kThreads, 0, stream>>>(fs, rg, op);
        } else if (in_bpp == 4 && out_bpp == 3) {
            if constexpr (Op::kPixelBytes == 4)
                vf_map_any_kernel<Op, 4, 3><<<grid, kThreads, 0, stream>>>(fs, rg, op);
        } else if (in_bpp == 8 && out_bpp == 8) {
            if constexpr (Op::kPixelBytes == 8)
                vf_map_any_kernel<Op, 8, 8><<<grid, kThreads, 0, stream>>>(fs, rg, op);
        } else {
            return cudaErrorInvalidValue;
        }
    }
    if (launches) *launches += 1;
    return cudaGetLastError();
}

inline HsvFilterParams make_filter_params(const HsvFilterArgs &a) {
    HsvFilterParams p;
    p.hue_shift = a.hue_shift;
    p.sat_mul = a.sat_mul;
    p.sat_off = a.sat_off;
    p.val_mul = a.val_mul;
    p.val_off = a.val_off;
    return p;
}

// AngleKind of a per-frame constant hue shift / offset (generic for NaN, inf, |d| > 360)
inline int angle_kind(float d) {
    if (d == 0.0f) return kAngleZero;
    if (d > 0.0f && d <= 360.0f) return kAngleNonNeg;
    if (d < 0.0f && d >= -360.0f) return kAngleNeg;
    return kAngleGeneric;
}

// The four colour-byte placements of the ten packed formats;
// 3-byte pixels are presented as [b0,b1,b2,0] and fall in the first or third.
#define VF_FOR_LAYOUT(lay, CALL)                                   \
    if (lay.r == 0 && lay.g == 1 && lay.b == 2) { CALL(0, 1, 2) }  \
    else if (lay.r == 1 && lay.g == 2 && lay.b == 3) { CALL(1, 2, 3) } \
    else if (lay.r == 2 && lay.g == 1 && lay.b == 0) { CALL(2, 1, 0) } \
    else if (lay.r == 3 && lay.g == 2 && lay.b == 1) { CALL(3, 2, 1) }

inline LutArgs make_lut_args(const DeviceLut &lut) {
    LutArgs L;
    L.lut3d = lut.lut3d;
    L.lut_rx = lut.lut3d_rx;
    L.lut_rg = lut.lut3d_rg;
    L.lut_baked = lut.lut3d_baked;
    L.lut1d = lut.lut1d;
    L.lut3d_d = lut.lut3d_d;
    L.k16_hi = lut.k16_hi;
    L.k16_lo = lut.k16_lo;
    L.bias_bits = VF_MAGIC_BITS;
    L.neg_zero = -0.0f;
    L.n = lut.size;
    L.sy = lut.size + 1;
    L.sz = (lut.size + 1) * (lut.size + 1);
    L.sm1 = (float)lut.size - 1.0f;  // imp.rs:411, 438
    for (int c = 0; c < 3; c++) L.scale[c] = lut.scale[c], L.offset[c] = lut.offset[c];
    return L;
}

// resolved path: 0 direct, 1 R-resampled, 2 1D, 3 RG-resampled, 4 baked, 5 tetrahedral, 6 nearest,
// 7 the 16-bit fast op (ColorLut64Op; needs its table; lut_path auto, or 4 to pin it).
// Auto, 8-bit: the table baked to native resolution when it exists for this interpolation (the
// ABI builds it on first use), else the RG-resampled, R-resampled and direct kernels in that order.
inline int resolve_lut_path(const DeviceLut &lut, int bits, int math_mode, int lut_path,
                            int interp = kInterpTrilinear) {
    if (lut.kind == 1) return 2;
    const bool baked_ok = bits == 8 && lut.lut3d_baked && lut.baked_interp == interp &&
                          (lut_path == kLutAuto || lut_path == kLutBaked);
    if (interp != kInterpTrilinear)  // no resampled tables: the weights are not separable
        return baked_ok ? 4 : interp == kInterpTetrahedral ? 5 : 6;
    if (bits == 16 && (lut_path == kLutAuto || lut_path == kLutBaked) && lut.lut3d_d && lut.coords16_ok && lut.identity_domain &&
        math_mode != kMathPlain)
        return 7;
    if (bits != 8 || lut_path == kLutDirect) return 0;
    if (baked_ok) return 4;
    const bool fast = math_mode != kMathPlain;
    if (lut_path == kLutResampledR) return lut.lut3d_rx ? 1 : 0;
    if (lut.lut3d_rg && fast) return 3;  // auto / RG / baked table unavailable
    return lut.lut3d_rx ? 1 : 0;
}

}  // namespace vf
