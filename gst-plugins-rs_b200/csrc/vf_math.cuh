// vf_math.cuh — per-pixel f32 arithmetic of colorlut / hsvfilter / hsvdetector
// for sm_100a, bit-compatible with the reference's Rust.
//
// Two families:
//   *_plain : literal translation (IEEE `/`, fmodf, branch ladder).  Obviously
//             equal to the reference; used as the in-library cross-check
//             ("hsv.math"=1) and for parameter ranges the fast path excludes.
//   *_fast  : the same values from shorter instruction sequences.  Every
//             shortcut is either proven exhaustively on the CPU
//             (tests/cpu_proofs/verify_math.c), or is an exact-arithmetic identity
//             argued in the comment beside it, and the whole RGB→HSV and
//             HSV→RGB maps are checked exhaustively over all 2^24 inputs on
//             the GPU against the oracle (tests/test_gpu_hsv.py).
//
// Build with -fmad=false: Rust never contracts a*b+c.  Where an
// FMA is wanted it is written explicitly (__fmaf_rn).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vf {

// ---------------------------------------------------------------------------
// constants
// ---------------------------------------------------------------------------
#ifndef VF_DIV4
#define VF_DIV4 1  // 0 = keep the Newton refinement of the reciprocal (A/B builds)
#endif
#define VF_MAGIC 8388608.0f          // 2^23: float whose ulp is 1
#define VF_MAGIC_BITS 0x4B000000u
#define VF_K255_HI 0x1.010102p-8f    // RN(1/255)
#define VF_K255_LO -0x1.fdfdfep-33f  // RN(1/255 - K255_HI)
#define VF_K65535_HI 0x1.0001p-16f   // RN(1/65535)
#define VF_K65535_LO 0x1.0001p-48f   // RN(1/65535 - K65535_HI)
#define VF_K60_HI 0x1.111112p-6f      // RN(1/60)
#define VF_K60_LO -0x1.dddddep-31f    // RN(1/60 - K60_HI)
#define VF_FLT_MIN 1.17549435e-38f

struct Hsv {
    float h, s, v;
};

// ---------------------------------------------------------------------------
// small exact helpers
// ---------------------------------------------------------------------------

// prmt.b32 with a register selector (no `& 0x7777`: callers keep bit 3 of every nibble clear).
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// byte `idx` (0..3) of px as an integer-valued float, without I2F:
// PRMT builds the bit pattern of 2^23 + byte, one FADD removes the bias (exact).
__device__ __forceinline__ float byte_to_float(uint32_t px, uint32_t idx) {
    return __uint_as_float(__byte_perm(px, VF_MAGIC_BITS, 0x7440u | idx)) - VF_MAGIC;
}

// RN(c / 255) for integer-valued c in [0,255]: 1/255 split into hi+lo so that the
// single rounding of the FMA sees c/255 to ~2^-48 (proven for all 256 inputs).
__device__ __forceinline__ float div255_exact(float c) {
    return __fmaf_rn(c, VF_K255_HI, __fmul_rn(c, VF_K255_LO));
}

// RN(c / 65535) for integer-valued c in [0,65535] (proven for all 65536 inputs).
__device__ __forceinline__ float div65535_exact(float c) {
    return __fmaf_rn(c, VF_K65535_HI, __fmul_rn(c, VF_K65535_LO));
}

// RN(h / 60) with the same hi/lo split (proven for every float in [2^-20, 720]; below
// 2^-20 any result in [0, 2^-19] leads to the same pixel, see to_rgb_fast).
__device__ __forceinline__ float div60_exact(float h) {
    return __fmaf_rn(h, VF_K60_HI, __fmul_rn(h, VF_K60_LO));
}

// RN(a / b) for the operand pairs RGB→HSV produces (differences of k/255 quotients:
// b in [1/255, 1], or FLT_MIN with a == 0): MUFU.RCP + one residual correction.  This is
// NOT a general IEEE division; its exactness on this domain is established by enumeration.
__device__ __forceinline__ float div_exact(float a, float b) {
    float y0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y0) : "f"(b));
#if VF_DIV4
    // One residual correction straight from the MUFU reciprocal (|rel err| <= 2^-23): exact
    // for every operand pair RGB->HSV can produce — checked over all 2^24 inputs on the device
    // against the oracle's (h,s,v) floats (tests/test_gpu_hsv.py::test_from_rgb_floats_exhaustive).
    float q0 = __fmul_rn(a, y0);
    float r = __fmaf_rn(-b, q0, a);
    return __fmaf_rn(r, y0, q0);
#else
    float e = __fmaf_rn(-b, y0, 1.0f);
    float y = __fmaf_rn(y0, e, y0);
    float q0 = __fmul_rn(a, y);
    float r = __fmaf_rn(-b, q0, a);
    return __fmaf_rn(r, y, q0);
#endif
}

// Rust inherent f32::clamp: NaN propagates.
__device__ __forceinline__ float rs_clamp(float v, float lo, float hi) {
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    return v;
}

// Rust `f32 as u8`: truncate, saturate, NaN → 0 (cvt.rzi.sat does exactly this).
__device__ __forceinline__ uint32_t rs_as_u8(float v) {
    uint32_t r;
    asm("cvt.rzi.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

__device__ __forceinline__ uint32_t rs_as_u16(float v) {
    uint32_t r;
    asm("cvt.rzi.sat.u16.f32 %0, %1;" : "=r"(r) : "f"(v));
    return r;
}

// ---------------------------------------------------------------------------
// RGB → HSV   (hsvutils.rs:44-84; from_bgr :88-128 is the same with r/b swapped)
// r8,g8,b8 are integer-valued floats in [0,255].
// ---------------------------------------------------------------------------

__device__ __forceinline__ Hsv from_rgb_plain(float r8, float g8, float b8) {
    float r = r8 / 255.0f, g = g8 / 255.0f, b = b8 / 255.0f;
    float value = fmaxf(r8, fmaxf(g8, b8)) / 255.0f;  // integer max, then divide (:49-53)
    float chroma = value - fminf(r8, fminf(g8, b8)) / 255.0f;
    float hue;
    if (chroma == 0.0f)
        hue = 0.0f;
    else if (fabsf(value - r) < 0.00001f)
        hue = 60.0f * ((g - b) / chroma);
    else if (fabsf(value - g) < 0.00001f)
        hue = 60.0f * (2.0f + ((b - r) / chroma));
    else if (fabsf(value - b) < 0.00001f)
        hue = 60.0f * (4.0f + ((r - g) / chroma));
    else
        hue = 0.0f;
    if (hue < 0.0f) hue += 360.0f;
    float sat = value == 0.0f ? 0.0f : chroma / value;
    Hsv o;
    o.h = fmodf(hue, 360.0f);
    o.s = rs_clamp(sat, 0.0f, 1.0f);
    o.v = rs_clamp(value, 0.0f, 1.0f);
    return o;
}

// Same values, ~40 instructions, no branches:
//  * c/255 by div255_exact; division is monotone, so max/min of the quotients equal
//    the quotients of the integer max/min.
//  * |value - x| < 1e-5 ⇔ x == value (distinct k/255 differ by ≥ 1/255).
//  * the three hue arms become one:  60 * (off + num/chroma) with off ∈ {0,2,4};
//    0 + q is exact and q is never -0 (num = a - b is +0 when a == b).
//  * chroma == 0 ⇒ num == 0, so dividing by max(chroma, FLT_MIN) yields the 0 the
//    reference assigns; likewise value == 0 ⇒ chroma == 0 for the saturation.
//  * hue ends in [0,360) so `% 360` is the identity; sat, value are already in [0,1].
__device__ __forceinline__ Hsv from_rgb_fast(float r8, float g8, float b8) {
    float r = div255_exact(r8), g = div255_exact(g8), b = div255_exact(b8);
    float value = fmaxf(r, fmaxf(g, b));
    float chroma = value - fminf(r, fminf(g, b));
    bool pr = (r == value), pg = (g == value);
    float na = pr ? g : (pg ? b : r);
    float nb = pr ? b : (pg ? r : g);
    float off = pr ? 0.0f : (pg ? 2.0f : 4.0f);
    float q = div_exact(na - nb, fmaxf(chroma, VF_FLT_MIN));
    float hue = 60.0f * (off + q);
    if (hue < 0.0f) hue += 360.0f;
    Hsv o;
    o.h = hue;
    o.s = div_exact(chroma, fmaxf(value, VF_FLT_MIN));
    o.v = value;
    return o;
}


// Same as from_rgb_fast with the three-way arm selection done by predicated FADDs on
// the FMA pipe instead of FSELs on the (half-rate) ALU pipe, and the zero-denominator
// guards as exact adds:  d + FLT_MIN == d for every d >= 1/255, and == FLT_MIN for d == 0.
// Returns hue in degrees in [0,360), s, v.
__device__ __forceinline__ Hsv from_rgb_fast2(float r8, float g8, float b8) {
    float r = div255_exact(r8), g = div255_exact(g8), b = div255_exact(b8);
    float value = fmaxf(r, fmaxf(g, b));
    float chroma = value - fminf(r, fminf(g, b));
    float dc = chroma + VF_FLT_MIN;
    float hue;
    asm("{\n\t"
        ".reg .pred pr, pg, pn;\n\t"
        ".reg .f32 num, y0, e, y, q0, rr, q, hp;\n\t"
        "setp.eq.f32 pr, %1, %4;\n\t"
        "setp.eq.f32 pg, %2, %4;\n\t"
        "sub.rn.f32 num, %1, %2;\n\t"          // B max: r - g
        "@pg sub.rn.f32 num, %3, %1;\n\t"      // G max: b - r
        "@pr sub.rn.f32 num, %2, %3;\n\t"      // R max: g - b (highest priority)
        "rcp.approx.ftz.f32 y0, %5;\n\t"
#if VF_DIV4
        "mov.f32 y, y0;\n\t"
#else
        "neg.f32 e, %5;\n\t"
        "fma.rn.f32 e, e, y0, 0f3F800000;\n\t"
        "fma.rn.f32 y, y0, e, y0;\n\t"
#endif
        "mul.rn.f32 q0, num, y;\n\t"
        "neg.f32 rr, %5;\n\t"
        "fma.rn.f32 rr, rr, q0, num;\n\t"
        "fma.rn.f32 q, rr, y, q0;\n\t"
        "add.rn.f32 hp, q, 0f40800000;\n\t"     // 4 + q
        "@pg add.rn.f32 hp, q, 0f40000000;\n\t" // 2 + q
        "@pr mov.f32 hp, q;\n\t"                // q (the reference has no 0 + here)
        "mul.rn.f32 %0, hp, 0f42700000;\n\t"    // * 60
        "setp.lt.f32 pn, %0, 0f00000000;\n\t"
        "@pn add.rn.f32 %0, %0, 0f43B40000;\n\t" // + 360
        "}"
        : "=f"(hue)
        : "f"(r), "f"(g), "f"(b), "f"(value), "f"(dc));
    Hsv o;
    o.h = hue;
    o.s = div_exact(chroma, value + VF_FLT_MIN);
    o.v = value;
    return o;
}

// ---------------------------------------------------------------------------
// HSV → RGB   (hsvutils.rs:132-163)
// ---------------------------------------------------------------------------

// Literal translation.  Returns r | g<<8 | b<<16.
__device__ __forceinline__ uint32_t to_rgb_plain(Hsv in) {
    float c = in.v * in.s;
    float hp = in.h / 60.0f;
    float x = c * (1.0f - fabsf(fmodf(hp, 2.0f) - 1.0f));
    float p0, p1, p2;
    if (hp < 0.0f) {
        p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;
    } else if (hp <= 1.0f) {
        p0 = c, p1 = x, p2 = 0.0f;
    } else if (hp <= 2.0f) {
        p0 = x, p1 = c, p2 = 0.0f;
    } else if (hp <= 3.0f) {
        p0 = 0.0f, p1 = c, p2 = x;
    } else if (hp <= 4.0f) {
        p0 = 0.0f, p1 = x, p2 = c;
    } else if (hp <= 5.0f) {
        p0 = x, p1 = 0.0f, p2 = c;
    } else if (hp <= 6.0f) {
        p0 = c, p1 = 0.0f, p2 = x;
    } else {
        p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;
    }
    float m = in.v - c;
    uint32_t r = rs_as_u8(rs_clamp((p0 + m) * 255.0f, 0.0f, 255.0f));
    uint32_t g = rs_as_u8(rs_clamp((p1 + m) * 255.0f, 0.0f, 255.0f));
    uint32_t b = rs_as_u8(rs_clamp((p2 + m) * 255.0f, 0.0f, 255.0f));
    return r | (g << 8) | (b << 16);
}

// Sector table entry for to_rgb_fast: `center` ∈ {1,3,5} is the odd integer the
// sector's triangle wave is centred on, `sel` the PRMT selector that assembles the
// output pixel from bytes {0: c+m, 1: x+m, 2: m} and the original pixel (4..7).
struct SectorEntry {
    float center;
    uint32_t sel;
};

// Requires h in [0,360] or NaN, and s, v in [0,1] (guaranteed by the filter step).
// `tab` is the 8-entry sector table (shared memory), `orig` the source pixel whose
// non-colour byte the selector may pick.  Returns the finished output pixel.
//  * k = ceil(h/60) ∈ 0..6 (7 for NaN) replaces the `<=` ladder: arm i of the
//    reference is exactly hp ∈ (i-1, i]; arms 0 and 1 are the same assignment.
//  * fmod(hp,2) - 1 = hp - center(k) in exact arithmetic except at hp ∈ {2,4,6} where
//    the reference gets |0 - 1| and we get |hp - (hp-1)| — both 1.  hp - 2j is exact
//    (Sterbenz), so one rounded subtraction of center equals the reference's two.
//  * with s,v ∈ [0,1]: c = RN(v*s) ≤ v, m = RN(v-c) ∈ [0,1], p+m ∈ [0,1], so the
//    reference's clamp(…,0,255) is the identity and `as u8` is floor: FADD.RM with
//    2^23 leaves floor(t) in the low mantissa byte.
//  * NaN hue: the reference's ladder falls to (0,0,0) → all channels m; entry 7.
//  * h < 2^-20 (outside div60_exact's proven range): hp ∈ [0, 2^-19], k ∈ {0,1},
//    t = |hp - 1| = 1 after rounding either way — same pixel as the exact quotient.
// FINITE_H: the caller guarantees h is not NaN, so k can come from F2I.CEIL (XU pipe)
// instead of FADD.RP + mask.
template <bool FINITE_H>
__device__ __forceinline__ uint32_t to_rgb_fast(Hsv in, const SectorEntry *tab, uint32_t orig) {
    float c = __fmul_rn(in.v, in.s);
    float hp = div60_exact(in.h);
    uint32_t k;
    if (FINITE_H)
        k = (uint32_t)__float2int_ru(hp);  // 0..6
    else
        k = __float_as_uint(__fadd_ru(hp, VF_MAGIC)) & 7u;  // low bits of 2^23 + ceil(hp); NaN → 7
    SectorEntry e = tab[k];
    float t = fabsf(hp - e.center);
    float x = __fmul_rn(c, 1.0f - t);
    float m = in.v - c;
    uint32_t A = __float_as_uint(__fadd_rd(__fmul_rn(c + m, 255.0f), VF_MAGIC));
    uint32_t B = __float_as_uint(__fadd_rd(__fmul_rn(x + m, 255.0f), VF_MAGIC));
    uint32_t C = __float_as_uint(__fadd_rd(__fmul_rn(m, 255.0f), VF_MAGIC));
    uint32_t ab = __byte_perm(A, B, 0x0040u);    // [A0, B0, ., .]
    uint32_t abc = __byte_perm(ab, C, 0x0410u);  // [A0, B0, C0, .]
    return prmt(abc, orig, e.sel);
}

// Angle-offset kinds, chosen on the host from the (per-frame constant) hue shift / offset d,
// for h in [0,360):
//   kAngleGeneric : any d (incl. NaN/inf)     → the reference's fmodf sequence
//   kAngleNonNeg  : 0 < d <= 360, u in (0,720] → fmod(u,360) = u - 360 iff u >= 360 (exact, Sterbenz)
//   kAngleNeg     : -360 <= d < 0, u in [-360,360) → only the reference's `if u < 0 { u += 360 }`
//   kAngleZero    : d == ±0                    → u = h, nothing to do
// (u = 720 → 360 instead of 0 and u = -360 → +0 instead of -0: hue 360 and 0 select different
// ladder arms with x = 0, i.e. the same pixel; for the detector |360-180| = |0-180|.)
enum AngleKind { kAngleGeneric = 0, kAngleNonNeg = 1, kAngleNeg = 2, kAngleZero = 3 };

template <int KIND>
__device__ __forceinline__ float add_angle(float h, float d) {
    if (KIND == kAngleZero) return h;
    float u = h + d;
    if (KIND == kAngleNonNeg) {
        asm("{\n\t.reg .pred p;\n\tsetp.ge.f32 p, %0, 0f43B40000;\n\t"
            "@p add.rn.f32 %0, %0, 0fC3B40000;\n\t}"
            : "+f"(u));
    } else if (KIND == kAngleNeg) {
        asm("{\n\t.reg .pred q;\n\tsetp.lt.f32 q, %0, 0f00000000;\n\t"
            "@q add.rn.f32 %0, %0, 0f43B40000;\n\t}"
            : "+f"(u));
    }
    return u;
}

// ---------------------------------------------------------------------------
// hsvfilter adjust (hsvfilter/imp.rs:102-115)
// ---------------------------------------------------------------------------

struct HsvFilterParams {
    float hue_shift, sat_mul, sat_off, val_mul, val_off;
};

__device__ __forceinline__ Hsv hsv_adjust_plain(Hsv a, const HsvFilterParams &p) {
    Hsv o;
    o.h = fmodf(a.h + p.hue_shift, 360.0f);
    if (o.h < 0.0f) o.h += 360.0f;
    o.s = fminf(fmaxf(p.sat_mul * a.s + p.sat_off, 0.0f), 1.0f);  // Clamp trait: NaN → 0
    o.v = fminf(fmaxf(p.val_mul * a.v + p.val_off, 0.0f), 1.0f);
    return o;
}

// add.sat clamps to [0,1] and maps NaN to 0, like max(0).min(1) of the Clamp trait.
template <int KIND>
__device__ __forceinline__ Hsv hsv_adjust_fast(Hsv a, const HsvFilterParams &p) {
    Hsv o;
    if (KIND == kAngleGeneric) {
        o.h = fmodf(a.h + p.hue_shift, 360.0f);
        if (o.h < 0.0f) o.h += 360.0f;
    } else {
        o.h = add_angle<KIND>(a.h, p.hue_shift);
    }
    o.s = __saturatef(__fadd_rn(__fmul_rn(p.sat_mul, a.s), p.sat_off));
    o.v = __saturatef(__fadd_rn(__fmul_rn(p.val_mul, a.v), p.val_off));
    return o;
}

// ---------------------------------------------------------------------------
// hsvdetector predicate (hsvdetector/imp.rs:141-153)
// ---------------------------------------------------------------------------

struct HsvDetectParams {
    float hue_off;  // 180 - hue_ref, computed once per frame as the reference does per pixel
    float hue_var, sat_ref, sat_var, val_ref, val_var;
};

__device__ __forceinline__ bool hsv_match_plain(Hsv a, const HsvDetectParams &p) {
    float sh = a.h + p.hue_off;
    if (sh < 0.0f) sh += 360.0f;
    sh = fmodf(sh, 360.0f);
    return fabsf(sh - 180.0f) <= p.hue_var && fabsf(a.s - p.sat_ref) <= p.sat_var &&
           fabsf(a.v - p.val_ref) <= p.val_var;
}

// imp.rs:141-148 computes (h + off), `+= 360 if < 0`, then `% 360`: with |off| <= 360 that is
// add_angle (for kAngleNeg the sum may round to exactly 360, where |360-180| = |0-180|).
template <int KIND>
__device__ __forceinline__ bool hsv_match_fast(Hsv a, const HsvDetectParams &p) {
    float sh;
    if (KIND == kAngleGeneric) {
        sh = a.h + p.hue_off;
        if (sh < 0.0f) sh += 360.0f;
        sh = fmodf(sh, 360.0f);
    } else {
        sh = add_angle<KIND>(a.h, p.hue_off);
    }
    return fabsf(sh - 180.0f) <= p.hue_var && fabsf(a.s - p.sat_ref) <= p.sat_var &&
           fabsf(a.v - p.val_ref) <= p.val_var;
}

// ---------------------------------------------------------------------------
// colorlut helpers (colorlut/imp.rs:471-543)
// ---------------------------------------------------------------------------

// float_to_u8 / float_to_u16 (imp.rs:537-543):  (v.clamp(0,1) * MAX).round() as uN
//  * __saturatef clamps and maps NaN → 0; the reference's NaN also ends as 0.
//  * y = w*MAX ≥ 0, so round-half-away = floor(y + 0.5).  y + 0.5 is rounded toward
//    zero (never reaches the next integer unless the exact sum does), then FADD.RM
//    with 2^23 leaves floor() in the low mantissa bits.
template <int BITS>
__device__ __forceinline__ uint32_t unit_to_code(float v) {
    const float mx = BITS == 8 ? 255.0f : 65535.0f;
    float y = __fmul_rn(__saturatef(v), mx);
    float f = __fadd_rd(__fadd_rz(y, 0.5f), VF_MAGIC);
    return __float_as_uint(f) & (BITS == 8 ? 0xFFu : 0xFFFFu);
}

// Same, returning the raw bits of 2^23 + code (callers pick the low bytes with PRMT).
// UNIT: v is known to be finite and (up to rounding) within [0,1], so the clamp is skipped;
// a last-ulp excursion still floors to the clamped code.
template <int BITS, bool UNIT = false>
__device__ __forceinline__ uint32_t unit_to_code_bits(float v) {
    const float mx = BITS == 8 ? 255.0f : 65535.0f;
    float y = __fmul_rn(UNIT ? v : __saturatef(v), mx);
    return __float_as_uint(__fadd_rd(__fadd_rz(y, 0.5f), VF_MAGIC));
}

template <int BITS>
__device__ __forceinline__ uint32_t unit_to_code_plain(float v) {
    const float mx = BITS == 8 ? 255.0f : 65535.0f;
    float y = roundf(rs_clamp(v, 0.0f, 1.0f) * mx);
    return BITS == 8 ? rs_as_u8(y) : rs_as_u16(y);
}

// ---- packed pairs of f32 (sm_100a: add / sub / fma .f32x2 -> FADD2 / FFMA2) --------------------
// The same IEEE round-to-nearest operations as the scalar ones, two per instruction.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ float lo2(f32x2 v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float hi2(f32x2 v) { return __uint_as_float((uint32_t)(v >> 32)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 c;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
    return c;
}
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 c;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(c) : "l"(a), "l"(b));
    return c;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
// a * b with its own rounding.  ptxas (12.9) contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2
// whatever -fmad says, which would drop a rounding the reference performs; a * b + (-0.0) is the
// same value as a * b in every case (signed zeros, infinities and NaN included), and with the -0.0
// arriving as kernel data there is nothing for the assembler to fold.
__device__ __forceinline__ f32x2 mulz2(f32x2 a, f32x2 b, float neg_zero) {
    return fma2(a, b, pk2(neg_zero, neg_zero));
}

__device__ __forceinline__ float lerp_ref(float a, float b, float t) {
    return __fadd_rn(a, __fmul_rn(__fsub_rn(b, a), t));  // a + (b - a) * t, three roundings
}

}  // namespace vf
