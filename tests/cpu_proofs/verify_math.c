/*
 * verify_math.c — exhaustive CPU proofs for the division shortcuts the CUDA
 * kernels use instead of IEEE `/` (gst-plugins-rs_b200/csrc/vf_math.cuh).
 * Each shortcut must equal the correctly rounded quotient for EVERY input in
 * the stated domain; fmaf() here is the same single-rounding FMA as FFMA.
 *
 *   q255(c)   = fma(c, K_hi, c*K_lo)                 c ∈ {0..255}      → c/255.0f
 *   q65535(c) = fma(c, K_hi, c*K_lo)                 c ∈ {0..65535}    → c/65535.0f
 *   q60(h)    = Markstein(h, 60)                     h ∈ [2^-20, 720]  → h/60.0f
 *   q60b(h)   = fma(h, K_hi, h*K_lo)                 h ∈ [2^-20, 720]  → h/60.0f  (2 ops)
 *
 * and for the float → code conversion of colorlut (unit_to_code, vf_math.cuh):
 *
 *   code(w) = low bits of RD(RZ(w*MAX + 0.5) + 2^23)   every float w ∈ [0, 1], MAX = 255 and 65535
 *                                                     → (w*MAX).round() as uN  (round half away)
 *
 * and for the triangle wave of HSV → RGB (to_rgb_fast, vf_math.cuh):
 *
 *   |hp - centre(ceil(hp))|, centre = 1, 3, 5       every float hp ∈ [0, 6]  → |fmod(hp, 2) - 1|
 *   ceil(hp) (0 counted as 1)                                                → the arm of the `<=` ladder
 *
 * Build: gcc -O2 -ffp-contract=off [-fopenmp] verify_math.c -lm ; exit code 0 = all proven.
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static float split_lo(double k, float *hi) {
    *hi = (float)k;
    return (float)(k - (double)*hi);
}

static inline float twoterm(float c, float hi, float lo) { return fmaf(c, hi, c * lo); }

/* RZ(y + 0.5) for y >= 0: the sum is exact in double; convert, then step back if RN went up */
static inline float add_half_rz(float y) {
    double d = (double)y + 0.5;
    float f = (float)d;
    if ((double)f > d) f = nextafterf(f, 0.0f);
    return f;
}

/* RD(f + 2^23) keeps floor(f) in the low mantissa bits for 0 <= f < 2^23 (ulp = 1 up there) */
static inline uint32_t low_bits_rd_magic(float f) { return (uint32_t)floorf(f); }

static inline float markstein(float a, float b, float rb) {
    float q0 = a * rb;
    float r = fmaf(-q0, b, a);
    return fmaf(r, rb, q0);
}

int main(void) {
    int bad = 0;
    float hi, lo;

    lo = split_lo(1.0 / 255.0, &hi);
    printf("K255  hi=%a lo=%a\n", hi, lo);
    for (int c = 0; c < 256; c++)
        if (twoterm((float)c, hi, lo) != (float)c / 255.0f) bad++, printf("q255 fail %d\n", c);

    lo = split_lo(1.0 / 65535.0, &hi);
    printf("K65535 hi=%a lo=%a\n", hi, lo);
    for (int c = 0; c < 65536; c++)
        if (twoterm((float)c, hi, lo) != (float)c / 65535.0f) bad++, printf("q65535 fail %d\n", c);

    /* h/60 over every float in [2^-20, 720] */
    float r60 = 1.0f / 60.0f;
    printf("R60 = %a\n", r60);
    uint32_t lo_bits, hi_bits;
    float lo_f = 0x1p-20f, hi_f = 720.0f;
    memcpy(&lo_bits, &lo_f, 4);
    memcpy(&hi_bits, &hi_f, 4);
    float hi60;
    float lo60 = split_lo(1.0 / 60.0, &hi60);
    printf("K60  hi=%a lo=%a\n", hi60, lo60);
    long bad60b = 0;
    long n60 = 0, bad60 = 0;
#pragma omp parallel for reduction(+ : bad60, bad60b, n60) schedule(static)
    for (uint32_t u = lo_bits; u <= hi_bits; u++) {
        float h;
        memcpy(&h, &u, 4);
        if (markstein(h, 60.0f, r60) != h / 60.0f) {
            if (bad60 < 5) printf("q60 fail %a\n", h);
            bad60++;
        }
        if (twoterm(h, hi60, lo60) != h / 60.0f) {
            if (bad60b < 5) printf("q60b fail %a\n", h);
            bad60b++;
        }
        n60++;
    }
    printf("q60b (two-term): %ld failures\n", bad60b);
    bad += bad60b != 0;
    printf("q60: %ld values, %ld failures\n", n60, bad60);
    bad += bad60 != 0;

    /* unit_to_code: every float in [0, 1] (sat() has clamped, NaN -> 0 on both sides) */
    for (int bits = 8; bits <= 16; bits += 8) {
        const float mx = bits == 8 ? 255.0f : 65535.0f;
        long badr = 0, nr = 0;
#pragma omp parallel for reduction(+ : badr, nr) schedule(static)
        for (uint32_t u = 0; u <= 0x3F800000u; u++) {
            float w;
            memcpy(&w, &u, 4);
            const float y = w * mx;
            if (low_bits_rd_magic(add_half_rz(y)) != (uint32_t)roundf(y)) {
                if (badr < 5) printf("round%d fail %a\n", bits, w);
                badr++;
            }
            nr++;
        }
        printf("round%d: %ld values, %ld failures\n", bits, nr, badr);
        bad += badr != 0;
    }

    /* to_rgb_fast: sector and triangle wave for every float hp in [0, 6] */
    {
        uint32_t six_bits;
        const float six = 6.0f;
        memcpy(&six_bits, &six, 4);
        long badt = 0, bada = 0, nt = 0;
#pragma omp parallel for reduction(+ : badt, bada, nt) schedule(static)
        for (uint32_t u = 0; u <= six_bits; u++) {
            float hp;
            memcpy(&hp, &u, 4);
            const int k = (int)ceilf(hp);
            const float centre = k >= 5 ? 5.0f : (k >= 3 ? 3.0f : 1.0f);
            if (fabsf(hp - centre) != fabsf(fmodf(hp, 2.0f) - 1.0f)) {
                if (badt < 5) printf("wave fail %a\n", hp);
                badt++;
            }
            /* hsvutils.rs:138-154: first arm i = 1..6 with hp <= i (hp >= 0 here) */
            int arm = 1;
            while (arm < 6 && !(hp <= (float)arm)) arm++;
            if ((k < 1 ? 1 : k) != arm) {
                if (bada < 5) printf("arm fail %a\n", hp);
                bada++;
            }
            nt++;
        }
        printf("wave: %ld values, %ld failures; arm: %ld failures\n", nt, badt, bada);
        bad += badt != 0 || bada != 0;
    }

    printf(bad ? "FAILED\n" : "ALL PROVEN\n");
    return bad != 0;
}
