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
 * Build: gcc -O2 -ffp-contract=off verify_math.c -lm ; exit code 0 = all proven.
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

    printf(bad ? "FAILED\n" : "ALL PROVEN\n");
    return bad != 0;
}
