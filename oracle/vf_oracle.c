/*
 * vf_oracle.c — CPU ORACLE (test infrastructure, see vf_oracle.h).
 *
 * Plain-C restatement of the reference's Rust arithmetic, operation for
 * operation.  Rust semantics preserved:
 *   - every f32 *, +, -, / is separately rounded  → build with -ffp-contract=off
 *   - `%` on f32                                   → fmodf
 *   - f32::round (half away from zero)             → roundf
 *   - `x as u8` / `as u16` / `as usize`            → saturating truncation, NaN → 0
 *   - inherent f32::clamp                          → compare/assign, NaN propagates
 *   - hsvutils::Clamp trait                        → fmaxf then fminf, NaN → bound
 *
 * Must not be compiled with -ffast-math or FMA contraction.
 */
#define _GNU_SOURCE
#include "vf_oracle.h"

#include <errno.h>
#include <limits.h>
#include <locale.h>
#include <math.h>
#include <pthread.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#if defined(__FAST_MATH__)
#error "the oracle must not be built with -ffast-math"
#endif

/* ------------------------------------------------------------------ */
/* Rust scalar semantics                                              */
/* ------------------------------------------------------------------ */

/* core::f32::clamp — `if self < min {min}; if self > max {max}`; NaN stays NaN */
static inline float rs_clamp(float v, float lo, float hi) {
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    return v;
}

/* hsvutils.rs:23-37 — local Clamp trait: self.max(lower).min(upper) */
static inline float trait_clamp(float v, float lo, float hi) { return fminf(fmaxf(v, lo), hi); }

/* Rust `f32 as u8`: truncate toward zero, saturate, NaN → 0 */
static inline uint8_t rs_as_u8(float v) {
    if (!(v > 0.0f)) return 0; /* NaN, negatives, zero */
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}

static inline uint16_t rs_as_u16(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 65535.0f) return 65535;
    return (uint16_t)v;
}

/* Rust `f32 as usize` */
static inline size_t rs_as_usize(float v) {
    if (!(v > 0.0f)) return 0;
    if (v >= 18446744073709551616.0f) return SIZE_MAX;
    return (size_t)v;
}

static inline size_t min_sz(size_t a, size_t b) { return a < b ? a : b; }

/* ------------------------------------------------------------------ */
/* .cube parser — video/colorlut/src/parser.rs:104-375                */
/* ------------------------------------------------------------------ */

#define LUT_1D_MIN_SIZE 2u     /* parser.rs:12 */
#define LUT_1D_MAX_SIZE 65536u /* parser.rs:13 */
#define LUT_3D_MIN_SIZE 2u     /* parser.rs:15 */
#define LUT_3D_MAX_SIZE 256u   /* parser.rs:16 */

static void set_err(char *err, size_t errlen, const char *fmt, ...) {
    if (!err || errlen == 0) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err, errlen, fmt, ap);
    va_end(ap);
}

/* Decode one UTF-8 scalar at p (< end).  Returns byte length, 0 if invalid. */
static int utf8_decode(const unsigned char *p, const unsigned char *end, uint32_t *cp) {
    unsigned char c = p[0];
    if (c < 0x80) {
        *cp = c;
        return 1;
    }
    int n;
    uint32_t v, min;
    if ((c & 0xE0) == 0xC0) {
        n = 2, v = c & 0x1F, min = 0x80;
    } else if ((c & 0xF0) == 0xE0) {
        n = 3, v = c & 0x0F, min = 0x800;
    } else if ((c & 0xF8) == 0xF0) {
        n = 4, v = c & 0x07, min = 0x10000;
    } else {
        return 0;
    }
    if (end - p < n) return 0;
    for (int i = 1; i < n; i++) {
        if ((p[i] & 0xC0) != 0x80) return 0;
        v = (v << 6) | (p[i] & 0x3F);
    }
    if (v < min || v > 0x10FFFF || (v >= 0xD800 && v <= 0xDFFF)) return 0;
    *cp = v;
    return n;
}

/* char::is_whitespace — Unicode White_Space (used by str::trim / split_whitespace) */
static int is_rust_ws(uint32_t cp) {
    return (cp >= 0x09 && cp <= 0x0D) || cp == 0x20 || cp == 0x85 || cp == 0xA0 || cp == 0x1680 ||
           (cp >= 0x2000 && cp <= 0x200A) || cp == 0x2028 || cp == 0x2029 || cp == 0x202F ||
           cp == 0x205F || cp == 0x3000;
}

typedef struct {
    const unsigned char *p, *end;
} tokiter;

/* split_whitespace().next(): returns 1 and [*tok, *tok + *len) or 0 */
static int next_token(tokiter *it, const unsigned char **tok, size_t *len) {
    uint32_t cp;
    while (it->p < it->end) {
        int n = utf8_decode(it->p, it->end, &cp);
        if (!is_rust_ws(cp)) break;
        it->p += n;
    }
    if (it->p >= it->end) return 0;
    *tok = it->p;
    while (it->p < it->end) {
        int n = utf8_decode(it->p, it->end, &cp);
        if (is_rust_ws(cp)) break;
        it->p += n;
    }
    *len = (size_t)(it->p - *tok);
    return 1;
}

static int ascii_ieq(const unsigned char *s, size_t n, const char *lit) {
    if (strlen(lit) != n) return 0;
    for (size_t i = 0; i < n; i++) {
        unsigned char c = s[i];
        if (c >= 'A' && c <= 'Z') c = (unsigned char)(c + 32);
        if (c != (unsigned char)lit[i]) return 0;
    }
    return 1;
}

/* <f32 as FromStr>::from_str grammar (core::num::dec2flt): [+-] then
 * inf|infinity|nan (any case) or digits [. digits] [e|E [+-] digits] with at
 * least one mantissa digit; whole token consumed.  Value: correctly rounded
 * (glibc strtof is, too).  Returns 1 on success. */
static int rs_parse_f32(const unsigned char *s, size_t n, float *out) {
    if (n == 0 || n > 4000) return 0;
    size_t i = 0;
    int neg = 0;
    if (s[0] == '+' || s[0] == '-') {
        neg = s[0] == '-';
        i = 1;
    }
    if (i >= n) return 0;
    if (ascii_ieq(s + i, n - i, "inf") || ascii_ieq(s + i, n - i, "infinity")) {
        *out = neg ? -INFINITY : INFINITY;
        return 1;
    }
    if (ascii_ieq(s + i, n - i, "nan")) {
        *out = NAN;
        return 1;
    }
    size_t digits = 0;
    while (i < n && s[i] >= '0' && s[i] <= '9') i++, digits++;
    if (i < n && s[i] == '.') {
        i++;
        while (i < n && s[i] >= '0' && s[i] <= '9') i++, digits++;
    }
    if (digits == 0) return 0;
    if (i < n && (s[i] == 'e' || s[i] == 'E')) {
        i++;
        if (i < n && (s[i] == '+' || s[i] == '-')) i++;
        size_t ed = 0;
        while (i < n && s[i] >= '0' && s[i] <= '9') i++, ed++;
        if (ed == 0) return 0;
    }
    if (i != n) return 0;
    char buf[4001];
    memcpy(buf, s, n);
    buf[n] = 0;
    /* Rust's from_str never looks at the process locale: convert under an explicit "C" locale
       (plain strtof would read "0.5" as 0 after setlocale(LC_ALL, "de_DE")) */
    static locale_t c_loc = (locale_t)0;
    if (!c_loc) c_loc = newlocale(LC_ALL_MASK, "C", (locale_t)0);
    char *end = NULL;
    *out = c_loc ? strtof_l(buf, &end, c_loc) : strtof(buf, &end);
    return end == buf + n;
}

/* `{:?}` of an f32 (core::fmt::float::float_to_general_debug): shortest digits that parse back to
 * the same f32; `d[.ddd]e[-]X` when 0 < |v| < 1e-4 or |v| >= 1e16, else plain decimal with at
 * least one fractional digit; "NaN", "inf", "-inf".  Digits by trying 1..9 significant digits. */
static void rs_debug_f32(float v, char *dst, size_t cap) {
    if (isnan(v)) {
        snprintf(dst, cap, "NaN");
        return;
    }
    if (isinf(v)) {
        snprintf(dst, cap, v < 0 ? "-inf" : "inf");
        return;
    }
    const char *sign = signbit(v) ? "-" : "";
    float a = fabsf(v);
    if (a == 0.0f) {
        snprintf(dst, cap, "%s0.0", sign);
        return;
    }
    char sci[64], digits[16];
    int e10 = 0;
    for (int prec = 0; prec < 9; prec++) {
        snprintf(sci, sizeof sci, "%.*e", prec, (double)a);
        for (char *p = sci; *p; p++)
            if (*p == ',') *p = '.'; /* a comma-decimal process locale */
        float back;
        if (rs_parse_f32((const unsigned char *)sci, strlen(sci), &back) && back == a) break;
    }
    size_t nd = 0;
    const char *p = sci;
    for (; *p && *p != 'e'; p++)
        if (*p != '.') digits[nd++] = *p;
    digits[nd] = 0;
    e10 = atoi(p + 1);
    while (nd > 1 && digits[nd - 1] == '0') digits[--nd] = 0;
    if (a < 1e-4f || a >= 1e16f) {
        if (nd > 1)
            snprintf(dst, cap, "%s%c.%se%d", sign, digits[0], digits + 1, e10);
        else
            snprintf(dst, cap, "%s%ce%d", sign, digits[0], e10);
    } else if (e10 < 0) {
        snprintf(dst, cap, "%s0.%.*s%s", sign, -e10 - 1, "000000000", digits);
    } else if ((size_t)e10 + 1 >= nd) {
        snprintf(dst, cap, "%s%s%.*s.0", sign, digits, (int)((size_t)e10 + 1 - nd), "0000000000000000");
    } else {
        snprintf(dst, cap, "%s%.*s.%s", sign, e10 + 1, digits, digits + e10 + 1);
    }
}

/* <usize as FromStr>::from_str: optional '+', ≥1 ASCII digits, overflow = error */
static int rs_parse_usize(const unsigned char *s, size_t n, size_t *out) {
    size_t i = 0;
    if (n == 0) return 0;
    if (s[0] == '+') i = 1;
    if (i >= n) return 0;
    unsigned long long v = 0;
    for (; i < n; i++) {
        if (s[i] < '0' || s[i] > '9') return 0;
        unsigned d = (unsigned)(s[i] - '0');
        if (v > (ULLONG_MAX - d) / 10ULL) return 0;
        v = v * 10ULL + d;
    }
    *out = (size_t)v;
    return 1;
}

enum { ST_HEADER = 0, ST_LUT1D = 1, ST_LUT3D = 3 }; /* parser.rs:97-102 ParseState */

typedef struct {
    int kind;
    size_t size;
    int have_data;
} pstate;

/* parser.rs:284-303 */
static int ensure_header(const pstate *st, size_t line_no, const char *line, char *err,
                         size_t errlen) {
    if (st->kind != ST_HEADER && st->have_data) {
        set_err(err, errlen, "Invalid LUT: Header found after LUT data at line %zu: %s", line_no,
                line);
        return 0;
    }
    return 1;
}

/* parser.rs:360-370 + 372-375 */
static int parse_next_f32(tokiter *it, float *out, size_t line_no, const char *line, char *err,
                          size_t errlen) {
    const unsigned char *tok;
    size_t len;
    if (!next_token(it, &tok, &len)) {
        set_err(err, errlen, "Invalid LUT: Invalid line %zu: %s", line_no, line);
        return 0;
    }
    if (!rs_parse_f32(tok, len, out)) {
        set_err(err, errlen, "Invalid LUT: Invalid float at line %zu: %s", line_no, line);
        return 0;
    }
    return 1;
}

/* parser.rs:320-336 */
static int parse_vec3(tokiter *it, float v[3], size_t line_no, const char *line, char *err,
                      size_t errlen) {
    for (int i = 0; i < 3; i++)
        if (!parse_next_f32(it, &v[i], line_no, line, err, errlen)) return 0;
    const unsigned char *tok;
    size_t len;
    if (next_token(it, &tok, &len)) {
        set_err(err, errlen, "Invalid LUT: Invalid line %zu: %s", line_no, line);
        return 0;
    }
    return 1;
}

/* parser.rs:338-358 */
static int parse_single_usize(tokiter *it, size_t *out, size_t line_no, const char *line,
                              char *err, size_t errlen) {
    const unsigned char *tok;
    size_t len;
    if (!next_token(it, &tok, &len)) {
        set_err(err, errlen, "Invalid LUT: Invalid line %zu: %s", line_no, line);
        return 0;
    }
    if (!rs_parse_usize(tok, len, out)) {
        set_err(err, errlen, "Invalid LUT: Invalid integer at line %zu: %s", line_no, line);
        return 0;
    }
    if (next_token(it, &tok, &len)) {
        set_err(err, errlen, "Invalid LUT: Invalid line %zu: %s", line_no, line);
        return 0;
    }
    return 1;
}

/* parser.rs:305-318 */
static int validate_lut_size(size_t size, size_t min, size_t max, size_t line_no, char *err,
                             size_t errlen) {
    if (size < min || size > max) {
        set_err(err, errlen, "Invalid LUT: Invalid LUT size %zu at line %zu, expected %zu..=%zu",
                size, line_no, min, max);
        return 0;
    }
    return 1;
}

static int tok_is(const unsigned char *tok, size_t len, const char *lit) {
    return strlen(lit) == len && memcmp(tok, lit, len) == 0;
}

/* parser.rs:110-281 CubeLut::parse */
int orc_cube_parse(const char *text, size_t len, orc_cube *out, char *err, size_t errlen) {
    memset(out, 0, sizeof(*out));
    const unsigned char *p = (const unsigned char *)text, *end = p + len;

    /* fs::read_to_string (parser.rs:106) fails on invalid UTF-8 → CubeParseError::Io */
    for (const unsigned char *q = p; q < end;) {
        uint32_t cp;
        int n = utf8_decode(q, end, &cp);
        if (n == 0) {
            set_err(err, errlen, "IO error: stream did not contain valid UTF-8");
            return 2;
        }
        q += n;
    }

    float domain_min[3] = {0.0f, 0.0f, 0.0f}; /* parser.rs:111 */
    float domain_max[3] = {1.0f, 1.0f, 1.0f}; /* parser.rs:112 */
    pstate st = {ST_HEADER, 0, 0};
    size_t cap = 0, count = 0;
    float *values = NULL; /* Vec<[f32;3]> */
    char *linebuf = NULL;
    int rc = 1;

    size_t line_no = 0;
    while (p < end) { /* str::lines(): split on '\n', strip one trailing '\r' */
        const unsigned char *nl = memchr(p, '\n', (size_t)(end - p));
        const unsigned char *ls = p, *le = nl ? nl : end;
        p = nl ? nl + 1 : end;
        line_no++;
        if (le > ls && le[-1] == '\r') le--;

        /* trim() */
        uint32_t cp;
        while (ls < le) {
            int n = utf8_decode(ls, le, &cp);
            if (!is_rust_ws(cp)) break;
            ls += n;
        }
        while (le > ls) {
            const unsigned char *q = le - 1;
            while (q > ls && (*q & 0xC0) == 0x80) q--;
            utf8_decode(q, le, &cp);
            if (!is_rust_ws(cp)) break;
            le = q;
        }
        if (ls == le || *ls == '#') continue; /* parser.rs:120-122 */

        free(linebuf);
        linebuf = (char *)malloc((size_t)(le - ls) + 1);
        memcpy(linebuf, ls, (size_t)(le - ls));
        linebuf[le - ls] = 0;

        tokiter it = {ls, le};
        const unsigned char *first;
        size_t flen;
        if (!next_token(&it, &first, &flen)) continue;

        if (tok_is(first, flen, "TITLE")) { /* parser.rs:132-134 */
            if (!ensure_header(&st, line_no, linebuf, err, errlen)) goto done;
        } else if (tok_is(first, flen, "DOMAIN_MIN")) { /* :135-138 */
            if (!ensure_header(&st, line_no, linebuf, err, errlen)) goto done;
            if (!parse_vec3(&it, domain_min, line_no, linebuf, err, errlen)) goto done;
        } else if (tok_is(first, flen, "DOMAIN_MAX")) { /* :139-142 */
            if (!ensure_header(&st, line_no, linebuf, err, errlen)) goto done;
            if (!parse_vec3(&it, domain_max, line_no, linebuf, err, errlen)) goto done;
        } else if (tok_is(first, flen, "LUT_1D_SIZE") || tok_is(first, flen, "LUT_3D_SIZE")) {
            /* :143-176 */
            int is1d = first[4] == '1';
            if (!ensure_header(&st, line_no, linebuf, err, errlen)) goto done;
            if (st.kind != ST_HEADER) {
                set_err(err, errlen, "Invalid LUT: Invalid LUT_%cD_SIZE at line %zu: %s",
                        is1d ? '1' : '3', line_no, linebuf);
                goto done;
            }
            size_t size;
            if (!parse_single_usize(&it, &size, line_no, linebuf, err, errlen)) goto done;
            if (!validate_lut_size(size, is1d ? LUT_1D_MIN_SIZE : LUT_3D_MIN_SIZE,
                                   is1d ? LUT_1D_MAX_SIZE : LUT_3D_MAX_SIZE, line_no, err, errlen))
                goto done;
            st.kind = is1d ? ST_LUT1D : ST_LUT3D;
            st.size = size;
            st.have_data = 0;
        } else { /* data line, parser.rs:177-201 */
            if (st.kind == ST_HEADER) {
                set_err(err, errlen, "Invalid LUT: LUT data found before LUT size at line %zu: %s",
                        line_no, linebuf);
                goto done;
            }
            st.have_data = 1;
            float v[3];
            if (!rs_parse_f32(first, flen, &v[0])) {
                set_err(err, errlen, "Invalid LUT: Invalid float at line %zu: %s", line_no,
                        linebuf);
                goto done;
            }
            if (!parse_next_f32(&it, &v[1], line_no, linebuf, err, errlen)) goto done;
            if (!parse_next_f32(&it, &v[2], line_no, linebuf, err, errlen)) goto done;
            const unsigned char *tok;
            size_t tl;
            if (next_token(&it, &tok, &tl)) {
                set_err(err, errlen, "Invalid LUT: Invalid line %zu: %s", line_no, linebuf);
                goto done;
            }
            if (count == cap) {
                cap = cap ? cap * 2 : 1024;
                values = (float *)realloc(values, cap * 3 * sizeof(float));
            }
            memcpy(values + count * 3, v, sizeof(v));
            count++;
        }
    }

    /* parser.rs:205-212 — NaN bounds pass this test, as in Rust */
    if (domain_min[0] >= domain_max[0] || domain_min[1] >= domain_max[1] ||
        domain_min[2] >= domain_max[2]) {
        char d[6][48]; /* Rust formats both arrays with {:?} */
        for (int c = 0; c < 3; c++) {
            rs_debug_f32(domain_min[c], d[c], sizeof d[c]);
            rs_debug_f32(domain_max[c], d[3 + c], sizeof d[c]);
        }
        set_err(err, errlen, "Invalid LUT: Invalid domain min [%s, %s, %s], max [%s, %s, %s]", d[0],
                d[1], d[2], d[3], d[4], d[5]);
        goto done;
    }

    if (st.kind == ST_HEADER) { /* :215-217 */
        set_err(err, errlen, "Invalid LUT: Missing LUT size");
        goto done;
    } else if (st.kind == ST_LUT1D) { /* :218-237 */
        if (count != st.size) {
            set_err(err, errlen, "Invalid LUT: Invalid 1D LUT value count, expected %zu, got %zu",
                    st.size, count);
            goto done;
        }
        out->kind = ORC_LUT_1D;
        out->size = (uint32_t)st.size;
        out->n_floats = 3 * st.size;
        out->data = (float *)malloc(out->n_floats * sizeof(float));
        for (size_t i = 0; i < st.size; i++) {
            out->data[i] = values[3 * i + 0];
            out->data[st.size + i] = values[3 * i + 1];
            out->data[2 * st.size + i] = values[3 * i + 2];
        }
    } else { /* :238-261 */
        size_t expected = st.size * st.size * st.size;
        if (count != expected) {
            set_err(err, errlen, "Invalid LUT: Invalid 3D LUT value count, expected %zu, got %zu",
                    expected, count);
            goto done;
        }
        out->kind = ORC_LUT_3D;
        out->size = (uint32_t)st.size;
        out->n_floats = 4 * expected;
        out->data = (float *)malloc(out->n_floats * sizeof(float));
        for (size_t i = 0; i < expected; i++) {
            out->data[4 * i + 0] = values[3 * i + 0];
            out->data[4 * i + 1] = values[3 * i + 1];
            out->data[4 * i + 2] = values[3 * i + 2];
            out->data[4 * i + 3] = 1.0f; /* :255 */
        }
    }

    for (int c = 0; c < 3; c++) { /* :264-274 */
        out->domain_scale[c] = 1.0f / (domain_max[c] - domain_min[c]);
        out->domain_offset[c] = -domain_min[c] * out->domain_scale[c];
    }
    rc = 0;

done:
    free(values);
    free(linebuf);
    return rc;
}

/* parser.rs:105-108 */
int orc_cube_parse_file(const char *path, orc_cube *out, char *err, size_t errlen) {
    memset(out, 0, sizeof(*out));
    FILE *f = fopen(path, "rb");
    if (!f) {
        set_err(err, errlen, "IO error: %s", strerror(errno));
        return 2;
    }
    size_t cap = 1 << 16, len = 0;
    char *buf = (char *)malloc(cap);
    for (;;) {
        size_t n = fread(buf + len, 1, cap - len, f);
        len += n;
        if (n == 0) break;
        if (len == cap) buf = (char *)realloc(buf, cap *= 2);
    }
    int io_err = ferror(f);
    fclose(f);
    if (io_err) {
        free(buf);
        set_err(err, errlen, "IO error: read failed");
        return 2;
    }
    int rc = orc_cube_parse(buf, len, out, err, errlen);
    free(buf);
    return rc;
}

void orc_cube_free(orc_cube *c) {
    if (!c) return;
    free(c->data);
    memset(c, 0, sizeof(*c));
}

/* ------------------------------------------------------------------ */
/* colorlut — video/colorlut/src/colorlut/imp.rs:226-543              */
/* ------------------------------------------------------------------ */

/* imp.rs:471-474 */
static inline float norm_comp(const orc_cube *lut, int c, uint8_t value) {
    float v = (float)value / 255.0f;
    return rs_clamp(v * lut->domain_scale[c] + lut->domain_offset[c], 0.0f, 1.0f);
}

/* imp.rs:476-479 */
static inline float norm_comp_u16(const orc_cube *lut, int c, uint16_t value) {
    float v = (float)value / 65535.0f;
    return rs_clamp(v * lut->domain_scale[c] + lut->domain_offset[c], 0.0f, 1.0f);
}

/* imp.rs:537-539 */
static inline uint8_t float_to_u8(float v) { return rs_as_u8(roundf(rs_clamp(v, 0.0f, 1.0f) * 255.0f)); }

/* imp.rs:541-543 */
static inline uint16_t float_to_u16(float v) {
    return rs_as_u16(roundf(rs_clamp(v, 0.0f, 1.0f) * 65535.0f));
}

/* imp.rs:482-490 — linear */
static inline float sample_1d(const float *lut, size_t len, float x) {
    size_t max_idx = len - 1;
    size_t x0 = min_sz(rs_as_usize(floorf(x)), max_idx);
    size_t x1 = min_sz(x0 + 1, max_idx);
    float t = x - (float)x0;
    return lut[x0] + (lut[x1] - lut[x0]) * t;
}

/* imp.rs:528-535 — all four lanes, as the reference computes them */
static inline void lerp4(const float *a, const float *b, float t, float *o) {
    o[0] = a[0] + (b[0] - a[0]) * t;
    o[1] = a[1] + (b[1] - a[1]) * t;
    o[2] = a[2] + (b[2] - a[2]) * t;
    o[3] = a[3] + (b[3] - a[3]) * t;
}

/* parser.rs:43-53 Lut3D::at */
static inline const float *lut_at(const orc_cube *lut, size_t x, size_t y, size_t z) {
    size_t n = lut->size;
    return lut->data + 4 * (x + y * n + z * n * n);
}

/* imp.rs:493-526 — trilinear */
static inline void sample_3d(const orc_cube *lut, float x, float y, float z, float out[4]) {
    size_t max_idx = (size_t)lut->size - 1;

    size_t x0 = min_sz(rs_as_usize(floorf(x)), max_idx);
    size_t y0 = min_sz(rs_as_usize(floorf(y)), max_idx);
    size_t z0 = min_sz(rs_as_usize(floorf(z)), max_idx);

    size_t x1 = min_sz(x0 + 1, max_idx);
    size_t y1 = min_sz(y0 + 1, max_idx);
    size_t z1 = min_sz(z0 + 1, max_idx);

    float tx = x - (float)x0;
    float ty = y - (float)y0;
    float tz = z - (float)z0;

    const float *c000 = lut_at(lut, x0, y0, z0);
    const float *c100 = lut_at(lut, x1, y0, z0);
    const float *c010 = lut_at(lut, x0, y1, z0);
    const float *c110 = lut_at(lut, x1, y1, z0);
    const float *c001 = lut_at(lut, x0, y0, z1);
    const float *c101 = lut_at(lut, x1, y0, z1);
    const float *c011 = lut_at(lut, x0, y1, z1);
    const float *c111 = lut_at(lut, x1, y1, z1);

    float c00[4], c10[4], c01[4], c11[4], c0[4], c1[4];
    lerp4(c000, c100, tx, c00);
    lerp4(c010, c110, tx, c10);
    lerp4(c001, c101, tx, c01);
    lerp4(c011, c111, tx, c11);

    lerp4(c00, c10, ty, c0);
    lerp4(c01, c11, ty, c1);

    lerp4(c0, c1, tz, out);
}

/* imp.rs:399-413 */
static inline uint8_t apply_1d(const orc_cube *lut, int c, uint8_t value) {
    const float *table = lut->data + (size_t)c * lut->size;
    float x = norm_comp(lut, c, value) * ((float)lut->size - 1.0f);
    return float_to_u8(sample_1d(table, lut->size, x));
}

/* imp.rs:415-429 */
static inline uint16_t apply_1d_u16(const orc_cube *lut, int c, uint16_t value) {
    const float *table = lut->data + (size_t)c * lut->size;
    float x = norm_comp_u16(lut, c, value) * ((float)lut->size - 1.0f);
    return float_to_u16(sample_1d(table, lut->size, x));
}

/* imp.rs:431-449 */
static inline void apply_3d(const orc_cube *lut, uint8_t r, uint8_t g, uint8_t b, uint8_t o[3]) {
    float sm1 = (float)lut->size - 1.0f;
    float x = norm_comp(lut, 0, r) * sm1;
    float y = norm_comp(lut, 1, g) * sm1;
    float z = norm_comp(lut, 2, b) * sm1;
    float out[4];
    sample_3d(lut, x, y, z, out);
    o[0] = float_to_u8(out[0]);
    o[1] = float_to_u8(out[1]);
    o[2] = float_to_u8(out[2]);
}

/* imp.rs:451-469 */
static inline void apply_3d_u16(const orc_cube *lut, uint16_t r, uint16_t g, uint16_t b,
                                uint16_t o[3]) {
    float sm1 = (float)lut->size - 1.0f;
    float x = norm_comp_u16(lut, 0, r) * sm1;
    float y = norm_comp_u16(lut, 1, g) * sm1;
    float z = norm_comp_u16(lut, 2, b) * sm1;
    float out[4];
    sample_3d(lut, x, y, z, out);
    o[0] = float_to_u16(out[0]);
    o[1] = float_to_u16(out[1]);
    o[2] = float_to_u16(out[2]);
}

void orc_colorlut_apply_u8(const orc_cube *lut, const uint8_t in[3], uint8_t out[3]) {
    if (lut->kind == ORC_LUT_1D) {
        for (int c = 0; c < 3; c++) out[c] = apply_1d(lut, c, in[c]);
    } else {
        apply_3d(lut, in[0], in[1], in[2], out);
    }
}

void orc_colorlut_apply_u16(const orc_cube *lut, const uint16_t in[3], uint16_t out[3]) {
    if (lut->kind == ORC_LUT_1D) {
        for (int c = 0; c < 3; c++) out[c] = apply_1d_u16(lut, c, in[c]);
    } else {
        apply_3d_u16(lut, in[0], in[1], in[2], out);
    }
}

static inline uint16_t bswap16(uint16_t v) { return (uint16_t)((v >> 8) | (v << 8)); }

/* imp.rs:237-294 (RGBA) and 307-397 (RGBA64 LE/BE).  Host is little-endian, so
 * u16::from_le/to_le are the identity and from_be/to_be swap bytes. */
int orc_colorlut_frame(const orc_cube *lut, const uint8_t *src, size_t src_stride, uint8_t *dst,
                       size_t dst_stride, uint32_t width, uint32_t height, int format) {
    if (!lut || !lut->data) return -1;
    if (format == ORC_FMT_RGBA) {
        size_t wb = (size_t)width * 4;
        for (uint32_t row = 0; row < height; row++) {
            const uint8_t *s = src + (size_t)row * src_stride;
            uint8_t *d = dst + (size_t)row * dst_stride;
            if (lut->kind == ORC_LUT_1D) { /* imp.rs:258-263 */
                for (size_t i = 0; i < wb; i += 4) {
                    for (int c = 0; c < 3; c++) d[i + c] = apply_1d(lut, c, s[i + c]);
                    d[i + 3] = s[i + 3];
                }
            } else { /* imp.rs:288-292 */
                for (size_t i = 0; i < wb; i += 4) {
                    apply_3d(lut, s[i], s[i + 1], s[i + 2], d + i);
                    d[i + 3] = s[i + 3];
                }
            }
        }
        return 0;
    }
    if (format == ORC_FMT_RGBA64_LE || format == ORC_FMT_RGBA64_BE) {
        int le = format == ORC_FMT_RGBA64_LE;
        if ((src_stride | dst_stride) & 1) return -1; /* as_slice_of::<u16>() needs whole u16s */
        size_t ss = src_stride / 2, ds = dst_stride / 2; /* imp.rs:315-316 */
        for (uint32_t row = 0; row < height; row++) {
            const uint16_t *s = (const uint16_t *)(const void *)src + (size_t)row * ss;
            uint16_t *d = (uint16_t *)(void *)dst + (size_t)row * ds;
            for (size_t i = 0; i < (size_t)width * 4; i += 4) {
                uint16_t in[3], o[3];
                for (int c = 0; c < 3; c++) in[c] = le ? s[i + c] : bswap16(s[i + c]);
                if (lut->kind == ORC_LUT_1D) {
                    for (int c = 0; c < 3; c++) o[c] = apply_1d_u16(lut, c, in[c]);
                } else {
                    apply_3d_u16(lut, in[0], in[1], in[2], o);
                }
                for (int c = 0; c < 3; c++) d[i + c] = le ? o[c] : bswap16(o[c]);
                d[i + 3] = s[i + 3]; /* imp.rs:345, 394 — alpha word copied raw */
            }
        }
        return 0;
    }
    return -1;
}

/* ------------------------------------------------------------------ */
/* EXTENSION — interpolation modes the reference does NOT have         */
/* ------------------------------------------------------------------ */
/* BASELINE.json names "nearest/trilinear/tetrahedral"; the reference implements trilinear only
 * (SURVEY.md F1).  These two modes therefore have no reference counterpart and no reference
 * parity claim: they restate the published algorithms (tetrahedral: the six-tetrahedra split of
 * the unit cell ordered by the fractional coordinates, as in Kasson et al. / FFmpeg's lut3d
 * filter; nearest: round half up per axis) in the reference's own conventions — coordinates
 * from norm_comp (imp.rs:471-479), i0/t split as in sample_3d (imp.rs:496-508), unfused f32,
 * float_to_u8/u16 on the way out.  The CUDA kernels are checked against THIS definition. */
static inline void sample_3d_tetrahedral(const orc_cube *lut, float x, float y, float z, float out[3]) {
    size_t max_idx = (size_t)lut->size - 1;
    size_t x0 = min_sz(rs_as_usize(floorf(x)), max_idx);
    size_t y0 = min_sz(rs_as_usize(floorf(y)), max_idx);
    size_t z0 = min_sz(rs_as_usize(floorf(z)), max_idx);
    size_t x1 = min_sz(x0 + 1, max_idx);
    size_t y1 = min_sz(y0 + 1, max_idx);
    size_t z1 = min_sz(z0 + 1, max_idx);
    float tx = x - (float)x0;
    float ty = y - (float)y0;
    float tz = z - (float)z0;

    const float *c000 = lut_at(lut, x0, y0, z0), *c111 = lut_at(lut, x1, y1, z1);
    const float *ca, *cb;
    float w0, wa, wb, w1;
    if (tx > ty) {
        if (ty > tz) { /* x > y > z */
            w0 = 1.0f - tx, wa = tx - ty, wb = ty - tz, w1 = tz;
            ca = lut_at(lut, x1, y0, z0), cb = lut_at(lut, x1, y1, z0);
        } else if (tx > tz) { /* x > z >= y */
            w0 = 1.0f - tx, wa = tx - tz, wb = tz - ty, w1 = ty;
            ca = lut_at(lut, x1, y0, z0), cb = lut_at(lut, x1, y0, z1);
        } else { /* z >= x > y */
            w0 = 1.0f - tz, wa = tz - tx, wb = tx - ty, w1 = ty;
            ca = lut_at(lut, x0, y0, z1), cb = lut_at(lut, x1, y0, z1);
        }
    } else {
        if (tz > ty) { /* z > y >= x */
            w0 = 1.0f - tz, wa = tz - ty, wb = ty - tx, w1 = tx;
            ca = lut_at(lut, x0, y0, z1), cb = lut_at(lut, x0, y1, z1);
        } else if (tz > tx) { /* y >= z > x */
            w0 = 1.0f - ty, wa = ty - tz, wb = tz - tx, w1 = tx;
            ca = lut_at(lut, x0, y1, z0), cb = lut_at(lut, x0, y1, z1);
        } else { /* y >= x >= z */
            w0 = 1.0f - ty, wa = ty - tx, wb = tx - tz, w1 = tz;
            ca = lut_at(lut, x0, y1, z0), cb = lut_at(lut, x1, y1, z0);
        }
    }
    for (int c = 0; c < 3; c++) {
        float acc = w0 * c000[c];
        acc = acc + wa * ca[c];
        acc = acc + wb * cb[c];
        out[c] = acc + w1 * c111[c];
    }
}

static inline void sample_3d_nearest(const orc_cube *lut, float x, float y, float z, float out[3]) {
    size_t max_idx = (size_t)lut->size - 1;
    size_t xi = min_sz(rs_as_usize(floorf(x + 0.5f)), max_idx);
    size_t yi = min_sz(rs_as_usize(floorf(y + 0.5f)), max_idx);
    size_t zi = min_sz(rs_as_usize(floorf(z + 0.5f)), max_idx);
    const float *c = lut_at(lut, xi, yi, zi);
    out[0] = c[0], out[1] = c[1], out[2] = c[2];
}

static inline void sample_3d_mode(const orc_cube *lut, float x, float y, float z, float out[3],
                                  int interpolation) {
    if (interpolation == ORC_INTERP_TETRAHEDRAL) {
        sample_3d_tetrahedral(lut, x, y, z, out);
    } else if (interpolation == ORC_INTERP_NEAREST) {
        sample_3d_nearest(lut, x, y, z, out);
    } else {
        float o4[4];
        sample_3d(lut, x, y, z, o4);
        out[0] = o4[0], out[1] = o4[1], out[2] = o4[2];
    }
}

int orc_colorlut_frame_ex(const orc_cube *lut, const uint8_t *src, size_t src_stride, uint8_t *dst,
                          size_t dst_stride, uint32_t width, uint32_t height, int format,
                          int interpolation) {
    if (!lut || !lut->data) return -1;
    if (interpolation < ORC_INTERP_TRILINEAR || interpolation > ORC_INTERP_NEAREST) return -1;
    if (lut->kind != ORC_LUT_3D || interpolation == ORC_INTERP_TRILINEAR) /* 1D LUTs stay linear */
        return orc_colorlut_frame(lut, src, src_stride, dst, dst_stride, width, height, format);
    const float sm1 = (float)lut->size - 1.0f;
    float o[3];
    if (format == ORC_FMT_RGBA) {
        for (uint32_t row = 0; row < height; row++) {
            const uint8_t *s = src + (size_t)row * src_stride;
            uint8_t *d = dst + (size_t)row * dst_stride;
            for (size_t i = 0; i < (size_t)width * 4; i += 4) {
                sample_3d_mode(lut, norm_comp(lut, 0, s[i]) * sm1, norm_comp(lut, 1, s[i + 1]) * sm1,
                               norm_comp(lut, 2, s[i + 2]) * sm1, o, interpolation);
                for (int c = 0; c < 3; c++) d[i + c] = float_to_u8(o[c]);
                d[i + 3] = s[i + 3];
            }
        }
        return 0;
    }
    if (format == ORC_FMT_RGBA64_LE || format == ORC_FMT_RGBA64_BE) {
        int le = format == ORC_FMT_RGBA64_LE;
        if ((src_stride | dst_stride) & 1) return -1;
        for (uint32_t row = 0; row < height; row++) {
            const uint16_t *s = (const uint16_t *)(const void *)(src + (size_t)row * src_stride);
            uint16_t *d = (uint16_t *)(void *)(dst + (size_t)row * dst_stride);
            for (size_t i = 0; i < (size_t)width * 4; i += 4) {
                uint16_t in[3];
                for (int c = 0; c < 3; c++) in[c] = le ? s[i + c] : bswap16(s[i + c]);
                sample_3d_mode(lut, norm_comp_u16(lut, 0, in[0]) * sm1,
                               norm_comp_u16(lut, 1, in[1]) * sm1, norm_comp_u16(lut, 2, in[2]) * sm1,
                               o, interpolation);
                for (int c = 0; c < 3; c++) {
                    uint16_t v = float_to_u16(o[c]);
                    d[i + c] = le ? v : bswap16(v);
                }
                d[i + 3] = s[i + 3];
            }
        }
        return 0;
    }
    return -1;
}

/* ------------------------------------------------------------------ */
/* hsvutils — video/hsv/src/hsvutils.rs                               */
/* ------------------------------------------------------------------ */

#define HSV_EPSILON 0.00001f /* hsvutils.rs:40 */

/* hsvutils.rs:44-84 with (r,g,b) already picked from the pixel */
static inline void hsv_from_channels(uint8_t rb, uint8_t gb, uint8_t bb, float hsv[3]) {
    float r = (float)rb / 255.0f;
    float g = (float)gb / 255.0f;
    float b = (float)bb / 255.0f;

    uint8_t mx = rb > gb ? rb : gb;
    mx = mx > bb ? mx : bb;
    uint8_t mn = rb < gb ? rb : gb;
    mn = mn < bb ? mn : bb;

    float value = (float)mx / 255.0f;
    float chroma = value - ((float)mn / 255.0f);

    float hue;
    if (chroma == 0.0f) {
        hue = 0.0f;
    } else if (fabsf(value - r) < HSV_EPSILON) {
        hue = 60.0f * ((g - b) / chroma);
    } else if (fabsf(value - g) < HSV_EPSILON) {
        hue = 60.0f * (2.0f + ((b - r) / chroma));
    } else if (fabsf(value - b) < HSV_EPSILON) {
        hue = 60.0f * (4.0f + ((r - g) / chroma));
    } else {
        hue = 0.0f;
    }

    if (hue < 0.0f) hue += 360.0f;

    float saturation = value == 0.0f ? 0.0f : chroma / value;

    hsv[0] = fmodf(hue, 360.0f);
    hsv[1] = rs_clamp(saturation, 0.0f, 1.0f);
    hsv[2] = rs_clamp(value, 0.0f, 1.0f);
}

void orc_hsv_from_rgb(const uint8_t p[3], float hsv[3]) { hsv_from_channels(p[0], p[1], p[2], hsv); }
/* hsvutils.rs:88-128 — bytes are B,G,R */
void orc_hsv_from_bgr(const uint8_t p[3], float hsv[3]) { hsv_from_channels(p[2], p[1], p[0], hsv); }

/* hsvutils.rs:132-163 — returns r,g,b */
static inline void hsv_to_channels(const float in_p[3], uint8_t rgb[3]) {
    float c = in_p[2] * in_p[1];
    float hue_prime = in_p[0] / 60.0f;

    float x = c * (1.0f - fabsf(fmodf(hue_prime, 2.0f) - 1.0f));

    float p0, p1, p2;
    if (hue_prime < 0.0f) {
        p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;
    } else if (hue_prime <= 1.0f) {
        p0 = c, p1 = x, p2 = 0.0f;
    } else if (hue_prime <= 2.0f) {
        p0 = x, p1 = c, p2 = 0.0f;
    } else if (hue_prime <= 3.0f) {
        p0 = 0.0f, p1 = c, p2 = x;
    } else if (hue_prime <= 4.0f) {
        p0 = 0.0f, p1 = x, p2 = c;
    } else if (hue_prime <= 5.0f) {
        p0 = x, p1 = 0.0f, p2 = c;
    } else if (hue_prime <= 6.0f) {
        p0 = c, p1 = 0.0f, p2 = x;
    } else {
        p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;
    }

    float m = in_p[2] - c;

    rgb[0] = rs_as_u8(rs_clamp((p0 + m) * 255.0f, 0.0f, 255.0f));
    rgb[1] = rs_as_u8(rs_clamp((p1 + m) * 255.0f, 0.0f, 255.0f));
    rgb[2] = rs_as_u8(rs_clamp((p2 + m) * 255.0f, 0.0f, 255.0f));
}

void orc_hsv_from_rgba_batch(const uint8_t *rgba, size_t n, float *hsv) {
    for (size_t i = 0; i < n; i++) hsv_from_channels(rgba[4 * i], rgba[4 * i + 1], rgba[4 * i + 2], hsv + 3 * i);
}

void orc_hsv_to_rgb(const float hsv[3], uint8_t out[3]) { hsv_to_channels(hsv, out); }

/* hsvutils.rs:167-198 — stored reversed */
void orc_hsv_to_bgr(const float hsv[3], uint8_t out[3]) {
    uint8_t rgb[3];
    hsv_to_channels(hsv, rgb);
    out[0] = rgb[2];
    out[1] = rgb[1];
    out[2] = rgb[0];
}

/* ------------------------------------------------------------------ */
/* Pixel layouts                                */
/* ------------------------------------------------------------------ */

typedef struct {
    int bpp;     /* pixel stride in bytes */
    int off;     /* first colour byte */
    int bgr;     /* colour order is B,G,R */
    int alpha_i; /* index of the alpha/x byte, -1 if none */
} pixfmt;

static int fmt_info(int format, pixfmt *f) {
    switch (format) {
    case ORC_FMT_RGBA:
    case ORC_FMT_RGBX: *f = (pixfmt){4, 0, 0, 3}; return 1;
    case ORC_FMT_XRGB:
    case ORC_FMT_ARGB: *f = (pixfmt){4, 1, 0, 0}; return 1;
    case ORC_FMT_BGRX:
    case ORC_FMT_BGRA: *f = (pixfmt){4, 0, 1, 3}; return 1;
    case ORC_FMT_XBGR:
    case ORC_FMT_ABGR: *f = (pixfmt){4, 1, 1, 0}; return 1;
    case ORC_FMT_RGB: *f = (pixfmt){3, 0, 0, -1}; return 1;
    case ORC_FMT_BGR: *f = (pixfmt){3, 0, 1, -1}; return 1;
    default: return 0;
    }
}

/* ------------------------------------------------------------------ */
/* hsvfilter — video/hsv/src/hsvfilter/imp.rs:76-120, 323-376          */
/* ------------------------------------------------------------------ */

int orc_hsvfilter_frame(uint8_t *data, size_t stride, uint32_t width, uint32_t height, int format,
                        const orc_hsvfilter_params *p) {
    pixfmt f;
    if (!fmt_info(format, &f)) return -1;
    const orc_hsvfilter_params s = *p; /* imp.rs:85 — settings copied once per frame */
    size_t line_bytes = (size_t)width * (size_t)f.bpp; /* imp.rs:94 */
    for (uint32_t row = 0; row < height; row++) {
        uint8_t *line = data + (size_t)row * stride;
        for (size_t i = 0; i < line_bytes; i += (size_t)f.bpp) {
            uint8_t *px = line + i + f.off;
            float hsv[3];
            if (f.bgr)
                orc_hsv_from_bgr(px, hsv);
            else
                orc_hsv_from_rgb(px, hsv);

            hsv[0] = fmodf(hsv[0] + s.hue_shift, 360.0f); /* imp.rs:102 */
            if (hsv[0] < 0.0f) hsv[0] += 360.0f;          /* imp.rs:103-105 */
            hsv[1] = trait_clamp(s.saturation_mul * hsv[1] + s.saturation_off, 0.0f, 1.0f);
            hsv[2] = trait_clamp(s.value_mul * hsv[2] + s.value_off, 0.0f, 1.0f);

            if (f.bgr)
                orc_hsv_to_bgr(hsv, px);
            else
                orc_hsv_to_rgb(hsv, px);
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* hsvdetector — video/hsv/src/hsvdetector/imp.rs:100-160, 423-707     */
/* ------------------------------------------------------------------ */

int orc_hsvdetector_frame(const uint8_t *in, size_t in_stride, int in_format, uint8_t *out,
                          size_t out_stride, int out_format, uint32_t width, uint32_t height,
                          const orc_hsvdetector_params *p) {
    pixfmt fi, fo;
    if (!fmt_info(in_format, &fi) || !fmt_info(out_format, &fo)) return -1;
    /* sink formats imp.rs:78-87, src formats imp.rs:89-96 */
    if (!(in_format == ORC_FMT_RGBX || in_format == ORC_FMT_XRGB || in_format == ORC_FMT_BGRX ||
          in_format == ORC_FMT_XBGR || in_format == ORC_FMT_RGB || in_format == ORC_FMT_BGR))
        return -1;
    if (!(out_format == ORC_FMT_RGBA || out_format == ORC_FMT_ARGB || out_format == ORC_FMT_BGRA ||
          out_format == ORC_FMT_ABGR))
        return -1;
    const orc_hsvdetector_params s = *p;
    for (uint32_t row = 0; row < height; row++) {
        const uint8_t *il = in + (size_t)row * in_stride;
        uint8_t *ol = out + (size_t)row * out_stride;
        for (uint32_t xpx = 0; xpx < width; xpx++) {
            const uint8_t *ip = il + (size_t)xpx * (size_t)fi.bpp + fi.off;
            uint8_t *op = ol + (size_t)xpx * 4;
            float hsv[3];
            uint8_t r, g, b;
            if (fi.bgr) {
                orc_hsv_from_bgr(ip, hsv);
                b = ip[0], g = ip[1], r = ip[2];
            } else {
                orc_hsv_from_rgb(ip, hsv);
                r = ip[0], g = ip[1], b = ip[2];
            }

            float ref_hue_offset = 180.0f - s.hue_ref; /* imp.rs:141 */
            float shifted_hue = hsv[0] + ref_hue_offset;
            if (shifted_hue < 0.0f) shifted_hue += 360.0f;
            shifted_hue = fmodf(shifted_hue, 360.0f);

            uint8_t val = (fabsf(shifted_hue - 180.0f) <= s.hue_var &&
                           fabsf(hsv[1] - s.saturation_ref) <= s.saturation_var &&
                           fabsf(hsv[2] - s.value_ref) <= s.value_var)
                              ? 255
                              : 0;

            uint8_t *oc = op + fo.off; /* colour bytes keep the input's R,G,B values */
            if (fo.bgr) {
                oc[0] = b, oc[1] = g, oc[2] = r;
            } else {
                oc[0] = r, oc[1] = g, oc[2] = b;
            }
            op[fo.alpha_i] = val;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* Frame-parallel drivers for bench.py only                            */
/* ------------------------------------------------------------------ */

typedef struct {
    int op; /* 0 colorlut, 1 hsvfilter, 2 hsvdetector */
    const orc_cube *lut;
    const uint8_t *const *src;
    uint8_t *const *dst;
    size_t n_frames, src_stride, dst_stride;
    uint32_t width, height;
    int in_format, out_format;
    const void *params;
    size_t next;
    pthread_mutex_t mu;
    int rc;
} mt_job;

static void *mt_worker(void *arg) {
    mt_job *j = (mt_job *)arg;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        size_t i = j->next++;
        pthread_mutex_unlock(&j->mu);
        if (i >= j->n_frames) break;
        int rc;
        if (j->op == 0)
            rc = orc_colorlut_frame(j->lut, j->src[i], j->src_stride, j->dst[i], j->dst_stride,
                                    j->width, j->height, j->in_format);
        else if (j->op == 1)
            rc = orc_hsvfilter_frame(j->dst[i], j->dst_stride, j->width, j->height, j->in_format,
                                     (const orc_hsvfilter_params *)j->params);
        else
            rc = orc_hsvdetector_frame(j->src[i], j->src_stride, j->in_format, j->dst[i],
                                       j->dst_stride, j->out_format, j->width, j->height,
                                       (const orc_hsvdetector_params *)j->params);
        if (rc) j->rc = rc;
    }
    return NULL;
}

static int mt_run(mt_job *j, int n_threads) {
    if (n_threads < 1) n_threads = 1;
    if (n_threads > 1024) n_threads = 1024;
    pthread_mutex_init(&j->mu, NULL);
    j->next = 0;
    j->rc = 0;
    if (n_threads == 1) {
        mt_worker(j);
    } else {
        pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)n_threads);
        int started = 0;
        for (int t = 0; t < n_threads; t++)
            if (pthread_create(&th[started], NULL, mt_worker, j) == 0) started++;
        if (started == 0) mt_worker(j);
        for (int t = 0; t < started; t++) pthread_join(th[t], NULL);
        free(th);
    }
    pthread_mutex_destroy(&j->mu);
    return j->rc;
}

int orc_colorlut_frames_mt(const orc_cube *lut, const uint8_t *const *src, uint8_t *const *dst,
                           size_t n_frames, size_t stride, uint32_t width, uint32_t height,
                           int format, int n_threads) {
    mt_job j = {0};
    j.op = 0, j.lut = lut, j.src = src, j.dst = dst, j.n_frames = n_frames;
    j.src_stride = j.dst_stride = stride, j.width = width, j.height = height, j.in_format = format;
    return mt_run(&j, n_threads);
}

int orc_hsvfilter_frames_mt(uint8_t *const *frames, size_t n_frames, size_t stride, uint32_t width,
                            uint32_t height, int format, const orc_hsvfilter_params *p,
                            int n_threads) {
    mt_job j = {0};
    j.op = 1, j.dst = frames, j.n_frames = n_frames, j.dst_stride = stride;
    j.width = width, j.height = height, j.in_format = format, j.params = p;
    return mt_run(&j, n_threads);
}

int orc_hsvdetector_frames_mt(const uint8_t *const *in, uint8_t *const *out, size_t n_frames,
                              size_t in_stride, int in_format, size_t out_stride, int out_format,
                              uint32_t width, uint32_t height, const orc_hsvdetector_params *p,
                              int n_threads) {
    mt_job j = {0};
    j.op = 2, j.src = in, j.dst = out, j.n_frames = n_frames, j.src_stride = in_stride;
    j.dst_stride = out_stride, j.width = width, j.height = height, j.in_format = in_format;
    j.out_format = out_format, j.params = p;
    return mt_run(&j, n_threads);
}
