// vf_cube_parser.cpp — Adobe .cube parser of the product library.
//
// Behavioural twin of CubeLut::parse (video/colorlut/src/parser.rs:104-375),
// including its error taxonomy and message texts, so that
// `colorlut location=…` accepts and rejects exactly the files the reference does.
// Written against std::string_view; shares no code with oracle/.
#include <algorithm>
#include <cerrno>
#include <charconv>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <string_view>
#include <vector>
#include <new>

#include "vf_internal.h"

namespace vf {
namespace {

using sv = std::string_view;

constexpr size_t k1dMin = 2, k1dMax = 65536;  // parser.rs:12-13
constexpr size_t k3dMin = 2, k3dMax = 256;    // parser.rs:15-16

// --- UTF-8 / Unicode White_Space (str::trim, split_whitespace, read_to_string) ---

// Length of the scalar starting at s[0] and its code point; 0 if malformed.
size_t decode_scalar(sv s, char32_t &cp) {
    auto b = [&](size_t i) { return (unsigned char)s[i]; };
    if (s.empty()) return 0;
    unsigned char c = b(0);
    if (c < 0x80) return cp = c, 1;
    size_t n = (c >> 5) == 0x6 ? 2 : (c >> 4) == 0xE ? 3 : (c >> 3) == 0x1E ? 4 : 0;
    if (n == 0 || s.size() < n) return 0;
    char32_t v = c & (0xFF >> (n + 1));
    for (size_t i = 1; i < n; i++) {
        if ((b(i) & 0xC0) != 0x80) return 0;
        v = (v << 6) | (b(i) & 0x3F);
    }
    static const char32_t lowest[5] = {0, 0, 0x80, 0x800, 0x10000};
    if (v < lowest[n] || v > 0x10FFFF || (v >= 0xD800 && v < 0xE000)) return 0;
    return cp = v, n;
}

bool is_space(char32_t c) {
    switch (c) {
    case 0x09: case 0x0A: case 0x0B: case 0x0C: case 0x0D: case 0x20: case 0x85: case 0xA0:
    case 0x1680: case 0x2028: case 0x2029: case 0x202F: case 0x205F: case 0x3000:
        return true;
    default:
        return c >= 0x2000 && c <= 0x200A;
    }
}

bool valid_utf8(sv s) {
    while (!s.empty()) {
        char32_t cp;
        size_t n = decode_scalar(s, cp);
        if (!n) return false;
        s.remove_prefix(n);
    }
    return true;
}

sv trim(sv s) {
    char32_t cp;
    for (size_t n; !s.empty() && (n = decode_scalar(s, cp)) && is_space(cp);) s.remove_prefix(n);
    while (!s.empty()) {
        size_t i = s.size() - 1;
        while (i > 0 && ((unsigned char)s[i] & 0xC0) == 0x80) i--;
        if (!decode_scalar(s.substr(i), cp) || !is_space(cp)) break;
        s.remove_suffix(s.size() - i);
    }
    return s;
}

// split_whitespace() cursor
struct Words {
    sv rest;
    bool next(sv &word) {
        char32_t cp;
        for (size_t n; !rest.empty() && (n = decode_scalar(rest, cp)) && is_space(cp);)
            rest.remove_prefix(n);
        if (rest.empty()) return false;
        size_t len = 0;
        while (len < rest.size()) {
            size_t n = decode_scalar(rest.substr(len), cp);
            if (is_space(cp)) break;
            len += n;
        }
        word = rest.substr(0, len);
        rest.remove_prefix(len);
        return true;
    }
};

// --- Rust FromStr grammars ---------------------------------------------------

// ASCII case-insensitive compare against a lower-case literal (never consults the C locale)
bool ieq(sv a, const char *lit) {
    size_t n = std::strlen(lit);
    if (a.size() != n) return false;
    for (size_t i = 0; i < n; i++) {
        char ch = a[i];
        if (ch >= 'A' && ch <= 'Z') ch = (char)(ch - 'A' + 'a');
        if (ch != lit[i]) return false;
    }
    return true;
}

// Decimal order of magnitude of a token already checked against the grammar below: > 0 when the
// value is >= 1.  Only used to tell overflow from underflow when std::from_chars reports
// result_out_of_range (Rust's dec2flt returns inf / 0 there, it never fails).
long long decimal_magnitude(sv body) {
    long long point_pos = -1, first_nonzero = -1, n_digits = 0, exp10 = 0;
    size_t i = 0;
    for (; i < body.size() && (body[i] | 0x20) != 'e'; i++) {
        if (body[i] == '.') {
            point_pos = n_digits;
            continue;
        }
        if (body[i] != '0' && first_nonzero < 0) first_nonzero = n_digits;
        n_digits++;
    }
    if (point_pos < 0) point_pos = n_digits;
    if (i < body.size()) {
        i++;
        bool eneg = false;
        if (i < body.size() && (body[i] == '+' || body[i] == '-')) eneg = body[i++] == '-';
        for (; i < body.size(); i++) exp10 = std::min<long long>(exp10 * 10 + (body[i] - '0'), 1000000000LL);
        if (eneg) exp10 = -exp10;
    }
    return point_pos - first_nonzero + exp10;  // digits before the point, counted from the first non-zero one
}

// <f32 as FromStr>: [+-](inf|infinity|nan | digits[.digits][(e|E)[+-]digits]), ≥1 mantissa digit
bool parse_f32(sv t, float &out) {
    sv body = t;
    bool neg = false;
    if (!body.empty() && (body[0] == '+' || body[0] == '-')) {
        neg = body[0] == '-';
        body.remove_prefix(1);
    }
    if (body.empty()) return false;
    if (ieq(body, "inf") || ieq(body, "infinity")) return out = neg ? -INFINITY : INFINITY, true;
    if (ieq(body, "nan")) return out = NAN, true;
    size_t i = 0, mant = 0;
    auto digits = [&]() {
        size_t k = 0;
        while (i < body.size() && body[i] >= '0' && body[i] <= '9') i++, k++;
        return k;
    };
    mant += digits();
    if (i < body.size() && body[i] == '.') i++, mant += digits();
    if (!mant) return false;
    if (i < body.size() && (body[i] | 0x20) == 'e') {
        i++;
        if (i < body.size() && (body[i] == '+' || body[i] == '-')) i++;
        if (!digits()) return false;
    }
    if (i != body.size()) return false;
    // std::from_chars is correctly rounded like dec2flt and, unlike strtof, ignores LC_NUMERIC
    // (a process that called setlocale(LC_ALL, "") under de_DE would read "0.5" as 0 with strtof).
    float v = 0.0f;
    const std::from_chars_result r = std::from_chars(body.data(), body.data() + body.size(), v);
    if (r.ec == std::errc::result_out_of_range)
        v = decimal_magnitude(body) > 0 ? INFINITY : 0.0f;
    else if (r.ec != std::errc() || r.ptr != body.data() + body.size())
        return false;  // unreachable: the grammar was checked above
    out = neg ? -v : v;
    return true;
}

// `{:?}` of an f32 as Rust prints it (core::fmt::float, float_to_general_debug): the shortest
// digits that round-trip; exponential form `1e-5` / `1.5e16` when 0 < |v| < 1e-4 or |v| >= 1e16,
// otherwise decimal with at least one fractional digit (`1.0`); `NaN`, `inf`, `-inf`.
std::string rust_debug_f32(float v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v < 0 ? "-inf" : "inf";
    std::string out = std::signbit(v) ? "-" : "";
    const float a = std::fabs(v);
    if (a == 0.0f) return out + "0.0";
    char buf[64];
    const std::to_chars_result r = std::to_chars(buf, buf + sizeof buf, a, std::chars_format::scientific);
    sv sci(buf, (size_t)(r.ptr - buf));  // d[.ddd]e[+-]XX, shortest round-trip digits
    const size_t epos = sci.find('e');
    std::string digits;
    for (char ch : sci.substr(0, epos))
        if (ch != '.') digits.push_back(ch);
    int exp10 = 0;
    std::from_chars(sci.data() + epos + (sci[epos + 1] == '+' ? 2 : 1), sci.data() + sci.size(), exp10);
    if (a < 1e-4f || a >= 1e16f) {
        out += digits[0];
        if (digits.size() > 1) out += "." + digits.substr(1);
        return out + "e" + std::to_string(exp10);
    }
    if (exp10 < 0) return out + "0." + std::string((size_t)(-exp10 - 1), '0') + digits;
    if ((size_t)exp10 + 1 >= digits.size())
        return out + digits + std::string((size_t)exp10 + 1 - digits.size(), '0') + ".0";
    return out + digits.substr(0, (size_t)exp10 + 1) + "." + digits.substr((size_t)exp10 + 1);
}

// <usize as FromStr>: [+]digits, no overflow
bool parse_usize(sv t, size_t &out) {
    if (!t.empty() && t[0] == '+') t.remove_prefix(1);
    if (t.empty()) return false;
    unsigned long long v = 0;
    for (char ch : t) {
        if (ch < '0' || ch > '9') return false;
        unsigned d = (unsigned)(ch - '0');
        if (v > (ULLONG_MAX - d) / 10) return false;
        v = v * 10 + d;
    }
    out = (size_t)v;
    return true;
}

struct Fail {
    std::string msg;
};

std::string at_line(const char *what, size_t line_no, sv line) {
    return std::string("Invalid LUT: ") + what + " " + std::to_string(line_no) + ": " +
           std::string(line);
}

float need_f32(Words &w, size_t line_no, sv line) {
    sv tok;
    if (!w.next(tok)) throw Fail{at_line("Invalid line", line_no, line)};
    float v;
    if (!parse_f32(tok, v)) throw Fail{at_line("Invalid float at line", line_no, line)};
    return v;
}

void need_end(Words &w, size_t line_no, sv line) {
    sv tok;
    if (w.next(tok)) throw Fail{at_line("Invalid line", line_no, line)};
}

}  // namespace

int parse_cube_text(const char *text, size_t len, CubeData &out, std::string &err) {
    sv all(text, len);
    if (!valid_utf8(all)) {
        err = "IO error: stream did not contain valid UTF-8";
        return 2;
    }
    try {
        float dmin[3] = {0, 0, 0}, dmax[3] = {1, 1, 1};
        enum { Header, Lut1D, Lut3D } state = Header;
        size_t size = 0;
        bool have_data = false;
        std::vector<float> rgb;  // 3 per data line

        auto header_ok = [&](size_t line_no, sv line) {  // parser.rs:284-303
            if (state != Header && have_data)
                throw Fail{at_line("Header found after LUT data at line", line_no, line)};
        };

        size_t line_no = 0;
        for (sv rest = all; !rest.empty();) {  // str::lines()
            size_t nl = rest.find('\n');
            sv raw = rest.substr(0, nl);
            rest = nl == sv::npos ? sv() : rest.substr(nl + 1);
            line_no++;
            if (!raw.empty() && raw.back() == '\r') raw.remove_suffix(1);
            sv line = trim(raw);
            if (line.empty() || line.front() == '#') continue;

            Words w{line};
            sv first;
            if (!w.next(first)) continue;

            if (first == "TITLE") {
                header_ok(line_no, line);
            } else if (first == "DOMAIN_MIN" || first == "DOMAIN_MAX") {
                header_ok(line_no, line);
                float *d = first == "DOMAIN_MIN" ? dmin : dmax;
                float v0 = need_f32(w, line_no, line), v1 = need_f32(w, line_no, line),
                      v2 = need_f32(w, line_no, line);
                need_end(w, line_no, line);
                d[0] = v0, d[1] = v1, d[2] = v2;
            } else if (first == "LUT_1D_SIZE" || first == "LUT_3D_SIZE") {
                const bool one_d = first[4] == '1';
                header_ok(line_no, line);
                if (state != Header)
                    throw Fail{at_line(one_d ? "Invalid LUT_1D_SIZE at line"
                                             : "Invalid LUT_3D_SIZE at line",
                                       line_no, line)};
                sv tok;
                if (!w.next(tok)) throw Fail{at_line("Invalid line", line_no, line)};
                size_t n;
                if (!parse_usize(tok, n))
                    throw Fail{at_line("Invalid integer at line", line_no, line)};
                need_end(w, line_no, line);
                const size_t lo = one_d ? k1dMin : k3dMin, hi = one_d ? k1dMax : k3dMax;
                if (n < lo || n > hi)
                    throw Fail{"Invalid LUT: Invalid LUT size " + std::to_string(n) + " at line " +
                               std::to_string(line_no) + ", expected " + std::to_string(lo) +
                               "..=" + std::to_string(hi)};
                state = one_d ? Lut1D : Lut3D;
                size = n;
                have_data = false;
            } else {  // data line (parser.rs:177-201)
                if (state == Header)
                    throw Fail{at_line("LUT data found before LUT size at line", line_no, line)};
                have_data = true;
                float r;
                if (!parse_f32(first, r))
                    throw Fail{at_line("Invalid float at line", line_no, line)};
                float g = need_f32(w, line_no, line), b = need_f32(w, line_no, line);
                need_end(w, line_no, line);
                rgb.push_back(r), rgb.push_back(g), rgb.push_back(b);
            }
        }

        // parser.rs:205-212 (comparisons with NaN are false, as in Rust)
        if (dmin[0] >= dmax[0] || dmin[1] >= dmax[1] || dmin[2] >= dmax[2]) {
            auto arr = [](const float *d) {  // `{:?}` of [f32; 3]
                return "[" + rust_debug_f32(d[0]) + ", " + rust_debug_f32(d[1]) + ", " +
                       rust_debug_f32(d[2]) + "]";
            };
            throw Fail{"Invalid LUT: Invalid domain min " + arr(dmin) + ", max " + arr(dmax)};
        }
        if (state == Header) throw Fail{"Invalid LUT: Missing LUT size"};

        const size_t got = rgb.size() / 3;
        out = CubeData();
        out.size = (uint32_t)size;
        if (state == Lut1D) {
            if (got != size)
                throw Fail{"Invalid LUT: Invalid 1D LUT value count, expected " +
                           std::to_string(size) + ", got " + std::to_string(got)};
            out.kind = 1;
            out.data.resize(3 * size);
            for (size_t i = 0; i < size; i++)
                for (int c = 0; c < 3; c++) out.data[(size_t)c * size + i] = rgb[3 * i + c];
        } else {
            const size_t want = size * size * size;
            if (got != want)
                throw Fail{"Invalid LUT: Invalid 3D LUT value count, expected " +
                           std::to_string(want) + ", got " + std::to_string(got)};
            out.kind = 3;
            out.data.resize(4 * want);
            for (size_t i = 0; i < want; i++) {
                out.data[4 * i + 0] = rgb[3 * i + 0];
                out.data[4 * i + 1] = rgb[3 * i + 1];
                out.data[4 * i + 2] = rgb[3 * i + 2];
                out.data[4 * i + 3] = 1.0f;  // parser.rs:255
            }
        }
        for (int c = 0; c < 3; c++) {  // parser.rs:264-274
            out.domain_scale[c] = 1.0f / (dmax[c] - dmin[c]);
            out.domain_offset[c] = -dmin[c] * out.domain_scale[c];
        }
        return 0;
    } catch (const Fail &f) {
        err = f.msg;
        return 1;
    } catch (const std::bad_alloc &) {
        err = "IO error: out of memory";
        return 2;
    }
}

int parse_cube_file(const char *path, CubeData &out, std::string &err) {
    std::FILE *f = std::fopen(path, "rb");
    if (!f) {
        err = std::string("IO error: ") + std::strerror(errno);
        return 2;
    }
    std::string text;
    char buf[1 << 16];
    size_t n;
    while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, n);
    const bool bad = std::ferror(f) != 0;
    std::fclose(f);
    if (bad) {
        err = "IO error: read failed";
        return 2;
    }
    return parse_cube_text(text.data(), text.size(), out, err);
}

}  // namespace vf
