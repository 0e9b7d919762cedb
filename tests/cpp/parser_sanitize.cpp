// parser_sanitize.cpp — the product's .cube parser (csrc/vf_cube_parser.cpp, twin of
// video/colorlut/src/parser.rs:104-375) under AddressSanitizer + UBSan on hostile input: a LUT
// file is untrusted data that `colorlut location=…` reads in `start`.  Deterministic generator;
// checks that every outcome is either a clean rejection or a well-formed LUT.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../gst-plugins-rs_b200/csrc/vf_internal.h"

static uint64_t g_state = 0x5EED0000ull;
static uint64_t next_u64() {  // SplitMix64
    uint64_t z = (g_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static size_t below(size_t n) { return n ? (size_t)(next_u64() % n) : 0; }
// mostly printable ASCII so that mutations reach the grammar, sometimes any byte (UTF-8 checks)
static char fuzz_byte() { return below(4) ? (char)(9 + below(118)) : (char)below(256); }

static std::string base_file(int which) {
    std::string s;
    if (which == 0) {
        s = "TITLE \"fuzz\"\nLUT_3D_SIZE 2\nDOMAIN_MIN 0 0 0\nDOMAIN_MAX 1 1 1\n";
        for (int i = 0; i < 8; i++) s += "0.25 0.5 0.75\n";
    } else if (which == 1) {
        s = "# comment\nLUT_1D_SIZE 3\n0 0 0\n0.5 0.5 0.5\n1 1 1\n";
    } else if (which == 2) {
        s = "LUT_3D_SIZE 3\n";
        for (int i = 0; i < 27; i++) s += std::to_string(i / 26.0) + " 0 1e-3\n";
    } else {
        s = "LUT_1D_SIZE 2\r\nDOMAIN_MIN -1 -1 -1\r\n\r\n0 0 0\r\n1 1 1\r\n";
    }
    return s;
}

static const char *const kTokens[] = {
    "nan", "NaN", "inf", "-inf", "infinity", "-0", "+1", "1e999", "-1e999", "1e-999", "0x10", "1_000",
    "1.", ".5", "1e", "e5", "--1", "1,0", "\xEF\xBB\xBF", "\xC2\xA0", "\xE2\x80\xA8", "\xE3\x80\x80",
    "\t", "\r", "\n\n", "\"", "#", "LUT_3D_SIZE", "LUT_1D_SIZE", "DOMAIN_MIN", "DOMAIN_MAX", "TITLE",
    "LUT_3D_INPUT_RANGE", "LUT_3D_SIZE 257", "LUT_3D_SIZE 1", "LUT_1D_SIZE 65537", "LUT_1D_SIZE 0",
    "LUT_3D_SIZE 4294967297", "LUT_3D_SIZE 18446744073709551616", "LUT_3D_SIZE -2", "LUT_3D_SIZE 2.0",
    "DOMAIN_MAX 0 0 0", "DOMAIN_MIN 1 1", "\xFF", "\xC0\x80", "\xED\xA0\x80", "\xF4\x90\x80\x80", "\0x"};

static std::string mutate(std::string s) {
    const int n = 1 + (int)below(4);
    for (int k = 0; k < n; k++) {
        const size_t pos = below(s.size() + 1);
        switch (below(7)) {
        case 0: if (!s.empty()) s[below(s.size())] = fuzz_byte(); break;
        case 1: s.insert(pos, 1, fuzz_byte()); break;
        case 2: if (!s.empty()) s.erase(below(s.size()), 1 + below(8)); break;
        case 3: s.insert(pos, kTokens[below(sizeof kTokens / sizeof *kTokens)]); break;
        case 4: s.insert(pos, " " + std::string(kTokens[below(sizeof kTokens / sizeof *kTokens)]) + " "); break;
        case 5: s = s.substr(0, pos); break;                       // truncated file
        default: s += s.substr(below(s.size() + 1)); break;        // duplicated tail (too many rows)
        }
    }
    return s;
}

int main(int argc, char **argv) {
    const int iterations = argc > 1 ? std::atoi(argv[1]) : 20000;
    int accepted = 0, rejected = 0, io = 0;
    for (int i = 0; i < iterations; i++) {
        std::string text;
        if (i % 5 == 4) {  // raw bytes
            text.resize(below(200));
            for (char &c : text) c = (char)below(256);
        } else {
            text = mutate(base_file(i % 4));
        }
        vf::CubeData cd;
        std::string err;
        const int rc = vf::parse_cube_text(text.data(), text.size(), cd, err);
        if (rc == 0) {
            accepted++;
            const size_t n = cd.size;
            const bool ok = (cd.kind == 1 && n >= 2 && n <= 65536 && cd.data.size() == 3 * n) ||
                            (cd.kind == 3 && n >= 2 && n <= 256 && cd.data.size() == 4 * n * n * n);
            if (!ok) {
                std::printf("FAILED: accepted LUT is malformed (kind %d size %u floats %zu) at %d\n",
                            cd.kind, cd.size, cd.data.size(), i);
                return 1;
            }
            for (int c = 0; c < 3; c++)  // parser.rs:264-274: scale = 1/(max-min) is finite or rejected
                if (std::isnan(cd.domain_scale[c]) && !std::isnan(cd.domain_offset[c])) {
                    std::printf("FAILED: NaN domain scale with a non-NaN offset at %d\n", i);
                    return 1;
                }
        } else if (rc == 1) {
            rejected++;
            if (err.empty()) {
                std::printf("FAILED: rejection without a message at %d\n", i);
                return 1;
            }
        } else if (rc == 2) {
            io++;  // invalid UTF-8 = the reference's read_to_string failure
        } else {
            std::printf("FAILED: unknown status %d at %d\n", rc, i);
            return 1;
        }
    }
    // an empty and a NULL-free zero-length input
    vf::CubeData cd;
    std::string err;
    if (vf::parse_cube_text("", 0, cd, err) == 0) {
        std::printf("FAILED: empty input accepted\n");
        return 1;
    }
    std::printf("parser_sanitize: ok (%d inputs: %d accepted, %d rejected, %d not UTF-8)\n", iterations,
                accepted, rejected, io);
    return accepted > 0 && rejected > 0 && io > 0 ? 0 : 1;
}
