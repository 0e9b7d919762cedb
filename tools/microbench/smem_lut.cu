// smem_lut.cu — two experiments the round-1 review asked for by name (VERDICT r1 tasks 1c and 1b),
// as stand-alone kernels so that the answer is a measurement and not an argument:
//
//  A. "LUT staged into shared memory when it fits" (the north star's wording).  A 3D LUT applied by
//     exact trilinear interpolation (the reference's operation order, unfused) from
//       glob_f4  : the pair-packed f32 table in global memory (L1/L2-resident), N = 17 and N = 33
//       smem_f4  : the same f32 table staged into shared memory by every CTA, N = 17 (18^3 x 16 B =
//                  93 KB, two CTAs per SM) — bit-exact by construction, verified against glob_f4
//       smem_u16 : N = 33 packed to unorm16 x 3 (6 B per entry = 215.6 KB, the only packing of a
//                  33^3 RGB table that fits 227 KB), one 1024-thread CTA per SM — NOT bit-exact:
//                  exact-match fraction and max difference against glob_f4 are reported
//     next to the product's answer for 8-bit frames, the table baked to 2^24 entries (one gather).
//
//  B. "hsvdetector as a blocked bit table": the detector's alpha as one bit per colour triple (2 MiB,
//     8x8x4 colours per 32-byte sector) against the 4-byte blocked table (64 MiB, 4x4x2 per line).
//
// Content classes as in gst-plugins-rs_b200/frames.py (bars, grad, noise = grad +-2, rand); 16 frames
// of 3840x2160 RGBA (1.06 GB in + out > L2); % of the measured HBM copy peak at 8 B per pixel.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>
#include <vector>

#define CK(x)                                                                      \
    do {                                                                           \
        cudaError_t e_ = (x);                                                      \
        if (e_ != cudaSuccess) {                                                   \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            fflush(stdout);                                                        \
            exit(1);                                                               \
        }                                                                          \
    } while (0)

static const double kPeak = 6548.5;
constexpr int W = 3840, H = 2160, NF = 16;
constexpr size_t kPixels = (size_t)W * H * NF;
constexpr size_t kUnits = kPixels / 4;  // 16-byte units
constexpr int kThreads = 256;

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16, x *= 0x7feb352dU, x ^= x >> 15, x *= 0x846ca68bU, x ^= x >> 16;
    return x;
}

// ---- content ---------------------------------------------------------------------------------
__global__ void gen_kernel(uint32_t *f, int cls, int frame) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    uint32_t r, g, b;
    if (cls == 0) {  // bars
        const uint32_t bars[7] = {0xBFBFBF, 0x00BFBF, 0xBFBF00, 0x00BF00, 0xBF00BF, 0x0000BF, 0xBF0000};
        if (y < H * 2 / 3) {
            uint32_t c = bars[min(x * 7 / W, 6)];
            r = c & 255, g = (c >> 8) & 255, b = c >> 16;
        } else {
            r = g = b = x * 255 / (W - 1);
        }
    } else if (cls == 3) {
        uint32_t h = hash32((uint32_t)(y * W + x) * 2654435761u + frame * 97u);
        r = h & 255, g = (h >> 8) & 255, b = (h >> 16) & 255;
    } else {
        r = x * 255 / (W - 1), g = y * 255 / (H - 1), b = (x + y) * 255 / (W + H - 2);
        if (cls == 2) {
            uint32_t h = hash32((uint32_t)(y * W + x) * 2654435761u + frame * 97u);
            int amp = 2, span = 2 * amp + 1;
            r = (uint32_t)min(255, max(0, (int)r + (int)(h % span) - amp));
            g = (uint32_t)min(255, max(0, (int)g + (int)((h >> 8) % span) - amp));
            b = (uint32_t)min(255, max(0, (int)b + (int)((h >> 16) % span) - amp));
        }
    }
    f[(size_t)y * W + x] = r | g << 8 | b << 16 | 0xFF000000u;
}

// ---- the reference's arithmetic, shared by host and device ----------------
// Built with -fmad=false: every operation below is rounded separately on the device; the host
// compiler targets baseline x86-64 (no FMA instructions).
struct Coord {
    uint32_t i0;
    float t;
};
// RN(c / 255) without a division, as the product computes it: 1/255 split into hi + lo, one FMA
// (checked against the division for all 256 codes at start-up)
#define K255_HI ((float)(1.0 / 255.0))
#define K255_LO ((float)(1.0 / 255.0 - (double)((float)(1.0 / 255.0))))
__host__ __device__ __forceinline__ float div255(float c) { return fmaf(c, K255_HI, c * K255_LO); }
__host__ __device__ __forceinline__ Coord coord(uint32_t code, uint32_t n) {
    const float v = div255((float)code);
    const float p = v * ((float)n - 1.0f);
    uint32_t i0 = (uint32_t)floorf(p);
    if (i0 > n - 1) i0 = n - 1;
    Coord c;
    c.i0 = i0;
    c.t = p - (float)i0;
    return c;
}
__host__ __device__ __forceinline__ float lerp_ref(float a, float b, float t) { return a + (b - a) * t; }
__host__ __device__ __forceinline__ uint32_t to_u8(float v) {
    v = fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f;
    return (uint32_t)roundf(v);
}

// synthetic .cube of SURVEY.md §8(d), entries rounded to six decimals as the text file holds them
static float six_decimals(double v) {
    char buf[32];
    snprintf(buf, sizeof buf, "%.6f", v < 0 ? 0.0 : (v > 1 ? 1.0 : v));
    return strtof(buf, nullptr);
}
static std::vector<float> make_lut(int n) {  // [z][y][x][3], x = R fastest
    std::vector<float> l((size_t)n * n * n * 3);
    const double pi = 3.14159265358979323846;
    for (int z = 0; z < n; z++)
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++) {
                const double r = (double)x / (n - 1), g = (double)y / (n - 1), b = (double)z / (n - 1);
                float *e = &l[(((size_t)z * n + y) * n + x) * 3];
                e[0] = six_decimals(pow(r, 0.8) * 0.9 + 0.1 * g);
                e[1] = six_decimals(0.5 - 0.45 * cos(pi * g) + 0.05 * b);
                e[2] = six_decimals(pow(b, 1.2) * 0.85 + 0.15 * r);
            }
    return l;
}

// CPU: sample_3d with clamped i1, as the reference indexes it
static uint32_t ref_px(const std::vector<float> &l, uint32_t n, uint32_t in) {
    const Coord cx = coord(in & 255u, n), cy = coord((in >> 8) & 255u, n), cz = coord((in >> 16) & 255u, n);
    const uint32_t x1 = cx.i0 + 1 < n ? cx.i0 + 1 : n - 1, y1 = cy.i0 + 1 < n ? cy.i0 + 1 : n - 1,
                   z1 = cz.i0 + 1 < n ? cz.i0 + 1 : n - 1;
    uint32_t out = in & 0xFF000000u;
    for (int c = 0; c < 3; c++) {
        auto at = [&](uint32_t x, uint32_t y, uint32_t z) { return l[(((size_t)z * n + y) * n + x) * 3 + c]; };
        const float c00 = lerp_ref(at(cx.i0, cy.i0, cz.i0), at(x1, cy.i0, cz.i0), cx.t);
        const float c10 = lerp_ref(at(cx.i0, y1, cz.i0), at(x1, y1, cz.i0), cx.t);
        const float c01 = lerp_ref(at(cx.i0, cy.i0, z1), at(x1, cy.i0, z1), cx.t);
        const float c11 = lerp_ref(at(cx.i0, y1, z1), at(x1, y1, z1), cx.t);
        const float c0 = lerp_ref(c00, c10, cy.t), c1 = lerp_ref(c01, c11, cy.t);
        out |= to_u8(lerp_ref(c0, c1, cz.t)) << (8 * c);
    }
    return out;
}

// Pair-packed, edge-replicated table of (n+1)^3 entries {R(x), R(x+1), G(x), B(x)} (the product's layout)
static std::vector<float4> pack_f4(const std::vector<float> &l, int n) {
    const int s = n + 1;
    std::vector<float4> t((size_t)s * s * s);
    auto at = [&](int x, int y, int z, int c) {
        x = x < n ? x : n - 1, y = y < n ? y : n - 1, z = z < n ? z : n - 1;
        return l[(((size_t)z * n + y) * n + x) * 3 + c];
    };
    for (int z = 0; z < s; z++)
        for (int y = 0; y < s; y++)
            for (int x = 0; x < s; x++)
                t[((size_t)z * s + y) * s + x] = make_float4(at(x, y, z, 0), at(x + 1, y, z, 0), at(x, y, z, 1), at(x, y, z, 2));
    return t;
}

// unorm16 x 3 per entry, n^3 entries + 16 bytes of padding (the last pair's window)
static std::vector<uint16_t> pack_u16(const std::vector<float> &l, int n) {
    std::vector<uint16_t> t((size_t)n * n * n * 3 + 8, 0);
    for (size_t i = 0; i < (size_t)n * n * n * 3; i++) {
        float v = l[i];
        v = v < 0 ? 0 : (v > 1 ? 1 : v);
        t[i] = (uint16_t)lrintf(v * 65535.0f);
    }
    return t;
}

// ---- A: interpolating kernels ---------------------------------------------------------------------
template <int N>
__device__ __forceinline__ uint32_t px_f4(const float4 *lut, uint32_t in) {
    constexpr uint32_t S = N + 1;
    const Coord cx = coord(in & 255u, N), cy = coord((in >> 8) & 255u, N), cz = coord((in >> 16) & 255u, N);
    const float4 *b = lut + (cx.i0 + cy.i0 * S + cz.i0 * S * S);
    float r[4], g[4], bl[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const float4 *e = b + (k & 1) * S + (k >> 1) * S * S;
        const float4 a = e[0];
        const float2 n2 = reinterpret_cast<const float2 *>(e + 1)[1];
        r[k] = lerp_ref(a.x, a.y, cx.t);
        g[k] = lerp_ref(a.z, n2.x, cx.t);
        bl[k] = lerp_ref(a.w, n2.y, cx.t);
    }
    const float ro = lerp_ref(lerp_ref(r[0], r[1], cy.t), lerp_ref(r[2], r[3], cy.t), cz.t);
    const float go = lerp_ref(lerp_ref(g[0], g[1], cy.t), lerp_ref(g[2], g[3], cy.t), cz.t);
    const float bo = lerp_ref(lerp_ref(bl[0], bl[1], cy.t), lerp_ref(bl[2], bl[3], cy.t), cz.t);
    return to_u8(ro) | to_u8(go) << 8 | to_u8(bo) << 16 | (in & 0xFF000000u);
}

template <int N, bool SMEM, int T>
__global__ void __launch_bounds__(T) lut_f4_kernel(const uint4 *in, uint4 *out, size_t units, const float4 *glut) {
    extern __shared__ float4 s_lut[];
    constexpr uint32_t E = (N + 1) * (N + 1) * (N + 1);
    if (SMEM) {
        for (uint32_t i = threadIdx.x; i < E; i += T) s_lut[i] = glut[i];
        __syncthreads();
    }
    for (size_t u0 = (size_t)blockIdx.x * T * 2 + threadIdx.x; u0 < units; u0 += (size_t)gridDim.x * T * 2) {
        uint4 v[2];
#pragma unroll
        for (int j = 0; j < 2; j++)
            if (u0 + j * T < units) v[j] = __ldcs(in + u0 + j * T);
#pragma unroll
        for (int j = 0; j < 2; j++)
            if (u0 + j * T < units) {
                uint4 o;
                if (SMEM) {
                    o.x = px_f4<N>(s_lut, v[j].x), o.y = px_f4<N>(s_lut, v[j].y);
                    o.z = px_f4<N>(s_lut, v[j].z), o.w = px_f4<N>(s_lut, v[j].w);
                } else {
                    o.x = px_f4<N>(glut, v[j].x), o.y = px_f4<N>(glut, v[j].y);
                    o.z = px_f4<N>(glut, v[j].z), o.w = px_f4<N>(glut, v[j].w);
                }
                __stcs(out + u0 + j * T, o);
            }
    }
}

// unorm16 x 3 entries in shared memory.  The x-pair of a corner row is 12 contiguous bytes at 6 * idx
// (2-byte aligned): four aligned words and a funnel shift by 0 or 16 bits.  Values are interpolated
// in table units (0..65535, lerps contracted to FMAs — this path is approximate anyway) and scaled
// once at the end.
constexpr int kU16Threads = 1024;
__device__ __forceinline__ float unb_lo(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7610u)) - 8388608.0f; }
__device__ __forceinline__ float unb_hi(uint32_t w) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7632u)) - 8388608.0f; }
__device__ __forceinline__ float lerp_fma(float a, float b, float t) { return __fmaf_rn(b - a, t, a); }

template <int N>
__device__ __forceinline__ uint32_t px_u16(const uint32_t *s_w, uint32_t in) {
    const Coord cx = coord(in & 255u, N), cy = coord((in >> 8) & 255u, N), cz = coord((in >> 16) & 255u, N);
    const uint32_t dy = cy.i0 < N - 1 ? N : 0u, dz = cz.i0 < N - 1 ? N * N : 0u;  // no room for a padded table
    const uint32_t idx = cx.i0 + cy.i0 * N + cz.i0 * N * N;
    float r[4], g[4], bl[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t byte = 6u * (idx + (k & 1) * dy + (k >> 1) * dz);
        const uint32_t *w = s_w + (byte >> 2);
        const uint32_t sh = (byte & 2u) * 8u;
        const uint32_t w0 = w[0], w1 = w[1], w2 = w[2], w3 = w[3];
        const uint32_t a = __funnelshift_r(w0, w1, sh), b = __funnelshift_r(w1, w2, sh), c = __funnelshift_r(w2, w3, sh);
        r[k] = lerp_fma(unb_lo(a), unb_hi(b), cx.t);   // a = {R0, G0}, b = {B0, R1}, c = {G1, B1}
        g[k] = lerp_fma(unb_hi(a), unb_lo(c), cx.t);
        bl[k] = lerp_fma(unb_lo(b), unb_hi(c), cx.t);
    }
    const float k = 1.0f / 65535.0f;
    const float ro = lerp_fma(lerp_fma(r[0], r[1], cy.t), lerp_fma(r[2], r[3], cy.t), cz.t) * k;
    const float go = lerp_fma(lerp_fma(g[0], g[1], cy.t), lerp_fma(g[2], g[3], cy.t), cz.t) * k;
    const float bo = lerp_fma(lerp_fma(bl[0], bl[1], cy.t), lerp_fma(bl[2], bl[3], cy.t), cz.t) * k;
    return to_u8(ro) | to_u8(go) << 8 | to_u8(bo) << 16 | (in & 0xFF000000u);
}

template <int N>
__global__ void __launch_bounds__(kU16Threads, 1) lut_u16_kernel(const uint4 *in, uint4 *out, size_t units, const uint32_t *gw) {
    extern __shared__ uint32_t s_w[];
    constexpr uint32_t WORDS = (N * N * N * 6 + 16) / 4;
    for (uint32_t i = threadIdx.x; i < WORDS; i += kU16Threads) s_w[i] = gw[i];
    __syncthreads();
    for (size_t u = (size_t)blockIdx.x * kU16Threads + threadIdx.x; u < units; u += (size_t)gridDim.x * kU16Threads) {
        const uint4 v = __ldcs(in + u);
        uint4 o;
        o.x = px_u16<N>(s_w, v.x), o.y = px_u16<N>(s_w, v.y), o.z = px_u16<N>(s_w, v.z), o.w = px_u16<N>(s_w, v.w);
        __stcs(out + u, o);
    }
}

// ---- the product's answer for 8-bit frames: 2^24-entry table, 4x4x2 colour blocks per line, 2-D tiles
__host__ __device__ __forceinline__ uint32_t swz7(uint32_t x) {
    uint32_t j = x & 0x00FE0003u;
    j |= (x >> 5) & 0x000007F8u;
    j |= (x & 0x000000FCu) << 9;
    j |= (x >> 14) & 0x00000004u;
    return j;
}
template <int N>
__global__ void bake_kernel(uint32_t *t, const float4 *glut) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    t[swz7(i)] = px_f4<N>(glut, i) & 0xFFFFFFu;
}
// alpha = the detector stand-in of experiment B: "greenish" colours match
__host__ __device__ __forceinline__ bool detect(uint32_t x) {
    const int r = x & 255, g = (x >> 8) & 255, b = (x >> 16) & 255;
    const int m = r < b ? r : b;
    return g > r && g > b && g - m > 40;
}
__global__ void detect_table_kernel(uint32_t *t) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    t[swz7(i)] = (i & 0xFFFFFFu) | (detect(i) ? 0xFF000000u : 0u);
}
// bit table: bit = [c2 0][c1 1..0][c0 1..0] within a word (4x4x2 colours), word within a 32-byte
// sector = [c2 1][c1 2][c0 2] (a sector = 8x8x4 colours), then [c2 7..2][c1 7..3][c0 7..3]
__host__ __device__ __forceinline__ uint32_t bit_index(uint32_t x) {
    const uint32_t c0 = x & 255u, c1 = (x >> 8) & 255u, c2 = (x >> 16) & 255u;
    const uint32_t lo = (c0 & 3u) | (c1 & 3u) << 2 | (c2 & 1u) << 4;
    const uint32_t w3 = ((c0 >> 2) & 1u) | ((c1 >> 2) & 1u) << 1 | ((c2 >> 1) & 1u) << 2;
    const uint32_t hi = (c0 >> 3) | (c1 >> 3) << 5 | (c2 >> 2) << 10;
    return lo | w3 << 5 | hi << 8;
}
__global__ void detect_bits_kernel(uint32_t *bits) {  // one thread per word
    const uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t v = 0;
    for (uint32_t b = 0; b < 32; b++) {
        // invert bit_index for (w, b)
        const uint32_t j = w << 5 | b;
        const uint32_t lo = j & 31u, w3 = (j >> 5) & 7u, hi = j >> 8;
        const uint32_t c0 = (lo & 3u) | (w3 & 1u) << 2 | (hi & 31u) << 3;
        const uint32_t c1 = ((lo >> 2) & 3u) | ((w3 >> 1) & 1u) << 2 | ((hi >> 5) & 31u) << 3;
        const uint32_t c2 = ((lo >> 4) & 1u) | ((w3 >> 2) & 1u) << 1 | (hi >> 10) << 2;
        if (detect(c0 | c1 << 8 | c2 << 16)) v |= 1u << b;
    }
    bits[w] = v;
}

// MODE 0: 4-byte blocked table, colour from the table and alpha from the frame (colorlut)
// MODE 1: 4-byte blocked table, the whole word from the table (hsvdetector through its function table)
// MODE 2: bit table, colour bytes from the frame, alpha from the bit
template <int MODE>
__global__ void __launch_bounds__(kThreads, 8)
    tile_kernel(const uint8_t *in, uint8_t *out, size_t frame_bytes, uint32_t units_per_row, uint32_t rows, long long stride,
                const uint32_t *table) {
    constexpr int TWU = 16, RPP = kThreads / TWU;
    const uint32_t ux = threadIdx.x % TWU, uy = threadIdx.x / TWU;
    const uint32_t x = blockIdx.x * TWU + ux;
    const uint32_t y0 = blockIdx.y * (RPP * 4) + uy;
    const uint8_t *src = in + (size_t)blockIdx.z * frame_bytes + (size_t)x * 16;
    uint8_t *dst = out + (size_t)blockIdx.z * frame_bytes + (size_t)x * 16;
    uint4 v[4];
    bool ok[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint32_t y = y0 + j * RPP;
        ok[j] = y < rows;  // 960 units per row = 60 full tile columns
        if (ok[j]) v[j] = __ldcs(reinterpret_cast<const uint4 *>(src + (size_t)y * stride));
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
        if (ok[j]) {
            const uint32_t p[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
            uint32_t o[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                if (MODE == 0) {
                    o[k] = __byte_perm(__ldg(table + swz7(p[k])), p[k], 0x7210u);
                } else if (MODE == 1) {
                    o[k] = __ldg(table + swz7(p[k]));
                } else {
                    const uint32_t j2 = bit_index(p[k]);
                    const uint32_t wv = __ldg(table + (j2 >> 5));
                    const uint32_t a = 0u - ((wv >> (j2 & 31u)) & 1u);  // 0 or 0xFFFFFFFF
                    o[k] = __byte_perm(p[k], a, 0x4210u);
                }
            }
            __stcs(reinterpret_cast<uint4 *>(dst + (size_t)(y0 + j * RPP) * stride), make_uint4(o[0], o[1], o[2], o[3]));
        }
}

// ---- harness ---------------------------------------------------------------------------------
static int g_warm = 3, g_iters = 10;  // SMEM_LUT_ONCE=1 (under ncu): every kernel once, random colours only
template <class F>
static float time_ms(F launch) {
    const int iters = g_iters;
    for (int i = 0; i < g_warm; i++) launch();
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; i++) launch();
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0), cudaEventDestroy(e1);
    return ms / iters;
}

// per-byte comparison of two frame sets: [0] pixels that differ, [1] largest byte difference
__global__ void diff_kernel(const uint32_t *a, const uint32_t *b, size_t n, unsigned long long *acc) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t x = a[i], y = b[i];
    if (x == y) return;
    atomicAdd(&acc[0], 1ull);
    unsigned long long m = 0;
    for (int c = 0; c < 4; c++) {
        const int d = (int)((x >> (8 * c)) & 255u) - (int)((y >> (8 * c)) & 255u);
        const unsigned long long ad = (unsigned long long)(d < 0 ? -d : d);
        m = ad > m ? ad : m;
    }
    atomicMax(&acc[1], m);
}

struct Diff {
    unsigned long long differ, maxdiff;
};
static Diff compare(const uint8_t *a, const uint8_t *b, unsigned long long *d_acc) {
    CK(cudaMemset(d_acc, 0, 16));
    diff_kernel<<<(unsigned)((kPixels + 255) / 256), 256>>>((const uint32_t *)a, (const uint32_t *)b, kPixels, d_acc);
    Diff d;
    CK(cudaMemcpy(&d, d_acc, 16, cudaMemcpyDeviceToHost));
    return d;
}

static double pct(float ms) { return (double)kPixels * 8.0 / (ms * 1e-3) / 1e9 / kPeak * 100.0; }

int main() {
    setvbuf(stdout, nullptr, _IOLBF, 0);
    const bool once = getenv("SMEM_LUT_ONCE") != nullptr;
    if (once) g_warm = 0, g_iters = 1;
    for (int c = 0; c < 256; c++)
        if (div255((float)c) != (float)c / 255.0f) {
            printf("div255(%d) differs from the division\n", c);
            return 1;
        }
    {   // host self-check: an identity LUT returns its input, a grid point its entry
        std::vector<float> id((size_t)17 * 17 * 17 * 3);
        for (int z = 0; z < 17; z++)
            for (int y = 0; y < 17; y++)
                for (int x = 0; x < 17; x++) {
                    float *e = &id[(((size_t)z * 17 + y) * 17 + x) * 3];
                    e[0] = six_decimals(x / 16.0), e[1] = six_decimals(y / 16.0), e[2] = six_decimals(z / 16.0);
                }
        for (uint32_t i = 0; i < (1u << 24); i += 4099u)
            if (ref_px(id, 17, i | 0xFF000000u) != (i | 0xFF000000u)) {
                printf("host reference: identity LUT changed colour %06x\n", i);
                return 1;
            }
        printf("host self-check ok (exact /255 split, identity LUT)\n");
    }
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("device: %s, %d SMs, %zu KB shared memory per block (opt-in)\n", prop.name, sms, prop.sharedMemPerBlockOptin / 1024);
    const size_t frame_bytes = (size_t)W * H * 4, total = frame_bytes * NF;
    uint8_t *d_in, *d_out, *d_ref;
    unsigned long long *d_acc;
    CK(cudaMalloc(&d_in, total));
    CK(cudaMalloc(&d_out, total));
    CK(cudaMalloc(&d_ref, total));
    CK(cudaMalloc(&d_acc, 16));

    // tables
    const std::vector<float> l17 = make_lut(17), l33 = make_lut(33);
    const std::vector<float4> f17 = pack_f4(l17, 17), f33 = pack_f4(l33, 33);
    const std::vector<uint16_t> u33 = pack_u16(l33, 33);
    float4 *d_f17, *d_f33;
    uint32_t *d_u33, *d_baked, *d_det, *d_bits;
    CK(cudaMalloc(&d_f17, f17.size() * 16));
    CK(cudaMalloc(&d_f33, f33.size() * 16));
    CK(cudaMalloc(&d_u33, u33.size() * 2));
    CK(cudaMalloc(&d_baked, (size_t)4 << 24));
    CK(cudaMalloc(&d_det, (size_t)4 << 24));
    CK(cudaMalloc(&d_bits, (size_t)2 << 20));
    CK(cudaMemcpy(d_f17, f17.data(), f17.size() * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_f33, f33.data(), f33.size() * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_u33, u33.data(), u33.size() * 2, cudaMemcpyHostToDevice));
    bake_kernel<33><<<(1 << 24) / 256, 256>>>(d_baked, d_f33);
    detect_table_kernel<<<(1 << 24) / 256, 256>>>(d_det);
    detect_bits_kernel<<<(1 << 19) / 256, 256>>>(d_bits);
    CK(cudaDeviceSynchronize());

    const size_t smem17 = f17.size() * 16, smem33 = u33.size() * 2;
    CK(cudaFuncSetAttribute(lut_f4_kernel<17, true, 256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem17));
    CK(cudaFuncSetAttribute(lut_f4_kernel<17, true, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem17));
    CK(cudaFuncSetAttribute(lut_u16_kernel<33>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem33));
    printf("shared memory: smem_f4 N=17 %zu B per CTA (2 CTAs/SM), smem_u16 N=33 %zu B per CTA (1 CTA/SM)\n", smem17, smem33);

    const uint32_t units_per_row = W / 4;
    const dim3 tgrid((units_per_row + 15) / 16, (H + 63) / 64, NF);
    const char *names[4] = {"bars", "grad", "noise", "rand"};
    printf("\n%% of the %.1f GB/s HBM copy peak at 8 B per pixel, %d frames of %dx%d RGBA per launch\n", kPeak, NF, W, H);
    printf("%-6s | %-8s %-8s %-8s | %-8s %-8s %-8s | %-9s %-8s\n", "", "glob17", "smem17", "smem17x2", "glob33", "smemu16", "baked33", "det 4B", "det bit");
    printf("(smem17: 2 CTAs x 256 threads per SM; smem17x2: 2 CTAs x 512 threads; smemu16: 1 CTA x 1024 threads)\n");
    for (int cls = once ? 3 : 0; cls < 4; cls++) {
        for (int f = 0; f < NF; f++)
            gen_kernel<<<dim3((W + 255) / 256, H), 256>>>((uint32_t *)(d_in + f * frame_bytes), cls, f);
        CK(cudaDeviceSynchronize());
        const uint4 *in = (const uint4 *)d_in;
        uint4 *out = (uint4 *)d_out, *ref = (uint4 *)d_ref;

        // N = 17: global f32 table (reference for smem_f4) — and its check against the CPU on frame 0's first megapixel
        lut_f4_kernel<17, false, 256><<<sms * 8, 256>>>(in, ref, kUnits, d_f17);
        CK(cudaDeviceSynchronize());
        {
            const size_t n = 1 << 20;
            std::vector<uint32_t> hi(n), ho(n);
            CK(cudaMemcpy(hi.data(), d_in, n * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(ho.data(), d_ref, n * 4, cudaMemcpyDeviceToHost));
            size_t bad = 0;
            for (size_t i = 0; i < n; i++) bad += ref_px(l17, 17, hi[i]) != ho[i];
            if (bad) printf("  !! %s: glob17 differs from the CPU reference on %zu of %zu pixels\n", names[cls], bad, n);
        }
        const float t_g17 = time_ms([&] { lut_f4_kernel<17, false, 256><<<sms * 8, 256>>>(in, ref, kUnits, d_f17); });
        const float t_s17 = time_ms([&] { lut_f4_kernel<17, true, 256><<<sms * 2, 256, smem17>>>(in, out, kUnits, d_f17); });
        CK(cudaGetLastError());
        const Diff d17 = compare(d_out, d_ref, d_acc);
        const float t_s17b = time_ms([&] { lut_f4_kernel<17, true, 512><<<sms * 2, 512, smem17>>>(in, out, kUnits, d_f17); });
        CK(cudaGetLastError());
        const Diff d17b = compare(d_out, d_ref, d_acc);

        // N = 33: global f32 table = the exact answer; packed table in shared memory; baked table
        lut_f4_kernel<33, false, 256><<<sms * 8, 256>>>(in, ref, kUnits, d_f33);
        CK(cudaDeviceSynchronize());
        {
            const size_t n = 1 << 20;
            std::vector<uint32_t> hi(n), ho(n);
            CK(cudaMemcpy(hi.data(), d_in, n * 4, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(ho.data(), d_ref, n * 4, cudaMemcpyDeviceToHost));
            size_t bad = 0;
            for (size_t i = 0; i < n; i++) bad += ref_px(l33, 33, hi[i]) != ho[i];
            if (bad) printf("  !! %s: glob33 differs from the CPU reference on %zu of %zu pixels\n", names[cls], bad, n);
        }
        const float t_g33 = time_ms([&] { lut_f4_kernel<33, false, 256><<<sms * 8, 256>>>(in, ref, kUnits, d_f33); });
        const float t_u16 = time_ms([&] { lut_u16_kernel<33><<<sms, kU16Threads, smem33>>>(in, out, kUnits, d_u33); });
        CK(cudaGetLastError());
        const Diff du = compare(d_out, d_ref, d_acc);
        const float t_bk = time_ms([&] {
            tile_kernel<0><<<tgrid, kThreads>>>(d_in, d_out, frame_bytes, units_per_row, H, (long long)W * 4, d_baked);
        });
        CK(cudaGetLastError());
        const Diff db = compare(d_out, d_ref, d_acc);

        // B: detector through the 4-byte table and through the bit table
        const float t_d4 = time_ms([&] {
            tile_kernel<1><<<tgrid, kThreads>>>(d_in, d_ref, frame_bytes, units_per_row, H, (long long)W * 4, d_det);
        });
        const float t_db = time_ms([&] {
            tile_kernel<2><<<tgrid, kThreads>>>(d_in, d_out, frame_bytes, units_per_row, H, (long long)W * 4, d_bits);
        });
        CK(cudaGetLastError());
        const Diff dd = compare(d_out, d_ref, d_acc);

        printf("%-6s | %6.1f %% %6.1f %% %6.1f %% | %6.1f %% %6.1f %% %6.1f %% | %7.1f %% %6.1f %%\n", names[cls], pct(t_g17), pct(t_s17),
               pct(t_s17b), pct(t_g33), pct(t_u16), pct(t_bk), pct(t_d4), pct(t_db));
        printf("         smem17 vs glob17: %llu pixels differ | smemu16 vs exact: %.4f %% exact, max diff %llu | baked vs exact: %llu differ | bit vs 4B detector: %llu differ\n",
               d17.differ + d17b.differ, 100.0 * (1.0 - (double)du.differ / (double)kPixels), du.maxdiff, db.differ, dd.differ);
    }
    return 0;
}
