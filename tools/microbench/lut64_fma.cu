// lut64_fma.cu — RGBA64 through a 3D LUT with every lerp contracted to one FMA:
// what would the relaxation of bit-exactness buy?  The library's own op (ColorLut64Op, compiled from
// csrc/ together with its table builder) against a copy whose 21 multiply-adds are FMAs:
//   a + d * tx            -> fma(d, tx, a)           (x-lerps on the delta table)
//   a + (b - a) * t       -> fma(b - a, t, a)        (y- and z-lerps)
// The library op is first checked against the library's direct kernel (all pixels equal), then the
// FMA copy against the library op (share of pixels that differ, largest difference in 16-bit codes).
// Content = the 8-bit classes of gst-plugins-rs_b200/frames.py widened by 257 (as bench.py does);
// 16 frames of 3840x2160 RGBA64_LE per launch; % of the measured HBM copy peak at 16 B per pixel.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -o lut64_fma lut64_fma.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../gst-plugins-rs_b200/csrc/vf_launch_colorlut.cu"

using namespace vf;

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

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16, x *= 0x7feb352dU, x ^= x >> 15, x *= 0x846ca68bU, x ^= x >> 16;
    return x;
}

__global__ void gen_kernel(uint2 *f, int cls, int frame) {
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
    if (x >= W) return;
    uint32_t r, g, b;
    if (cls == 0) {
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
    f[(size_t)y * W + x] = make_uint2(r * 257u | (g * 257u) << 16, b * 257u | 0xFFFF0000u);
}
// every 16-bit code on every axis somewhere: full-range random words (a second exactness pass)
__global__ void gen_random16_kernel(uint2 *f, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) f[i] = make_uint2(hash32((uint32_t)i * 2u + 1u), hash32((uint32_t)i * 2u + 0x9E3779B9u));
}

// ---- the library's ColorLut64Op with contracted lerps (generated from csrc/vf_ops.cuh: only the
// lines marked below differ) -------------------------------------------------------------------------
namespace vf {
template <bool BE, bool POW2, bool UNIT, int S>
struct ColorLut64FmaOp {
    static constexpr int kPixelBytes = 8;
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
        const f32x2 c0 = fma2(pk2(p.a0.z, p.a0.w), tx, pk2(p.a0.x, p.a0.y));   // a + d * t in one rounding
        const f32x2 c1 = fma2(pk2(p.a1.z, p.a1.w), tx, pk2(p.a1.x, p.a1.y));
        const f32x2 cb = fma2(pk2(p.b.z, p.b.w), tx, pk2(p.b.x, p.b.y));       // (B(y0), B(y0+1))
        rg = fma2(sub2(c1, c0), pk2(c.ty, c.ty), c0);
        b = __fmaf_rn(hi2(cb) - lo2(cb), c.ty, lo2(cb));
    }
    __device__ __forceinline__ uint2 eval(const Plane &p0, const Plane &p1, const Cell &c, uint32_t in_y) const {
        f32x2 rg0, rg1;
        float b0, b1;
        plane_xy(p0, c, rg0, b0);
        plane_xy(p1, c, rg1, b1);
        const f32x2 rg = fma2(sub2(rg1, rg0), pk2(c.tz, c.tz), rg0);
        const float bo = __fmaf_rn(b1 - b0, c.tz, b0);
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

}  // namespace vf

// ---- harness ---------------------------------------------------------------------------------
static float six_decimals(double v) {
    char buf[32];
    snprintf(buf, sizeof buf, "%.6f", v < 0 ? 0.0 : (v > 1 ? 1.0 : v));
    return strtof(buf, nullptr);
}
// synthetic .cube of SURVEY.md §8(d) as the padded pair-packed table {R(x), R(x+1), G(x), B(x)}
static std::vector<float4> make_lut(int n) {
    const double pi = 3.14159265358979323846;
    std::vector<float> l((size_t)n * n * n * 3);
    for (int z = 0; z < n; z++)
        for (int y = 0; y < n; y++)
            for (int x = 0; x < n; x++) {
                const double r = (double)x / (n - 1), g = (double)y / (n - 1), b = (double)z / (n - 1);
                float *e = &l[(((size_t)z * n + y) * n + x) * 3];
                e[0] = six_decimals(pow(r, 0.8) * 0.9 + 0.1 * g);
                e[1] = six_decimals(0.5 - 0.45 * cos(pi * g) + 0.05 * b);
                e[2] = six_decimals(pow(b, 1.2) * 0.85 + 0.15 * r);
            }
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

template <class F>
static float time_ms(F launch, int iters = 10) {
    for (int i = 0; i < 3; i++) launch();
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

// [0] pixels that differ, [1] largest difference of a 16-bit word
__global__ void diff_kernel(const uint2 *a, const uint2 *b, size_t n, unsigned long long *acc) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint2 x = a[i], y = b[i];
    if (x.x == y.x && x.y == y.y) return;
    atomicAdd(&acc[0], 1ull);
    const uint32_t xs[4] = {x.x & 0xFFFFu, x.x >> 16, x.y & 0xFFFFu, x.y >> 16};
    const uint32_t ys[4] = {y.x & 0xFFFFu, y.x >> 16, y.y & 0xFFFFu, y.y >> 16};
    unsigned long long m = 0;
    for (int c = 0; c < 4; c++) {
        const unsigned long long d = xs[c] > ys[c] ? xs[c] - ys[c] : ys[c] - xs[c];
        m = d > m ? d : m;
    }
    atomicMax(&acc[1], m);
}
struct Diff {
    unsigned long long differ, maxdiff;
};
static Diff compare(const void *a, const void *b, size_t n, unsigned long long *d_acc) {
    CK(cudaMemset(d_acc, 0, 16));
    diff_kernel<<<(unsigned)((n + 255) / 256), 256>>>((const uint2 *)a, (const uint2 *)b, n, d_acc);
    Diff d;
    CK(cudaMemcpy(&d, d_acc, 16, cudaMemcpyDeviceToHost));
    return d;
}
static double pct(float ms) { return (double)kPixels * 16.0 / (ms * 1e-3) / 1e9 / kPeak * 100.0; }

int main() {
    setvbuf(stdout, nullptr, _IOLBF, 0);
    constexpr int N = 33;
    const size_t frame_bytes = (size_t)W * H * 8, total = frame_bytes * NF;
    uint8_t *d_in, *d_out, *d_ref;
    unsigned long long *d_acc;
    CK(cudaMalloc(&d_in, total));
    CK(cudaMalloc(&d_out, total));
    CK(cudaMalloc(&d_ref, total));
    CK(cudaMalloc(&d_acc, 16));

    // what b200vf_colorlut_set_lut + ensure_lut64 (vf_abi.cpp) set up for a 33^3 LUT with the default domain
    const std::vector<float4> packed = make_lut(N);
    DeviceLut lut;
    lut.kind = 3, lut.size = N, lut.identity_domain = true, lut.unit_range = true;
    CK(cudaMalloc((void **)&lut.lut3d, packed.size() * 16));
    CK(cudaMemcpy(lut.lut3d, packed.data(), packed.size() * 16, cudaMemcpyHostToDevice));
    lut.sm1_pow2 = true;  // N - 1 = 32
    const double k = (double)(N - 1) / 65535.0;
    lut.k16_hi = (float)k, lut.k16_lo = (float)(k - (double)lut.k16_hi);
    lut.coords16_ok = true;  // vouched for below: the op must equal the direct kernel on every pixel
    lut.lut3d_d_stride = 65;
    CK(cudaMalloc((void **)&lut.lut3d_d, (size_t)65 * 65 * (N + 1) * 32));
    CK(cudaMemset(lut.lut3d_d, 0, (size_t)65 * 65 * (N + 1) * 32));
    CK(launch_build_lut64(0, lut, nullptr));
    CK(cudaDeviceSynchronize());

    ColorLutOp<16, false, true, true, 0> direct;
    direct.L = make_lut_args(lut);
    ColorLut64Op<false, true, true, 65> lib;
    lib.L = make_lut_args(lut);
    ColorLut64FmaOp<false, true, true, 65> fma;
    fma.L = make_lut_args(lut);

    auto run = [&](auto &op, uint8_t *out) {
        FrameSet fs;
        for (int f = 0; f < NF; f++) fs.in[f] = d_in + f * frame_bytes, fs.out[f] = out + f * frame_bytes;
        Geom g{(long long)W * 8, (long long)W * 8, W, H};
        CK(launch_map(0, fs, NF, g, 8, 8, op, nullptr));
    };

    // exactness on full-range random 16-bit words
    gen_random16_kernel<<<(unsigned)((kPixels + 255) / 256), 256>>>((uint2 *)d_in, kPixels);
    run(direct, d_ref);
    run(lib, d_out);
    const Diff dl = compare(d_out, d_ref, kPixels, d_acc);
    run(fma, d_ref);
    const Diff df = compare(d_ref, d_out, kPixels, d_acc);
    printf("%zu random RGBA64 pixels, 33^3 LUT: library op vs the direct kernel: %llu pixels differ; FMA copy vs library op: "
           "%llu pixels differ (%.4f %%), largest difference %llu code of 65535\n",
           kPixels, dl.differ, df.differ, 100.0 * (double)df.differ / (double)kPixels, df.maxdiff);

    const char *names[4] = {"bars", "grad", "noise", "rand"};
    printf("\n%% of the %.1f GB/s HBM copy peak at 16 B per pixel, %d frames of %dx%d RGBA64_LE per launch\n", kPeak, NF, W, H);
    printf("%-6s | %-9s %-9s %-9s | FMA copy vs library op: pixels differing, largest difference\n", "", "direct", "library", "FMA copy");
    for (int cls = 0; cls < 4; cls++) {
        for (int f = 0; f < NF; f++)
            gen_kernel<<<dim3((W + 255) / 256, H), 256>>>((uint2 *)(d_in + f * frame_bytes), cls, f);
        CK(cudaDeviceSynchronize());
        const float td = time_ms([&] { run(direct, d_ref); });
        const float tl = time_ms([&] { run(lib, d_ref); });
        const float tf = time_ms([&] { run(fma, d_out); });
        const Diff d = compare(d_out, d_ref, kPixels, d_acc);
        printf("%-6s | %7.1f %% %7.1f %% %7.1f %% | %llu (%.4f %%), %llu\n", names[cls], pct(td), pct(tl), pct(tf), d.differ,
               100.0 * (double)d.differ / (double)kPixels, d.maxdiff);
    }
    return 0;
}
