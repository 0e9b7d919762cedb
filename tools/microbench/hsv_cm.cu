// hsv_cm.cu — VERDICT r1 task 1, last paragraph: "s', v', c = v'·s' and m = v' − c depend on (max, min)
// only → a 64 K-entry {c, m} table per settings change removes both div_exacts of the saturation
// path and the adjust step; measure."
//
// The library's own compute kernel (vf_map_vec_kernel<HsvFilterFastOp>, included from csrc/) against
// two variants of it that fetch what depends on (max, min) from a table filled with the library's
// own instruction sequence:
//   cm8  : entry {c, m} (8 bytes, 512 KB)
//   cm16 : entry {c, m, A | C << 8, -} with the two finished output codes that do not depend on the
//          hue, A = floor((c + m) * 255) and C = floor(m * 255) (16 bytes, 1 MB)
// Every variant must reproduce the library kernel's bytes on all 2^24 colours and on the timed
// frames.  Content classes as in gst-plugins-rs_b200/frames.py; 16 frames of 3840x2160 RGBA per
// launch; % of the measured HBM copy peak at 8 B per pixel.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -prec-div=true -ftz=false -o hsv_cm hsv_cm.cu
#include <cstdio>
#include <cstdlib>

#include "../../gst-plugins-rs_b200/csrc/vf_ops.cuh"

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

__global__ void gen_kernel(uint32_t *f, int cls, int frame) {
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
    f[(size_t)y * W + x] = r | g << 8 | b << 16 | 0xFF000000u;
}
__global__ void all_colours_kernel(uint32_t *f) {  // 4096 x 4096: every colour triple once
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    f[i] = i | 0x5A000000u;
}

// ---- the (max, min) table, filled with the library's own sequence ----------------------------------
struct CmEntry16 {
    float c, m;
    uint32_t ac, pad;
};

// entry [mx * 256 + mn], mn <= mx: from_rgb_fast2's value / chroma / saturation, hsv_adjust_fast's
// s' and v', to_rgb_fast's c, m and the two hue-independent output codes
__global__ void build_cm_kernel(CmEntry16 *t16, float2 *t8, HsvFilterParams p) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t mx = i >> 8, mn = i & 255u;
    CmEntry16 e = {0.0f, 0.0f, 0u, 0u};
    if (mn <= mx) {
        const float value = div255_exact((float)mx);
        const float chroma = value - div255_exact((float)mn);
        Hsv a;
        a.h = 0.0f;
        a.s = div_exact(chroma, value + VF_FLT_MIN);
        a.v = value;
        const Hsv b = hsv_adjust_fast<kAngleZero>(a, p);
        const float c = __fmul_rn(b.v, b.s);
        const float m = b.v - c;
        const uint32_t A = __float_as_uint(__fadd_rd(__fmul_rn(c + m, 255.0f), VF_MAGIC));
        const uint32_t C = __float_as_uint(__fadd_rd(__fmul_rn(m, 255.0f), VF_MAGIC));
        e.c = c, e.m = m, e.ac = (A & 255u) | (C & 255u) << 8;
    }
    t16[i] = e;
    t8[i] = make_float2(e.c, e.m);
}

// hue of from_rgb_fast2 (vf_math.cuh), from the quotients and value / chroma + FLT_MIN
__device__ __forceinline__ float hue_of(float r, float g, float b, float value, float dc) {
    float hue;
    asm("{\n\t"
        ".reg .pred pr, pg, pn;\n\t"
        ".reg .f32 num, y0, q0, rr, q, hp;\n\t"
        "setp.eq.f32 pr, %1, %4;\n\t"
        "setp.eq.f32 pg, %2, %4;\n\t"
        "sub.rn.f32 num, %1, %2;\n\t"
        "@pg sub.rn.f32 num, %3, %1;\n\t"
        "@pr sub.rn.f32 num, %2, %3;\n\t"
        "rcp.approx.ftz.f32 y0, %5;\n\t"
        "mul.rn.f32 q0, num, y0;\n\t"
        "neg.f32 rr, %5;\n\t"
        "fma.rn.f32 rr, rr, q0, num;\n\t"
        "fma.rn.f32 q, rr, y0, q0;\n\t"
        "add.rn.f32 hp, q, 0f40800000;\n\t"
        "@pg add.rn.f32 hp, q, 0f40000000;\n\t"
        "@pr mov.f32 hp, q;\n\t"
        "mul.rn.f32 %0, hp, 0f42700000;\n\t"
        "setp.lt.f32 pn, %0, 0f00000000;\n\t"
        "@pn add.rn.f32 %0, %0, 0f43B40000;\n\t"
        "}"
        : "=f"(hue)
        : "f"(r), "f"(g), "f"(b), "f"(value), "f"(dc));
    return hue;
}

template <int KIND, int RI, int GI, int BI, bool E16, int MINB = 0>
struct HsvFilterCmOp {
    static constexpr int kPixelBytes = 4;
    static constexpr int kMinBlocks = MINB;  // register budget: 6 CTAs per SM = 40 registers (unbounded: 44 = 5 CTAs; 8 CTAs = 32 registers spills)
    HsvFilterParams p;  // hue_shift only; the rest is in the table
    const void *table;

    __device__ __forceinline__ void init(TabEntry *tab) const {
        HsvFilterFastOp<KIND, RI, GI, BI> base;
        base.init(tab);
    }

    __device__ __forceinline__ uint32_t px(uint32_t in, const TabEntry *tab) const {
        const float r8 = byte_to_float(in, RI), g8 = byte_to_float(in, GI), b8 = byte_to_float(in, BI);
        const float mx8 = fmaxf(r8, fmaxf(g8, b8)), mn8 = fminf(r8, fminf(g8, b8));
        const uint32_t idx = __float2uint_rz(__fmaf_rn(mx8, 256.0f, mn8));
        const float r = div255_exact(r8), g = div255_exact(g8), b = div255_exact(b8);
        // division is monotone: the quotient of the max is the max of the quotients
        const float value = div255_exact(mx8);
        const float chroma = value - div255_exact(mn8);
        const float hue = hue_of(r, g, b, value, chroma + VF_FLT_MIN);
        float h;
        if (KIND == kAngleGeneric) {
            h = fmodf(hue + p.hue_shift, 360.0f);
            if (h < 0.0f) h += 360.0f;
        } else {
            h = add_angle<KIND>(hue, p.hue_shift);
        }
        const float hp = div60_exact(h);
        const uint32_t k = KIND != kAngleGeneric ? (uint32_t)__float2int_ru(hp)
                                                 : (__float_as_uint(__fadd_ru(hp, VF_MAGIC)) & 7u);
        const SectorEntry e = tab[k];
        const float t = fabsf(hp - e.center);
        uint32_t abc;
        if (E16) {
            const uint4 q = __ldg(reinterpret_cast<const uint4 *>(table) + idx);
            const float c = __uint_as_float(q.x), m = __uint_as_float(q.y);
            const float x = __fmul_rn(c, 1.0f - t);
            const uint32_t B = __float_as_uint(__fadd_rd(__fmul_rn(x + m, 255.0f), VF_MAGIC));
            abc = __byte_perm(q.z, B, 0x0140u);  // [A, B, C, .]
        } else {
            const float2 q = __ldg(reinterpret_cast<const float2 *>(table) + idx);
            const float c = q.x, m = q.y;
            const float x = __fmul_rn(c, 1.0f - t);
            const uint32_t A = __float_as_uint(__fadd_rd(__fmul_rn(c + m, 255.0f), VF_MAGIC));
            const uint32_t B = __float_as_uint(__fadd_rd(__fmul_rn(x + m, 255.0f), VF_MAGIC));
            const uint32_t C = __float_as_uint(__fadd_rd(__fmul_rn(m, 255.0f), VF_MAGIC));
            abc = __byte_perm(__byte_perm(A, B, 0x0040u), C, 0x0410u);
        }
        return prmt(abc, in, e.sel);
    }
};

// ---- harness ---------------------------------------------------------------------------------
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

__global__ void diff_kernel(const uint32_t *a, const uint32_t *b, size_t n, unsigned long long *cnt) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && a[i] != b[i]) atomicAdd(cnt, 1ull);
}
static unsigned long long differ(const void *a, const void *b, size_t n, unsigned long long *d_cnt) {
    CK(cudaMemset(d_cnt, 0, 8));
    diff_kernel<<<(unsigned)((n + 255) / 256), 256>>>((const uint32_t *)a, (const uint32_t *)b, n, d_cnt);
    unsigned long long h;
    CK(cudaMemcpy(&h, d_cnt, 8, cudaMemcpyDeviceToHost));
    return h;
}
static double pct(float ms) { return (double)kPixels * 8.0 / (ms * 1e-3) / 1e9 / kPeak * 100.0; }

int main() {
    setvbuf(stdout, nullptr, _IOLBF, 0);
    const size_t frame_bytes = (size_t)W * H * 4, total = frame_bytes * NF;
    uint8_t *d_in, *d_out, *d_ref;
    unsigned long long *d_cnt;
    CmEntry16 *d_t16;
    float2 *d_t8;
    CK(cudaMalloc(&d_in, total));
    CK(cudaMalloc(&d_out, total));
    CK(cudaMalloc(&d_ref, total));
    CK(cudaMalloc(&d_cnt, 8));
    CK(cudaMalloc(&d_t16, sizeof(CmEntry16) << 16));
    CK(cudaMalloc(&d_t8, sizeof(float2) << 16));

    // BASELINE configs[1] settings: hue-shift 37.5 (kAngleNonNeg), saturation x 1.2 + 0.05, value x 0.9 + 0.02
    const HsvFilterArgs args = {37.5f, 1.2f, 0.05f, 0.9f, 0.02f};
    const HsvFilterParams p = make_filter_params(args);
    build_cm_kernel<<<(1 << 16) / 256, 256>>>(d_t16, d_t8, p);
    CK(cudaDeviceSynchronize());

    HsvFilterFastOp<kAngleNonNeg, 0, 1, 2> base;
    base.p = p;
    HsvFilterCmOp<kAngleNonNeg, 0, 1, 2, false> cm8;
    cm8.p = p, cm8.table = d_t8;
    HsvFilterCmOp<kAngleNonNeg, 0, 1, 2, true> cm16;
    cm16.p = p, cm16.table = d_t16;
    HsvFilterCmOp<kAngleNonNeg, 0, 1, 2, true, 6> cm16o;
    cm16o.p = p, cm16o.table = d_t16;

    auto run = [&](auto &op, const uint8_t *in, uint8_t *out, int n, uint32_t w, uint32_t h) {
        FrameSet fs;
        for (int f = 0; f < n; f++) fs.in[f] = in + (size_t)f * w * h * 4, fs.out[f] = out + (size_t)f * w * h * 4;
        Geom g{(long long)w * 4, (long long)w * 4, w, h};
        CK(launch_map(0, fs, n, g, 4, 4, op, nullptr));
    };

    // exactness on every colour triple (4096 x 4096 frame; fits the 16-frame buffers)
    all_colours_kernel<<<(1 << 24) / 256, 256>>>((uint32_t *)d_in);
    run(base, d_in, d_ref, 1, 4096, 4096);
    run(cm8, d_in, d_out, 1, 4096, 4096);
    const unsigned long long bad8 = differ(d_out, d_ref, (size_t)1 << 24, d_cnt);
    run(cm16, d_in, d_out, 1, 4096, 4096);
    const unsigned long long bad16 = differ(d_out, d_ref, (size_t)1 << 24, d_cnt);
    run(cm16o, d_in, d_out, 1, 4096, 4096);
    const unsigned long long bad16o = differ(d_out, d_ref, (size_t)1 << 24, d_cnt);
    const unsigned long long moved = differ(d_in, d_ref, (size_t)1 << 24, d_cnt);
    printf("all 2^24 colours: cm8 differs from the library kernel on %llu, cm16 on %llu, cm16 at 6 CTAs/SM on %llu (the filter changes %llu of them)\n",
           bad8, bad16, bad16o, moved);

    const char *names[4] = {"bars", "grad", "noise", "rand"};
    printf("\n%% of the %.1f GB/s HBM copy peak at 8 B per pixel, %d frames of %dx%d RGBA per launch\n", kPeak, NF, W, H);
    printf("%-6s | %-9s %-9s %-9s %-9s | pixels differing from the library kernel\n", "", "library", "cm8", "cm16", "cm16@6");
    for (int cls = 0; cls < 4; cls++) {
        for (int f = 0; f < NF; f++)
            gen_kernel<<<dim3((W + 255) / 256, H), 256>>>((uint32_t *)(d_in + f * frame_bytes), cls, f);
        CK(cudaDeviceSynchronize());
        const float t0 = time_ms([&] { run(base, d_in, d_ref, NF, W, H); });
        const float t8 = time_ms([&] { run(cm8, d_in, d_out, NF, W, H); });
        const unsigned long long b8 = differ(d_out, d_ref, kPixels, d_cnt);
        const float t16 = time_ms([&] { run(cm16, d_in, d_out, NF, W, H); });
        const unsigned long long b16 = differ(d_out, d_ref, kPixels, d_cnt);
        const float t16o = time_ms([&] { run(cm16o, d_in, d_out, NF, W, H); });
        const unsigned long long b16o = differ(d_out, d_ref, kPixels, d_cnt);
        printf("%-6s | %7.1f %% %7.1f %% %7.1f %% %7.1f %% | %llu, %llu, %llu\n", names[cls], pct(t0), pct(t8), pct(t16), pct(t16o), b8, b16, b16o);
    }
    return 0;
}
