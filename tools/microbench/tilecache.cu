// tilecache.cu — experiment behind the "tile-staged table" kernel: how should a
// 2^24-entry function table (baked colorlut / hsvfilter / hsvdetector / chain) be applied to frames
// whose colours are locally coherent but noisy?
//
//   flat   : today's kernel — frame flattened to one row, one LDG.32 gather per pixel
//   tile   : 2-D tile per CTA, the same global gather (L1 sees a small colour footprint)
//   tileblk: tile + table stored in 4x4x2 colour blocks per 128-byte line
//   cache  : 2-D tile, colour bounding box of the tile by min/max reduction, that slice of the table
//            staged into shared memory, pixels looked up with IADD + IDP4A + LDS; tiles whose box
//            does not fit fall back to the global gather
//
// Content classes as in gst-plugins-rs_b200/frames.py (bars, grad, noise = grad +-2, rand), 16 frames
// of 3840x2160 RGBA (1.06 GB in + out > L2).  Prints % of the measured HBM copy peak and verifies
// every variant against the flat kernel's output.
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
            exit(1);                                                               \
        }                                                                          \
    } while (0)

static const double kPeak = 6548.5;
constexpr int W = 3840, H = 2160, NF = 16;
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

// Table laid out in colour blocks of 2^A x 2^B x 2^C (r x g x b) entries per 128-byte line.
template <int A, int B, int C>
__host__ __device__ __forceinline__ uint32_t swz(uint32_t x) {
    static_assert(A + B + C == 5, "one line");
    if (A == 5) return x & 0xFFFFFFu;
    const uint32_t r = x & 255u, g = (x >> 8) & 255u, b = (x >> 16) & 255u;
    const uint32_t lo = (r & ((1u << A) - 1)) | (g & ((1u << B) - 1)) << A | (b & ((1u << C) - 1)) << (A + B);
    const uint32_t hi = (r >> A) | (g >> B) << (8 - A) | (b >> C) << (16 - A - B);
    return lo | hi << 5;
}
// (8,4,1) blocks by rotating the 7-bit field [g1 g0 r7..r3] two places: 4 ALU ops
__device__ __forceinline__ uint32_t swz_rot320(uint32_t x) {
    const uint32_t t1 = x << 2, t2 = x >> 5;
    uint32_t rot;
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(rot) : "r"(0x3E0u), "r"(t1), "r"(t2));  // M ? t1 : t2
    uint32_t o;
    asm("lop3.b32 %0, %1, %2, %3, 0xCA;" : "=r"(o) : "r"(0x3F8u), "r"(rot), "r"(x));
    return o & 0xFFFFFFu;
}

template <int A, int B, int C>
__global__ void table_kernel(uint32_t *t) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    t[swz<A, B, C>(i)] = hash32(i) & 0xFFFFFFu;
}

// ---- flat (today's vf_map_vec_kernel<ColorLutBakedOp>) ------------------------------------------
__device__ __forceinline__ uint32_t px_flat(uint32_t in, const uint32_t *table) {
    return __byte_perm(__ldg(table + (in & 0xFFFFFFu)), in, 0x7210u);
}
template <int A, int B, int C, bool ROT>
__device__ __forceinline__ uint32_t px_blk(uint32_t in, const uint32_t *table) {
    const uint32_t j = ROT ? swz_rot320(in) : swz<A, B, C>(in);
    return __byte_perm(__ldg(table + j), in, 0x7210u);
}

__global__ void __launch_bounds__(kThreads) flat_kernel(const uint4 *in, uint4 *out, uint32_t units_per_frame,
                                                        const uint32_t *table) {
    const uint4 *src = in + (size_t)blockIdx.z * units_per_frame;
    uint4 *dst = out + (size_t)blockIdx.z * units_per_frame;
    constexpr uint32_t kTile = kThreads * 4;
    const uint32_t tiles = (units_per_frame + kTile - 1) / kTile;
    for (uint32_t seg = blockIdx.x; seg < tiles; seg += gridDim.x) {
        const uint32_t u0 = seg * kTile + threadIdx.x;
        uint4 v[4];
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (u0 + j * kThreads < units_per_frame) v[j] = __ldcs(src + u0 + j * kThreads);
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (u0 + j * kThreads < units_per_frame) {
                uint4 o;
                o.x = px_flat(v[j].x, table), o.y = px_flat(v[j].y, table);
                o.z = px_flat(v[j].z, table), o.w = px_flat(v[j].w, table);
                __stcs(dst + u0 + j * kThreads, o);
            }
    }
}

// ---- tile kernels --------------------------------------------------------------------------------
// TWU = 16-byte units per tile row (16 -> 64 pixels); tile height = 4 * 256 / TWU rows.
// MODE 0: global gather, 1: global gather from the blocked table, 2: shared-memory staged slice
struct Box {
    uint32_t mn[3], mx[3];
    uint32_t pad[2];
};

template <int TWU, int MODE, int CACHE, bool PERSIST, int A = 5, int B = 0, int C = 0, bool ROT = false>
__global__ void __launch_bounds__(kThreads, MODE == 2 ? 6 : 8)
    tile_kernel(const uint8_t *in, uint8_t *out, size_t frame_bytes, uint32_t units_per_row, uint32_t rows,
                long long stride, uint32_t tiles_x, uint32_t tiles_y, const uint32_t *table) {
    constexpr int RPP = kThreads / TWU;  // rows per pass
    constexpr int TH = RPP * 4;
    __shared__ uint32_t cache[MODE == 2 ? CACHE : 1];
    __shared__ Box box[2];
    const uint32_t ux = threadIdx.x % TWU, uy = threadIdx.x / TWU;
    const uint32_t n_tiles = tiles_x * tiles_y;
    if (MODE == 2 && threadIdx.x < 16) {
        uint32_t *b = reinterpret_cast<uint32_t *>(box);
        b[threadIdx.x] = (threadIdx.x & 7) < 3 ? 0xFFFFFFFFu : 0u;
    }
    if (MODE == 2) __syncthreads();
    uint32_t par = 0;
    for (uint32_t t = blockIdx.x; t < n_tiles; t += PERSIST ? gridDim.x : n_tiles) {
        const uint32_t tyi = t / tiles_x, txi = t - tyi * tiles_x;
        const uint32_t x = txi * TWU + ux;
        const uint32_t y0 = tyi * TH + uy;
        const uint8_t *src = in + (size_t)blockIdx.z * frame_bytes + (size_t)x * 16;
        uint8_t *dst = out + (size_t)blockIdx.z * frame_bytes + (size_t)x * 16;
        uint4 v[4];
        bool ok[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t y = y0 + j * RPP;
            ok[j] = x < units_per_row && y < rows;
            if (ok[j]) v[j] = __ldcs(reinterpret_cast<const uint4 *>(src + (size_t)y * stride));
        }
        bool cached = false;
        uint32_t org = 0, wts = 0;
        if (MODE == 2) {
            // colour bounding box: (r,b) and (g,a) as u16x2 lanes
            uint32_t mn_rb = 0x00FF00FFu, mx_rb = 0, mn_g = 0x00FF00FFu, mx_g = 0;
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (ok[j]) {
                    const uint32_t p[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const uint32_t rb = __byte_perm(p[k], 0, 0x4240u), ga = __byte_perm(p[k], 0, 0x4341u);
                        mn_rb = __vminu2(mn_rb, rb), mx_rb = __vmaxu2(mx_rb, rb);
                        mn_g = __vminu2(mn_g, ga), mx_g = __vmaxu2(mx_g, ga);
                    }
                }
            const uint32_t rmin = __reduce_min_sync(0xFFFFFFFFu, mn_rb & 0xFFFFu);
            const uint32_t bmin = __reduce_min_sync(0xFFFFFFFFu, mn_rb >> 16);
            const uint32_t gmin = __reduce_min_sync(0xFFFFFFFFu, mn_g & 0xFFFFu);
            const uint32_t rmax = __reduce_max_sync(0xFFFFFFFFu, mx_rb & 0xFFFFu);
            const uint32_t bmax = __reduce_max_sync(0xFFFFFFFFu, mx_rb >> 16);
            const uint32_t gmax = __reduce_max_sync(0xFFFFFFFFu, mx_g & 0xFFFFu);
            Box &bx = box[par];
            if ((threadIdx.x & 31) == 0) {
                atomicMin(&bx.mn[0], rmin), atomicMin(&bx.mn[1], gmin), atomicMin(&bx.mn[2], bmin);
                atomicMax(&bx.mx[0], rmax), atomicMax(&bx.mx[1], gmax), atomicMax(&bx.mx[2], bmax);
            }
            if (threadIdx.x < 8) {  // reset the other box for the next tile
                uint32_t *b = reinterpret_cast<uint32_t *>(&box[par ^ 1]);
                b[threadIdx.x] = threadIdx.x < 3 ? 0xFFFFFFFFu : 0u;
            }
            __syncthreads();
            const uint32_t r0 = bx.mn[0], g0 = bx.mn[1], b0 = bx.mn[2];
            const uint32_t nr = bx.mx[0] - r0 + 1, ng = bx.mx[1] - g0 + 1, nb = bx.mx[2] - b0 + 1;
            // axis order: the widest channel is the slowest axis; weights of the other two <= 255
            uint32_t n1, n2, s1, s2, s3, w_r, w_g, w_b;
            if (nb >= nr && nb >= ng)
                n1 = nr, n2 = ng, s1 = 0, s2 = 8, s3 = 16, w_r = 1, w_g = nr, w_b = nr * ng;
            else if (ng >= nr)
                n1 = nr, n2 = nb, s1 = 0, s2 = 16, s3 = 8, w_r = 1, w_b = nr, w_g = nr * nb;
            else
                n1 = ng, n2 = nb, s1 = 8, s2 = 16, s3 = 0, w_g = 1, w_b = ng, w_r = ng * nb;
            const uint32_t vol = nr * ng * nb;
            cached = vol <= (uint32_t)CACHE && n1 * n2 <= 255u;
            if (cached) {
                org = r0 | g0 << 8 | b0 << 16;
                wts = w_r | w_g << 8 | w_b << 16;
                const uint32_t inv1 = 0xFFFFFFFFu / n1 + 1, inv2 = 0xFFFFFFFFu / n2 + 1;  // exact: e * n < 2^32
                uint32_t *c = cache;
                for (uint32_t e = threadIdx.x; e < vol; e += kThreads) {
                    const uint32_t q1 = n1 == 1 ? e : __umulhi(e, inv1), d1 = e - q1 * n1;
                    const uint32_t q2 = n2 == 1 ? q1 : __umulhi(q1, inv2), d2 = q1 - q2 * n2;
                    const uint32_t col = org + (d1 << s1) + (d2 << s2) + (q2 << s3);
                    c[e] = __ldg(table + col);
                }
            }
            __syncthreads();
            par ^= 1;
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (ok[j]) {
                const uint32_t p[4] = {v[j].x, v[j].y, v[j].z, v[j].w};
                uint32_t o[4];
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (MODE == 2 && cached)
                        o[k] = __byte_perm(cache[__dp4a(p[k] - org, wts, 0u)], p[k], 0x7210u);
                    else if (MODE == 1)
                        o[k] = px_blk<A, B, C, ROT>(p[k], table);
                    else
                        o[k] = px_flat(p[k], table);
                }
                __stcs(reinterpret_cast<uint4 *>(dst + (size_t)(y0 + j * RPP) * stride),
                       make_uint4(o[0], o[1], o[2], o[3]));
            }
        if (MODE == 2 && PERSIST) __syncthreads();  // lookups done before the next tile's fill
    }
}

// ---- production candidate: dynamic tile scheduler + 7-op swizzle ------------------------------------
// Table index [c2 7..1][c0 7..2][c1 7..2][c1 1..0][c2 0][c0 1..0]: a 4 x 4 x 2 colour block per 128-byte line with
// three field moves (c1 as a whole, c0's high bits, c2's low bit); the alpha byte is masked by the way.
__host__ __device__ __forceinline__ uint32_t swz7(uint32_t x) {
    uint32_t j = x & 0x00FE0003u;               // c2 high, c0 low stay
    j |= (x >> 5) & 0x000007F8u;                // c1: 10-15 -> 5-10, 8-9 -> 3-4
    j |= (x & 0x000000FCu) << 9;                // c0 high: 2-7 -> 11-16
    j |= (x >> 14) & 0x00000004u;               // c2 low: 16 -> 2
    return j;
}
__global__ void table_swz7_kernel(uint32_t *t) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    t[swz7(i)] = hash32(i) & 0xFFFFFFu;
}

// SCHED 0: one tile per CTA, 1: static grid-stride, 2: dynamic (atomic counter, next tile prefetched)
template <int TWU, int SCHED>
__global__ void __launch_bounds__(kThreads, 8)
    tile_prod_kernel(const uint8_t *in, uint8_t *out, size_t frame_bytes, uint32_t units_per_row, uint32_t rows,
                     long long stride, uint32_t tiles_x, uint32_t tiles_per_frame, uint32_t n_tiles,
                     const uint32_t *table, uint32_t *counter) {
    constexpr int RPP = kThreads / TWU;
    constexpr int TH = RPP * 4;
    __shared__ uint32_t s_next[2];
    const uint32_t ux = threadIdx.x % TWU, uy = threadIdx.x / TWU;
    uint32_t par = 0;
    for (uint32_t t = blockIdx.x; t < n_tiles;) {
        if (SCHED == 2 && threadIdx.x == 0) s_next[par] = atomicAdd(counter, 1u) + gridDim.x;
        const uint32_t f = t / tiles_per_frame, tf = t - f * tiles_per_frame;
        const uint32_t tyi = tf / tiles_x, txi = tf - tyi * tiles_x;
        const uint32_t x = txi * TWU + ux;
        const uint32_t y0 = tyi * TH + uy;
        const uint8_t *src = in + (size_t)f * frame_bytes + (size_t)x * 16;
        uint8_t *dst = out + (size_t)f * frame_bytes + (size_t)x * 16;
        uint4 v[4];
        bool ok[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t y = y0 + j * RPP;
            ok[j] = x < units_per_row && y < rows;
            if (ok[j]) v[j] = __ldcs(reinterpret_cast<const uint4 *>(src + (size_t)y * stride));
        }
#pragma unroll
        for (int j = 0; j < 4; j++)
            if (ok[j]) {
                uint4 o;
                o.x = __byte_perm(__ldg(table + swz7(v[j].x)), v[j].x, 0x7210u);
                o.y = __byte_perm(__ldg(table + swz7(v[j].y)), v[j].y, 0x7210u);
                o.z = __byte_perm(__ldg(table + swz7(v[j].z)), v[j].z, 0x7210u);
                o.w = __byte_perm(__ldg(table + swz7(v[j].w)), v[j].w, 0x7210u);
                __stcs(reinterpret_cast<uint4 *>(dst + (size_t)(y0 + j * RPP) * stride), o);
            }
        if (SCHED == 0) break;
        if (SCHED == 1) t += gridDim.x;
        if (SCHED == 2) {
            __syncthreads();
            t = s_next[par];
            par ^= 1;
        }
    }
}

// ---- harness ---------------------------------------------------------------------------------
struct Bufs {
    uint8_t *in, *out, *ref;
    uint32_t *table, *tblk;
};

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
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && a[i] != b[i]) atomicAdd(cnt, 1ull);
}

static unsigned long long *g_cnt;
static void report(const char *name, const char *cls, float ms, const Bufs &b, bool blocked_ref = false) {
    const size_t n = (size_t)W * H * NF;
    CK(cudaMemset(g_cnt, 0, 8));
    diff_kernel<<<(unsigned)((n + 255) / 256), 256>>>((const uint32_t *)b.out, (const uint32_t *)b.ref, n, g_cnt);
    unsigned long long bad = 0;
    CK(cudaMemcpy(&bad, g_cnt, 8, cudaMemcpyDeviceToHost));
    const double gbs = 8.0 * n / (ms * 1e-3) / 1e9;
    printf("%-6s %-34s %8.3f ms  %7.1f GB/s  %5.1f %% of HBM peak  %s\n", cls, name, ms, gbs, gbs / kPeak * 100,
           bad ? "MISMATCH" : "ok");
    (void)blocked_ref;
    CK(cudaMemset(b.out, 0, n * 4));
}

template <int TWU, int MODE, int CACHE, bool PERSIST, int A = 5, int B = 0, int C = 0, bool ROT = false>
static void run_tile(const char *name, const char *cls, const Bufs &b, int ctas_per_sm = 64) {
    constexpr int TH = kThreads / TWU * 4;
    const uint32_t upr = W / 4, tiles_x = (upr + TWU - 1) / TWU, tiles_y = (H + TH - 1) / TH;
    const uint32_t nt = tiles_x * tiles_y;
    const uint32_t gx = PERSIST ? std::min<uint32_t>(nt, std::max(1, 148 * ctas_per_sm / NF)) : nt;
    dim3 grid(gx, 1, NF);
    const uint32_t *tab = b.table;
    if (MODE == 1) {
        table_kernel<A, B, C><<<(1u << 24) / 256, 256>>>(b.tblk);
        tab = b.tblk;
    }
    float ms = time_ms([&] {
        tile_kernel<TWU, MODE, CACHE, PERSIST, A, B, C, ROT><<<grid, kThreads>>>(
            b.in, b.out, (size_t)W * H * 4, upr, H, (long long)W * 4, tiles_x, tiles_y, tab);
    });
    CK(cudaGetLastError());
    report(name, cls, ms, b);
}

template <int TWU, int SCHED>
static void run_prod(const char *name, const char *cls, const Bufs &b, int ctas_per_sm, uint32_t *counter) {
    constexpr int TH = kThreads / TWU * 4;
    const uint32_t upr = W / 4, tiles_x = (upr + TWU - 1) / TWU, tiles_y = (H + TH - 1) / TH;
    const uint32_t tpf = tiles_x * tiles_y, nt = tpf * NF;
    const uint32_t gx = SCHED == 0 ? nt : std::min<uint32_t>(nt, 148 * ctas_per_sm);
    table_swz7_kernel<<<(1u << 24) / 256, 256>>>(b.tblk);
    float ms = time_ms([&] {
        if (SCHED == 2) cudaMemsetAsync(counter, 0, 4);
        tile_prod_kernel<TWU, SCHED><<<gx, kThreads>>>(b.in, b.out, (size_t)W * H * 4, upr, H, (long long)W * 4,
                                                        tiles_x, tpf, nt, b.tblk, counter);
    });
    CK(cudaGetLastError());
    report(name, cls, ms, b);
}

int main(int argc, char **argv) {
    Bufs b;
    const size_t fb = (size_t)W * H * 4;
    CK(cudaMalloc(&b.in, fb * NF));
    CK(cudaMalloc(&b.out, fb * NF));
    CK(cudaMalloc(&b.ref, fb * NF));
    CK(cudaMalloc(&b.table, 4u << 24));
    CK(cudaMalloc(&b.tblk, 4u << 24));
    CK(cudaMalloc(&g_cnt, 8));
    uint32_t *counter;
    CK(cudaMalloc(&counter, 4));
    table_kernel<5, 0, 0><<<(1u << 24) / 256, 256>>>(b.table);
    const char *names[4] = {"bars", "grad", "noise", "rand"};
    const bool full = argc > 1;
    for (int cls = 0; cls < 4; cls++) {
        for (int f = 0; f < NF; f++)
            gen_kernel<<<dim3((W + 255) / 256, H), 256>>>((uint32_t *)(b.in + fb * f), cls, f);
        CK(cudaDeviceSynchronize());
        const uint32_t upf = W * H / 4;
        dim3 grid(148 * 64 / NF, 1, NF);
        flat_kernel<<<grid, kThreads>>>((const uint4 *)b.in, (uint4 *)b.ref, upf, b.table);
        CK(cudaDeviceSynchronize());
        float ms = time_ms([&] { flat_kernel<<<grid, kThreads>>>((const uint4 *)b.in, (uint4 *)b.out, upf, b.table); });
        report("flat (round-1 kernel)", names[cls], ms, b);
        const char *c = names[cls];
        run_tile<16, 0, 4096, false>("tile 64x64, natural table", c, b);
        run_tile<32, 0, 4096, false>("tile 128x32, natural table", c, b);
        run_tile<64, 0, 4096, false>("tile 256x16, natural table", c, b);
        run_tile<16, 0, 4096, true>("tile 64x64, natural, persistent x64", c, b);
#define SHAPE(A, B, C)                                                                            \
    {                                                                                             \
        run_tile<16, 1, 4096, false, A, B, C>("tile 64x64, blocks 2^(" #A "," #B "," #C ")", c, b);  \
        run_tile<32, 1, 4096, false, A, B, C>("tile 128x32, blocks 2^(" #A "," #B "," #C ")", c, b); \
    }
        SHAPE(2, 2, 1)
        if (full) SHAPE(3, 2, 0)
        if (full) SHAPE(2, 3, 0)
        if (full) SHAPE(3, 1, 1)
        SHAPE(2, 1, 2)
        SHAPE(1, 2, 2)
        if (full) SHAPE(3, 0, 2)
        if (full) SHAPE(4, 1, 0)
        if (full) SHAPE(1, 1, 3)
        if (full) SHAPE(0, 2, 3)
        run_tile<16, 1, 4096, false, 3, 2, 0, true>("tile 64x64, blocks 2^(3,2,0) rot", c, b);
        run_tile<32, 1, 4096, false, 3, 2, 0, true>("tile 128x32, blocks 2^(3,2,0) rot", c, b);
        run_tile<64, 1, 4096, false, 3, 2, 0, true>("tile 256x16, blocks 2^(3,2,0) rot", c, b);
        run_tile<16, 1, 4096, true, 2, 2, 1>("tile 64x64, blocks 2^(2,2,1) persistent x64", c, b);
        run_prod<16, 0>("prod 64x64 swz7, 1 tile/CTA", c, b, 0, counter);
        run_prod<16, 1>("prod 64x64 swz7, static x16", c, b, 16, counter);
        run_prod<16, 1>("prod 64x64 swz7, static x32", c, b, 32, counter);
        run_prod<16, 1>("prod 64x64 swz7, static x64", c, b, 64, counter);
        run_prod<16, 1>("prod 64x64 swz7, static x128", c, b, 128, counter);
        run_prod<16, 2>("prod 64x64 swz7, dynamic x8", c, b, 8, counter);
        run_prod<16, 2>("prod 64x64 swz7, dynamic x16", c, b, 16, counter);
        run_prod<32, 2>("prod 128x32 swz7, dynamic x8", c, b, 8, counter);
        run_prod<32, 0>("prod 128x32 swz7, 1 tile/CTA", c, b, 0, counter);
        run_prod<8, 2>("prod 32x128 swz7, dynamic x8", c, b, 8, counter);
        if (full) {  // the shared-memory staged slice (kept as the measured negative result)
            run_tile<16, 2, 4096, false>("smem slice 64x64 4096e", c, b);
            run_tile<32, 2, 4096, false>("smem slice 128x32 4096e", c, b);
            run_tile<16, 2, 8192, false>("smem slice 64x64 8192e", c, b);
            run_tile<16, 2, 4096, true>("smem slice 64x64 4096e persistent x64", c, b);
        }
    }
    return 0;
}
