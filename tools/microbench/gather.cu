// gather.cu — L1 cost of 16-byte gathers (LDG.128) as a function of the address pattern across a
// warp, for DESIGN.md's analysis of the colorlut kernel.  Each thread issues ITERS dependent-free
// loads from a small L1-resident table; reports SM cycles per warp-level LDG.128.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 2048

// pattern: entry index for lane l at iteration it
//  0: all lanes the same entry            1: consecutive entries (coalesced 512 B)
//  2: groups of 4 lanes share an entry    3: floor(l * 0.27) (grad-like, ~9 entries)
//  4: like 3 but the run starts at a random offset within the line (straddles)
//  5: 32 entries, each in a different 128 B line
//  6: 2 lines: lanes alternate between two entries 4 KB apart
template <int P, int VEC>
__global__ void __launch_bounds__(256) k(const float4 *tab, float *out, uint32_t mask) {
    const uint32_t lane = threadIdx.x & 31;
    float acc = 0.f;
    uint32_t base = (blockIdx.x * 37u + (threadIdx.x >> 5) * 11u) & mask;
    for (int it = 0; it < ITERS; it++) {
        uint32_t e;
        if (P == 0) e = 0;
        if (P == 1) e = lane;
        if (P == 2) e = lane >> 2;
        if (P == 3) e = (lane * 69u) >> 8;
        if (P == 4) e = ((lane * 69u) >> 8) + (it & 7);
        if (P == 5) e = lane * 8;
        if (P == 6) e = (lane & 1) * 256;
        const float4 *p = tab + ((base + e + (uint32_t)it * 8u) & mask);
        if (VEC == 16) {
            float4 v = __ldg(p);
            acc += v.x + v.y + v.z;
        } else if (VEC == 8) {
            float2 v = __ldg(reinterpret_cast<const float2 *>(p));
            acc += v.x + v.y;
        } else {
            acc += __ldg(reinterpret_cast<const float *>(p));
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

template <int P, int VEC>
void run(const char *name, const float4 *tab, float *out) {
    const int blocks = 148 * 8;
    const uint32_t mask = 4095;  // 4096 entries = 64 KB: L1-resident
    k<P, VEC><<<blocks, 256>>>(tab, out, mask);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<P, VEC><<<blocks, 256>>>(tab, out, mask);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warp_loads_per_sm = 8.0 * 8 * ITERS;  // 8 CTAs/SM * 8 warps
    double cycles = ms * 1e-3 * 1.965e9;
    printf("%-46s LDG.%-3d %7.3f ms  %6.2f SM-cycles per warp-level load\n", name, VEC * 8, ms,
           cycles / warp_loads_per_sm);
}

int main() {
    float4 *tab;
    float *out;
    cudaMalloc(&tab, 4096 * sizeof(float4));
    cudaMemset(tab, 0, 4096 * sizeof(float4));
    cudaMalloc(&out, 4);
    run<0, 16>("all lanes one entry", tab, out);
    run<1, 16>("32 consecutive entries (512 B)", tab, out);
    run<2, 16>("8 consecutive entries, 4 lanes each (128 B)", tab, out);
    run<3, 16>("grad-like: ~9 consecutive entries", tab, out);
    run<4, 16>("grad-like, sliding start (line straddles)", tab, out);
    run<5, 16>("32 entries in 32 different lines", tab, out);
    run<6, 16>("2 entries 4 KB apart, alternating lanes", tab, out);
    run<0, 8>("all lanes one entry", tab, out);
    run<1, 8>("32 consecutive 16 B slots, 8 B each", tab, out);
    run<3, 8>("grad-like", tab, out);
    run<5, 8>("32 different lines", tab, out);
    run<0, 4>("all lanes one entry", tab, out);
    run<1, 4>("32 consecutive 16 B slots, 4 B each", tab, out);
    run<3, 4>("grad-like", tab, out);
    run<5, 4>("32 different lines", tab, out);
    return 0;
}
