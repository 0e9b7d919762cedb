// coissue.cu — how packed f32x2 ops share issue/dispatch with ALU-pipe ops on sm_100a.
// Each test runs NF packed/scalar FMA-pipe ops and NA ALU-pipe ops per chain-iteration over
// 8 independent chains and reports cycles per chain-iteration per SMSP (8 warps/SMSP).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 2048
#define CH 8

// KIND: 0 = FFMA2, 1 = scalar FFMA, 2 = FADD2, 3 = scalar FADD
// AK: 0 = PRMT, 1 = LOP3 (cross-chain), 2 = FMNMX (cross-chain), 3 = FSETP+FSEL, 4 = MUFU.RCP, 5 = LDS
template <int KIND, int NF, int AK, int NA>
__global__ void __launch_bounds__(256) k(float *out, float a0, float b0, uint32_t sel) {
    __shared__ float sm[256];
    sm[threadIdx.x] = a0;
    __syncthreads();
    float x[CH], z[CH];
    float2 y[CH];
    uint32_t u[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) {
        x[i] = a0 + threadIdx.x * 1e-3f + i;
        z[i] = x[i] * 0.5f;
        y[i] = make_float2(x[i], x[i] + 1.f);
        u[i] = __float_as_uint(x[i]) | 1u;
    }
    float b = b0;
    float2 b2 = make_float2(b0, b0 * 1.0001f);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
#pragma unroll
            for (int f = 0; f < NF; f++) {
                if (KIND == 0) y[i] = __ffma2_rn(y[i], b2, b2);
                if (KIND == 1) x[i] = __fmaf_rn(x[i], b, a0);
                if (KIND == 2) y[i] = __fadd2_rn(y[i], b2);
                if (KIND == 3) x[i] = __fadd_rn(x[i], b);
            }
#pragma unroll
            for (int a = 0; a < NA; a++) {
                if (AK == 0) u[i] = __byte_perm(u[i], sel, 0x2103u);
                if (AK == 1) u[i] = u[i] ^ (u[(i + 1) % CH] & sel);
                if (AK == 2) z[i] = fmaxf(z[i], -z[(i + 1) % CH]);
                if (AK == 3) z[i] = (z[(i + 1) % CH] > b) ? z[i] : -z[i];
                if (AK == 4) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(z[i]));
                if (AK == 5) z[i] = sm[(__float_as_uint(z[i]) + it) & 255];
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += x[i] + z[i] + y[i].x + y[i].y + __uint_as_float(u[i]);
    if (s == 123.456f) out[0] = s;
}

template <int KIND, int NF, int AK, int NA>
void run(const char *name) {
    float *d;
    cudaMalloc(&d, 4);
    int blocks = 148 * 8;
    k<KIND, NF, AK, NA><<<blocks, 256>>>(d, 1.0f, 0.999f, 0x3210u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<KIND, NF, AK, NA><<<blocks, 256>>>(d, 1.0f, 0.999f, 0x3210u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    // warps per SMSP: 8 blocks/SM * 8 warps / 4 = 16; chain-iterations per SMSP:
    double chain_iters = 16.0 * ITERS * CH;
    double clocks = ms * 1e-3 * 1.965e9;  // assumes max clock (short kernels)
    printf("%-34s %8.3f ms  %6.2f cycles per (%d F + %d A) per SMSP\n", name, ms,
           clocks / chain_iters, NF, NA);
    cudaFree(d);
}

int main() {
    run<0, 1, 0, 0>("FFMA2 x1");
    run<1, 1, 0, 0>("FFMA x1");
    run<1, 0, 0, 1>("PRMT x1");
    run<1, 0, 1, 1>("LOP3 x1");
    run<1, 0, 2, 1>("FMNMX x1");
    run<1, 0, 3, 1>("FSETP+FSEL x1");
    run<1, 0, 4, 1>("MUFU.RCP x1");
    run<1, 0, 5, 1>("LDS(+LOP/IADD) x1");
    run<0, 1, 0, 1>("FFMA2 + PRMT");
    run<0, 2, 0, 1>("2 FFMA2 + PRMT");
    run<0, 1, 0, 2>("FFMA2 + 2 PRMT");
    run<1, 2, 0, 1>("2 FFMA + PRMT");
    run<1, 4, 0, 1>("4 FFMA + PRMT");
    run<1, 2, 0, 2>("2 FFMA + 2 PRMT");
    run<0, 1, 1, 1>("FFMA2 + LOP3");
    run<0, 1, 2, 1>("FFMA2 + FMNMX");
    run<1, 2, 2, 1>("2 FFMA + FMNMX");
    run<0, 1, 3, 1>("FFMA2 + FSETP+FSEL");
    run<0, 1, 4, 1>("FFMA2 + MUFU");
    run<1, 2, 4, 1>("2 FFMA + MUFU");
    run<2, 1, 0, 1>("FADD2 + PRMT");
    run<3, 2, 0, 1>("2 FADD + PRMT");
    run<0, 2, 2, 2>("2 FFMA2 + 2 FMNMX");
    run<1, 4, 2, 2>("4 FFMA + 2 FMNMX");
    return 0;
}
