// pipes.cu — issue/pipe throughput microbenchmark for the instruction mix of the HSV /
// LUT kernels on sm_100a (development aid; informs DESIGN.md's per-pixel budgets).
// Reports warp-instructions per clock per SM for each op, measured with full occupancy.
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

#define ITERS 4096
#define CHAINS 8

template <int OP>
__global__ void __launch_bounds__(256) k(float *out, float a0, float b0, uint32_t sel) {
    float x[CHAINS];
    float2 y[CHAINS];
    uint32_t u[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) {
        x[i] = a0 + threadIdx.x * 1e-3f + i;
        y[i] = make_float2(x[i], x[i] + 1.f);
        u[i] = __float_as_uint(x[i]);
    }
    float b = b0;
    float2 b2 = make_float2(b0, b0 * 1.0001f);
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (OP == 0) x[i] = __fmaf_rn(x[i], b, a0);
            if (OP == 1) y[i] = __ffma2_rn(y[i], b2, b2);
            if (OP == 2) x[i] = __fadd_rn(x[i], b);
            if (OP == 3) y[i] = __fadd2_rn(y[i], b2);
            if (OP == 4) x[i] = __fmul_rn(x[i], b);
            if (OP == 5) y[i] = __fmul2_rn(y[i], b2);
            if (OP == 6) x[i] = (x[i] > b) ? a0 : x[i] + 0.f;  // FSETP+FSEL-ish
            if (OP == 7) u[i] = __byte_perm(u[i], sel, 0x2103u);
            if (OP == 8) x[i] = fmaxf(x[i], b);
            if (OP == 9) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
            if (OP == 10) x[i] = __fadd_rd(x[i], b);
            if (OP == 11) x[i] = __saturatef(__fadd_rn(x[i], b));
            if (OP == 12) {  // mixed: one FFMA2 + one PRMT per chain (co-issue test)
                y[i] = __ffma2_rn(y[i], b2, b2);
                u[i] = __byte_perm(u[i], sel, 0x2103u);
            }
            if (OP == 13) {  // mixed: FFMA + PRMT
                x[i] = __fmaf_rn(x[i], b, a0);
                u[i] = __byte_perm(u[i], sel, 0x2103u);
            }
            if (OP == 14) u[i] = (u[i] & sel) ^ 0x55u;          // LOP3
            if (OP == 15) x[i] = (float)(int)x[i];              // F2I+I2F
            if (OP == 16) {  // FFMA2 + FSEL-like ALU pair
                y[i] = __ffma2_rn(y[i], b2, b2);
                x[i] = fmaxf(x[i], b);
            }
            if (OP == 17) y[i] = __fadd2_rd(y[i], b2);
            if (OP == 18) x[i] = fmaxf(x[i], fminf(b, a0 + x[i]));  // FMNMX3-ish
            if (OP == 19) {  // 2 FFMA2 + 1 PRMT + 1 FMNMX
                y[i] = __ffma2_rn(y[i], b2, b2);
                y[i] = __fmul2_rn(y[i], b2);
                u[i] = __byte_perm(u[i], sel, 0x2103u);
                x[i] = fmaxf(x[i], b);
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) s += x[i] + y[i].x + y[i].y + __uint_as_float(u[i]);
    if (s == 123.456f) out[0] = s;
}

template <int OP>
void run(const char *name, int instr_per_chain_iter, int flops_lanes) {
    float *d;
    cudaMalloc(&d, 4);
    int blocks = 148 * 8;
    k<OP><<<blocks, 256>>>(d, 1.0f, 0.999f, 0x3210u);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, 256>>>(d, 1.0f, 0.999f, 0x3210u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warp_instr = (double)blocks * 8 * ITERS * CHAINS * instr_per_chain_iter;
    int clk_khz;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double clocks = ms * 1e-3 * clk_khz * 1e3;
    printf("%-28s %8.3f ms  %6.2f warp-instr/clk/SM (at %d MHz nominal)  %s\n", name, ms,
           warp_instr / clocks / 148.0, clk_khz / 1000, flops_lanes ? "" : "");
    cudaFree(d);
}

int main() {
    run<0>("FFMA", 1, 1);
    run<1>("FFMA2", 1, 2);
    run<2>("FADD", 1, 1);
    run<3>("FADD2", 1, 2);
    run<4>("FMUL", 1, 1);
    run<5>("FMUL2", 1, 2);
    run<6>("FSETP+FSEL(+FADD)", 3, 0);
    run<7>("PRMT", 1, 0);
    run<8>("FMNMX", 1, 0);
    run<9>("MUFU.RCP", 1, 0);
    run<10>("FADD.RM", 1, 0);
    run<11>("FADD.SAT", 1, 0);
    run<12>("FFMA2+PRMT", 2, 0);
    run<13>("FFMA+PRMT", 2, 0);
    run<14>("LOP3", 1, 0);
    run<15>("F2I+I2F", 2, 0);
    run<16>("FFMA2+FMNMX", 2, 0);
    run<17>("FADD2.RM", 1, 0);
    run<18>("FMNMX3(+FADD)", 2, 0);
    run<19>("2xF2x2+PRMT+FMNMX", 4, 0);
    return 0;
}
