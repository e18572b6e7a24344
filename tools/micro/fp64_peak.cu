// Microbenchmark: sustained DFMA / DMUL+DFMA throughput per SM on this GPU (experiment, not product code).
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int MODE>
__global__ void __launch_bounds__(128, 4) k(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (MODE == 0) x[i] = fma(x[i], a, b);                 // 1 DFMA
            else { double t = x[i] * b; x[i] = fma(x[i], a, t); }  // DMUL + DFMA (the REAL gate pattern)
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    if (s == 12345.678) out[0] = s;
}

template <int ILP, int MODE>
void run(const char* name, int blocks_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* d; cudaMalloc(&d, 8);
    int iters = 20000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<ILP, MODE><<<sms * blocks_per_sm, 128>>>(d, 100, 1.0000001, 1e-9);
    cudaEventRecord(e0);
    k<ILP, MODE><<<sms * blocks_per_sm, 128>>>(d, iters, 1.0000001, 1e-9);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double inst = (double)sms * blocks_per_sm * 4 /*warps*/ * iters * ILP * (MODE ? 2 : 1);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("%s ILP=%d blocks/SM=%d: %.3f ms, %.3f fp64 warp-instr/clk/SM (at %d MHz nominal), %.2f TFLOP/s (FMA=2)\n", name, ILP, blocks_per_sm, ms,
           inst / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000, inst * 32 * (MODE ? 1.5 : 2.0) / (ms * 1e-3) / 1e12);
}


// ---- FP64 tensor-core (DMMA) peak: mma.sync m8n8k4 and m16n8k16 (tcgen05 has no FP64 kind) ----
template <int ILP, int SHAPE>
__global__ void __launch_bounds__(128, 4) k_dmma(double* out, int iters) {
    double c[ILP][4];
#pragma unroll
    for (int i = 0; i < ILP; i++) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = 0.5; c[i][2] = 0.25; c[i][3] = 0.125; }
    const double a0 = 1.0000001, b0 = 1e-9;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) {
            if (SHAPE == 0) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a0), "d"(b0));
            } else {
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(c[i][0]), "+d"(c[i][1]), "+d"(c[i][2]), "+d"(c[i][3])
                             : "d"(a0), "d"(a0), "d"(a0), "d"(a0), "d"(a0), "d"(a0), "d"(a0), "d"(a0), "d"(b0), "d"(b0), "d"(b0), "d"(b0));
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    if (s == 12345.678) out[0] = s;
}

template <int ILP, int SHAPE>
void run_dmma(const char* name, int blocks_per_sm) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double* d; cudaMalloc(&d, 8);
    int iters = 4000;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_dmma<ILP, SHAPE><<<sms * blocks_per_sm, 128>>>(d, 100);
    cudaEventRecord(e0);
    k_dmma<ILP, SHAPE><<<sms * blocks_per_sm, 128>>>(d, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flop_per_instr = SHAPE == 0 ? 2.0 * 8 * 8 * 4 : 2.0 * 16 * 8 * 16;
    double inst = (double)sms * blocks_per_sm * 4 * iters * ILP;
    printf("%s ILP=%d blocks/SM=%d: %.3f ms, %.2f TFLOP/s (%.3f warp-mma/clk/SM at 1.8 GHz)\n", name, ILP, blocks_per_sm, ms,
           inst * flop_per_instr / (ms * 1e-3) / 1e12, inst / (ms * 1e-3) / sms / 1.8e9);
}

int main() {
    run_dmma<8, 0>("dmma m8n8k4", 4);
    run_dmma<4, 0>("dmma m8n8k4", 2);
    run_dmma<8, 1>("dmma m16n8k16", 4);
    run_dmma<4, 1>("dmma m16n8k16", 2);
    run<8, 0>("dfma", 4);
    run<16, 0>("dfma", 4);
    run<8, 0>("dfma", 1);
    run<16, 1>("dmul+dfma", 4);
    run<8, 1>("dmul+dfma", 2);
    return 0;
}
