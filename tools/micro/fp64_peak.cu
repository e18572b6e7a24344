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

int main() {
    run<8, 0>("dfma", 4);
    run<16, 0>("dfma", 4);
    run<8, 0>("dfma", 1);
    run<16, 1>("dmul+dfma", 4);
    run<8, 1>("dmul+dfma", 2);
    return 0;
}
