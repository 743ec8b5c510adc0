// micro-benchmark: FFMA vs FFMA2 (packed fp32x2) issue throughput, alone and mixed with ALU/XU work
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b) {
    float2 x[8];
    for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    float2 A = make_float2(a, a), B = make_float2(b, b);
    unsigned u = threadIdx.x;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); }
            if (MODE == 1) { x[i] = __ffma2_rn(x[i], A, B); }
            if (MODE == 2) { x[i] = __ffma2_rn(x[i], A, B); u = (u ^ (u >> 3)) + 0x9E3779B9u; }            // + 2 ALU ops per FFMA2
            if (MODE == 4) { x[i] = __ffma2_rn(x[i], A, B); if (i & 1) u = (u ^ (u >> 3)); }                           // + 1/2 ALU op per FFMA2
            if (MODE == 5) { float sc = __uint_as_float(u | 0x3F000000u); x[i] = __ffma2_rn(x[i], make_float2(sc, sc), B); }   // scalar-broadcast operand (R.F32)
            if (MODE == 3) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); u = (u ^ (u >> 3)) + 0x9E3779B9u; }
        }
    }
    float r = 0; for (int i = 0; i < 8; i++) r += x[i].x + x[i].y;
    if (r == 123.456f || u == 77) out[0] = r;
}
template <int MODE> void run(const char* name) {
    float* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 1 << 14, blocks = 148 * 8;
    k<MODE><<<blocks, 256>>>(d, 256, 0.999f, 0.001f);
    float best = 1e9;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(d, iters, 0.999f, 0.001f); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double fma = 16.0 * iters * 256.0 * blocks;
    printf("%-28s %.3f ms  %.2f TFLOP/s (fp32 fma flops)\n", name, best, 2 * fma / best / 1e9);
}
int main() { run<0>("FFMA x2"); run<1>("FFMA2"); run<2>("FFMA2 + 2 ALU"); run<3>("FFMA x2 + 2 ALU"); run<4>("FFMA2 + 0.5 ALU"); run<5>("FFMA2 bcast operand"); return 0; }
