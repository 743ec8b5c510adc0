// micro-benchmark: issue rate of the packed fp32x2 instructions by operand form (which of them cost one FMA-pipe slot, which two),
// and of the instruction pairs the gathering loop is built from. 8 independent chains per thread, 8 CTAs of 256 threads per SM.
// Prints warp-instructions per SM sub-partition and cycle (1.0 = one instruction per cycle per scheduler).
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float a, float b, unsigned seed) {
    float2 x[8];
    a *= 1.0f + (threadIdx.x & 1) * 1e-7f; b *= 1.0f + (threadIdx.x & 2) * 1e-7f;      // per-thread values: register operands, not uniform registers
    for (int i = 0; i < 8; i++) x[i] = make_float2(threadIdx.x + i, threadIdx.x - i);
    const float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 0.999f);
    unsigned u = threadIdx.x ^ seed;
    float s = a;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); }                 // 2 x FFMA (3 registers)
            if (MODE == 1) x[i] = __ffma2_rn(x[i], A, B);                                                  // FFMA2, three 64-bit operands
            if (MODE == 2) x[i] = __ffma2_rn(x[i], make_float2(s, s), B);                                  // FFMA2, one scalar-broadcast operand
            if (MODE == 3) x[i] = __fmul2_rn(x[i], A);                                                     // FMUL2
            if (MODE == 4) x[i] = __fadd2_rn(x[i], B);                                                     // FADD2
            if (MODE == 5) x[i] = __fadd2_rn(x[i], make_float2(-1.5f, -1.5f));                             // FADD2 immediate
            if (MODE == 6) { x[i].x = fmaf(x[i].x, 0.999f, 0.001f); x[i].y = fmaf(x[i].y, 0.999f, 0.001f); }   // 2 x FFMA immediate
            if (MODE == 7) { x[i].x = fminf(x[i].x, a); x[i].y = fmaxf(x[i].y, b); }                       // 2 x FMNMX (ALU pipe)
            if (MODE == 8) { x[i] = __ffma2_rn(x[i], A, B); u = (u ^ (u >> 3)) + 0x9E3779B9u; }            // FFMA2 + 2 ALU
            if (MODE == 9) { x[i] = __ffma2_rn(x[i], A, B); x[(i + 1) & 7].x = fminf(x[(i + 1) & 7].x, a); }   // FFMA2 + FMNMX
            if (MODE == 10) { x[i] = __fmul2_rn(x[i], make_float2(s, s)); }                                 // FMUL2 with scalar-broadcast
            if (MODE == 11) { x[i].x = x[i].x * a; x[i].y = x[i].y * b; }                                   // 2 x FMUL
            if (MODE == 12) { x[i].x = x[i].x + a; x[i].y = x[i].y + b; }                                   // 2 x FADD
            if (MODE == 13) { if (i < 4) x[i] = __ffma2_rn(x[i], A, B); else { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); } }   // 4 FFMA2 + 8 FFMA per trip
            if (MODE == 14) { if (i < 6) x[i] = __ffma2_rn(x[i], A, B); else { x[i].x = fmaf(x[i].x, a, b); x[i].y = fmaf(x[i].y, a, b); } }   // 6 FFMA2 + 4 FFMA per trip
        }
        if (MODE == 2 || MODE == 10) s = __uint_as_float((__float_as_uint(s) & 0x3FFFFFFFu) | 0x3F000000u);
    }
    float r = 0; for (int i = 0; i < 8; i++) r += x[i].x + x[i].y;
    if (r == 123.456f || u == 77) out[0] = r + s;
}
static int g_sms = 148; static double g_ghz = 1.965;
template <int MODE> void run(const char* name, double inst_per_iter) {
    float* d; cudaMalloc(&d, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int iters = 1 << 13, blocks = g_sms * 8;
    k<MODE><<<blocks, 256>>>(d, 256, 0.999f, 0.001f, 1u);
    float best = 1e9;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(d, iters, 0.999f, 0.001f, 1u); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    const double warp_inst = inst_per_iter * iters * 8.0 * blocks;          // 8 warps per CTA
    const double cycles = best * 1e-3 * g_ghz * 1e9;
    printf("%-40s %.3f ms   %.3f warp-inst / cycle / SMSP\n", name, best, warp_inst / cycles / (g_sms * 4.0));
    cudaFree(d);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0); g_sms = p.multiProcessorCount; g_ghz = p.clockRate * 1e-6;
    printf("%s, %d SMs, %.3f GHz nominal\n", p.name, g_sms, g_ghz);
    run<0>("2 x FFMA (reg)", 16); run<6>("2 x FFMA (imm)", 16); run<11>("2 x FMUL", 16); run<12>("2 x FADD", 16);
    run<1>("FFMA2 (3 x 64-bit)", 8); run<2>("FFMA2 (scalar-broadcast operand)", 8); run<3>("FMUL2", 8); run<10>("FMUL2 (scalar-broadcast)", 8);
    run<4>("FADD2", 8); run<5>("FADD2 (imm)", 8); run<7>("2 x FMNMX", 16); run<8>("FFMA2 + 2 ALU", 24); run<9>("FFMA2 + FMNMX", 16);
    run<13>("4 FFMA2 + 8 FFMA (flop-equal mix)", 12); run<14>("6 FFMA2 + 4 FFMA", 10);
    return 0;
}
