// Exhaustive check (all 2^32 float bit patterns) of a three-instruction division by a constant against IEEE division:
//   c = RN(1/b);  q0 = RN(x*c);  r = RN(fma(-b, q0, x));  q = RN(fma(r, c, q0))        (Markstein's sequence)
// Prints, per divisor, how many inputs disagree with __fdiv_rn(x, b) and the magnitude range in which they fall.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o divcheck divcheck.cu ; run on a GPU.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void check(float b, float c, unsigned long long* bad, unsigned int* lo_exp, unsigned int* hi_exp) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32); i += stride) {
        const float x = __uint_as_float((unsigned int)i);
        const float want = __fdiv_rn(x, b);
        const float q0 = __fmul_rn(x, c);
        const float r = __fmaf_rn(-b, q0, x);
        const float q = __fmaf_rn(r, c, q0);
        const bool same = (__float_as_uint(q) == __float_as_uint(want)) || (q != q && want != want);
        if (!same) {
            atomicAdd(bad, 1ull);
            const unsigned int e = ((unsigned int)i >> 23) & 0xFF;
            atomicMin(lo_exp, e); atomicMax(hi_exp, e);
        }
    }
}
int main() {
    const float divisors[] = {2.f, 3.f, 4.f, 5.f, 6.f, 7.f, 8.f, 9.f, 10.f, 11.f, 12.f, 13.f, 14.f, 15.f, 16.f, 3.14f};
    unsigned long long* bad; unsigned int *lo, *hi;
    cudaMalloc(&bad, 8); cudaMalloc(&lo, 4); cudaMalloc(&hi, 4);
    for (float b : divisors) {
        unsigned long long z = 0; unsigned int l = 255, h = 0;
        cudaMemcpy(bad, &z, 8, cudaMemcpyHostToDevice); cudaMemcpy(lo, &l, 4, cudaMemcpyHostToDevice); cudaMemcpy(hi, &h, 4, cudaMemcpyHostToDevice);
        const float c = 1.0f / b;     // host IEEE division: correctly rounded reciprocal
        check<<<148 * 16, 256>>>(b, c, bad, lo, hi);
        cudaDeviceSynchronize();
        cudaMemcpy(&z, bad, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&l, lo, 4, cudaMemcpyDeviceToHost); cudaMemcpy(&h, hi, 4, cudaMemcpyDeviceToHost);
        printf("b = %-6g mismatches %llu  biased exponent range of x [%u, %u]\n", b, z, l, h);
    }
    return 0;
}
