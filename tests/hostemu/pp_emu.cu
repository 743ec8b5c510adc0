// pp_emu.cu -- TEST INFRASTRUCTURE. Compiles the product's pull-push phases (csrc/pullpush_core.cuh) for the HOST and runs
// the 256 "threads" of each phase in a loop, one 64x64 ISM at a time, so that the arithmetic the CUDA kernel executes can be
// compared with the oracle bit for bit on a machine without a GPU (tests/test_pullpush_hostemu.py). Never linked into
// libmlgi.so; the product has no CPU path.
#include <vector>
#include "../../many-lights-gi_b200/csrc/pullpush_core.cuh"
#include "../../many-lights-gi_b200/csrc/common.cuh"

using namespace mlgi::pp;

extern "C" __attribute__((visibility("default")))
void pp_emu(const uint32_t* atlas, int S, float zFar, float ismPixelSize, uint16_t* ism) {
    const int A = 64 * S;
#pragma omp parallel
    {
        std::vector<Shared> sh(1);
        Shared& s = sh[0];
#pragma omp for schedule(dynamic, 4)
        for (int tile = 0; tile < S * S; tile++) {
            const int gx0 = (tile % S) * 64, gy0 = (tile / S) * 64;
            for (int r = 0; r < 64; r++)
                for (int c = 0; c < 64; c++) s.L0[r * 64 + c] = atlas[(size_t)(gy0 + r) * A + gx0 + c];
#define ALL(call) for (int tid = 0; tid < kThreads; tid++) { call; }
            ALL(phase_tables(s, tid, zFar, ismPixelSize, gx0, gy0))
            ALL(phase_pull1_classify(s, tid))
            ALL(phase_pull1_dense(s, tid, zFar, ismPixelSize, gx0, gy0))
            for (int level = 2; level <= 6; level++) ALL(phase_pull(s, tid, level, gx0, gy0))
            for (int level = 5; level >= 2; level--) ALL(phase_push_top(s, tid, level))
            ALL(phase_push1(s, tid))
            ALL(phase_push0(s, tid))
#undef ALL
            for (int r = 0; r < 64; r++)
                for (int c = 0; c < 64; c++) ism[(size_t)(gy0 + r) * A + gx0 + c] = (uint16_t)s.L0[r * 64 + c];
        }
    }
}

// unpack_ufloat_fast (the branch-free decode the blur / de-interleave kernels use) against unpack_ufloat_bits (the reference
// decode the oracle mirrors), for every 11-bit and 10-bit code: number of codes whose fp32 bit patterns differ
extern "C" __attribute__((visibility("default")))
int ufloat_fast_mismatches() {
    int bad = 0;
    for (uint32_t v = 0; v < 2048; v++) { float f = mlgi::unpack_ufloat_fast<6>(v); uint32_t b; memcpy(&b, &f, 4); bad += b != mlgi::unpack_ufloat_bits<6>(v); }
    for (uint32_t v = 0; v < 1024; v++) { float f = mlgi::unpack_ufloat_fast<5>(v); uint32_t b; memcpy(&b, &f, 4); bad += b != mlgi::unpack_ufloat_bits<5>(v); }
    return bad;
}
