// idct_core.cuh — the arithmetic of the IDCT stage, shared by k2_idct.cu and the fused IDCT + output kernel
// (k23_fused.cu): libjpeg's jidctint.c islow butterfly (CONST_BITS 13, PASS1_BITS 2), bit-exact with libjpeg-turbo's
// jpeg_idct_islow on encoder-produced data, as BASELINE.json mandates. The reference delegates this to the VCN
// fixed-function engine (src/rocjpeg_vaapi_decoder.cpp:677-689, 816-828).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rjb {
namespace idct {

// zig-zag position -> natural (row-major) position
__device__ __constant__ const uint8_t kZigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                                     12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                                     35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                                     58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// One 8-point pass of the islow IDCT. `bias` (rounding, and in the second pass the +128
// level shift) is added to the even part once instead of to each of the 8 outputs.
template <int SHIFT>
__device__ __forceinline__ void Islow8(const int (&i)[8], int (&o)[8], int bias) {
    int z1 = (i[2] + i[6]) * 4433;
    const int t2 = z1 - i[6] * 15137;
    const int t3 = z1 + i[2] * 6270;
    const int t0 = (i[0] + i[4]) * 8192 + bias;
    const int t1 = (i[0] - i[4]) * 8192 + bias;
    const int t10 = t0 + t3, t13 = t0 - t3, t11 = t1 + t2, t12 = t1 - t2;
    int a = i[7], b = i[5], c = i[3], d = i[1];
    z1 = a + d;
    int z2 = b + c, z3 = a + c, z4 = b + d;
    const int z5 = (z3 + z4) * 9633;
    a *= 2446; b *= 16819; c *= 25172; d *= 12299;
    z1 *= -7373; z2 *= -20995;
    z3 = z3 * -16069 + z5;
    z4 = z4 * -3196 + z5;
    a += z1 + z3; b += z2 + z4; c += z2 + z3; d += z1 + z4;
    o[0] = (t10 + d) >> SHIFT; o[7] = (t10 - d) >> SHIFT;
    o[1] = (t11 + c) >> SHIFT; o[6] = (t11 - c) >> SHIFT;
    o[2] = (t12 + b) >> SHIFT; o[5] = (t12 - b) >> SHIFT;
    o[3] = (t13 + a) >> SHIFT; o[4] = (t13 - a) >> SHIFT;
}

// d = sat_u8(v0) | sat_u8(v1) << 8 | sat_u8(v2) << 16 | sat_u8(v3) << 24
__device__ __forceinline__ uint32_t PackSat4(int v0, int v1, int v2, int v3) {
    uint32_t hi, d;
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(hi) : "r"(v3), "r"(v2), "r"(0));
    asm("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(v1), "r"(v0), "r"(hi));
    return d;
}

}  // namespace idct
}  // namespace rjb
