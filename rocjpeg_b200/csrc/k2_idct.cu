// k2_idct.cu — fused dequantise + integer "islow" 8x8 inverse DCT (K2) for sm_100a.
//
// Replaces the dequantisation + IDCT the reference delegates to the VCN
// fixed-function engine (src/rocjpeg_vaapi_decoder.cpp:677-689, 816-828). The
// arithmetic is libjpeg's jidctint.c islow (CONST_BITS 13, PASS1_BITS 2), which
// BASELINE.json mandates: column pass on coef*quant descaled by 11 bits, row pass
// descaled by 18 bits, +128, clamp — bit-exact with libjpeg-turbo's
// jpeg_idct_islow on encoder-produced data (the zero-column shortcuts in
// libjpeg are arithmetically identical to the general path, so none are taken).
//
// HBM-bound integer work on CUDA cores: 128 B read + 64 B written per block.
// Mapping: 8 threads per block (one 16-byte coefficient row each -> perfectly
// coalesced 128 B per block), 32 horizontally adjacent blocks of one component
// per CTA, transposes through padded shared memory (bank-conflict free), one
// 8-byte store per thread, i.e. full 32-byte sectors per block row.
#include <cuda_runtime.h>

#include "stages.h"

namespace rjb {
namespace {

constexpr int kBlocksPerCta = 32;
constexpr int kThreads = kBlocksPerCta * 8;

__device__ __forceinline__ uint32_t UpperIndexK2(const uint32_t* a, uint32_t n, uint32_t v) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) <= v) lo = mid; else hi = mid;
    }
    return lo;
}

// One 8-point pass of the islow IDCT (same butterfly for both passes, only the
// descale differs): even part from i0,i2,i4,i6, odd part from i1,i3,i5,i7.
template <int SHIFT>
__device__ __forceinline__ void Islow8(const int (&i)[8], int (&o)[8]) {
    constexpr int kRound = 1 << (SHIFT - 1);
    int z1 = (i[2] + i[6]) * 4433;
    const int t2 = z1 - i[6] * 15137;
    const int t3 = z1 + i[2] * 6270;
    const int t0 = (i[0] + i[4]) * 8192;
    const int t1 = (i[0] - i[4]) * 8192;
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
    o[0] = (t10 + d + kRound) >> SHIFT; o[7] = (t10 - d + kRound) >> SHIFT;
    o[1] = (t11 + c + kRound) >> SHIFT; o[6] = (t11 - c + kRound) >> SHIFT;
    o[2] = (t12 + b + kRound) >> SHIFT; o[5] = (t12 - b + kRound) >> SHIFT;
    o[3] = (t13 + a + kRound) >> SHIFT; o[4] = (t13 - a + kRound) >> SHIFT;
}

__device__ __forceinline__ int Clamp255(int v) { return min(max(v, 0), 255); }

__global__ void __launch_bounds__(kThreads) k2_idct(K2Args a) {
    // [block][row][col] with row stride 9 and block stride 72 words: both the
    // row-wise and the column-wise access of a warp hit 32 distinct banks.
    __shared__ int ws[kBlocksPerCta * 72];
    __shared__ uint32_t s_img;
    const int tid = threadIdx.x;
    if (tid == 0) s_img = UpperIndexK2(a.img_tile0, uint32_t(a.nimages), blockIdx.x);
    __syncthreads();
    const ImageDesc& im = a.images[s_img];
    uint32_t t = blockIdx.x - a.img_tile0[s_img];
    // tile -> (component, block row, first block column)
    int comp = 0;
    uint32_t tiles_x = 0;
    for (; comp < im.ncomp; comp++) {
        tiles_x = (uint32_t(im.blocks_w[comp]) + kBlocksPerCta - 1) / kBlocksPerCta;
        const uint32_t n = tiles_x * uint32_t(im.blocks_h[comp]);
        if (t < n) break;
        t -= n;
    }
    if (comp >= im.ncomp) return;
    const int by = int(t / tiles_x);
    const int b = tid >> 3, j = tid & 7;
    const int bx = int(t % tiles_x) * kBlocksPerCta + b;
    const bool valid = bx < im.blocks_w[comp];
    int* my = ws + b * 72;
    if (valid) {
        const int H = im.ncomp == 1 ? 1 : im.hs[comp], V = im.ncomp == 1 ? 1 : im.vs[comp];
        const uint32_t mcu = uint32_t(by / V) * uint32_t(im.mcus_x) + uint32_t(bx / H);
        const uint32_t k = uint32_t(im.comp_first_blk[comp] + (by % V) * H + (bx % H));
        const size_t blk = size_t(im.blk0) + size_t(mcu) * im.bpm + k;
        const uint4 cq = __ldg(reinterpret_cast<const uint4*>(a.coef + blk * 64) + j);
        const uint4 qq = __ldg(reinterpret_cast<const uint4*>(a.qtables + size_t(im.qt_index[comp]) * 64) + j);
        const uint32_t cw[4] = {cq.x, cq.y, cq.z, cq.w}, qw[4] = {qq.x, qq.y, qq.z, qq.w};
        int* row = my + j * 9;
#pragma unroll
        for (int w = 0; w < 4; w++) {
            row[2 * w] = int(int16_t(cw[w] & 0xFFFFu)) * int(qw[w] & 0xFFFFu);
            row[2 * w + 1] = int(int16_t(cw[w] >> 16)) * int(qw[w] >> 16);
        }
    }
    __syncwarp();
    int in[8], out[8];
    if (valid) {
#pragma unroll
        for (int r = 0; r < 8; r++) in[r] = my[r * 9 + j];   // column j
        Islow8<11>(in, out);
#pragma unroll
        for (int r = 0; r < 8; r++) my[r * 9 + j] = out[r];
    }
    __syncwarp();
    if (valid) {
#pragma unroll
        for (int c = 0; c < 8; c++) in[c] = my[j * 9 + c];   // row j
        Islow8<18>(in, out);
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int c = 0; c < 4; c++) {
            lo |= uint32_t(Clamp255(out[c] + 128)) << (8 * c);
            hi |= uint32_t(Clamp255(out[c + 4] + 128)) << (8 * c);
        }
        uint8_t* dst = a.planes + im.plane_off[comp] + size_t(by * 8 + j) * im.plane_pitch[comp] + size_t(bx) * 8;
        *reinterpret_cast<uint2*>(dst) = make_uint2(lo, hi);
    }
}

}  // namespace

cudaError_t LaunchK2Idct(const K2Args& a, cudaStream_t stream) {
    if (a.total_tiles == 0) return cudaSuccess;
    k2_idct<<<a.total_tiles, kThreads, 0, stream>>>(a);
    return cudaGetLastError();
}

}  // namespace rjb
