// k23_fused.cu — dequantise + islow IDCT + chroma upsampling + colour conversion in ONE kernel (K2 + K3 fused), sm_100a.
//
// For ROCJPEG_OUTPUT_RGB / RGB_PLANAR of a whole picture the decoded component planes are an intermediate nobody asks
// for: K2 writes them to HBM only for K3 to read them back (a quarter of both kernels' traffic on the 500x375 batch).
// Here a CTA owns one MCU row of a 256-pixel-wide strip of one picture: it expands and transforms the strip's blocks
// (the VCN stage of the reference, src/rocjpeg_vaapi_decoder.cpp:677-689) into shared-memory planes and converts them
// to the caller's pixels right there (the reference's ColorConvertToRGB[Planar] kernels, src/rocjpeg_decoder.cpp:450-557,
// src/rocjpeg_hip_kernels.cpp:52-2029). Algorithmic bytes = coefficient entries + block records read, pixels written.
// The arithmetic is the two stages' own: idct_core.cuh and k3_rows.cuh, bit-exact with the un-fused path (which keeps
// serving crops, the planar formats the IDCT stage writes directly, the interleaved NATIVE surfaces and the test taps).
#include <cuda_runtime.h>

#include "idct_core.cuh"
#include "k3_rows.cuh"
#include "stages.h"

namespace rjb {
namespace {

using namespace idct;
using namespace k3;

constexpr int kStripW = 256;          // luma samples per strip (= kK3TileW: the row routines cover it with 8 per lane)
constexpr int kFThreads = 256;        // 32 blocks x 8 threads per IDCT batch; 8 warps for the rows
constexpr int kMaxBlocks = 128;       // blocks of one strip: 32 MCUs x 4 (4:4:0) or 16 MCUs x 8 ((2x2, 1x2, 1x2) 4:2:2)
constexpr int kLumaBytes = 16 * kStripW, kChromaBytes = 8 * kStripW;   // plane capacities: 16 rows of luma; 8 x 256 or 16 x 128 of chroma
static_assert(kStripW == kTileW, "the row routines assume 8 samples per lane over the strip");

__device__ __forceinline__ uint32_t UpperIndexF(const uint32_t* a, uint32_t n, uint32_t v) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) <= v) lo = mid; else hi = mid;
    }
    return lo;
}

__global__ void __launch_bounds__(kFThreads, 6) k23_fused(K23Args a) {
    PdlEntry();
    constexpr int kRS = 12, kBS = 104;   // workspace strides as in k2_idct.cu (bank-conflict free column / row access)
    __shared__ __align__(16) int ws[32 * kBS];
    __shared__ __align__(16) uint8_t s_pl[kLumaBytes + 2 * kChromaBytes];
    __shared__ __align__(16) uint8_t s_buf[kFThreads / 32][kRowBuf];
    __shared__ uint32_t s_tab[3][64];            // per component and zig-zag code: workspace byte offset | quantiser step << 16
    __shared__ uint32_t s_first[kMaxBlocks];
    __shared__ uint16_t s_count[kMaxBlocks], s_off[kMaxBlocks];
    __shared__ int16_t s_dc[kMaxBlocks];
    __shared__ uint8_t s_comp[kMaxBlocks];
    __shared__ K3Job s_job;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // every thread reads the picture's record itself (uniform loads, one cache line): no thread-0 prologue
    const FusedImage& fi = a.fused[__ldg(a.tile_img + blockIdx.x)];
    const uint32_t t = blockIdx.x - fi.tile0;
    const int mrow = int(t / fi.tiles_x), tx = int(t - uint32_t(mrow) * fi.tiles_x);
    const int m0 = tx * fi.mpt, nm = min(fi.mpt, fi.mcus_x - m0);
    const int ncomp = fi.ncomp;
    const int nbw0 = nm * fi.H[0], nbw1 = ncomp > 1 ? nm * fi.H[1] : 0, nbw2 = ncomp > 2 ? nm * fi.H[2] : 0;
    const int cnt0 = nbw0 * fi.V[0], cnt1 = nbw1 * (ncomp > 1 ? fi.V[1] : 0), cnt2 = nbw2 * (ncomp > 2 ? fi.V[2] : 0);
    const int nblocks = cnt0 + cnt1 + cnt2;
    const uint32_t* entries = a.entries + fi.ent0;
    const BlockRec* rec = a.blk_rec + fi.blk0;
    const int xt = tx * kStripW, rows = 8 * fi.vmax, y0t = mrow * rows;
    if (tid == 0) {
        // the row routines see the strip's shared-memory planes through pointers whose origin is the picture's (0, 0)
        K3Job j;
        for (int c = 0; c < 3; c++) {
            const int cc = c < ncomp ? c : 0;
            const int shx = cc ? fi.sx : 0, shy = cc ? fi.sy : 0;
            j.pitch[c] = fi.pitch[cc];
            j.p[c] = s_pl + fi.base[cc] - (size_t(y0t >> shy) * fi.pitch[cc] + size_t(xt >> shx));
            j.dst[c] = fi.dst[c];
            j.dpitch[c] = fi.dpitch;
        }
        j.dst[3] = nullptr;
        j.dpitch[3] = 0;
        j.W = fi.width; j.H = fi.height; j.x0 = 0; j.y0 = 0; j.css = fi.css; j.fmt = fi.fmt;
        j.xt = xt;
        j.ty = mrow;
        j.nx = min(kStripW, fi.width - xt);
        s_job = j;
    }
    // quantiser / offset tables per component
    for (int idx = tid; idx < ncomp * 64; idx += kFThreads) {
        const int c = idx >> 6, code = idx & 63;   // entries carry position + 1 (huff_core.cuh)
        const int nat = kZigzag[(code + 63) & 63];
        s_tab[c][code] = uint32_t(((nat >> 3) * kRS + (nat & 7)) * 4) | (uint32_t(__ldg(a.qtables + size_t(fi.qidx[c]) * 64 + nat)) << 16);
    }
    // per block of the strip: its record (where its entries lie, integrated DC) and where its samples go
    if (tid < nblocks) {
        int g = tid, c = 0, nbw = nbw0;
        if (g >= cnt0) { g -= cnt0; c = 1; nbw = nbw1; if (g >= cnt1) { g -= cnt1; c = 2; nbw = nbw2; } }
        const int v = g / nbw, bx = g - v * nbw;
        const int H = fi.H[c];
        const size_t blk = size_t(mrow * fi.mcus_x + m0 + (bx >> fi.hshift[c])) * size_t(fi.bpm) + size_t(fi.first_blk[c] + v * H + (bx & (H - 1)));
        const uint2 r = __ldg(reinterpret_cast<const uint2*>(rec + blk));
        uint32_t e0 = blk ? __ldg(&rec[blk - 1].end) : 0u, e1 = r.x;
        if (e0 == kNoEntry || e1 == kNoEntry || e1 < e0 || e1 - e0 > 0xFFFFu || e1 > fi.ent_cap) e1 = e0 = 0;   // never decoded (as k2_idct.cu)
        s_first[tid] = e0;
        s_count[tid] = uint16_t(e1 - e0);
        s_dc[tid] = int16_t(r.y & 0xFFFFu);
        s_comp[tid] = uint8_t(c);
        s_off[tid] = uint16_t(fi.base[c] + uint32_t(v * 8) * fi.pitch[c] + uint32_t(bx * 8));
    }
    __syncthreads();
    // ---- IDCT: 32 blocks at a time, 8 threads per block (k2_idct.cu's inner loop, samples to shared memory) ----
    const int b = tid >> 3, jj = tid & 7;
    int* my = ws + b * kBS;
    const uint32_t my_sa = uint32_t(__cvta_generic_to_shared(my));
#pragma unroll 1
    for (int g0 = 0; g0 < nblocks; g0 += 32) {
        const int g = g0 + b;
        const bool valid = g < nblocks;
        const int c = valid ? s_comp[g] : 0;
        const uint32_t n = valid ? s_count[g] : 0u;
        const uint32_t* ep = entries + (valid ? s_first[g] : 0u) + uint32_t(jj);
        if (valid) {
            int4* row = reinterpret_cast<int4*>(my + jj * kRS);
            row[0] = make_int4(0, 0, 0, 0);
            row[1] = make_int4(0, 0, 0, 0);
        }
        __syncwarp();
        const uint32_t tab_sa = uint32_t(__cvta_generic_to_shared(&s_tab[c][0]));
        auto put = [&](uint32_t en) {
            uint32_t t;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(t) : "r"(tab_sa + ((en >> 14) & 0xFCu)));
            const int v = int(int16_t(en & 0xFFFFu)) * int(t >> 16);
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(my_sa + (t & 0xFFFFu)), "r"(v) : "memory");
        };
        {
            const uint32_t k = uint32_t(jj);
            uint32_t e[4];
#pragma unroll
            for (int u = 0; u < 4; u++) e[u] = (k + 8u * u < n) ? __ldg(ep + 8 * u) : 0u;
#pragma unroll
            for (int u = 0; u < 4; u++)
                if (k + 8u * u < n) put(e[u]);
        }
        for (uint32_t k = uint32_t(jj) + 32u; k < n; k += 8) put(__ldg(ep + (k - uint32_t(jj))));
        __syncwarp();
        if (valid && jj == 0) my[0] = int(s_dc[g]) * int(s_tab[c][1] >> 16);   // integrated DC replaces any DC-difference entry
        __syncwarp();
        int in[8], out[8];
        if (valid) {
#pragma unroll
            for (int r = 0; r < 8; r++) in[r] = my[r * kRS + jj];   // column jj
            Islow8<11>(in, out, 1 << 10);
#pragma unroll
            for (int r = 0; r < 8; r++) my[r * kRS + jj] = out[r];
        }
        __syncwarp();
        if (valid) {
            const int4 lo = *reinterpret_cast<const int4*>(my + jj * kRS), hi = *reinterpret_cast<const int4*>(my + jj * kRS + 4);   // row jj
            in[0] = lo.x; in[1] = lo.y; in[2] = lo.z; in[3] = lo.w; in[4] = hi.x; in[5] = hi.y; in[6] = hi.z; in[7] = hi.w;
            Islow8<18>(in, out, (1 << 17) + (128 << 18));
            *reinterpret_cast<uint2*>(s_pl + s_off[g] + uint32_t(jj) * fi.pitch[c]) =
                make_uint2(PackSat4(out[0], out[1], out[2], out[3]), PackSat4(out[4], out[5], out[6], out[7]));
        }
        __syncwarp();
    }
    __syncthreads();
    // ---- rows: upsample + convert + store (k3_output.cu's RGB rows, planes in shared memory) ----
    const K3Job& j = s_job;
    const int fmt = j.fmt, css = j.css;
    if (j.nx <= 0) return;
    if (fmt == FMT_RGB && (j.dst[0] == nullptr || j.dpitch[0] == 0)) return;
    if (fmt == FMT_RGB_PLANAR && (!j.dst[0] || !j.dst[1] || !j.dst[2] || j.dpitch[0] == 0)) return;
    const int sx = css == CSS_411 ? 2 : (css == CSS_422 || css == CSS_420) ? 1 : 0;
    const int sy = (css == CSS_440 || css == CSS_420) ? 1 : 0;
    const bool gray = (css == CSS_400);
    const uintptr_t bases = fmt == FMT_RGB ? reinterpret_cast<uintptr_t>(j.dst[0]) + size_t(j.xt) * 3
                                           : (reinterpret_cast<uintptr_t>(j.dst[0]) | reinterpret_cast<uintptr_t>(j.dst[1]) |
                                              reinterpret_cast<uintptr_t>(j.dst[2])) + size_t(j.xt);
    uint8_t* buf = s_buf[warp];
    for (int r = warp; r < rows; r += kFThreads / 32) {
        const int y = y0t + r;
        if (y >= j.H) break;
        const bool dst_ok = ((bases | (size_t(y) * j.dpitch[0])) & 3) == 0;
        if (dst_ok) {
            if (sx == 2) RowRgbFast<2, true>(j, sy, gray, y, lane); else if (sx) RowRgbFast<1, true>(j, sy, gray, y, lane); else RowRgbFast<0, true>(j, sy, gray, y, lane);
        } else {
            if (sx == 2) RowRgbStaged<2, true>(j, sy, gray, buf, y, lane); else if (sx) RowRgbStaged<1, true>(j, sy, gray, buf, y, lane);
            else RowRgbStaged<0, true>(j, sy, gray, buf, y, lane);
        }
    }
}

}  // namespace

cudaError_t LaunchK23Fused(const K23Args& a, cudaStream_t stream) {
    if (a.total_tiles == 0) return cudaSuccess;
    return LaunchPdl(k23_fused, dim3(a.total_tiles), dim3(kFThreads), 0, stream, a);
}

cudaError_t PreloadK23() {
    cudaFuncAttributes at;
    return cudaFuncGetAttributes(&at, k23_fused);
}

}  // namespace rjb
