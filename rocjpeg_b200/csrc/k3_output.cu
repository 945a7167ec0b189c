// k3_output.cu — fused chroma upsampling + colour conversion + output layout (K3), sm_100a.
//
// One kernel replaces the reference's 14 post-processing kernels and its
// CopyChannel memcpys (src/rocjpeg_hip_kernels.cpp:52-2233,
// src/rocjpeg_decoder.cpp:372-636): it reads the decoded component planes at
// their coded resolution and writes the caller's buffers directly in the
// requested RocJpegOutputFormat, for every image of a batch in one launch.
//
// Arithmetic is the reference's, bit for bit:
//   * nearest-neighbour chroma replication (hip_kernels.cpp:585-617 for 4:4:0,
//     :947-954 for 4:2:2, :1389-1429 for 4:2:0);
//   * BT.709 full-range fmaf chains (hip_kernels.cpp:76-89):
//       R = fmaf(1.5748, V-128, Y); G = fmaf(-0.4681, V-128, fmaf(-0.1873, U-128, Y));
//       B = fmaf(1.8556, U-128, Y);
//   * saturating round-to-nearest-even float->u8 pack (hipPack, :25-30; the
//     rounding of v_cvt_pk_u8_f32 is this repository's documented convention);
//   * 4:0:0 replicates Y into R, G, B (hip_kernels.cpp:1915-1927, 1986-1991).
// Layout semantics (NATIVE surfaces, planar chroma sizes, pitch rules, ROI)
// follow src/rocjpeg_decoder.cpp:143-180 and :372-636; see orc_convert in
// oracle/jpeg_oracle.c for the ROI convention and the three reference defects
// that are deliberately not reproduced.
//
// Unlike the reference kernels (8 px x 2 rows per thread, no tail guard, reliant
// on allocation slack — samples/rocjpeg_samples_utils.h:387-392) every store is
// exact-bounds. Destination rows may start at any byte (pitch = 3*width is the
// samples' choice): each warp assembles one output row segment in shared memory
// at the destination's 16-byte phase and writes it with 128-bit stores in the
// aligned body, byte stores only in the unaligned head and tail.
#include <cuda_runtime.h>

#include "stages.h"

namespace rjb {
namespace {

constexpr int kTileW = kK3TileW;   // luma samples per tile row (8 per lane)
constexpr int kTileH = kK3TileH;   // rows per tile: each of the 8 warps walks kTileH / 8 of them
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kChanBuf = kTileW + 32;      // one planar channel row + alignment phase
constexpr int kRowBuf = 3 * kChanBuf;      // >= 3 * kTileW + 32 (packed RGB row)

__device__ __forceinline__ uint32_t UpperIndexK3(const uint32_t* a, uint32_t n, uint32_t v) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) <= v) lo = mid; else hi = mid;
    }
    return lo;
}

// hipPack convention: saturating round-to-nearest-even float -> u8 (one F2I on sm_100a).
__device__ __forceinline__ uint32_t PackU8(float f) {
    uint32_t r;
    asm("cvt.rni.sat.u8.f32 %0, %1;" : "=r"(r) : "f"(f));
    return r;
}
__device__ __forceinline__ uint32_t Pack4(uint32_t b0, uint32_t b1, uint32_t b2, uint32_t b3) {
    return __byte_perm(__byte_perm(b0, b1, 0x0040), __byte_perm(b2, b3, 0x0040), 0x5410);
}
__device__ __forceinline__ float ByteF(uint32_t w, int i) { return float((w >> (8 * i)) & 0xFFu); }

// Write `n` bytes staged at buf[phase .. phase+n) to dst (dst & 15 == phase).
__device__ __forceinline__ void FlushRow(const uint8_t* buf, int phase, uint8_t* dst, int n, int lane) {
    int head = (16 - phase) & 15;
    if (head > n) head = n;
    if (lane < head) dst[lane] = buf[phase + lane];
    const int nvec = (n - head) >> 4;
    const uint4* s = reinterpret_cast<const uint4*>(buf + phase + head);
    uint4* d = reinterpret_cast<uint4*>(dst + head);
    for (int v = lane; v < nvec; v += 32) d[v] = s[v];
    const int done = head + (nvec << 4);
    if (done + lane < n) dst[done + lane] = buf[phase + done + lane];
}

// Load 8 consecutive samples of a plane row starting at column x (any alignment) as two
// little-endian words.
__device__ __forceinline__ uint2 Load8(const uint8_t* row, int x) {
    const uint8_t* p = row + x;
    if ((reinterpret_cast<uintptr_t>(p) & 7) == 0) return __ldg(reinterpret_cast<const uint2*>(p));
    uint32_t b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) b[i] = __ldg(p + i);
    return make_uint2(Pack4(b[0], b[1], b[2], b[3]), Pack4(b[4], b[5], b[6], b[7]));
}

// Stage `n` bytes produced 8 per lane (two words) at buf[phase + 8*lane ..): word stores when
// the phase allows, byte stores otherwise and in the ragged last lane.
__device__ __forceinline__ void Stage8(uint8_t* buf, int phase, int lane, int n, uint2 v) {
    const int i0 = lane * 8;
    if (i0 >= n) return;
    uint8_t* o = buf + phase + i0;
    if ((phase & 3) == 0 && i0 + 8 <= n) {
        *reinterpret_cast<uint32_t*>(o) = v.x;
        *reinterpret_cast<uint32_t*>(o + 4) = v.y;
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++)
            if (i0 + i < n) o[i] = uint8_t(((i < 4 ? v.x : v.y) >> (8 * (i & 3))) & 0xFFu);
    }
}

// Copy `n` bytes of one plane row (columns x .. x+n) into the caller's row.
__device__ __forceinline__ void CopyRow(uint8_t* buf, const uint8_t* src_row, int x, uint8_t* dst, int n, int lane) {
    // fast case (warp-uniform): aligned plane loads, word-aligned destination -> registers only
    if (((reinterpret_cast<uintptr_t>(src_row) + size_t(x)) & 7) == 0 && (reinterpret_cast<uintptr_t>(dst) & 3) == 0) {
        const int m = n - lane * 8;
        if (m > 0) {
            const uint2 v = __ldg(reinterpret_cast<const uint2*>(src_row + x) + lane);
            uint8_t* d = dst + lane * 8;
            if (m >= 8) {
                if ((reinterpret_cast<uintptr_t>(dst) & 7) == 0) {
                    *reinterpret_cast<uint2*>(d) = v;
                } else {
                    reinterpret_cast<uint32_t*>(d)[0] = v.x;
                    reinterpret_cast<uint32_t*>(d)[1] = v.y;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 8; i++)
                    if (i < m) d[i] = uint8_t(((i < 4 ? v.x : v.y) >> (8 * (i & 3))) & 0xFFu);
            }
        }
        return;
    }
    const int phase = int(reinterpret_cast<uintptr_t>(dst) & 15);
    // planes are MCU-padded (and the arena has slack): reading up to 7 samples past n is in bounds
    if (lane * 8 < n) Stage8(buf, phase, lane, n, Load8(src_row, x + lane * 8));
    __syncwarp();
    FlushRow(buf, phase, dst, n, lane);
    __syncwarp();
}

// Everything a CTA needs to know about its tile, resolved once by thread 0 (the kernel was
// dominated by per-thread setup when each warp did a single row: profiles/r01b_*).
struct K3Job {
    const uint8_t* p[3];
    uint32_t pitch[3];
    uint8_t* dst[4];
    uint32_t dpitch[4];
    int W, H, x0, y0, css, fmt, xt, nx, ty;
};

// One output row segment of RGB / RGB_PLANAR (nx pixels starting at column xt of row y).
__device__ __forceinline__ void RowRgb(const K3Job& j, int sx, int sy, bool gray, uint8_t* buf, int y, int lane) {
    const int nx = j.nx, xt = j.xt;
    const int Y = j.y0 + y;
    const int X = j.x0 + xt + lane * 8;    // first luma column of this lane
    const bool have = lane * 8 < nx;
    uint2 R = make_uint2(0, 0), G = R, B = R;   // 8 packed bytes per channel
    if (have) {
        const uint2 yy = Load8(j.p[0] + size_t(Y) * j.pitch[0], X);
        if (gray) {
            R = G = B = yy;   // hip_kernels.cpp:1915-1927
        } else {
            const uint8_t* urow = j.p[1] + size_t(Y >> sy) * j.pitch[1];
            const uint8_t* vrow = j.p[2] + size_t(Y >> sy) * j.pitch[2];
            // chroma bytes per luma pixel (nearest neighbour): `pair` = the 4 bytes of uu.x/vv.x
            // each serve two pixels (aligned 4:2:x fast path), otherwise one byte per pixel.
            uint2 uu, vv;
            const bool pair = (sx == 1) && ((X & 7) == 0);
            if (sx == 0) {
                uu = Load8(urow, X);
                vv = Load8(vrow, X);
            } else if (pair) {
                uu = make_uint2(__ldg(reinterpret_cast<const uint32_t*>(urow + (X >> 1))), 0);
                vv = make_uint2(__ldg(reinterpret_cast<const uint32_t*>(vrow + (X >> 1))), 0);
            } else {
                uint32_t ub[8], vb[8];
#pragma unroll
                for (int i = 0; i < 8; i++) {
                    ub[i] = __ldg(urow + ((X + i) >> sx));
                    vb[i] = __ldg(vrow + ((X + i) >> sx));
                }
                uu = make_uint2(Pack4(ub[0], ub[1], ub[2], ub[3]), Pack4(ub[4], ub[5], ub[6], ub[7]));
                vv = make_uint2(Pack4(vb[0], vb[1], vb[2], vb[3]), Pack4(vb[4], vb[5], vb[6], vb[7]));
            }
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t yw = h ? yy.y : yy.x;
                uint32_t r[4], g[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float fy = ByteF(yw, i);
                    float fu, fv;
                    if (pair) {
                        fu = ByteF(uu.x, 2 * h + (i >> 1)) - 128.0f;
                        fv = ByteF(vv.x, 2 * h + (i >> 1)) - 128.0f;
                    } else {
                        fu = ByteF(h ? uu.y : uu.x, i) - 128.0f;
                        fv = ByteF(h ? vv.y : vv.x, i) - 128.0f;
                    }
                    r[i] = PackU8(fmaf(1.5748f, fv, fy));
                    g[i] = PackU8(fmaf(-0.4681f, fv, fmaf(-0.1873f, fu, fy)));
                    b[i] = PackU8(fmaf(1.8556f, fu, fy));
                }
                const uint32_t rw = Pack4(r[0], r[1], r[2], r[3]), gw = Pack4(g[0], g[1], g[2], g[3]),
                               bw = Pack4(b[0], b[1], b[2], b[3]);
                if (h) { R.y = rw; G.y = gw; B.y = bw; } else { R.x = rw; G.x = gw; B.x = bw; }
            }
        }
    }
    if (j.fmt == FMT_RGB) {
        uint8_t* dst = j.dst[0] + size_t(y) * j.dpitch[0] + size_t(xt) * 3;
        const int phase = int(reinterpret_cast<uintptr_t>(dst) & 15);
        if (have) {
            // interleave R,G,B bytes: 8 pixels -> 6 words
            uint32_t w[6];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const uint32_t r = h ? R.y : R.x, g = h ? G.y : G.x, b = h ? B.y : B.x;
                w[3 * h + 0] = __byte_perm(__byte_perm(r, g, 0x1040), b, 0x3410);   // r0 g0 b0 r1
                w[3 * h + 1] = __byte_perm(__byte_perm(g, b, 0x2051), r, 0x3610);   // g1 b1 r2 g2
                w[3 * h + 2] = __byte_perm(__byte_perm(b, r, 0x3702), g, 0x3720);   // b2 r3 g3 b3
            }
            uint8_t* o = buf + phase + lane * 24;
            if ((phase & 3) == 0 && lane * 8 + 8 <= nx) {
#pragma unroll
                for (int k = 0; k < 6; k++) reinterpret_cast<uint32_t*>(o)[k] = w[k];
            } else {
#pragma unroll
                for (int k = 0; k < 24; k++)
                    if (lane * 8 + k / 3 < nx) o[k] = uint8_t((w[k >> 2] >> (8 * (k & 3))) & 0xFFu);
            }
        }
        __syncwarp();
        FlushRow(buf, phase, dst, nx * 3, lane);
    } else {
        // all three planes use pitch[0] (src/rocjpeg_decoder.cpp:526-544)
        uint8_t* dst[3];
        int phase[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
            dst[c] = j.dst[c] + size_t(y) * j.dpitch[0] + xt;
            phase[c] = int(reinterpret_cast<uintptr_t>(dst[c]) & 15);
            Stage8(buf + c * kChanBuf, phase[c], lane, nx, c == 0 ? R : c == 1 ? G : B);
        }
        __syncwarp();
#pragma unroll
        for (int c = 0; c < 3; c++) FlushRow(buf + c * kChanBuf, phase[c], dst[c], nx, lane);
    }
    __syncwarp();
}

// ---- fast RGB rows ------------------------------------------------------------------------
// The general row routine above stages every row in shared memory so that any destination
// alignment gets 128-bit stores; it costs ~65 instructions per pixel (profiles/r01c_*), four times
// the arithmetic. When the tile's first plane column is a multiple of 8 (always, unless a crop
// starts at an odd multiple) and the destination row is at least 4-byte aligned, a lane can load
// its 8 samples per plane with one aligned load each, convert in registers and store words
// directly: no staging, no per-byte work.

// byte k of w as float, minus `bias`, exactly: the byte is dropped into the mantissa of 2^23
__device__ __forceinline__ float ByteToFloat(uint32_t w, uint32_t sel, float magic) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, sel)) - magic;
}

struct Rgb8 {
    uint2 r, g, b;   // 8 packed bytes per channel
};

// 8 pixels: yy = 8 luma bytes; chroma bytes per pixel given by (uw, vw) words and the byte
// selectors in csel (pixel i uses chroma byte csel[i] of the pair {lo, hi}).
template <int SX>
__device__ __forceinline__ Rgb8 Convert8(uint2 yy, uint2 uu, uint2 vv) {
    constexpr float kY = 8388608.0f, kC = 8388608.0f + 128.0f;
    float fu[8], fv[8];
    if (SX == 0) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            fu[i] = ByteToFloat(uu.x, 0x7650 + i, kC);
            fu[4 + i] = ByteToFloat(uu.y, 0x7650 + i, kC);
            fv[i] = ByteToFloat(vv.x, 0x7650 + i, kC);
            fv[4 + i] = ByteToFloat(vv.y, 0x7650 + i, kC);
        }
    } else if (SX == 1) {
#pragma unroll
        for (int i = 0; i < 4; i++) {   // chroma byte i serves pixels 2i and 2i+1
            fu[2 * i] = fu[2 * i + 1] = ByteToFloat(uu.x, 0x7650 + i, kC);
            fv[2 * i] = fv[2 * i + 1] = ByteToFloat(vv.x, 0x7650 + i, kC);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; i++) {   // 4:1:1: chroma byte i serves pixels 4i .. 4i+3
            fu[i] = ByteToFloat(uu.x, 0x7650 + (i >> 2), kC);
            fv[i] = ByteToFloat(vv.x, 0x7650 + (i >> 2), kC);
        }
    }
    uint32_t r[8], g[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const float fy = ByteToFloat(i < 4 ? yy.x : yy.y, 0x7650 + (i & 3), kY);
        r[i] = PackU8(fmaf(1.5748f, fv[i], fy));
        g[i] = PackU8(fmaf(-0.4681f, fv[i], fmaf(-0.1873f, fu[i], fy)));
        b[i] = PackU8(fmaf(1.8556f, fu[i], fy));
    }
    Rgb8 o;
    o.r = make_uint2(Pack4(r[0], r[1], r[2], r[3]), Pack4(r[4], r[5], r[6], r[7]));
    o.g = make_uint2(Pack4(g[0], g[1], g[2], g[3]), Pack4(g[4], g[5], g[6], g[7]));
    o.b = make_uint2(Pack4(b[0], b[1], b[2], b[3]), Pack4(b[4], b[5], b[6], b[7]));
    return o;
}

__device__ __forceinline__ void StoreBytes(uint8_t* d, uint2 v, int n) {   // first n (< 8) bytes of v
#pragma unroll
    for (int i = 0; i < 8; i++)
        if (i < n) d[i] = uint8_t(((i < 4 ? v.x : v.y) >> (8 * (i & 3))) & 0xFFu);
}
__device__ __forceinline__ void Store8(uint8_t* d, uint2 v, bool al8) {   // d is 4-byte aligned at least
    if (al8) {
        *reinterpret_cast<uint2*>(d) = v;
    } else {
        reinterpret_cast<uint32_t*>(d)[0] = v.x;
        reinterpret_cast<uint32_t*>(d)[1] = v.y;
    }
}

// One output row segment (nx pixels from column xt of output row y), RGB or RGB_PLANAR.
// Preconditions (checked by the caller, warp-uniform): (x0 + xt) % 8 == 0, destination row(s) 4-byte aligned.
template <int SX>
__device__ __forceinline__ void RowRgbFast(const K3Job& j, int sy, bool gray, int y, int lane) {
    const int nx = j.nx, xt = j.xt;
    const int n = nx - lane * 8;   // pixels this lane owns (may be <= 0 or < 8 at the right edge)
    if (n <= 0) return;
    const int Y = j.y0 + y;
    const int X = j.x0 + xt + lane * 8;
    const uint2 yy = __ldg(reinterpret_cast<const uint2*>(j.p[0] + size_t(Y) * j.pitch[0] + X));
    Rgb8 o;
    if (gray) {
        o.r = o.g = o.b = yy;   // hip_kernels.cpp:1915-1927
    } else {
        const uint8_t* urow = j.p[1] + size_t(Y >> sy) * j.pitch[1];
        const uint8_t* vrow = j.p[2] + size_t(Y >> sy) * j.pitch[2];
        uint2 uu, vv;
        if (SX == 0) {
            uu = __ldg(reinterpret_cast<const uint2*>(urow + X));
            vv = __ldg(reinterpret_cast<const uint2*>(vrow + X));
        } else if (SX == 1) {
            uu = make_uint2(__ldg(reinterpret_cast<const uint32_t*>(urow + (X >> 1))), 0u);
            vv = make_uint2(__ldg(reinterpret_cast<const uint32_t*>(vrow + (X >> 1))), 0u);
        } else {
            uu = make_uint2(__ldg(reinterpret_cast<const uint16_t*>(urow + (X >> 2))), 0u);
            vv = make_uint2(__ldg(reinterpret_cast<const uint16_t*>(vrow + (X >> 2))), 0u);
        }
        o = Convert8<SX>(yy, uu, vv);
    }
    if (j.fmt == FMT_RGB) {
        uint8_t* d = j.dst[0] + size_t(y) * j.dpitch[0] + size_t(xt + lane * 8) * 3;
        // interleave R,G,B bytes: 8 pixels -> 6 words
        uint32_t w[6];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const uint32_t r = h ? o.r.y : o.r.x, g = h ? o.g.y : o.g.x, b = h ? o.b.y : o.b.x;
            w[3 * h + 0] = __byte_perm(__byte_perm(r, g, 0x1040), b, 0x3410);   // r0 g0 b0 r1
            w[3 * h + 1] = __byte_perm(__byte_perm(g, b, 0x2051), r, 0x3610);   // g1 b1 r2 g2
            w[3 * h + 2] = __byte_perm(__byte_perm(b, r, 0x3702), g, 0x3720);   // b2 r3 g3 b3
        }
        if (n >= 8) {
            if ((reinterpret_cast<uintptr_t>(d) & 7) == 0) {
#pragma unroll
                for (int k = 0; k < 3; k++) reinterpret_cast<uint2*>(d)[k] = make_uint2(w[2 * k], w[2 * k + 1]);
            } else {
#pragma unroll
                for (int k = 0; k < 6; k++) reinterpret_cast<uint32_t*>(d)[k] = w[k];
            }
        } else {
#pragma unroll
            for (int k = 0; k < 24; k++)
                if (k < 3 * n) d[k] = uint8_t((w[k >> 2] >> (8 * (k & 3))) & 0xFFu);
        }
    } else {
        // all three planes use pitch[0] (src/rocjpeg_decoder.cpp:526-544)
        const size_t off = size_t(y) * j.dpitch[0] + size_t(xt + lane * 8);
        uint8_t* d0 = j.dst[0] + off;
        uint8_t* d1 = j.dst[1] + off;
        uint8_t* d2 = j.dst[2] + off;
        if (n >= 8) {
            const bool al8 = ((reinterpret_cast<uintptr_t>(d0) | reinterpret_cast<uintptr_t>(d1) | reinterpret_cast<uintptr_t>(d2)) & 7) == 0;
            Store8(d0, o.r, al8);
            Store8(d1, o.g, al8);
            Store8(d2, o.b, al8);
        } else {
            StoreBytes(d0, o.r, n);
            StoreBytes(d1, o.g, n);
            StoreBytes(d2, o.b, n);
        }
    }
}

__global__ void __launch_bounds__(kThreads) k3_output(K3Args a) {
    PdlEntry();
    __shared__ __align__(16) uint8_t s_buf[kWarps][kRowBuf];
    __shared__ K3Job s_job;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) {
        const uint32_t img = UpperIndexK3(a.img_tile0, uint32_t(a.nimages), blockIdx.x);
        const ImageDesc& im = a.images[img];
        const OutputDesc& od = a.outputs[img];
        const uint32_t t = blockIdx.x - a.img_tile0[img];
        K3Job j;
        for (int c = 0; c < 3; c++) {
            j.p[c] = a.planes + im.plane_off[c < im.ncomp ? c : 0];
            j.pitch[c] = im.plane_pitch[c < im.ncomp ? c : 0];
        }
        for (int c = 0; c < 4; c++) {
            j.dst[c] = od.dst[c];
            j.dpitch[c] = od.dst_pitch[c];
        }
        j.W = od.w; j.H = od.h; j.x0 = od.x0; j.y0 = od.y0; j.css = im.css; j.fmt = od.fmt;
        j.xt = int(t % od.tiles_x) * kTileW;
        j.ty = int(t / od.tiles_x);
        j.nx = min(kTileW, od.w - j.xt);
        s_job = j;
    }
    __syncthreads();
    const K3Job& j = s_job;
    const int W = j.W, H = j.H, x0 = j.x0, y0 = j.y0, css = j.css, fmt = j.fmt;
    const int sx = css == CSS_411 ? 2 : (css == CSS_422 || css == CSS_420) ? 1 : 0;   // chroma shift: 4:1:1 has one chroma sample per four pixels
    const int sy = (css == CSS_440 || css == CSS_420) ? 1 : 0;
    const bool gray = (css == CSS_400);
    uint8_t* buf = s_buf[warp];
    const int xt = j.xt, nx = j.nx, ty = j.ty;
    if (nx <= 0) return;

    if (fmt == FMT_RGB || fmt == FMT_RGB_PLANAR) {
        if (fmt == FMT_RGB && (j.dst[0] == nullptr || j.dpitch[0] == 0)) return;
        if (fmt == FMT_RGB_PLANAR && (!j.dst[0] || !j.dst[1] || !j.dst[2] || j.dpitch[0] == 0)) return;
        // fast rows need aligned plane loads and word-aligned destination rows (warp-uniform tests)
        const bool src_ok = ((x0 + xt) & 7) == 0;
        const uintptr_t bases = fmt == FMT_RGB ? reinterpret_cast<uintptr_t>(j.dst[0]) + size_t(xt) * 3
                                               : (reinterpret_cast<uintptr_t>(j.dst[0]) | reinterpret_cast<uintptr_t>(j.dst[1]) |
                                                  reinterpret_cast<uintptr_t>(j.dst[2])) + size_t(xt);
        for (int r = warp; r < kTileH; r += kWarps) {
            const int y = ty * kTileH + r;
            if (y >= H) break;
            const bool dst_ok = ((bases | (size_t(y) * j.dpitch[0])) & 3) == 0;
            if (src_ok && dst_ok) {
                if (sx == 2) RowRgbFast<2>(j, sy, gray, y, lane); else if (sx) RowRgbFast<1>(j, sy, gray, y, lane); else RowRgbFast<0>(j, sy, gray, y, lane);
            } else {
                RowRgb(j, sx, sy, gray, buf, y, lane);
            }
        }
        return;
    }

    // ---- NATIVE / YUV_PLANAR / Y: plane copies and re-interleaving ----
    const bool native = (fmt == FMT_NATIVE);
    if (native && css == CSS_422) {
        // packed YUYV: byte k of the surface row is Y[k>>1] (k even), U[k>>2] (k%4==1), V[k>>2] (k%4==3);
        // the caller's row starts at surface byte 2*left (src/rocjpeg_decoder.cpp:384-388)
        if (j.dst[0] == nullptr || j.dpitch[0] == 0) return;
        for (int r = warp; r < kTileH; r += kWarps) {
            const int y = ty * kTileH + r;
            if (y >= H) break;
            const int Y = y0 + y;
            uint8_t* dst = j.dst[0] + size_t(y) * j.dpitch[0] + size_t(xt) * 2;
            const int phase = int(reinterpret_cast<uintptr_t>(dst) & 15);
            const uint8_t* yrow = j.p[0] + size_t(Y) * j.pitch[0];
            const uint8_t* urow = j.p[1] + size_t(Y) * j.pitch[1];
            const uint8_t* vrow = j.p[2] + size_t(Y) * j.pitch[2];
            const int n = nx * 2;
            for (int b = lane; b < n; b += 32) {
                const int k = 2 * (x0 + xt) + b;
                uint8_t v;
                if ((k & 1) == 0) v = __ldg(yrow + (k >> 1));
                else v = (k & 2) ? __ldg(vrow + (k >> 2)) : __ldg(urow + (k >> 2));
                buf[phase + b] = v;
            }
            __syncwarp();
            FlushRow(buf, phase, dst, n, lane);
            __syncwarp();
        }
        return;
    }
    // luma (every remaining format writes channel 0 = Y)
    if (j.dst[0] != nullptr && j.dpitch[0] != 0) {
        for (int r = warp; r < kTileH; r += kWarps) {
            const int y = ty * kTileH + r;
            if (y >= H) break;
            CopyRow(buf, j.p[0] + size_t(y0 + y) * j.pitch[0], x0 + xt, j.dst[0] + size_t(y) * j.dpitch[0] + xt, nx, lane);
        }
    }
    if (fmt == FMT_Y || gray) return;
    if (native && css == CSS_420) {
        // interleaved UV rows: surface byte k is U[k>>1] (k even) / V[k>>1] (k odd); the caller's row
        // starts at surface byte `left` of chroma row top>>1 (src/rocjpeg_decoder.cpp:380-388)
        if (j.dst[1] == nullptr || j.dpitch[1] == 0) return;
        const int rows = H >> 1;
        for (int r = warp; r < kTileH / 2; r += kWarps) {
            const int cy = ty * (kTileH / 2) + r;
            if (cy >= rows) break;
            uint8_t* dst = j.dst[1] + size_t(cy) * j.dpitch[1] + xt;
            const int phase = int(reinterpret_cast<uintptr_t>(dst) & 15);
            const uint8_t* urow = j.p[1] + size_t((y0 >> 1) + cy) * j.pitch[1];
            const uint8_t* vrow = j.p[2] + size_t((y0 >> 1) + cy) * j.pitch[2];
            for (int b = lane; b < nx; b += 32) {
                const int k = x0 + xt + b;
                buf[phase + b] = (k & 1) ? __ldg(vrow + (k >> 1)) : __ldg(urow + (k >> 1));
            }
            __syncwarp();
            FlushRow(buf, phase, dst, nx, lane);
            __syncwarp();
        }
        return;
    }
    // planar chroma at the coded resolution: (W>>sx) x (H>>sy) starting at (left>>sx, top>>sy).
    // 4:4:4 / 4:4:0 honour each channel's pitch; 4:2:2 / 4:2:0 use pitch[1] for both
    // (src/rocjpeg_decoder.cpp:589-590, 596-597, 600-601).
    const int cw = W >> sx, ch = H >> sy;
    const int rows_per_tile = kTileH >> sy, cols_per_tile = kTileW >> sx;
    const int cxt = (xt >> sx);
    const int cn = min(cols_per_tile, cw - cxt);
    if (cn <= 0) return;
    for (int task = warp; task < 2 * rows_per_tile; task += kWarps) {
        const int c = 1 + (task >= rows_per_tile ? 1 : 0);
        const int cy = ty * rows_per_tile + (task >= rows_per_tile ? task - rows_per_tile : task);
        const uint32_t pitch = (sx == 0) ? j.dpitch[c] : j.dpitch[1];
        if (cy >= ch || j.dst[c] == nullptr || pitch == 0) continue;
        CopyRow(buf, j.p[c] + size_t((y0 >> sy) + cy) * j.pitch[c], (x0 >> sx) + cxt, j.dst[c] + size_t(cy) * pitch + cxt, cn, lane);
    }
}

}  // namespace

cudaError_t LaunchK3Output(const K3Args& a, cudaStream_t stream) {
    if (a.total_tiles == 0) return cudaSuccess;
    return LaunchPdl(k3_output, dim3(a.total_tiles), dim3(kThreads), 0, stream, a);
}

// Forces the module holding this stage's kernels onto the device (CUDA loads lazily: the first launch
// of every kernel would otherwise pay for it inside the first decode call).
cudaError_t PreloadK3() {
    cudaFuncAttributes at;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&at, k3_output);
    return e;
}

}  // namespace rjb
