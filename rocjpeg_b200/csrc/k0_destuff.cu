// k0_destuff.cu — GPU pass over the raw entropy-coded bytes (K0), sm_100a.
//
// Takes over, for a whole batch per launch, the per-byte work of the reference's parser - the
// byte-serial search for FF D9 (RocJpegStreamParser::ParseEOI, src/rocjpeg_parser.cpp:400-416) -
// and what the VCN engine does with the slice it is handed (src/rocjpeg_vaapi_decoder.cpp:681:
// the hardware strips byte stuffing and restart markers itself). The host parser stops at the
// end of the SOS header; the bytes behind it are uploaded untouched and this pass
//   * finds the end of the slice (first FF D9),
//   * removes byte stuffing (FF 00 -> FF), fill bytes (FF FF) and restart markers (FF D0..D7),
//   * discovers the restart intervals and writes the batch's segment table (SegmentDesc),
//   * writes the destuffed bytes of interval k at SegmentStart(r_k, k, S) of the image's clean
//     stream (device_types.h) followed by 16 zero bytes - the layout K1 decodes from.
// Rules for damaged streams are those of the host restatement (jpeg_parser.cpp:
// ExtractEntropyData), which is also the expected value in the tests: any other marker inside an
// interval ends that interval's data; intervals beyond the frame's count are dropped; a lone FF
// at the end carries no data.
//
// Formulation. Whether byte p is kept, ends an interval, ends the slice or kills the interval
// depends on bytes p-1, p, p+1 only. Where it goes depends on a prefix over everything before it:
//   (restart markers so far, kept bytes since the last one, position behind the last one,
//    interval dead, slice ended)
// which composes associatively (Combine below), so three launches do it: per-tile reduction,
// one scan over the tiles of each image, per-tile scan + scatter. A tile is 4 KiB of one image,
// one 16-byte vector per thread. The bytes are read twice from L2 (the batch's raw bytes were
// written there a moment ago by the upload) and written once: HBM-bound, ~3 bytes moved per byte.
#include <cuda_runtime.h>

#include "k0_core.cuh"
#include "stages.h"

namespace rjb {
namespace {

using namespace k0;

constexpr int kThreads = 256;
constexpr int kTileBytes = kK0TileBytes;
static_assert(kTileBytes == kThreads * 16, "one 16-byte vector per thread");

__device__ __forceinline__ Elem ShflUp(const Elem& e, int d) {
    Elem r;
    r.nrst = __shfl_up_sync(0xFFFFFFFFu, e.nrst, d);
    r.tail = __shfl_up_sync(0xFFFFFFFFu, e.tail, d);
    r.last_r = __shfl_up_sync(0xFFFFFFFFu, e.last_r, d);
    r.flags = __shfl_up_sync(0xFFFFFFFFu, e.flags, d);
    return r;
}

// Loads and classifies the piece at byte offset `off` of the image's uploaded bytes (off = multiple of 16).
__device__ __forceinline__ Piece LoadPiece(const uint8_t* raw, const ImageDesc& im, uint64_t off) {
    const int64_t len = int64_t(im.raw_len), pos0 = int64_t(off) - int64_t(im.raw_skip);
    uint32_t w[4] = {0u, 0u, 0u, 0u}, prev = 0u, next = 0xFFu;
    if (PieceOverlaps(pos0, len)) {
        const uint8_t* base = raw + im.raw_off + off;
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(base));
        w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
        if (pos0 > 0) prev = __ldg(base - 1);
        if (pos0 + 16 < len) next = __ldg(base + 16);
    }
    return ClassifyPiece(w, prev, next, pos0, len);
}

// largest i in [0, n) with a[i] <= v
__device__ __forceinline__ uint32_t UpperIndexK0(const uint32_t* a, uint32_t n, uint32_t v) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) <= v) lo = mid; else hi = mid;
    }
    return lo;
}

// Scan of `e` over the CTA: *excl = exclusive value of the thread, *total = CTA total (when asked for).
__device__ __forceinline__ void CtaScan(const Elem& e, Elem* warp_tot /* shared [kThreads / 32] */, Elem* excl, Elem* total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Elem inc = e;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const Elem p = ShflUp(inc, d);
        if (lane >= d) inc = Combine(p, inc);
    }
    if (lane == 31) warp_tot[warp] = inc;
    __syncthreads();
    Elem pre{0u, 0u, 0u, 0u};
    for (int w = 0; w < warp; w++) pre = Combine(pre, warp_tot[w]);
    Elem ex = ShflUp(inc, 1);
    if (lane == 0) ex = Elem{0u, 0u, 0u, 0u};
    *excl = Combine(pre, ex);
    if (total) {
        Elem t = pre;
        for (int w = warp; w < kThreads / 32; w++) t = Combine(t, warp_tot[w]);
        *total = t;
    }
}

// ---------------------------------------------------------------- k0_reduce: one element per tile

__global__ void __launch_bounds__(kThreads) k0_reduce(K0Args a) {
    PdlEntry();
    __shared__ Elem s_warp[kThreads / 32];
    const uint32_t tile = blockIdx.x;
    const uint32_t img = UpperIndexK0(a.img_tile0, uint32_t(a.nimages), tile);
    const ImageDesc& im = a.images[img];
    const uint64_t off = uint64_t(tile - im.k0_tile0) * kTileBytes + uint64_t(threadIdx.x) * 16u;
    const Piece pc = LoadPiece(a.raw, im, off);
    Elem ex, tot;
    CtaScan(PieceElem(pc), s_warp, &ex, &tot);
    if (threadIdx.x == 0) a.tile_sum[tile] = make_uint4(tot.nrst, tot.tail, tot.last_r, tot.flags);
}

// ---------------------------------------------------------------- k0_scan: exclusive scan over an image's tiles

__global__ void __launch_bounds__(kThreads) k0_scan(K0Args a) {
    PdlEntry();
    __shared__ Elem s_warp[kThreads / 32];
    const uint32_t img = blockIdx.x;
    const uint32_t t0 = a.img_tile0[img], t1 = a.img_tile0[img + 1];
    Elem carry{0u, 0u, 0u, 0u};
    for (uint32_t base = t0; base < t1; base += kThreads) {
        const uint32_t k = base + threadIdx.x;
        Elem e{0u, 0u, 0u, 0u};
        if (k < t1) {
            const uint4 q = a.tile_sum[k];
            e = Elem{q.x, q.y, q.z, q.w};
        }
        Elem ex, tot;
        __syncthreads();   // the previous chunk's readers of s_warp are done
        CtaScan(e, s_warp, &ex, &tot);
        ex = Combine(carry, ex);
        if (k < t1) a.tile_carry[k] = make_uint4(ex.nrst, ex.tail, ex.last_r, ex.flags);
        carry = Combine(carry, tot);
    }
}

// ---------------------------------------------------------------- k0_apply: scatter + segment table

struct DevMem {
    const K0Args& a;
    const ImageDesc& im;
    uint32_t img;
    uint32_t* fill_from;   // shared: first restart interval the bytes do not contain
    __device__ __forceinline__ SegmentDesc& Segment(uint32_t k) const { return a.segments[im.seg0 + k]; }
    __device__ __forceinline__ uint8_t* Clean() const { return a.clean + im.data_off; }
    __device__ __forceinline__ void Finish(const ScanStatus& st, uint32_t from) const {
        a.status[img] = st;
        *fill_from = from;
    }
};

template <int S>
__global__ void __launch_bounds__(kThreads) k0_apply(K0Args a) {
    PdlEntry();
    __shared__ Elem s_warp[kThreads / 32];
    __shared__ __align__(16) uint8_t s_stage[kThreads / 32][512 + 32];
    __shared__ uint32_t s_fill_from;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    const uint32_t img = UpperIndexK0(a.img_tile0, uint32_t(a.nimages), tile);
    const ImageDesc& im = a.images[img];
    const uint64_t off = uint64_t(tile - im.k0_tile0) * kTileBytes + uint64_t(tid) * 16u;
    const Piece pc = LoadPiece(a.raw, im, off);
    const Elem mine = PieceElem(pc);
    Elem ex;
    if (tid == 0) s_fill_from = 0xFFFFFFFFu;
    CtaScan(mine, s_warp, &ex, nullptr);
    {
        const uint4 q = a.tile_carry[tile];
        ex = Combine(Elem{q.x, q.y, q.z, q.w}, ex);
    }
    DevMem mem{a, im, img, &s_fill_from};
    const Placer<DevMem> pl{im, uint32_t(S), mem};
    const bool ended_before = (ex.flags & kEnded) != 0;

    // Fast path, decided per warp: no marker of any kind in the warp's 512 bytes and the slice has not ended -
    // every lane's kept bytes go to one contiguous run. They are gathered in shared memory at the run's 16-byte
    // phase and leave as 128-bit stores (a byte store per kept byte otherwise).
    const bool simple = !ended_before && (pc.rst | pc.oth | pc.eoi) == 0u;
    if (__all_sync(0xFFFFFFFFu, simple)) {
        const uint32_t n = Popc(pc.keep);
        const uint32_t k0 = __shfl_sync(0xFFFFFFFFu, ex.nrst, 0), r0 = __shfl_sync(0xFFFFFFFFu, ex.last_r, 0),
                       c0 = __shfl_sync(0xFFFFFFFFu, ex.tail, 0), f0 = __shfl_sync(0xFFFFFFFFu, ex.flags, 0);
        uint32_t incl = n;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t p = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += p;
        }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        if (!(f0 & kDead) && pl.Wanted(k0) && total) {
            uint8_t* dst = pl.Dst(r0, k0) + c0;
            const uint32_t phase = uint32_t(reinterpret_cast<uintptr_t>(dst) & 15u);
            uint8_t* st = s_stage[warp] + phase + (incl - n);
            uint32_t keep = pc.keep;
            if (keep == 0xFFFFu && ((phase + incl - n) & 3u) == 0u) {
                reinterpret_cast<uint32_t*>(st)[0] = pc.w[0];
                reinterpret_cast<uint32_t*>(st)[1] = pc.w[1];
                reinterpret_cast<uint32_t*>(st)[2] = pc.w[2];
                reinterpret_cast<uint32_t*>(st)[3] = pc.w[3];
            } else {
                while (keep) {
                    const uint32_t i = LowestBit(keep);
                    keep &= keep - 1u;
                    *st++ = uint8_t(ByteOf(pc, i));
                }
            }
            __syncwarp();
            const uint8_t* buf = s_stage[warp] + phase;
            uint32_t head = (16u - phase) & 15u;
            if (head > total) head = total;
            if (uint32_t(lane) < head) dst[lane] = buf[lane];
            const uint32_t nvec = (total - head) >> 4;
            for (uint32_t v = uint32_t(lane); v < nvec; v += 32)
                reinterpret_cast<uint4*>(dst + head)[v] = reinterpret_cast<const uint4*>(buf + head)[v];
            const uint32_t done = head + (nvec << 4);
            if (done + uint32_t(lane) < total) dst[done + lane] = buf[done + lane];
        }
    } else if (!ended_before && pc.any) {
        WalkPiece(pc, ex, mine, pl);
    }
    FinishPiece(pc, ex, mine, pl, tile == im.k0_tile0 && tid == 0);
    __syncthreads();
    const uint32_t from = s_fill_from;
    if (from != 0xFFFFFFFFu)
        for (uint32_t q = from + uint32_t(tid); q < im.nseg; q += kThreads) pl.Missing(q);
}

}  // namespace

cudaError_t LaunchK0Destuff(const K0Args& a, cudaStream_t stream) {
    if (a.total_tiles == 0) return cudaSuccess;
    cudaError_t e = LaunchPdl(k0_reduce, dim3(a.total_tiles), dim3(kThreads), 0, stream, a);
    if (e == cudaSuccess) e = LaunchPdl(k0_scan, dim3(a.nimages), dim3(kThreads), 0, stream, a);
    if (e != cudaSuccess) return e;
    switch (a.sub_bytes) {
        case 32: return LaunchPdl(k0_apply<32>, dim3(a.total_tiles), dim3(kThreads), 0, stream, a);
        case 64: return LaunchPdl(k0_apply<64>, dim3(a.total_tiles), dim3(kThreads), 0, stream, a);
        case 128: return LaunchPdl(k0_apply<128>, dim3(a.total_tiles), dim3(kThreads), 0, stream, a);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t PreloadK0() {
    cudaFuncAttributes at;
    cudaError_t e = cudaFuncGetAttributes(&at, k0_reduce);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&at, k0_scan);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&at, k0_apply<32>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&at, k0_apply<64>);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&at, k0_apply<128>);
    return e;
}

}  // namespace rjb
