// k0_core.cuh — the per-piece logic of the destuffing pass (K0), shared by the CUDA kernels
// (k0_destuff.cu) and the host model of their schedule (tests/k1_model.cpp: k0_model_destuff).
//
// Replaces the byte-serial FF D9 search of the reference's parser (src/rocjpeg_parser.cpp:400-416)
// and the destuffing the VCN engine does on the slice it is handed. A *piece* is 16 consecutive
// bytes of an image's uploaded data, one thread. Whether byte p is kept, ends a restart interval,
// ends the slice or kills the interval depends on bytes p-1, p, p+1 only (ClassifyPiece). Where it
// goes depends on a prefix over everything before it,
//   (restart markers so far, kept bytes since the last one, raw position behind the last one,
//    interval dead, slice ended),
// which composes associatively (Combine), so a scan distributes it; WalkPiece then places the
// piece's bytes and writes the segment table entries the piece is responsible for.
#pragma once
#include <stdint.h>

#include "device_types.h"

#ifdef __CUDACC__
#define RJB_K0_HD __host__ __device__ __forceinline__
#else
#define RJB_K0_HD inline
#endif

namespace rjb {
namespace k0 {

RJB_K0_HD uint32_t Popc(uint32_t v) {
#ifdef __CUDA_ARCH__
    return uint32_t(__popc(v));
#else
    return uint32_t(__builtin_popcount(v));
#endif
}
RJB_K0_HD uint32_t LowestBit(uint32_t v) {   // index of the lowest set bit, v != 0
#ifdef __CUDA_ARCH__
    return uint32_t(__ffs(int(v))) - 1u;
#else
    return uint32_t(__builtin_ctz(v));
#endif
}
RJB_K0_HD uint32_t BitsBelow(uint32_t i) { return i >= 32u ? 0xFFFFFFFFu : (1u << i) - 1u; }

// Prefix element. flags: kDead = a stray marker was met since the last restart marker (the interval carries
// no more data), kEnded = FF D9 met (everything behind is ignored), kStray = a stray marker was met anywhere.
struct Elem {
    uint32_t nrst, tail, last_r, flags;
};
constexpr uint32_t kDead = 1u, kEnded = 2u, kStray = 4u;

RJB_K0_HD Elem Combine(const Elem& a, const Elem& b) {
    if (a.flags & kEnded) return a;
    Elem r;
    r.nrst = a.nrst + b.nrst;
    if (b.nrst) {
        r.tail = b.tail;
        r.last_r = b.last_r;
        r.flags = b.flags | (a.flags & kStray);
    } else {
        r.tail = (a.flags & kDead) ? a.tail : a.tail + b.tail;
        r.last_r = a.last_r;
        r.flags = a.flags | b.flags;
    }
    return r;
}

// bit i of the result = byte i of the 16 (little-endian words) equals the byte `pattern` repeats, under `and_mask`.
// Exact per-byte zero test of x = (w & and_mask) ^ pattern (no carries across bytes), then the four top bits
// gathered by one multiply.
RJB_K0_HD uint32_t EqMask16(const uint32_t (&w)[4], uint32_t and_mask, uint32_t pattern) {
    uint32_t m = 0;
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int q = 0; q < 4; q++) {
        const uint32_t x = (w[q] & and_mask) ^ pattern;
        const uint32_t z = ~(((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) & 0x80808080u;   // 0x80 in every zero byte of x
        m |= ((z * 0x00204081u) >> 28) << (4 * q);
    }
    return m;
}

// Classification of a piece: bit i of a mask = byte i.
struct Piece {
    uint32_t w[4];    // the bytes (those outside the scan replaced: 00 in front, FF behind)
    uint32_t keep;    // data byte (after destuffing)
    uint32_t rst;     // FF of a restart marker
    uint32_t oth;     // FF of any other marker (not RSTn, not EOI): the interval carries no more data
    uint32_t eoi;     // FF of FF D9
    int64_t pos0;     // scan position of byte 0 (negative inside the leading skip)
    bool any;         // the piece overlaps the scan
};

RJB_K0_HD uint32_t ByteOf(const Piece& pc, uint32_t i) { return (pc.w[i >> 2] >> (8 * (i & 3))) & 0xFFu; }

// The piece's kept bytes, in order, to dst (any alignment); returns their number. Unrolled: the words stay in registers.
RJB_K0_HD uint32_t CompactPiece(const Piece& pc, uint8_t* dst) {
    uint32_t n = 0;
#ifdef __CUDACC__
#pragma unroll
#endif
    for (int i = 0; i < 16; i++)
        if ((pc.keep >> i) & 1u) dst[n++] = uint8_t(pc.w[i >> 2] >> (8 * (i & 3)));
    return n;
}

// Does the piece at scan position pos0 overlap the scan [0, len)?
RJB_K0_HD bool PieceOverlaps(int64_t pos0, int64_t len) { return pos0 + 16 > 0 && pos0 < len; }

// `w`: the 16 bytes as loaded; prev / next: the bytes before and after (only read when inside the scan).
RJB_K0_HD Piece ClassifyPiece(const uint32_t (&w_in)[4], uint32_t prev, uint32_t next, int64_t pos0, int64_t len) {
    Piece pc;
    pc.pos0 = pos0;
    pc.any = PieceOverlaps(pos0, len);
    pc.keep = pc.rst = pc.oth = pc.eoi = 0u;
    for (int q = 0; q < 4; q++) pc.w[q] = w_in[q];
    if (!pc.any) return pc;
    const int64_t lo = pos0 < 0 ? -pos0 : 0;                  // first byte of the piece inside the scan
    const int64_t hi = len - pos0 < 16 ? len - pos0 : 16;     // one past the last
    // scan start: the byte before is not FF; scan end: the byte after is FF, so that a lone FF at the end is a fill byte
    if (pos0 <= 0) prev = 0x00u;
    if (pos0 + 16 >= len) next = 0xFFu;
    if (lo > 0 || hi < 16) {
        for (int i = 0; i < 16; i++) {
            if (i < lo) pc.w[i >> 2] &= ~(0xFFu << (8 * (i & 3)));
            if (i >= hi) pc.w[i >> 2] |= 0xFFu << (8 * (i & 3));
        }
    }
    const uint32_t valid = BitsBelow(uint32_t(hi)) & ~BitsBelow(uint32_t(lo)) & 0xFFFFu;
    const uint32_t F = EqMask16(pc.w, 0xFFFFFFFFu, 0xFFFFFFFFu);
    if (F == 0u && prev != 0xFFu) {   // no FF in or right before the piece (94 % of them): every byte is data
        pc.keep = valid;
        return pc;
    }
    const uint32_t Z = EqMask16(pc.w, 0xFFFFFFFFu, 0u), D = EqMask16(pc.w, 0xF8F8F8F8u, 0xD0D0D0D0u), E = EqMask16(pc.w, 0xFFFFFFFFu, 0xD9D9D9D9u);
    const uint32_t Fn = (F >> 1) | (next == 0xFFu ? 0x8000u : 0u), Zn = (Z >> 1) | (next == 0x00u ? 0x8000u : 0u),
                   Dn = (D >> 1) | ((next & 0xF8u) == 0xD0u ? 0x8000u : 0u), En = (E >> 1) | (next == 0xD9u ? 0x8000u : 0u);
    const uint32_t prevF = ((F << 1) | (prev == 0xFFu ? 1u : 0u)) & 0xFFFFu;
    pc.keep = ((~F & ~prevF) | (F & Zn)) & valid;
    pc.rst = F & Dn & valid;
    pc.eoi = F & En & valid;
    pc.oth = F & ~Zn & ~Fn & ~Dn & ~En & valid;
    return pc;
}

// The piece's prefix element, as if nothing preceded it.
RJB_K0_HD Elem PieceElem(const Piece& pc) {
    Elem e{0u, 0u, 0u, 0u};
    uint32_t ev = pc.rst | pc.oth | pc.eoi;
    if (!ev) {
        e.tail = Popc(pc.keep);
        return e;
    }
    uint32_t start = 0, cnt = 0;
    bool dead = false;
    while (ev) {
        const uint32_t i = LowestBit(ev);
        ev &= ev - 1u;
        if (!dead) cnt += Popc(pc.keep & BitsBelow(i) & ~BitsBelow(start));
        if ((pc.eoi >> i) & 1u) {
            e.tail = cnt;
            e.flags |= kEnded | (dead ? kDead : 0u);
            return e;
        }
        if ((pc.rst >> i) & 1u) {
            e.nrst++;
            e.last_r = uint32_t(pc.pos0 + int64_t(i) + 2);
            cnt = 0;
            dead = false;
        } else {
            dead = true;
            e.flags |= kStray;
        }
        start = i + 1u;
    }
    if (!dead) cnt += Popc(pc.keep & ~BitsBelow(start));
    e.tail = cnt;
    if (dead) e.flags |= kDead;
    return e;
}

// Placement of one image's restart intervals. `Mem` provides Segment(k) -> SegmentDesc&, Clean() -> uint8_t*
// (the image's clean stream), Flag(bits) (OR into the image's status flags) and Finish(segments seen, scan size,
// first missing interval).
template <class Mem>
struct Placer {
    const ImageDesc& im;
    uint32_t S;
    Mem& mem;
    RJB_K0_HD bool Wanted(uint32_t k) const { return k < im.nseg && k >= im.seg_keep_lo && k <= im.seg_keep_hi; }
    RJB_K0_HD uint8_t* Dst(uint32_t r, uint32_t k) const { return mem.Clean() + SegmentStart(r, k, S); }
    // interval k begins at raw position r
    RJB_K0_HD void Open(uint32_t r, uint32_t k) const {
        if (k >= im.nseg) return;
        SegmentDesc& sd = mem.Segment(k);
        const uint64_t c = SegmentStart(r, k, S);
        sd.data_off = im.data_off + c;
        sd.sub0 = im.sub0 + uint32_t(c / S);
        const uint32_t ri = im.restart_interval > 0 ? uint32_t(im.restart_interval) : uint32_t(im.total_mcus);
        const uint64_t m0 = uint64_t(k) * ri;
        const uint64_t left = m0 >= uint64_t(im.total_mcus) ? 0u : uint64_t(im.total_mcus) - m0;
        const uint64_t cnt = left < uint64_t(ri) ? left : uint64_t(ri);
        sd.blk_first = uint32_t(m0 * uint64_t(im.bpm));
        sd.blk_count = uint32_t(cnt * uint64_t(im.bpm));
    }
    // interval k, begun at raw position r, ended with n clean bytes
    RJB_K0_HD void Close(uint32_t r, uint32_t k, uint32_t n) const {
        if (k >= im.nseg) return;
        const bool w = Wanted(k);
        mem.Segment(k).nbytes = w ? n : 0u;   // outside the region of interest: no subsequences, its blocks stay "never decoded"
        if (w && n == 0u) {
            const uint32_t ri = im.restart_interval > 0 ? uint32_t(im.restart_interval) : uint32_t(im.total_mcus);
            if (uint64_t(k) * ri < uint64_t(im.total_mcus)) mem.Flag(kScanEmptyInterval);
        }
        if (w) {
            uint8_t* z = Dst(r, k) + n;
            for (int i = 0; i < 16; i++) z[i] = 0;   // the bit reader may look 16 bytes ahead
        }
    }
    // the slice ended (FF D9 at raw position `scan_size`, or the end of the buffer) inside interval k
    RJB_K0_HD void End(uint32_t r, uint32_t k, uint32_t n, uint32_t scan_size, uint32_t flags) const {
        Close(r, k, n);
        mem.Flag(flags | (k >= im.nseg ? kScanExtraRestarts : 0u));
        mem.Finish(k + 1u, scan_size, k + 1u);
    }
    // intervals the bytes do not contain (truncated file): empty, sorted behind every real subsequence
    RJB_K0_HD void Missing(uint32_t k) const {
        SegmentDesc& sd = mem.Segment(k);
        sd.data_off = im.data_off;
        sd.nbytes = 0;
        sd.sub0 = im.sub0 + im.nsub;
        sd.blk_first = 0;
        sd.blk_count = 0;
    }
};

// General path of the scatter: walks the piece's markers in order, stores its kept bytes one by one and
// opens / closes the restart intervals whose markers lie in the piece. `ex` = prefix of everything before
// the piece (slice not ended), `mine` = PieceElem(pc).
template <class Mem>
RJB_K0_HD void WalkPiece(const Piece& pc, const Elem& ex, const Elem& mine, const Placer<Mem>& pl) {
    uint32_t k = ex.nrst, r = ex.last_r, cnt = ex.tail;
    bool dead = (ex.flags & kDead) != 0;
    uint32_t rest = pc.rst | pc.oth | pc.eoi, start = 0;
    auto emit = [&](uint32_t upto) {   // kept bytes in [start, upto)
        uint32_t m = pc.keep & BitsBelow(upto) & ~BitsBelow(start);
        if (dead || !m) return;
        if (pl.Wanted(k)) {
            uint8_t* dst = pl.Dst(r, k) + cnt;
#ifdef __CUDACC__
#pragma unroll
#endif
            for (int i = 0; i < 16; i++)
                if ((m >> i) & 1u) {
                    *dst++ = uint8_t(pc.w[i >> 2] >> (8 * (i & 3)));
                    cnt++;
                }
        } else {
            cnt += Popc(m);
        }
    };
    while (rest) {
        const uint32_t i = LowestBit(rest);
        rest &= rest - 1u;
        emit(i);
        if ((pc.eoi >> i) & 1u) {   // first FF D9: the slice ends here
            pl.End(r, k, cnt, uint32_t(pc.pos0 + int64_t(i)), ((ex.flags | mine.flags) & kStray) ? kScanStrayMarker : 0u);
            return;
        }
        if ((pc.rst >> i) & 1u) {
            pl.Close(r, k, cnt);
            k++;
            r = uint32_t(pc.pos0 + int64_t(i) + 2);
            cnt = 0;
            dead = false;
            pl.Open(r, k);
        } else {
            dead = true;
        }
        start = i + 1u;
    }
    emit(16u);
}

// Everything a piece does besides the warp-level fast path of the kernel: its walk and, for the piece that
// holds the last byte of a slice without FF D9 (or the first piece of an empty one), the closing of the
// last interval.
template <class Mem>
RJB_K0_HD void FinishPiece(const Piece& pc, const Elem& ex, const Elem& mine, const Placer<Mem>& pl, bool first_piece) {
    if (first_piece) pl.Open(0u, 0u);   // (the kernel opens interval 0 itself and passes false)
    if (pl.im.raw_len == 0u) {
        if (first_piece) pl.End(0u, 0u, 0u, 0u, kScanNoEoi);
        return;
    }
    if ((ex.flags & kEnded) || !pc.any) return;
    if (!(mine.flags & kEnded) && pc.pos0 + 16 >= int64_t(pl.im.raw_len)) {
        const Elem inc = Combine(ex, mine);
        pl.End(inc.last_r, inc.nrst, inc.tail, pl.im.raw_len, kScanNoEoi | ((inc.flags & kStray) ? kScanStrayMarker : 0u));
    }
}

}  // namespace k0
}  // namespace rjb
