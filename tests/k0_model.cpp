// k0_model.cpp — CPU model of the K0 (destuffing) schedule (TEST TOOL, not product, not oracle).
//
// Runs the product's own parser (csrc/jpeg_parser.cpp) for the header and the product's own per-piece
// logic (csrc/k0_core.cuh, compiled for the host) through a serial emulation of what k0_destuff.cu does
// in parallel: 16-byte pieces, tiles of 16 KiB, Kogge-Stone warp scans + warp totals inside a tile (the kernels
// fold four pieces per thread first: another bracketing of the same associative composition)
// (k0_reduce / k0_apply), a chunked scan over the tiles of the image (k0_scan), then the scatter walk of
// every piece and the segment table. The no-GPU suite compares its result with the host restatement of
// the same rules (StreamParser::host_scan); the CUDA kernels themselves - including their warp-level
// fast path, which this model does not have - are checked by the -m gpu tests against the same.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "jpeg_parser.h"
#include "k0_core.cuh"

using namespace rjb;
using namespace rjb::k0;

namespace {
constexpr int kThreads = 1024, kTile = 16384;   // the kernels' tile: 256 threads x four 16-byte pieces

// the kernels' CtaScan: inclusive Kogge-Stone per warp of 32, then the warp totals
void CtaScanModel(const std::vector<Elem>& e, std::vector<Elem>* excl, Elem* total) {
    const size_t n = e.size();   // kThreads
    std::vector<Elem> inc(e);
    for (size_t w0 = 0; w0 < n; w0 += 32)
        for (int d = 1; d < 32; d <<= 1) {
            std::vector<Elem> prev(inc.begin() + long(w0), inc.begin() + long(w0) + 32);
            for (int l = d; l < 32; l++) inc[w0 + size_t(l)] = Combine(prev[size_t(l - d)], prev[size_t(l)]);
        }
    excl->assign(n, Elem{0, 0, 0, 0});
    Elem pre{0, 0, 0, 0};
    for (size_t w0 = 0; w0 < n; w0 += 32) {
        for (int l = 0; l < 32; l++) (*excl)[w0 + size_t(l)] = Combine(pre, l ? inc[w0 + size_t(l) - 1] : Elem{0, 0, 0, 0});
        pre = Combine(pre, inc[w0 + 31]);
    }
    *total = pre;
}

struct HostMem {
    std::vector<SegmentDesc>* segs;
    uint8_t* clean;
    ScanStatus* status;
    uint32_t* fill_from;
    SegmentDesc& Segment(uint32_t k) const { return (*segs)[k]; }
    uint8_t* Clean() const { return clean; }
    void Flag(uint32_t bits) const { status->flags |= bits; }
    void Finish(uint32_t segments_seen, uint32_t scan_size, uint32_t from) const {
        status->segments_seen = segments_seen;
        status->scan_size = scan_size;
        *fill_from = from;
    }
};
}  // namespace

// Destuffs the scan of `data` as the device would with subsequence size S, the scan starting `skip` bytes into a
// 16-byte vector, restart intervals [keep_lo, keep_hi] wanted. Outputs per restart interval: clean length, offset in
// the clean stream, first subsequence; the clean stream itself (clean_cap bytes available, pre-filled by the caller);
// status4 = {segments seen, scan size, flags, nseg}. Returns 0, or a negative RocJpegStatus.
extern "C" int k0_model_destuff(const uint8_t* data, size_t len, int S, int skip, uint32_t keep_lo, uint32_t keep_hi, uint32_t* seg_nbytes,
                                uint64_t* seg_off, uint32_t* seg_sub0, size_t seg_cap, uint8_t* clean, size_t clean_cap, uint32_t* status4) {
    StreamParser parser;
    if (!parser.Parse(data, len)) return -3;
    const ParsedJpeg& p = parser.parsed();
    if (p.support_status != 0) return p.support_status;
    const RawScan& rs = parser.raw();
    ImageDesc im = {};
    im.restart_interval = p.restart_interval;
    im.total_mcus = p.mcus_x * p.mcus_y;
    im.bpm = p.bpm;
    im.raw_skip = uint32_t(skip & 15);
    im.raw_len = rs.nbytes;
    im.nseg = p.nseg;
    im.seg_keep_lo = keep_lo;
    im.seg_keep_hi = keep_hi;
    im.data_off = 0;
    im.sub0 = 0;
    const uint64_t cap = CleanCapacity(rs.nbytes, p.nseg, uint32_t(S));
    im.nsub = uint32_t(cap / uint32_t(S));
    if (cap > clean_cap || p.nseg > seg_cap) return -2;
    // the uploaded bytes: junk in front (the 16-byte vector the scan starts in) and behind
    const size_t up = (size_t(im.raw_skip) + rs.nbytes + 15) / 16 * 16;
    std::vector<uint8_t> raw(up + 32, 0xFF);
    for (size_t i = 0; i < im.raw_skip; i++) raw[i] = uint8_t(0xFF - (i & 1) * 0x27);   // FF D8 FF D8 ...: must be ignored
    std::memcpy(raw.data() + im.raw_skip, rs.host, rs.nbytes);
    for (size_t i = size_t(im.raw_skip) + rs.nbytes; i < raw.size(); i++) raw[i] = (i & 1) ? 0xD9 : 0x00;
    const size_t ntiles = std::max<size_t>(1, (size_t(im.raw_skip) + rs.nbytes + kTile - 1) / kTile);
    auto load = [&](size_t off) {
        const int64_t pos0 = int64_t(off) - int64_t(im.raw_skip), L = int64_t(im.raw_len);
        uint32_t w[4] = {0, 0, 0, 0}, prev = 0, next = 0xFF;
        if (PieceOverlaps(pos0, L)) {
            std::memcpy(w, raw.data() + off, 16);
            if (pos0 > 0) prev = raw[off - 1];
            if (pos0 + 16 < L) next = raw[off + 16];
        }
        return ClassifyPiece(w, prev, next, pos0, L);
    };
    // k0_reduce
    std::vector<Elem> tile_sum(ntiles), tile_carry(ntiles);
    std::vector<Elem> e(kThreads), ex;
    for (size_t t = 0; t < ntiles; t++) {
        for (int i = 0; i < kThreads; i++) e[size_t(i)] = PieceElem(load(t * kTile + size_t(i) * 16));
        CtaScanModel(e, &ex, &tile_sum[t]);
    }
    // k0_scan
    Elem carry{0, 0, 0, 0};
    for (size_t base = 0; base < ntiles; base += kThreads) {
        for (int i = 0; i < kThreads; i++) e[size_t(i)] = base + size_t(i) < ntiles ? tile_sum[base + size_t(i)] : Elem{0, 0, 0, 0};
        Elem tot;
        CtaScanModel(e, &ex, &tot);
        for (int i = 0; i < kThreads && base + size_t(i) < ntiles; i++) tile_carry[base + size_t(i)] = Combine(carry, ex[size_t(i)]);
        carry = Combine(carry, tot);
    }
    // k0_apply
    std::vector<SegmentDesc> segs(p.nseg, SegmentDesc{0xDEADBEEFull, 0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu, 0xDEADBEEFu});
    ScanStatus status = {0xFFFFFFFFu, 0, 0, 0};   // flags start at zero as the tile reduction leaves them
    uint32_t fill_from = 0xFFFFFFFFu;
    HostMem mem{&segs, clean, &status, &fill_from};
    const Placer<HostMem> pl{im, uint32_t(S), mem};
    for (size_t t = 0; t < ntiles; t++) {
        std::vector<Piece> pcs;
        for (int i = 0; i < kThreads; i++) {
            pcs.push_back(load(t * kTile + size_t(i) * 16));
            e[size_t(i)] = PieceElem(pcs.back());
        }
        Elem tot;
        CtaScanModel(e, &ex, &tot);
        for (int i = 0; i < kThreads; i++) {
            const Elem x = Combine(tile_carry[t], ex[size_t(i)]);
            if (!(x.flags & kEnded) && pcs[size_t(i)].any) WalkPiece(pcs[size_t(i)], x, e[size_t(i)], pl);
            FinishPiece(pcs[size_t(i)], x, e[size_t(i)], pl, t == 0 && i == 0);
        }
    }
    if (fill_from == 0xFFFFFFFFu) return -8;   // nobody closed the slice: a bug in the schedule
    for (uint32_t q = fill_from; q < im.nseg; q++) pl.Missing(q);
    for (uint32_t k = 0; k < p.nseg; k++) {
        seg_nbytes[k] = segs[k].nbytes;
        seg_off[k] = segs[k].data_off;
        seg_sub0[k] = segs[k].sub0;
    }
    status4[0] = status.segments_seen;
    status4[1] = status.scan_size;
    status4[2] = status.flags;
    status4[3] = p.nseg;
    return 0;
}

// Serial emulation of the warp-level run logic of k0_apply (chunks of four pieces per thread, 32 chunks per warp, the
// staged runs of chunks without markers): checks the invariants the kernel's shared-memory staging relies on. Returns
// the number of violations (0 = fine); details of the first few go to `msg`.
extern "C" int k0_model_check_runs(const uint8_t* data, size_t len, int S, int skip, char* msg, size_t msg_cap) {
    StreamParser parser;
    if (!parser.Parse(data, len)) return -3;
    const ParsedJpeg& p = parser.parsed();
    const RawScan& rs = parser.raw();
    const int64_t L = int64_t(rs.nbytes);
    const size_t up = (size_t(skip & 15) + rs.nbytes + 15) / 16 * 16;
    std::vector<uint8_t> raw(up + 128, 0xFF);
    std::memcpy(raw.data() + (skip & 15), rs.host, rs.nbytes);
    auto load = [&](size_t off) {
        const int64_t pos0 = int64_t(off) - int64_t(skip & 15);
        uint32_t w[4] = {0, 0, 0, 0}, prev = 0, next = 0xFF;
        if (PieceOverlaps(pos0, L)) {
            std::memcpy(w, raw.data() + off, 16);
            if (pos0 > 0) prev = raw[off - 1];
            if (pos0 + 16 < L) next = raw[off + 16];
        }
        return ClassifyPiece(w, prev, next, pos0, L);
    };
    const size_t nchunks = ((size_t(skip & 15) + rs.nbytes + 16383) / 16384) * 256;
    struct Ch { Elem e, ex; bool plain, overlaps; };
    std::vector<Ch> ch(nchunks);
    Elem run{0, 0, 0, 0};
    for (size_t t = 0; t < nchunks; t++) {
        Ch& c = ch[t];
        c.e = Elem{0, 0, 0, 0};
        c.plain = true;
        const int64_t pos0 = int64_t(t * 64) - int64_t(skip & 15);
        c.overlaps = pos0 + 64 > 0 && pos0 < L;
        for (int j = 0; j < 4; j++) {
            const Piece pc = load(t * 64 + size_t(j) * 16);
            if (!pc.any) continue;
            c.e = Combine(c.e, PieceElem(pc));
            if (pc.rst | pc.oth | pc.eoi) c.plain = false;
        }
        c.ex = run;
        run = Combine(run, c.e);
    }
    int bad = 0;
    size_t mo = 0;
    auto say = [&](const char* what, size_t warp, int lane, uint32_t a, uint32_t b, uint32_t c3) {
        if (bad++ < 6 && mo + 160 < msg_cap)
            mo += size_t(snprintf(msg + mo, msg_cap - mo, "%s: warp %zu lane %d a=%u b=%u c=%u; ", what, warp, lane, a, b, c3));
    };
    for (size_t w0 = 0; w0 < nchunks; w0 += 32) {
        uint32_t todo = 0;
        bool staged[32];
        for (int l = 0; l < 32; l++) {
            const Ch& c = ch[w0 + size_t(l)];
            staged[l] = c.plain && c.overlaps && !(c.ex.flags & kEnded) && !(c.ex.flags & kDead) && c.e.tail != 0 && c.ex.nrst < p.nseg;
            if (staged[l]) todo |= 1u << l;
        }
        int guard = 0;
        while (todo) {
            if (++guard > 40) { say("loop does not end", w0 / 32, 0, todo, 0, 0); break; }
            const int leader = __builtin_ctz(todo);
            const Ch& cl = ch[w0 + size_t(leader)];
            uint32_t members = 0;
            for (int l = 0; l < 32; l++)
                if (staged[l] && ch[w0 + size_t(l)].ex.nrst == cl.ex.nrst) members |= 1u << l;
            if (!(members & (1u << leader))) say("leader not a member", w0 / 32, leader, members, todo, 0);
            const int last = 31 - __builtin_clz(members);
            uint32_t expect = cl.ex.tail;
            for (int l = 0; l < 32; l++) {
                if (!(members & (1u << l))) continue;
                const Ch& c = ch[w0 + size_t(l)];
                if (c.ex.tail != expect) say("member bytes not contiguous", w0 / 32, l, c.ex.tail, expect, members);
                if (c.ex.last_r != cl.ex.last_r) say("member of another interval start", w0 / 32, l, c.ex.last_r, cl.ex.last_r, members);
                expect = c.ex.tail + c.e.tail;
            }
            const Ch& cz = ch[w0 + size_t(last)];
            const uint32_t total = cz.ex.tail + cz.e.tail - cl.ex.tail;
            if (total > 2048u) say("run longer than the staging buffer", w0 / 32, last, total, cz.ex.tail, cl.ex.tail);
            todo &= ~members;
        }
    }
    return bad;
}

// The kernels' CtaScan over 256 chunk elements, literally: per warp either the integer fast path (every chunk plain)
// or Kogge-Stone with Combine; then either the all-plain CTA shortcut or the warp totals folded with Combine.
static void CtaScanKernelModel(const std::vector<Elem>& e, const std::vector<char>& plain, std::vector<Elem>* excl, Elem* total) {
    std::vector<Elem> inc(e), wt(8), exl(256);
    bool cta_plain = true;
    for (int w = 0; w < 8; w++) {
        bool wp = true;
        for (int l = 0; l < 32; l++) wp = wp && plain[size_t(w * 32 + l)];
        cta_plain = cta_plain && wp;
        if (wp) {
            uint32_t t = 0;
            for (int l = 0; l < 32; l++) { t += e[size_t(w * 32 + l)].tail; inc[size_t(w * 32 + l)].tail = t; }
        } else {
            for (int d = 1; d < 32; d <<= 1) {
                std::vector<Elem> prev(inc.begin() + w * 32, inc.begin() + w * 32 + 32);
                for (int l = d; l < 32; l++) inc[size_t(w * 32 + l)] = Combine(prev[size_t(l - d)], prev[size_t(l)]);
            }
        }
        wt[size_t(w)] = inc[size_t(w * 32 + 31)];
        for (int l = 0; l < 32; l++) exl[size_t(w * 32 + l)] = l ? inc[size_t(w * 32 + l - 1)] : Elem{0, 0, 0, 0};
    }
    excl->assign(256, Elem{0, 0, 0, 0});
    if (cta_plain) {
        uint32_t pre = 0;
        for (int w = 0; w < 8; w++) {
            for (int l = 0; l < 32; l++) { Elem x = exl[size_t(w * 32 + l)]; x.tail += pre; (*excl)[size_t(w * 32 + l)] = x; }
            pre += wt[size_t(w)].tail;
        }
        *total = Elem{0, pre, 0, 0};
        return;
    }
    Elem pre{0, 0, 0, 0};
    for (int w = 0; w < 8; w++) {
        for (int l = 0; l < 32; l++) (*excl)[size_t(w * 32 + l)] = Combine(pre, exl[size_t(w * 32 + l)]);
        pre = Combine(pre, wt[size_t(w)]);
    }
    *total = pre;
}

// Compares, for every 64-byte chunk of the scan, the prefix the kernels' scan bracketing gives with the sequential one.
extern "C" int k0_model_check_scan(const uint8_t* data, size_t len, int skip, char* msg, size_t msg_cap) {
    StreamParser parser;
    if (!parser.Parse(data, len)) return -3;
    const RawScan& rs = parser.raw();
    const int64_t L = int64_t(rs.nbytes);
    const size_t up = (size_t(skip & 15) + rs.nbytes + 15) / 16 * 16;
    std::vector<uint8_t> raw(up + 128, 0xFF);
    std::memcpy(raw.data() + (skip & 15), rs.host, rs.nbytes);
    auto load = [&](size_t off) {
        const int64_t pos0 = int64_t(off) - int64_t(skip & 15);
        uint32_t w[4] = {0, 0, 0, 0}, prev = 0, next = 0xFF;
        if (PieceOverlaps(pos0, L)) {
            std::memcpy(w, raw.data() + off, 16);
            if (pos0 > 0) prev = raw[off - 1];
            if (pos0 + 16 < L) next = raw[off + 16];
        }
        return ClassifyPiece(w, prev, next, pos0, L);
    };
    const size_t ntiles = std::max<size_t>(1, (size_t(skip & 15) + rs.nbytes + 16383) / 16384);
    std::vector<Elem> e(256), ex;
    std::vector<char> plain(256);
    std::vector<Elem> tile_sum(ntiles), tile_carry(ntiles);
    auto chunk = [&](size_t t, int i, char* pl) {
        Elem c{0, 0, 0, 0};
        *pl = 1;
        for (int j = 0; j < 4; j++) {
            const Piece pc = load(t * 16384 + size_t(i) * 64 + size_t(j) * 16);
            if (!pc.any) continue;
            c = Combine(c, PieceElem(pc));
            if (pc.rst | pc.oth | pc.eoi) *pl = 0;
        }
        return c;
    };
    for (size_t t = 0; t < ntiles; t++) {
        for (int i = 0; i < 256; i++) e[size_t(i)] = chunk(t, i, &plain[size_t(i)]);
        CtaScanKernelModel(e, plain, &ex, &tile_sum[t]);
    }
    Elem carry{0, 0, 0, 0};
    for (size_t base = 0; base < ntiles; base += 256) {
        std::vector<char> np(256, 0);
        for (int i = 0; i < 256; i++) e[size_t(i)] = base + size_t(i) < ntiles ? tile_sum[base + size_t(i)] : Elem{0, 0, 0, 0};
        Elem tot;
        CtaScanKernelModel(e, np, &ex, &tot);
        for (int i = 0; i < 256 && base + size_t(i) < ntiles; i++) tile_carry[base + size_t(i)] = Combine(carry, ex[size_t(i)]);
        carry = Combine(carry, tot);
    }
    int bad = 0;
    size_t mo = 0;
    Elem run{0, 0, 0, 0};
    for (size_t t = 0; t < ntiles; t++) {
        for (int i = 0; i < 256; i++) e[size_t(i)] = chunk(t, i, &plain[size_t(i)]);
        Elem tot;
        CtaScanKernelModel(e, plain, &ex, &tot);
        for (int i = 0; i < 256; i++) {
            const Elem x = Combine(tile_carry[t], ex[size_t(i)]);
            const bool same = (run.flags & kEnded) ? (x.flags & kEnded) != 0 : (x.nrst == run.nrst && x.tail == run.tail && x.last_r == run.last_r && x.flags == run.flags);
            if (!same && bad++ < 4 && mo + 200 < msg_cap)
                mo += size_t(snprintf(msg + mo, msg_cap - mo, "tile %zu chunk %d: kernel {%u %u %u %x} sequential {%u %u %u %x} plain %d; ", t, i, x.nrst, x.tail,
                                      x.last_r, x.flags, run.nrst, run.tail, run.last_r, run.flags, int(plain[size_t(i)])));
            run = Combine(run, e[size_t(i)]);
        }
    }
    return bad;
}
