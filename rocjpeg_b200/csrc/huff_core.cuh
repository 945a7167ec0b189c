// huff_core.cuh — the sequential heart of the entropy-decode stage (K1).
//
// Baseline Huffman decoding restated from ITU-T T.81 (F.2.2: DECODE, RECEIVE,
// EXTEND, the DC-difference and AC run/size procedures). The reference has no
// counterpart: it hands the entropy-coded slice to the VCN fixed-function block
// (src/rocjpeg_vaapi_decoder.cpp:677-689).
//
// The input is the parser's clean bitstream (jpeg_parser.h): byte stuffing and
// restart markers already removed, so a bit position is a plain integer and
// any thread can start reading at any bit. The decoder state that must agree
// for two decoders to produce identical symbols from a point on is just
//     (bit position, block index inside the MCU, zig-zag index inside the block)
// — DC prediction is not part of it because DC differences are written and
// integrated by a later prefix sum. That is what makes the speculative,
// self-synchronising schedule of k1_huffman.cu possible.
//
// Compiled for the device by nvcc and, with RJB_HD empty, for the host by the
// CPU model of the schedule in tests/k1_model.cpp.
#pragma once
#include <stdint.h>

#include "device_types.h"

#ifdef __CUDACC__
#define RJB_HD __host__ __device__ __forceinline__
#else
#define RJB_HD inline
#endif

namespace rjb {

// Packed decoder state exchanged between neighbouring subsequences.
//   bits 0..5   overflow: bits consumed past the subsequence boundary (0..63)
//   bits 6..9   block index inside the MCU
//   bits 10..15 zig-zag index of the next coefficient (0 = a DC symbol is next)
//   bits 16..31 number of blocks completed inside the subsequence
RJB_HD uint32_t PackState(uint32_t overflow, int c, int z, uint32_t nb) {
    return (overflow & 63u) | (uint32_t(c) << 6) | (uint32_t(z) << 10) | (nb << 16);
}
RJB_HD uint32_t StateOverflow(uint32_t s) { return s & 63u; }
RJB_HD int StateC(uint32_t s) { return int((s >> 6) & 15u); }
RJB_HD int StateZ(uint32_t s) { return int((s >> 10) & 63u); }
RJB_HD uint32_t StateBlocks(uint32_t s) { return s >> 16; }
RJB_HD uint32_t StateKey(uint32_t s) { return s & 0xFFFFu; }   // the part that must agree

RJB_HD uint32_t ByteSwap32(uint32_t w) {
#ifdef __CUDA_ARCH__
    return __byte_perm(w, 0, 0x0123);
#else
    return (w >> 24) | ((w >> 8) & 0xFF00u) | ((w << 8) & 0xFF0000u) | (w << 24);
#endif
}

struct NullSink {
    RJB_HD void Dc(uint32_t, int) const {}
    RJB_HD void Entry(int, int) const {}
    RJB_HD void EndBlock(uint32_t) const {}
};

// Coefficient entry of the sparse coefficient stream K1 writes and K2 expands: one 32-bit word
// per symbol that carries magnitude bits (every non-zero AC coefficient; DC symbols with a
// non-zero difference too, which K2 ignores in favour of the integrated DC).
//   bits 0..15   quantised value, int16
//   bits 16..21  zig-zag index AFTER the symbol = position + 1 (mod 64: 0 stands for position 63)
//   bits 22..31  don't care (K1 leaves decoder state bits there)
// kPadEntry fills the gaps between two threads' runs: value 0 at position 0, which K2
// overwrites with the integrated DC.
RJB_HD uint32_t MakeCoefEntry(int pos, int val) { return (uint32_t(val) & 0xFFFFu) | (uint32_t((pos + 1) & 63) << 16); }
RJB_HD int CoefEntryPos(uint32_t en) { return int(((en >> 16) - 1u) & 63u); }
constexpr uint32_t kPadEntry = 1u << 16;

// Symbol entry of the two-level tables (HuffLutSet::fast / ::sub), 32 bits, laid out so that ONE
// integer add advances the whole per-thread decoder state of k1_huffman.cu
// (acc = bit position | entries produced << 11 | zig-zag index << 21):
//   bits 0..4    code length + SSSS = bits the whole symbol consumes (1..31)
//   bits 5..10   0 (bit 10 set = link, see below)
//   bit  11      1 when the symbol carries magnitude bits (SSSS != 0): it produces a coefficient entry
//   bits 21..27  how far the symbol advances the zig-zag index: 1 for a DC symbol, RRRR + 1 for an
//                AC coefficient or ZRL (15 + 1), 64 for end-of-block (so "z += advance; z >= 64" is
//                the only block-end test)
//   bits 28..31  SSSS, number of magnitude bits that follow the code (adds harmlessly above z)
RJB_HD uint32_t MakeEntry(uint32_t len, uint32_t sym, bool is_ac) {
    const uint32_t s = sym & 15u;
    const uint32_t adv = !is_ac ? 1u : (sym == 0 ? 64u : (sym >> 4) + 1u);
    return (len + s) | (s ? (1u << 11) : 0u) | (adv << 21) | (s << 28);
}
RJB_HD uint32_t EntryBits(uint32_t e) { return e & 31u; }
RJB_HD uint32_t EntrySize(uint32_t e) { return e >> 28; }
RJB_HD int EntryAdvance(uint32_t e) { return int((e >> 21) & 127u); }
// Link entry (first level only, code longer than kFastBits): bits 0..9 = 0, bit 10 = 1 (added to
// the decoder state it raises the same "look closer" flag as the end of the subsequence does),
// bits 11..14 = x, the number of index bits of the sub-table (1..16-kFastBits; 0 = resolve by the
// canonical search), bits 16..31 = index of the sub-table's first entry in HuffLutSet::sub.
RJB_HD uint32_t MakeLink(uint32_t x, uint32_t first) { return (1u << 10) | (x << 11) | (first << 16); }
RJB_HD bool IsLink(uint32_t e) { return (e & (1u << 10)) != 0; }
RJB_HD uint32_t LinkBits(uint32_t e) { return (e >> 11) & 15u; }
RJB_HD uint32_t LinkFirst(uint32_t e) { return e >> 16; }

// Canonical search for codes longer than kFastBits (or invalid ones: 16 bits consumed, symbol 0 —
// T.81 leaves this undefined; the oracle does the same).
RJB_HD uint32_t SlowEntry(const HuffLutSet* lut, uint32_t tab, uint32_t v16) {
    for (int l = kFastBits + 1; l <= 16; l++) {
        if (v16 < lut->upper[tab][l])
            return MakeEntry(uint32_t(l), lut->vals[tab][(int32_t(v16 >> (16 - l)) + lut->valoff[tab][l]) & 255], tab >= uint32_t(kHuffIds));
    }
    return MakeEntry(16, 0, tab >= uint32_t(kHuffIds));
}

RJB_HD uint32_t FunnelLeft(uint32_t hi, uint32_t lo, uint32_t k) {
#ifdef __CUDA_ARCH__
    return __funnelshift_l(lo, hi, k);
#else
    k &= 31u;
    return k ? (hi << k) | (lo >> (32 - k)) : hi;
#endif
}

// Which Huffman tables (HuffLutSet slots) block c of the MCU uses.
struct TableSel {
    uint8_t dc[kMaxBlocksPerMcu], ac[kMaxBlocksPerMcu];
};
RJB_HD uint32_t DcTab(const TableSel& t, int c) { return t.dc[c]; }
RJB_HD uint32_t AcTab(const TableSel& t, int c) { return t.ac[c]; }
RJB_HD TableSel MakeTableSel(const uint8_t* mcu_dc, const uint8_t* mcu_ac, int bpm) {
    TableSel t = {};
    for (int c = 0; c < bpm; c++) {
        t.dc[c] = mcu_dc[c];
        t.ac[c] = mcu_ac[c];
    }
    return t;
}

// Three-word window over the big-endian stream: w0/w1 hold the bits being decoded, w2 is
// fetched one word ahead so the load never sits on the symbol-to-symbol dependency chain.
struct BitWindow {
    uint32_t w0, w1, w2, wi;
    template <class Loader>
    RJB_HD void Init(const Loader& load, uint32_t p) {
        wi = p >> 5;
        w0 = load(wi);
        w1 = load(wi + 1);
        w2 = load(wi + 2);
    }
    RJB_HD uint32_t Peek(uint32_t p) const { return FunnelLeft(w0, w1, p); }   // next 32 bits (shift taken mod 32)
    template <class Loader>
    RJB_HD void Advance(const Loader& load, uint32_t p_new) {   // a symbol is at most 31 bits: one word at most
        if ((p_new >> 5) != wi) {
            wi++;
            w0 = w1;
            w1 = w2;
            w2 = load(wi + 2);
        }
    }
};

// Two-level lookup of the symbol whose code starts at the top of `win` in table `tab` (0..3).
RJB_HD uint32_t LookupSymbol(const HuffLutSet* lut, uint32_t tab, uint32_t win) {
    uint32_t e = lut->fast[tab][win >> (32 - kFastBits)];
    if (IsLink(e)) {
        const uint32_t x = LinkBits(e);
        if (x == 0) e = SlowEntry(lut, tab, win >> 16);
        else e = lut->sub[LinkFirst(e) + ((win << kFastBits) >> (32u - x))];
    }
    return e;
}

// RECEIVE + EXTEND (T.81 F.2.2.1) for an entry `e` whose symbol starts at the top of `win`.
RJB_HD int SymbolValue(uint32_t e, uint32_t win) {
    const uint32_t s = EntrySize(e), tot = EntryBits(e);
    if (s == 0) return 0;
    const uint32_t extra = (win << (tot - s)) >> (32 - s);
    return (extra < (1u << (s - 1))) ? int(extra) - int((1u << s) - 1u) : int(extra);
}

// Decode symbols that START in [p, end_bit). `load(i)` returns the i-th 32-bit word of the
// subsequence in BIG-endian order (bit 31 = first bit of the stream); the backing store is
// zero-padded at least 16 bytes past end_bit.
//   WRITE = false: only the state is tracked (speculation / synchronisation).
//   WRITE = true : `sink` receives Dc(blk, diff) for every DC symbol, Entry(pos, val) for every
//                  symbol with magnitude bits, EndBlock(blk) when a block completes; stops early
//                  at blk_limit. (The CUDA kernels run their own lean loops over the same
//                  primitives; this form serves the host model.)
// On return p >= end_bit (or the block limit was reached); c, z, nb (blocks completed), nnz
// (entries produced), blk updated.
template <bool WRITE, class Loader, class Sink>
RJB_HD void DecodeSpan(const Loader& load, const HuffLutSet* lut, TableSel sel, int bpm, uint32_t& p, uint32_t end_bit, int& c,
                       int& z, uint32_t& nb, uint32_t& nnz, uint32_t& blk, uint32_t blk_limit, Sink& sink) {
    if (p >= end_bit) return;
    BitWindow bw;
    bw.Init(load, p);
    uint32_t dc_tab = DcTab(sel, c), ac_tab = AcTab(sel, c);
    while (p < end_bit) {
        if (WRITE && blk >= blk_limit) break;
        const uint32_t win = bw.Peek(p);
        const uint32_t e = LookupSymbol(lut, (z == 0) ? dc_tab : ac_tab, win);
        const int adv = EntryAdvance(e);
        if (EntrySize(e)) {
            nnz++;
            if (WRITE) sink.Entry((z + adv - 1) & 63, SymbolValue(e, win));
        }
        if (WRITE && z == 0) sink.Dc(blk, SymbolValue(e, win));
        z += adv;                                              // EOB advances by 64, ZRL by 16
        p += EntryBits(e);
        bw.Advance(load, p);
        if (z >= 64) {
            if (WRITE) sink.EndBlock(blk);
            z = 0;
            nb++;
            blk++;
            c = (c + 1 == bpm) ? 0 : c + 1;
            dc_tab = DcTab(sel, c);
            ac_tab = AcTab(sel, c);
        }
    }
}

}  // namespace rjb
