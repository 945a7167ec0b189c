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
    RJB_HD void Ac(uint32_t, int, int) const {}
};

// Decode symbols that START in [p, end_bit). Words come from `load(i)`: the i-th
// little-endian 32-bit word counted from the subsequence's first byte (the
// backing store is zero-padded at least 16 bytes past end_bit).
//   WRITE = false: only the state is tracked (speculation / synchronisation).
//   WRITE = true : coefficients go to `sink`; stops early at blk_limit.
// On return p >= end_bit (or the block limit was reached); c, z, nb, blk updated.
template <bool WRITE, class Loader, class Sink>
RJB_HD void DecodeSpan(const Loader& load, const HuffLutSet* lut, const uint8_t* mcu_dc, const uint8_t* mcu_ac, int bpm,
                       uint32_t& p, uint32_t end_bit, int& c, int& z, uint32_t& nb, uint32_t& blk, uint32_t blk_limit,
                       Sink& sink) {
    if (p >= end_bit) return;
    uint32_t wi = p >> 5;
    uint64_t bb = uint64_t(ByteSwap32(load(wi))) << (32 + (p & 31u));
    int bc = 32 - int(p & 31u);
    wi++;
    while (p < end_bit) {
        if (WRITE && blk >= blk_limit) break;
        if (bc <= 32) {
            bb |= uint64_t(ByteSwap32(load(wi))) << (32 - bc);
            bc += 32;
            wi++;
        }
        const int tab = (z == 0) ? mcu_dc[c] : mcu_ac[c];
        const uint32_t v16 = uint32_t(bb >> 48);
        uint32_t e = lut->fast[tab][v16 >> (16 - kFastBits)];
        uint32_t len = e >> 8, sym = e & 0xFFu;
        if (len == 0) {   // code longer than the first-level table (or invalid)
            len = 16;
            sym = 0;
            for (int l = kFastBits + 1; l <= 16; l++) {
                if (v16 < lut->upper[tab][l]) {
                    len = uint32_t(l);
                    sym = lut->vals[tab][(int32_t(v16 >> (16 - l)) + lut->valoff[tab][l]) & 255];
                    break;
                }
            }
        }
        bb <<= len;
        const uint32_t s = sym & 15u;
        int val = 0;
        if (s) {   // RECEIVE + EXTEND (T.81 F.2.2.1)
            const uint32_t extra = uint32_t(bb >> (64 - s));
            bb <<= s;
            val = (extra < (1u << (s - 1))) ? int(extra) - int((1u << s) - 1u) : int(extra);
        }
        bc -= int(len + s);
        p += len + s;
        if (z == 0) {
            if (WRITE) sink.Dc(blk, val);
            z = 1;
        } else {
            const uint32_t r = sym >> 4;
            if (s == 0) {
                z = (r == 15) ? z + 16 : 64;   // ZRL / EOB
            } else {
                z += int(r);
                if (WRITE && z < 64) sink.Ac(blk, z, val);
                z++;
            }
        }
        if (z >= 64) {
            z = 0;
            nb++;
            blk++;
            c = (c + 1 == bpm) ? 0 : c + 1;
        }
    }
}

}  // namespace rjb
