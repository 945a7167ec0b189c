// k1_huffman.cu — entropy-decode stage (K1) for sm_100a.
//
// Replaces the Huffman decoding the reference delegates to AMD's VCN
// fixed-function JPEG engine (vaRenderPicture/vaEndPicture at
// src/rocjpeg_vaapi_decoder.cpp:677-689 and :816-828). One launch covers every
// image of a batch.
//
// Work decomposition
//   segment      = one restart interval (or the whole scan when DRI is absent):
//                  byte-aligned start, known decoder state, known first block.
//   subsequence  = S consecutive bytes of a segment (S = 32/64/128), one thread.
//   CTA          = kK1Threads consecutive subsequences of ONE image; the image's
//                  Huffman tables and the CTA's bytes are staged in shared memory.
//
// Schedule (self-synchronising parallel Huffman decoding)
//   k1_sync round 0   every thread decodes its subsequence from a guessed state
//                     (exact for the first subsequence of a segment) and records
//                     its end state; then, inside the CTA, every thread whose
//                     predecessor's end state differs from the state it started
//                     from re-decodes, until nothing changes. Huffman streams
//                     re-synchronise after a few symbols, so this is typically
//                     two decodes per thread.
//   k1_sync round r>0 repairs what crossed CTA boundaries (thread 0 of a CTA had
//                     no predecessor state in round 0). A CTA whose incoming state
//                     is unchanged exits at once. The host checks the counter of
//                     the last round; a non-zero value triggers more rounds
//                     (correctness never depends on the stream synchronising).
//   k1_write          block positions = segmented prefix sums of the per-thread
//                     block counts (CTA scan + look-back over CTA partials), then
//                     the final decode assembles each block in shared memory and
//                     stores it as one whole 128-byte line of int16 coefficients
//                     (natural order), plus one DC difference per block.
//   dc_sums/dc_apply  per-component, per-restart-interval prefix sum of the DC
//                     differences; the absolute DC replaces the difference in the
//                     compact per-block DC array that K2 reads.
#include <cuda_runtime.h>

#include "huff_core.cuh"
#include "stages.h"

namespace rjb {

__constant__ uint8_t c_zigzag[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                                     12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                                     35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                                     58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

namespace {

constexpr int T = kK1Threads;
constexpr uint32_t kNoState = 0xFFFFFFFFu;

// largest i in [0, n) with a[i] <= v (a is non-decreasing, a[0] <= v)
__device__ __forceinline__ uint32_t UpperIndex(const uint32_t* a, uint32_t n, uint32_t v) {
    uint32_t lo = 0, hi = n;
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(a + mid) <= v) lo = mid; else hi = mid;
    }
    return lo;
}

struct SmemLoader {
    const uint32_t* base;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const { return base[i]; }
};
// Write pass: words of the thread's own slot come from shared memory; past it (finishing an
// owned block beyond the subsequence) they come straight from the scan arena.
struct SlotOrGlobalLoader {
    const uint32_t* slot;
    const uint32_t* gbase;
    uint32_t slot_words;
    __device__ __forceinline__ uint32_t operator()(uint32_t i) const {
        return i < slot_words ? slot[i] : ByteSwap32(__ldg(gbase + i));
    }
};

constexpr int kBlkBufBytes = 144;   // 128 B block + 16: 16-byte aligned rows, quarter-warps conflict-free

// Per-thread description of its subsequence.
struct Sub {
    bool active;       // maps to real data
    bool first;        // first subsequence of its segment (state known exactly)
    bool last;         // last subsequence of its segment
    uint32_t seg;      // global segment index
    uint32_t end_bit;  // bits of entropy-coded data inside the subsequence
    uint64_t start;    // byte offset of the subsequence in the scan arena
};

template <int S>
__device__ __forceinline__ Sub Locate(const K1Args& a, const ImageDesc& im, uint32_t g, bool use_cache) {
    Sub s;
    s.active = g < im.sub0 + im.nsub;
    s.first = s.last = false;
    s.seg = 0;
    s.end_bit = 0;
    s.start = 0;
    if (!s.active) return s;
    uint32_t k;
    if (use_cache) {
        k = a.sub_seg[g];
    } else {
        // segments of this image: [seg0, seg0 + nseg), ordered by sub0
        uint32_t lo = 0, hi = im.nseg;
        while (hi - lo > 1) {
            uint32_t mid = (lo + hi) >> 1;
            if (__ldg(&a.segments[im.seg0 + mid].sub0) <= g) lo = mid; else hi = mid;
        }
        k = im.seg0 + lo;
        a.sub_seg[g] = k;
    }
    const SegmentDesc sd = a.segments[k];
    const uint32_t j = g - sd.sub0;
    const uint32_t nchunks = (sd.nbytes + S - 1) / S;
    s.seg = k;
    s.first = (j == 0);
    s.last = (j + 1 >= nchunks);
    const uint32_t off = j * S;
    const uint32_t remain = sd.nbytes > off ? sd.nbytes - off : 0;
    s.end_bit = (remain < uint32_t(S) ? remain : uint32_t(S)) * 8u;
    s.start = sd.data_off + off;
    return s;
}

// Shared-memory image of one CTA's working set.
template <int S>
struct K1Smem {
    static constexpr int kSlotWords = (S + 16) / 4;      // subsequence + 16 bytes of look-ahead
    static constexpr int kSlotStride = kSlotWords + 1;   // odd stride: conflict-free when lanes read the same word index
    static constexpr int kSlotVecs = (S + 16) / 16;
    HuffLutSet lut;
    uint32_t words[T * kSlotStride];
    uint64_t start[T];
    uint32_t state[T];
    uint8_t mcu_dc[16], mcu_ac[16];
    uint8_t zigzag[64];
    uint32_t scratch[40];
};

template <int S>
struct K1WriteSmem {
    K1Smem<S> k;
    __align__(16) unsigned char blkbuf[T * kBlkBufBytes];
};

template <int S>
__device__ __forceinline__ void StageCta(K1Smem<S>& sm, const K1Args& a, const ImageDesc& im, const Sub& me) {
    const int tid = threadIdx.x;
    sm.start[tid] = me.active ? me.start : ~0ull;
    // Huffman tables of this image
    {
        const uint4* src = reinterpret_cast<const uint4*>(a.luts + im.lut_set);
        uint4* dst = reinterpret_cast<uint4*>(&sm.lut);
        for (int i = tid; i < int(sizeof(HuffLutSet) / 16); i += T) dst[i] = __ldg(src + i);
    }
    if (tid < 16) {
        sm.mcu_dc[tid] = tid < kMaxBlocksPerMcu ? im.mcu_dc[tid] : 0;
        sm.mcu_ac[tid] = tid < kMaxBlocksPerMcu ? im.mcu_ac[tid] : 2;
    }
    if (tid < 64) sm.zigzag[tid] = c_zigzag[tid];
    __syncthreads();
    constexpr int V = K1Smem<S>::kSlotVecs;
    for (int idx = tid; idx < T * V; idx += T) {
        const int slot = idx / V, v = idx - slot * V;
        const uint64_t st = sm.start[slot];
        if (st == ~0ull) continue;
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(a.scan + st) + v);
        uint32_t* w = sm.words + slot * K1Smem<S>::kSlotStride + v * 4;
        // stored big-endian: bit 31 of a word is the first bit of the stream (huff_core.cuh)
        w[0] = ByteSwap32(q.x); w[1] = ByteSwap32(q.y); w[2] = ByteSwap32(q.z); w[3] = ByteSwap32(q.w);
    }
    __syncthreads();
}

// ---------------------------------------------------------------- k1_sync

template <int S>
__global__ void __launch_bounds__(T) k1_sync(K1Args a, int round) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    K1Smem<S>& sm = *reinterpret_cast<K1Smem<S>*>(smem_raw);
    const int tid = threadIdx.x;
    const uint32_t cta = blockIdx.x;
    const uint32_t img = UpperIndex(a.img_cta0, uint32_t(a.nimages), cta);
    const ImageDesc& im = a.images[img];
    const uint32_t g = cta * T + tid;
    const Sub me = Locate<S>(a, im, g, round > 0);

    uint32_t my_used = 0, out = 0, old_out = kNoState;
    if (round > 0) {
        // Does anything entering this CTA differ from what it was decoded with?
        int need0 = 0;
        if (tid == 0 && me.active && !me.first) need0 = (StateKey(a.state[g - 1]) != a.used[g]);
        if (!__syncthreads_or(need0)) return;
        if (me.active) {
            my_used = a.used[g];
            out = a.state[g];
            old_out = out;
        }
    }
    StageCta<S>(sm, a, im, me);
    const SmemLoader loader{sm.words + tid * K1Smem<S>::kSlotStride};
    NullSink sink;

    const TableSel sel = MakeTableSel(sm.mcu_dc, sm.mcu_ac, im.bpm);
    const int bpm = im.bpm;
    auto decode_from = [&](uint32_t key) {
        uint32_t p = StateOverflow(key), nb = 0, blk = 0;
        int c = StateC(key), z = StateZ(key);
        DecodeSpan<false>(loader, &sm.lut, sel, bpm, p, me.end_bit, c, z, nb, blk, 0xFFFFFFFFu, sink);
        const uint32_t over = p > me.end_bit ? p - me.end_bit : 0;
        return PackState(over, c, z, nb > 0xFFFFu ? 0xFFFFu : nb);
    };

    uint32_t ndecodes = 0;
    if (round == 0 && me.active) {
        // guess for a mid-segment start: aligned on a symbol, first block of an MCU, DC next
        my_used = 0;
        out = decode_from(0);
        ndecodes++;
    }
    sm.state[tid] = out;
    // CTA-local fix-up: re-decode while the predecessor's end state is not the state used.
    for (int iter = 0; iter < T + 1; iter++) {
        __syncthreads();
        uint32_t in = my_used;
        if (me.active && !me.first) {
            if (tid > 0) in = StateKey(sm.state[tid - 1]);
            else if (round > 0) in = StateKey(a.state[g - 1]);
        }
        const int need = me.active && !me.first && in != my_used;
        if (!__syncthreads_or(need)) break;
        if (need) {
            out = decode_from(in);
            my_used = in;
            sm.state[tid] = out;
            ndecodes++;
        }
    }
    if (me.active) {
        a.state[g] = out;
        a.used[g] = my_used;
    }
    // A change of the state handed to the next CTA means that CTA must look again.
    const bool hands_over = me.active && !me.last && (tid == T - 1);
    if (hands_over && (old_out == kNoState || StateKey(old_out) != StateKey(out))) atomicAdd(&a.counters[round], 1u);
    if (ndecodes) atomicAdd(&a.counters[kMaxSyncRounds + round], ndecodes);

    // CTA partial for the block-position scan: (contains a segment start, blocks after the last start)
    uint32_t* red = sm.scratch;
    if (tid == 0) { red[0] = 0; red[1] = 0; }
    __syncthreads();
    if (me.active && me.first) atomicMax(&red[0], uint32_t(tid) + 1u);
    __syncthreads();
    const uint32_t last_first = red[0];   // 0 = none, else tid + 1
    if (me.active && (last_first == 0 || uint32_t(tid) + 1u >= last_first)) atomicAdd(&red[1], StateBlocks(out));
    __syncthreads();
    if (tid == 0) a.cta_partial[cta] = make_uint2(last_first != 0 ? 1u : 0u, red[1]);
}

// ---------------------------------------------------------------- k1_write

template <int S>
__global__ void __launch_bounds__(T) k1_write(K1Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    K1WriteSmem<S>& wsm = *reinterpret_cast<K1WriteSmem<S>*>(smem_raw);
    K1Smem<S>& sm = wsm.k;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t cta = blockIdx.x;
    const uint32_t img = UpperIndex(a.img_cta0, uint32_t(a.nimages), cta);
    const ImageDesc& im = a.images[img];
    const uint32_t g = cta * T + tid;
    const Sub me = Locate<S>(a, im, g, true);
    {   // zero this thread's block buffer (only ever touched by its owner)
        uint4* zb = reinterpret_cast<uint4*>(wsm.blkbuf + tid * kBlkBufBytes);
#pragma unroll
        for (int i = 0; i < kBlkBufBytes / 16; i++) zb[i] = make_uint4(0, 0, 0, 0);
    }
    StageCta<S>(sm, a, im, me);

    const uint32_t st = me.active ? a.state[g] : 0;
    const uint32_t nb = StateBlocks(st);
    // inclusive segmented scan of nb over the CTA (a segment start resets the sum)
    uint32_t v = nb;
    uint32_t f = (me.active && me.first) ? 1u : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t pv = __shfl_up_sync(0xFFFFFFFFu, v, d);
        const uint32_t pf = __shfl_up_sync(0xFFFFFFFFu, f, d);
        if (lane >= d) {
            if (!f) v += pv;
            f |= pf;
        }
    }
    uint32_t* wsum = sm.scratch;        // [4] warp totals (value of the open segment at warp end)
    uint32_t* wflag = sm.scratch + 4;   // [4] warp contains a start
    uint32_t* carry_s = sm.scratch + 8; // look-back result
    if (lane == 31) { wsum[warp] = v; wflag[warp] = f; }
    // look-back over previous CTAs of the same image (warp 0), in parallel with the scan above
    if (warp == 0) {
        uint32_t carry = 0;
        const uint32_t first_cta = a.img_cta0[img];
        int64_t k = int64_t(cta) - 1;
        bool done = false;
        while (!done && k >= int64_t(first_cta)) {
            const int64_t idx = k - lane;
            uint2 part = make_uint2(0u, 0u);
            if (idx >= int64_t(first_cta)) part = a.cta_partial[idx];
            const uint32_t flagged = __ballot_sync(0xFFFFFFFFu, part.x != 0);
            const int stop = flagged ? __ffs(flagged) - 1 : 31;   // nearest CTA (smallest lane) holding a start
            uint32_t contrib = (lane <= stop) ? part.y : 0u;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xFFFFFFFFu, contrib, d);
            carry += contrib;
            done = flagged != 0;
            k -= 32;
        }
        if (lane == 0) carry_s[0] = carry;
    }
    __syncthreads();
    // add the totals of earlier warps while the segment is still open
    uint32_t add = 0;
    bool open = (f == 0);   // no segment start at or before this thread inside its warp
    for (int w = warp - 1; w >= 0 && open; w--) {
        add += wsum[w];
        if (wflag[w]) open = false;
    }
    if (open) add += carry_s[0];
    const uint32_t incl = v + add;
    const uint32_t excl = (me.active && me.first) ? 0u : incl - nb;

    // ---- final decode. Every block is assembled in this thread's shared-memory buffer and
    // leaves as eight 128-bit stores (whole 128-byte lines: the coefficient arena needs no
    // clearing and sees no partial-sector read-modify-write). A block belongs to the thread that
    // decodes its DC symbol: that thread keeps decoding past the end of its subsequence until the
    // block is complete; a thread that starts inside a block stays silent (`live` false) until
    // the first block boundary. The warp runs "decode to the next block end" / "flush" in
    // lock-step so the 24-instruction flush is not replayed for every divergent lane.
    SegmentDesc sd = {};
    if (me.active) sd = a.segments[me.seg];
    uint32_t key = 0;
    if (me.active && !me.first) key = StateKey(a.state[g - 1]);
    uint32_t p = StateOverflow(key);
    int c = StateC(key), z = StateZ(key);
    uint32_t blk = sd.blk_first + excl;
    const uint32_t limit = sd.blk_first + sd.blk_count;
    const uint32_t seg_end_bit = me.active ? (sd.nbytes - uint32_t(me.start - sd.data_off)) * 8u : 0u;
    int16_t* coef = a.coef + size_t(im.blk0) * 64;
    int16_t* dcdiff = a.dcdiff + im.blk0;
    int16_t* buf = reinterpret_cast<int16_t*>(wsm.blkbuf + tid * kBlkBufBytes);
    const TableSel sel = MakeTableSel(sm.mcu_dc, sm.mcu_ac, im.bpm);
    const int bpm = im.bpm;
    const SlotOrGlobalLoader loader{sm.words + tid * K1Smem<S>::kSlotStride,
                                    reinterpret_cast<const uint32_t*>(a.scan + (me.active ? me.start : 0)),
                                    uint32_t(K1Smem<S>::kSlotWords)};
    auto flush = [&](uint32_t b) {
        uint4* src = reinterpret_cast<uint4*>(buf);
        uint4* dst = reinterpret_cast<uint4*>(coef + size_t(b) * 64);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            dst[i] = src[i];
            src[i] = make_uint4(0, 0, 0, 0);
        }
    };
    bool live = (z == 0);
    bool finishing = false;                 // past the subsequence, completing an owned block
    uint32_t stop = me.end_bit;
    bool done = !me.active || p >= stop || blk >= limit;
    BitWindow bw;
    bw.Init(loader, done ? 0u : p);
    uint32_t dc_off = DcOffset(sel, c), ac_off = AcOffset(sel, c);
    for (;;) {
        bool ended = false;
        while (!done && !ended) {
            const uint32_t win = bw.Peek(p);
            const uint32_t e = LookupSymbol(&sm.lut, (z == 0) ? dc_off : ac_off, win);
            const int adv = EntryAdvance(e);
            if (live) {
                const int val = SymbolValue(e, win);
                if (z == 0) dcdiff[blk] = int16_t(val);
                else if (EntrySize(e) && z + adv <= 64) buf[sm.zigzag[z + adv - 1]] = int16_t(val);
            }
            z += adv;
            p += EntryBits(e);
            bw.Advance(loader, p);
            if (z >= 64) {
                ended = true;
            } else if (p >= stop) {
                if (!finishing && live && p < seg_end_bit) {
                    finishing = true;       // own the unfinished block: follow it into the next subsequence(s)
                    stop = seg_end_bit;
                } else {
                    done = true;
                }
            }
        }
        if (!__any_sync(0xFFFFFFFFu, ended)) break;
        if (ended) {
            if (live) flush(blk);
            live = true;
            z = 0;
            blk++;
            c = (c + 1 == bpm) ? 0 : c + 1;
            dc_off = DcOffset(sel, c);
            ac_off = AcOffset(sel, c);
            if (finishing || p >= stop || blk >= limit) done = true;
        }
    }
    // Damaged / truncated data only: when the interval's data ends under this thread's hands,
    // what the interval still owes is written as zero blocks (the arena is never cleared, so
    // every block must be stored by someone).
    if (me.active && live && (me.last || p >= seg_end_bit) && blk < limit) {
        if (z != 0) {
            flush(blk);
            blk++;
        }
        for (; blk < limit; blk++) {
            flush(blk);
            dcdiff[blk] = 0;
        }
    }
}

// ---------------------------------------------------------------- DC prediction

// Does MCU range [m0, m1) of an image contain a predictor reset? Returns the last one, or -1.
__device__ __forceinline__ int64_t LastReset(int64_t m0, int64_t m1, int ri) {
    if (m1 <= m0) return -1;
    if (ri <= 0) return m0 == 0 ? 0 : -1;
    const int64_t last = ((m1 - 1) / ri) * ri;
    return last >= m0 ? last : -1;
}

__global__ void __launch_bounds__(kDcTileMcus) dc_sums(K1Args a) {
    __shared__ int red[3];
    const int tid = threadIdx.x;
    const uint32_t tile = blockIdx.x;
    const uint32_t img = UpperIndex(a.img_dctile0, uint32_t(a.nimages), tile);
    const ImageDesc& im = a.images[img];
    const int64_t m0 = int64_t(tile - a.img_dctile0[img]) * kDcTileMcus;
    const int64_t m1 = min(m0 + kDcTileMcus, int64_t(im.total_mcus));
    const int64_t reset = LastReset(m0, m1, im.restart_interval);
    const int64_t m = m0 + tid;
    if (tid < 3) red[tid] = 0;
    __syncthreads();
    int s[3] = {0, 0, 0};
    if (m < m1 && m >= reset) {
        const int16_t* d = a.dcdiff + im.blk0 + m * im.bpm;
        for (int k = 0; k < im.bpm; k++) s[im.mcu_comp[k]] += d[k];
    }
#pragma unroll
    for (int c = 0; c < 3; c++) {
        int v = s[c];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
        if ((tid & 31) == 0 && v) atomicAdd(&red[c], v);
    }
    __syncthreads();
    if (tid == 0) a.dc_partial[tile] = make_int3(red[0], red[1], red[2]);
}

__global__ void __launch_bounds__(kDcTileMcus) dc_apply(K1Args a) {
    __shared__ int wsum[8][3];
    __shared__ int wflag[8];
    __shared__ int carry_s[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t tile = blockIdx.x;
    const uint32_t img = UpperIndex(a.img_dctile0, uint32_t(a.nimages), tile);
    const ImageDesc& im = a.images[img];
    const uint32_t tile_first = a.img_dctile0[img];
    const int64_t m0 = int64_t(tile - tile_first) * kDcTileMcus;
    const int64_t m1 = min(m0 + kDcTileMcus, int64_t(im.total_mcus));
    const int ri = im.restart_interval;
    const int64_t m = m0 + tid;
    const bool active = m < m1;
    const bool reset_here = active && (ri > 0 ? (m % ri) == 0 : m == 0);

    int diffs[kMaxBlocksPerMcu];
    int s[3] = {0, 0, 0};
    if (active) {
        const int16_t* d = a.dcdiff + im.blk0 + m * im.bpm;
        for (int k = 0; k < im.bpm; k++) {
            diffs[k] = d[k];
            s[im.mcu_comp[k]] += diffs[k];
        }
    }
    // inclusive segmented scan of the per-MCU sums
    int v0 = s[0], v1 = s[1], v2 = s[2];
    uint32_t f = reset_here ? 1u : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int p0 = __shfl_up_sync(0xFFFFFFFFu, v0, d);
        const int p1 = __shfl_up_sync(0xFFFFFFFFu, v1, d);
        const int p2 = __shfl_up_sync(0xFFFFFFFFu, v2, d);
        const uint32_t pf = __shfl_up_sync(0xFFFFFFFFu, f, d);
        if (lane >= d) {
            if (!f) { v0 += p0; v1 += p1; v2 += p2; }
            f |= pf;
        }
    }
    if (lane == 31) { wsum[warp][0] = v0; wsum[warp][1] = v1; wsum[warp][2] = v2; wflag[warp] = int(f); }
    if (warp == 0) {
        // look-back over earlier tiles of the image until one that contains a reset (tile 0 always does)
        int c0 = 0, c1 = 0, c2 = 0;
        int64_t k = int64_t(tile) - 1;
        bool done = (tile == tile_first);
        while (!done && k >= int64_t(tile_first)) {
            const int64_t idx = k - lane;
            int3 part = make_int3(0, 0, 0);
            bool has_reset = false;
            if (idx >= int64_t(tile_first)) {
                part = a.dc_partial[idx];
                const int64_t t0 = (idx - tile_first) * kDcTileMcus;
                has_reset = LastReset(t0, min(t0 + kDcTileMcus, int64_t(im.total_mcus)), ri) >= 0;
            }
            const uint32_t flagged = __ballot_sync(0xFFFFFFFFu, has_reset);
            const int stop = flagged ? __ffs(flagged) - 1 : 31;
            int q0 = lane <= stop ? part.x : 0, q1 = lane <= stop ? part.y : 0, q2 = lane <= stop ? part.z : 0;
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                q0 += __shfl_xor_sync(0xFFFFFFFFu, q0, d);
                q1 += __shfl_xor_sync(0xFFFFFFFFu, q1, d);
                q2 += __shfl_xor_sync(0xFFFFFFFFu, q2, d);
            }
            c0 += q0; c1 += q1; c2 += q2;
            done = flagged != 0;
            k -= 32;
        }
        if (lane == 0) { carry_s[0] = c0; carry_s[1] = c1; carry_s[2] = c2; }
    }
    __syncthreads();
    bool open = (f == 0);
    int a0 = 0, a1 = 0, a2 = 0;
    for (int w = warp - 1; w >= 0 && open; w--) {
        a0 += wsum[w][0]; a1 += wsum[w][1]; a2 += wsum[w][2];
        if (wflag[w]) open = false;
    }
    if (open) { a0 += carry_s[0]; a1 += carry_s[1]; a2 += carry_s[2]; }
    if (!active) return;
    // exclusive prefix = predictor values entering this MCU
    int pred[3] = {v0 + a0 - s[0], v1 + a1 - s[1], v2 + a2 - s[2]};
    if (reset_here) pred[0] = pred[1] = pred[2] = 0;
    // absolute DC replaces the difference in the compact per-block array (K2 reads it from
    // there: one coalesced 2-byte load per block instead of a scattered store per block here)
    int16_t* out = a.dcdiff + im.blk0 + m * im.bpm;
    for (int k = 0; k < im.bpm; k++) {
        const int comp = im.mcu_comp[k];
        pred[comp] += diffs[k];
        out[k] = int16_t(pred[comp]);
    }
}

// ---------------------------------------------------------------- gather

constexpr int kGatherChunk = 16384;
constexpr int kGatherCtas = 64;   // PCIe-bound: a few CTAs saturate the link; the rest of the GPU stays free
                                  // for the kernels of the other pipeline lanes

__global__ void __launch_bounds__(256) gather_scans(const GatherItem* items, int nitems, uint32_t total_chunks, uint8_t* arena) {
    __shared__ int s_item;
    for (uint32_t chunk = blockIdx.x; chunk < total_chunks; chunk += gridDim.x) {
        __syncthreads();
        if (threadIdx.x == 0) {
            int lo = 0, hi = nitems;
            while (hi - lo > 1) {
                int mid = (lo + hi) >> 1;
                if (items[mid].chunk0 <= chunk) lo = mid; else hi = mid;
            }
            s_item = lo;
        }
        __syncthreads();
        const GatherItem it = items[s_item];
        const uint32_t off = (chunk - it.chunk0) * kGatherChunk;
        const uint32_t n = min(uint32_t(kGatherChunk), it.nbytes - off);
        const uint4* src = reinterpret_cast<const uint4*>(it.src + off);
        uint4* dst = reinterpret_cast<uint4*>(arena + it.dst_off + off);
        for (uint32_t i = threadIdx.x; i < n / 16; i += blockDim.x) dst[i] = src[i];
    }
}

template <int S>
cudaError_t SyncImpl(const K1Args& a, int round, cudaStream_t stream) {
    if (round >= 0) {
        static_assert(sizeof(K1Smem<S>) <= 48 * 1024, "k1_sync shared memory exceeds the default limit");
        k1_sync<S><<<a.total_ctas, T, sizeof(K1Smem<S>), stream>>>(a, round);
    } else {
        // above the 48 KiB default for S = 128: opt in (per device; cheap, so done on every launch)
        cudaError_t e = cudaFuncSetAttribute(k1_write<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(sizeof(K1WriteSmem<S>)));
        if (e != cudaSuccess) return e;
        k1_write<S><<<a.total_ctas, T, sizeof(K1WriteSmem<S>), stream>>>(a);
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t LaunchK1Sync(const K1Args& a, int round, cudaStream_t stream) {
    if (a.total_ctas == 0) return cudaSuccess;
    switch (a.sub_bytes) {
        case 32: return SyncImpl<32>(a, round, stream);
        case 64: return SyncImpl<64>(a, round, stream);
        case 128: return SyncImpl<128>(a, round, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t LaunchK1Write(const K1Args& a, cudaStream_t stream) {
    if (a.total_ctas == 0) return cudaSuccess;
    switch (a.sub_bytes) {
        case 32: return SyncImpl<32>(a, -1, stream);
        case 64: return SyncImpl<64>(a, -1, stream);
        case 128: return SyncImpl<128>(a, -1, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t LaunchDcScan(const K1Args& a, cudaStream_t stream) {
    if (a.total_dc_tiles == 0) return cudaSuccess;
    dc_sums<<<a.total_dc_tiles, kDcTileMcus, 0, stream>>>(a);
    dc_apply<<<a.total_dc_tiles, kDcTileMcus, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t LaunchGather(const GatherItem* items, int nitems, uint32_t total_chunks, uint8_t* arena, cudaStream_t stream) {
    if (total_chunks == 0) return cudaSuccess;
    gather_scans<<<min(total_chunks, uint32_t(kGatherCtas)), 256, 0, stream>>>(items, nitems, total_chunks, arena);
    return cudaGetLastError();
}

}  // namespace rjb
