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

// ---- lean shared-memory primitives for the two hot loops ---------------------------------
// Both K1 loops are instruction-issue bound (profiles/r01b_*), so they are written against raw
// 32-bit shared-memory addresses: no generic-address arithmetic, no window registers to rotate
// (two extra LDS per symbol cost latency, which the other warps hide, but no issue slots for
// bookkeeping), table select and bit position as the only loop-carried scalars.
__device__ __forceinline__ uint32_t SharedAddr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t Lds32(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t Lds16(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ uint32_t Lds8(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void Sts16(uint32_t a, uint32_t v) { asm volatile("st.shared.u16 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }

// Next 32 bits of the stream at bit position p of the slot at shared address `slot`.
__device__ __forceinline__ uint32_t PeekBits(uint32_t slot, uint32_t p) {
    const uint32_t a = slot + ((p >> 5) << 2);
    return __funnelshift_l(Lds32(a + 4), Lds32(a), p);
}
// Byte offsets of the Huffman tables of block c inside HuffLutSet::fast.
__device__ __forceinline__ uint32_t DcBytes(TableSel t, int c) { return ((t.dc_mask >> c) & 1u) << (kFastBits + 1); }
__device__ __forceinline__ uint32_t AcBytes(TableSel t, int c) { return (2u + ((t.ac_mask >> c) & 1u)) << (kFastBits + 1); }

// Count-only decode of one subsequence from state `key` (speculation / synchronisation):
// returns the packed end state and block count.
__device__ __forceinline__ uint32_t DecodeCount(uint32_t slot, uint32_t lut_sa, const HuffLutSet* lut, TableSel sel, int bpm,
                                                uint32_t key, uint32_t end_bit) {
    uint32_t p = StateOverflow(key), nb = 0;
    int c = StateC(key), z = StateZ(key);
    uint32_t dc_off = DcBytes(sel, c), ac_off = AcBytes(sel, c);
    uint32_t off = (z == 0) ? dc_off : ac_off;
    while (p < end_bit) {
        const uint32_t win = PeekBits(slot, p);
        uint32_t e = Lds16(lut_sa + off + ((win >> (31 - kFastBits)) & (2u * kFastSize - 2u)));
        if (e == 0) e = SlowEntry(lut, off >> (kFastBits + 1), win >> 16);
        p += EntryBits(e);
        z += EntryAdvance(e);
        off = ac_off;
        if (z >= 64) {
            z = 0;
            nb++;
            c = (c + 1 == bpm) ? 0 : c + 1;
            dc_off = DcBytes(sel, c);
            ac_off = AcBytes(sel, c);
            off = dc_off;
        }
    }
    const uint32_t over = p > end_bit ? p - end_bit : 0;
    return PackState(over, c, z, nb > 0xFFFFu ? 0xFFFFu : nb);
}

constexpr int kBlkBufBytes = 144;    // one staged block: 128 B + 16 (keeps 16-byte alignment)
constexpr int kLaneBufBytes = 2 * kBlkBufBytes + 8;   // two blocks per lane; 74-word lane stride spreads the banks
constexpr int kFlushEvery = 8;       // symbols decoded per lane between two cooperative flushes

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
    uint32_t used[T];
    uint32_t endbit[T];
    uint32_t queue[T];
    uint8_t mcu_dc[16], mcu_ac[16];
    uint8_t zigzag[64];
    uint32_t scratch[40];
};

template <int S>
struct K1WriteSmem {
    K1Smem<S> k;
    __align__(16) unsigned char blkbuf[T * kLaneBufBytes];
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
    const TableSel sel = MakeTableSel(sm.mcu_dc, sm.mcu_ac, im.bpm);
    const int bpm = im.bpm;
    const uint32_t slots_sa = SharedAddr(sm.words);
    const uint32_t lut_sa = SharedAddr(&sm.lut.fast[0][0]);
    auto decode = [&](uint32_t sub, uint32_t key) {
        return DecodeCount(slots_sa + sub * uint32_t(K1Smem<S>::kSlotStride * 4), lut_sa, &sm.lut, sel, bpm, key, sm.endbit[sub]);
    };

    uint32_t ndecodes = 0;
    sm.endbit[tid] = me.end_bit;
    if (tid == 0) sm.scratch[0] = 0;   // work-queue length
    __syncthreads();
    if (round == 0 && me.active) {
        // guess for a mid-segment start: aligned on a symbol, first block of an MCU, DC next
        my_used = 0;
        out = decode(uint32_t(tid), 0);
        ndecodes++;
    }
    sm.state[tid] = out;
    sm.used[tid] = my_used;
    // state entering the CTA (snapshot: the neighbouring CTA may still be changing it; a change is
    // caught by the boundary counter and the next round)
    uint32_t in0 = my_used;
    if (round > 0 && tid == 0 && me.active && !me.first) in0 = StateKey(a.state[g - 1]);
    const bool can_redo = me.active && !me.first;
    const int lane = tid & 31;
    // CTA-local fix-up: re-decode while the predecessor's end state is not the state used. The
    // subsequences that need it are compacted into a queue so that the re-decodes occupy as few
    // warps as possible (a warp with one busy lane costs as many issue slots as a full one).
    for (int iter = 0; iter < T + 1; iter++) {
        __syncthreads();
        const uint32_t in = (tid > 0) ? StateKey(sm.state[tid - 1]) : in0;
        const bool need = can_redo && in != sm.used[tid];
        const uint32_t mask = __ballot_sync(0xFFFFFFFFu, need);
        uint32_t base = 0;
        if (lane == 0 && mask) base = atomicAdd(&sm.scratch[0], uint32_t(__popc(mask)));
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        if (need) sm.queue[base + __popc(mask & ((1u << lane) - 1u))] = (in << 16) | uint32_t(tid);
        __syncthreads();
        const uint32_t n = sm.scratch[0];
        if (n == 0) break;
        uint32_t item = 0, res = 0;
        if (uint32_t(tid) < n) {
            item = sm.queue[tid];
            res = decode(item & 0xFFFFu, item >> 16);
            ndecodes++;
        }
        __syncthreads();
        if (uint32_t(tid) < n) {
            sm.state[item & 0xFFFFu] = res;
            sm.used[item & 0xFFFFu] = item >> 16;
        }
        if (tid == 0) sm.scratch[0] = 0;
    }
    __syncthreads();
    out = sm.state[tid];
    my_used = sm.used[tid];
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
    {   // zero the block buffers
        uint32_t* zb = reinterpret_cast<uint32_t*>(wsm.blkbuf);
        for (int i = tid; i < T * kLaneBufBytes / 4; i += T) zb[i] = 0;
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

    // ---- final decode. Every block is assembled in shared memory and leaves as one whole
    // 128-byte line (the coefficient arena needs no clearing and sees no partial-sector
    // read-modify-write). A block belongs to the thread that decodes its DC symbol: that thread
    // keeps decoding past the end of its subsequence until the block is complete; a thread that
    // starts inside a block stays silent (`live` false) until the first block boundary.
    //
    // The warp advances in lock-step, one symbol per lane per step, so all 32 lanes execute the
    // same instruction stream. Each lane owns two block buffers: a finished block is parked
    // (`pending`) while the lane continues in the other buffer; every kFlushEvery steps the
    // warp stores all parked blocks cooperatively — 32 lanes x 4 bytes = one coalesced line per
    // block. A lane that finishes a second block before the rendezvous simply waits for it.
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
    unsigned char* mybufs = wsm.blkbuf + tid * kLaneBufBytes;
    const TableSel sel = MakeTableSel(sm.mcu_dc, sm.mcu_ac, im.bpm);
    const int bpm = im.bpm;
    const uint32_t slot_sa = SharedAddr(sm.words + tid * K1Smem<S>::kSlotStride);
    const uint32_t lut_sa = SharedAddr(&sm.lut.fast[0][0]);
    const uint32_t zz_sa = SharedAddr(sm.zigzag);
    const uint32_t bufs_sa = SharedAddr(mybufs);
    const uint32_t* gwords = reinterpret_cast<const uint32_t*>(a.scan + (me.active ? me.start : 0));
    bool live = (z == 0);
    bool finishing = false;                 // past the subsequence, completing an owned block
    uint32_t stop = me.end_bit;
    bool done = !me.active || p >= stop || blk >= limit;
    uint32_t cur_sa = bufs_sa;              // buffer being filled
    bool pending = false;                   // the other buffer holds a finished block
    uint32_t pend_blk = 0;
    uint32_t pend_sa = 0;
    uint32_t dc_off = DcBytes(sel, c), ac_off = AcBytes(sel, c);
    uint32_t off = (z == 0) ? dc_off : ac_off;
    for (;;) {
#pragma unroll 1
        for (int step = 0; step < kFlushEvery; step++) {
            if (!done && z < 64) {
                const uint32_t wi = p >> 5;
                uint32_t w0, w1;
                if (wi + 1 < uint32_t(K1Smem<S>::kSlotWords)) {
                    w0 = Lds32(slot_sa + (wi << 2));
                    w1 = Lds32(slot_sa + (wi << 2) + 4);
                } else {   // finishing an owned block beyond the staged slot: straight from the arena
                    w0 = ByteSwap32(__ldg(gwords + wi));
                    w1 = ByteSwap32(__ldg(gwords + wi + 1));
                }
                const uint32_t win = __funnelshift_l(w1, w0, p);
                uint32_t e = Lds16(lut_sa + off + ((win >> (31 - kFastBits)) & (2u * kFastSize - 2u)));
                if (e == 0) e = SlowEntry(&sm.lut, off >> (kFastBits + 1), win >> 16);
                const uint32_t sz = EntrySize(e), bits = EntryBits(e);
                // RECEIVE + EXTEND (T.81 F.2.2.1): magnitude bits moved to the top of t
                const uint32_t t = win << (bits - sz);
                const uint32_t extra = __funnelshift_rc(t, 0u, 32u - sz);
                const int val = int(extra) - ((int(t) < 0) ? 0 : int((1u << sz) - 1u));
                const int z_old = z;
                z += EntryAdvance(e);
                p += bits;
                if (live) {
                    if (z_old == 0) dcdiff[blk] = int16_t(val);
                    else if (sz != 0 && z <= 64) Sts16(cur_sa + 2u * Lds8(zz_sa + uint32_t(z) - 1u), uint32_t(val));
                }
                off = ac_off;
                if (z < 64 && p >= stop) {
                    if (!finishing && live && p < seg_end_bit) {
                        finishing = true;   // own the unfinished block: follow it into the next subsequence(s)
                        stop = seg_end_bit;
                    } else {
                        done = true;
                    }
                }
            }
            if (!done && z >= 64 && !(live && pending)) {   // block complete and a buffer is free
                if (live) {
                    pending = true;
                    pend_blk = blk;
                    pend_sa = cur_sa;
                    cur_sa = bufs_sa + (cur_sa == bufs_sa ? uint32_t(kBlkBufBytes) : 0u);
                }
                live = true;
                z = 0;
                blk++;
                c = (c + 1 == bpm) ? 0 : c + 1;
                dc_off = DcBytes(sel, c);
                ac_off = AcBytes(sel, c);
                off = dc_off;
                if (finishing || p >= stop || blk >= limit) done = true;
            }
        }
        // rendezvous: store every parked block, one coalesced 128-byte line each
        __syncwarp();
        uint32_t m = __ballot_sync(0xFFFFFFFFu, pending);
        while (m) {
            const int L = __ffs(m) - 1;
            m &= m - 1;
            const uint32_t b = __shfl_sync(0xFFFFFFFFu, pend_blk, L);
            const uint32_t src = __shfl_sync(0xFFFFFFFFu, pend_sa, L) + 4u * uint32_t(lane);
            const uint32_t v = Lds32(src);
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(src), "r"(0u) : "memory");
            reinterpret_cast<uint32_t*>(coef + size_t(b) * 64)[lane] = v;
        }
        pending = false;
        __syncwarp();
        if (!__any_sync(0xFFFFFFFFu, !done)) break;
    }
    // Damaged / truncated data only: when the interval's data ends under this thread's hands,
    // what the interval still owes is written as zero blocks (the arena is never cleared, so
    // every block must be stored by someone).
    if (me.active && live && (me.last || p >= seg_end_bit) && blk < limit) {
        const uint32_t* src = reinterpret_cast<const uint32_t*>(mybufs + (cur_sa - bufs_sa));
        if (z != 0) {
            uint32_t* dst = reinterpret_cast<uint32_t*>(coef + size_t(blk) * 64);
            for (int i = 0; i < 32; i++) dst[i] = src[i];
            blk++;
        }
        for (; blk < limit; blk++) {
            uint4* dst = reinterpret_cast<uint4*>(coef + size_t(blk) * 64);
            for (int i = 0; i < 8; i++) dst[i] = make_uint4(0, 0, 0, 0);
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
