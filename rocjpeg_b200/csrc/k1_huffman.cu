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
//                     block counts, entry offsets = prefix sums of the per-thread
//                     entry counts (CTA scan + look-back over CTA partials); the
//                     final decode then writes the image's sparse coefficient stream
//                     (one 32-bit (position, int16 value) entry per non-zero
//                     coefficient), the index of every block's first entry, and one
//                     DC difference per block.
//   dc_sums/dc_apply  per-component, per-restart-interval prefix sum of the DC
//                     differences; the absolute DC replaces the difference in the
//                     compact per-block DC array that K2 reads.
#include <cuda_runtime.h>

#include <cstddef>

#include "huff_core.cuh"
#include "stages.h"

namespace rjb {

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

// Codes longer than the first-level table, without a search loop and without generic loads: the six
// length thresholds are fetched at once, the code length is a count of comparisons
// (huff_core.cuh: SlowEntry is the same computation written as a loop, for the host model).
// With 32 lanes per warp some lane takes this path in roughly every second step, so its latency
// is paid by the whole warp.
__device__ __forceinline__ uint32_t SlowEntryShared(uint32_t lutset_sa, uint32_t tab, uint32_t v16) {
    const uint32_t up = lutset_sa + uint32_t(offsetof(HuffLutSet, upper)) + tab * 68u + 4u * (kFastBits + 1);
    static_assert(kFastBits == 10, "six lengths (11..16) are searched");
    const uint32_t u11 = Lds32(up), u12 = Lds32(up + 4), u13 = Lds32(up + 8), u14 = Lds32(up + 12), u15 = Lds32(up + 16),
                   u16 = Lds32(up + 20);
    const uint32_t len = 11u + (v16 >= u11) + (v16 >= u12) + (v16 >= u13) + (v16 >= u14) + (v16 >= u15);
    if (v16 >= u16) return MakeEntry(16, 0, tab >= 2);   // invalid code: 16 bits consumed, symbol 0
    const int32_t voff = int32_t(Lds32(lutset_sa + uint32_t(offsetof(HuffLutSet, valoff)) + tab * 68u + 4u * len));
    const uint32_t sym = Lds8(lutset_sa + uint32_t(offsetof(HuffLutSet, vals)) + tab * 256u + (uint32_t(int32_t(v16 >> (16u - len)) + voff) & 255u));
    return MakeEntry(len, sym, tab >= 2);
}

// Next 32 bits of the stream at bit position p of the slot at shared address `slot`.
__device__ __forceinline__ uint32_t PeekBits(uint32_t slot, uint32_t p) {
    const uint32_t a = slot + ((p >> 5) << 2);
    return __funnelshift_l(Lds32(a + 4), Lds32(a), p);
}
// Byte offsets of the Huffman tables of block c inside HuffLutSet::fast.
__device__ __forceinline__ uint32_t DcBytes(TableSel t, int c) { return ((t.dc_mask >> c) & 1u) << (kFastBits + 1); }
__device__ __forceinline__ uint32_t AcBytes(TableSel t, int c) { return (2u + ((t.ac_mask >> c) & 1u)) << (kFastBits + 1); }

// Count-only decode of one subsequence from state `key` (speculation / synchronisation):
// returns the packed end state + block count, and the number of coefficient entries in *nnz.
__device__ __forceinline__ uint32_t DecodeCount(uint32_t slot, uint32_t lut_sa, TableSel sel, int bpm, uint32_t key,
                                                uint32_t end_bit, uint32_t* nnz_out) {
    uint32_t p = StateOverflow(key), nb = 0, nnz = 0;
    int c = StateC(key), z = StateZ(key);
    uint32_t dc_off = DcBytes(sel, c), ac_off = AcBytes(sel, c);
    uint32_t off = (z == 0) ? dc_off : ac_off;
    while (p < end_bit) {
        const uint32_t win = PeekBits(slot, p);
        uint32_t e = Lds16(lut_sa + off + ((win >> (31 - kFastBits)) & (2u * kFastSize - 2u)));
        if (e == 0) e = SlowEntryShared(lut_sa, off >> (kFastBits + 1), win >> 16);
        p += EntryBits(e);
        z += EntryAdvance(e);
        nnz += (e & (15u << 5)) ? 1u : 0u;
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
    *nnz_out = (nnz + 3u) & ~3u;   // a thread's run of the entry stream is padded to whole 16-byte stores
    const uint32_t over = p > end_bit ? p - end_bit : 0;
    return PackState(over, c, z, nb > 0xFFFFu ? 0xFFFFu : nb);
}

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
    uint32_t nnz[T];
    uint32_t endbit[T];
    uint32_t queue[T];
    uint8_t mcu_dc[16], mcu_ac[16];
    uint32_t scratch[40];
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
    auto decode = [&](uint32_t sub, uint32_t key) {   // result state returned, entry count left in sm.nnz[sub]
        uint32_t n;
        const uint32_t st = DecodeCount(slots_sa + sub * uint32_t(K1Smem<S>::kSlotStride * 4), lut_sa, sel, bpm, key, sm.endbit[sub], &n);
        sm.nnz[sub] = n;
        return st;
    };

    uint32_t ndecodes = 0;
    sm.endbit[tid] = me.end_bit;
    sm.nnz[tid] = (round > 0 && me.active) ? a.nnz[g] : 0u;
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
    const uint32_t my_nnz = sm.nnz[tid];
    if (me.active) {
        a.state[g] = out;
        a.used[g] = my_used;
        a.nnz[g] = my_nnz;
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
    // entries produced by the CTA (plain sum: entry offsets run through the whole image)
    if (tid == 0) red[2] = 0;
    __syncthreads();
    if (me.active && my_nnz) atomicAdd(&red[2], my_nnz);
    __syncthreads();
    if (tid == 0) a.cta_entries[cta] = red[2];
}

// ---------------------------------------------------------------- k1_write

template <int S>
__global__ void __launch_bounds__(T) k1_write(K1Args a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    K1Smem<S>& sm = *reinterpret_cast<K1Smem<S>*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t cta = blockIdx.x;
    const uint32_t img = UpperIndex(a.img_cta0, uint32_t(a.nimages), cta);
    const ImageDesc& im = a.images[img];
    const uint32_t g = cta * T + tid;
    const Sub me = Locate<S>(a, im, g, true);
    StageCta<S>(sm, a, im, me);

    const uint32_t st = me.active ? a.state[g] : 0;
    const uint32_t nb = StateBlocks(st);
    const uint32_t my_nnz = me.active ? a.nnz[g] : 0;
    // Two scans over the CTA: block positions (segmented: a segment start resets the count) and
    // entry offsets (plain, they run through the whole image).
    uint32_t v = nb, e = my_nnz;
    uint32_t f = (me.active && me.first) ? 1u : 0u;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t pv = __shfl_up_sync(0xFFFFFFFFu, v, d);
        const uint32_t pf = __shfl_up_sync(0xFFFFFFFFu, f, d);
        const uint32_t pe = __shfl_up_sync(0xFFFFFFFFu, e, d);
        if (lane >= d) {
            if (!f) v += pv;
            f |= pf;
            e += pe;
        }
    }
    uint32_t* wsum = sm.scratch;          // [4] warp totals (value of the open segment at warp end)
    uint32_t* wflag = sm.scratch + 4;     // [4] warp contains a start
    uint32_t* wents = sm.scratch + 8;     // [4] warp entry totals
    uint32_t* carry_s = sm.scratch + 12;  // [2] look-back results: blocks, entries
    if (lane == 31) { wsum[warp] = v; wflag[warp] = f; wents[warp] = e; }
    // look-back over previous CTAs of the same image (warp 0)
    if (warp == 0) {
        uint32_t carry = 0, ecarry = 0;
        const uint32_t first_cta = a.img_cta0[img];
        bool done = false;
        for (int64_t k = int64_t(cta) - 1; k >= int64_t(first_cta); k -= 32) {
            const int64_t idx = k - lane;
            uint2 part = make_uint2(0u, 0u);
            uint32_t ents = 0;
            if (idx >= int64_t(first_cta)) {
                part = a.cta_partial[idx];
                ents = a.cta_entries[idx];
            }
            if (!done) {
                const uint32_t flagged = __ballot_sync(0xFFFFFFFFu, part.x != 0);
                const int stop = flagged ? __ffs(flagged) - 1 : 31;   // nearest CTA (smallest lane) holding a start
                uint32_t contrib = (lane <= stop) ? part.y : 0u;
#pragma unroll
                for (int d = 16; d > 0; d >>= 1) contrib += __shfl_xor_sync(0xFFFFFFFFu, contrib, d);
                carry += contrib;
                done = flagged != 0;
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) ents += __shfl_xor_sync(0xFFFFFFFFu, ents, d);
            ecarry += ents;
        }
        if (lane == 0) { carry_s[0] = carry; carry_s[1] = ecarry; }
    }
    __syncthreads();
    uint32_t add = 0, eadd = carry_s[1];
    bool open = (f == 0);   // no segment start at or before this thread inside its warp
    for (int w = warp - 1; w >= 0; w--) {
        eadd += wents[w];
        if (open) {
            add += wsum[w];
            if (wflag[w]) open = false;
        }
    }
    if (open) add += carry_s[0];
    const uint32_t excl = (me.active && me.first) ? 0u : v + add - nb;
    uint32_t n = e + eadd - my_nnz;   // this thread's first entry, relative to the image

    if (!me.active) return;
    // ---- final decode: every symbol with magnitude bits becomes one 32-bit entry of the
    // image's coefficient stream, written at its final position (each thread owns a contiguous
    // run of the stream, so consecutive stores of a lane fill whole sectors in L2); every block
    // end records where the next block's entries begin; DC differences go to the compact
    // per-block array. No clearing, no read-modify-write, no ownership hand-over: a thread
    // decodes exactly the symbols that start inside its subsequence, as in the counting passes.
    const SegmentDesc sd = a.segments[me.seg];
    uint32_t key = 0;
    if (!me.first) key = StateKey(a.state[g - 1]);
    uint32_t p = StateOverflow(key);
    int c = StateC(key), z = StateZ(key);
    uint32_t blk = sd.blk_first + excl;
    const uint32_t limit = sd.blk_first + sd.blk_count;
    uint32_t* entries = a.entries + im.ent0;
    uint32_t* blk_ent = a.blk_ent + 2 * im.blk0;      // (first, end) entry index of every block
    int16_t* dcdiff = a.dcdiff + im.blk0;
    const uint32_t ent_cap = im.ent_cap;
    const TableSel sel = MakeTableSel(sm.mcu_dc, sm.mcu_ac, im.bpm);
    const int bpm = im.bpm;
    const uint32_t slot_sa = SharedAddr(sm.words + tid * K1Smem<S>::kSlotStride);
    const uint32_t lut_sa = SharedAddr(&sm.lut.fast[0][0]);
    const uint32_t end_bit = me.end_bit;
    if (me.first && blk < limit) blk_ent[2 * blk] = n;   // a restart interval's first block starts at this thread's first entry
    uint32_t dc_off = DcBytes(sel, c), ac_off = AcBytes(sel, c);
    uint32_t off = (z == 0) ? dc_off : ac_off;
    uint32_t q0 = 0, q1 = 0, q2 = 0, q3 = 0, cnt = 0;   // pending entries of the current 16-byte group
    while (p < end_bit && blk < limit) {
        const uint32_t win = PeekBits(slot_sa, p);
        uint32_t en = Lds16(lut_sa + off + ((win >> (31 - kFastBits)) & (2u * kFastSize - 2u)));
        if (en == 0) en = SlowEntryShared(lut_sa, off >> (kFastBits + 1), win >> 16);
        const uint32_t sz = EntrySize(en), bits = EntryBits(en);
        // RECEIVE + EXTEND (T.81 F.2.2.1): magnitude bits moved to the top of t
        const uint32_t t = win << (bits - sz);
        const uint32_t extra = __funnelshift_rc(t, 0u, 32u - sz);
        const int val = int(extra) - ((int(t) < 0) ? 0 : int((1u << sz) - 1u));
        if (z == 0) dcdiff[blk] = int16_t(val);
        z += EntryAdvance(en);
        p += bits;
        if (sz != 0) {
            // four entries are collected in registers (oldest in q0) and leave as one 128-bit store:
            // scattered 4-byte stores were throttling the L1 store path (profiles/r01c_*)
            q0 = q1; q1 = q2; q2 = q3;
            q3 = MakeCoefEntry((z - 1) & 63, val);
            if (++cnt == 4) {
                if (n + 4 <= ent_cap) *reinterpret_cast<uint4*>(entries + n) = make_uint4(q0, q1, q2, q3);
                n += 4;
                cnt = 0;
            }
        }
        off = ac_off;
        if (z >= 64) {
            z = 0;
            blk++;
            // [first, end) of the finished block's entries, and the first entry of the next block of the
            // interval (across a restart boundary the next interval's first thread records its own
            // start: the counting pass may have taken the padding bits that end an interval for one
            // more symbol, so that thread's entries can begin one slot further on)
            blk_ent[2 * (blk - 1) + 1] = n + cnt;
            if (blk < limit) blk_ent[2 * blk] = n + cnt;
            c = (c + 1 == bpm) ? 0 : c + 1;
            dc_off = DcBytes(sel, c);
            ac_off = AcBytes(sel, c);
            off = dc_off;
        }
    }
    if (cnt) {   // last, partial group: padded with zero entries (position 0, which K2 overwrites with the DC anyway)
        for (; cnt < 4; cnt++) { q0 = q1; q1 = q2; q2 = q3; q3 = 0; }
        if (n + 4 <= ent_cap) *reinterpret_cast<uint4*>(entries + n) = make_uint4(q0, q1, q2, q3);
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
    static_assert(sizeof(K1Smem<S>) <= 48 * 1024, "K1 shared memory exceeds the default limit");
    if (round >= 0) k1_sync<S><<<a.total_ctas, T, sizeof(K1Smem<S>), stream>>>(a, round);
    else k1_write<S><<<a.total_ctas, T, sizeof(K1Smem<S>), stream>>>(a);
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
