// k1_model.cpp — CPU model of the K1 schedule (TEST TOOL, not product, not oracle).
//
// Runs the product's own parser (csrc/jpeg_parser.cpp) and the product's own
// sequential decode core (csrc/huff_core.cuh, compiled for the host) through a
// serial emulation of what k1_huffman.cu does in parallel: subsequences of S
// bytes, CTAs of T subsequences, speculative round 0 with CTA-local Jacobi
// fix-up (with the CTA halo), cross-CTA rounds, CTA partials + prefix, write pass, DC scan.
// It lets the no-GPU test suite check the *algorithm* (state packing,
// convergence logic, block positions, DC integration) against the oracle; the
// CUDA kernels themselves are checked by the -m gpu tests.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "huff_core.cuh"
#include "jpeg_parser.h"

using namespace rjb;

namespace {
const uint8_t kZig[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
                          41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
                          30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

struct HostLoader {
    const uint8_t* base;
    uint32_t operator()(uint32_t i) const {
        uint32_t w;
        std::memcpy(&w, base + size_t(i) * 4, 4);
        return ByteSwap32(w);   // the core reads big-endian words
    }
};
// Mirrors k1_write in k1_huffman.cu: a sparse entry stream per image + per-block end-of-entries index.
struct HostSink {
    uint32_t* entries;
    uint32_t* blk_end;
    int16_t* dcdiff;
    uint32_t n;          // next entry index
    uint32_t cap;
    void Dc(uint32_t blk, int v) { dcdiff[blk] = int16_t(v); }
    void Entry(int pos, int v) {
        if (n < cap) entries[n] = MakeCoefEntry(pos, v);
        n++;
    }
    void EndBlock(uint32_t blk) { blk_end[blk] = n; }
};
struct SubInfo {
    uint32_t seg;
    bool first, last;
    uint32_t end_bit;
    size_t start;
};
}  // namespace

struct K1ModelStats {
    uint32_t rounds;            // k1_sync launches needed (>= 2)
    uint32_t decodes[8];        // decodes per round
    uint32_t max_local_iters;   // deepest CTA-local fix-up loop
    uint32_t nsub, nctas;
};

extern "C" int k1_model_decode(const uint8_t* data, size_t len, int S, int T, int16_t* out, size_t count, K1ModelStats* stats) {
    StreamParser parser;
    if (!parser.Parse(data, len)) return -3;
    const ParsedJpeg& p = parser.parsed();
    if (p.support_status != 0) return p.support_status;
    const HostScan& hscan = parser.host_scan();   // the host restatement of the destuffing pass (the device runs k0_destuff.cu)
    const uint8_t* clean = hscan.clean.data();
    const int bpm = p.bpm;
    uint8_t mcu_comp[kMaxBlocksPerMcu], mcu_dc[kMaxBlocksPerMcu], mcu_ac[kMaxBlocksPerMcu];
    int comp_first[3] = {0, 0, 0};
    int k = 0;
    for (int c = 0; c < p.ncomp; c++) {
        int H = p.ncomp == 1 ? 1 : p.hs[c], V = p.ncomp == 1 ? 1 : p.vs[c];
        comp_first[c] = k;
        for (int b = 0; b < H * V; b++, k++) {
            mcu_comp[k] = uint8_t(c);
            mcu_dc[k] = uint8_t(p.td[c]);
            mcu_ac[k] = uint8_t(kHuffIds + p.ta[c]);
        }
    }
    const TableSel sel = MakeTableSel(mcu_dc, mcu_ac, bpm);
    const uint32_t total_mcus = uint32_t(p.mcus_x) * uint32_t(p.mcus_y);
    const uint32_t ri = p.restart_interval > 0 ? uint32_t(p.restart_interval) : total_mcus;
    // subsequences
    std::vector<SubInfo> subs;
    std::vector<uint32_t> seg_blk_first, seg_blk_count;
    for (size_t s = 0; s < hscan.segments.size(); s++) {
        const Segment& sg = hscan.segments[s];
        uint32_t n = (sg.nbytes + S - 1) / S;
        uint64_t mf = uint64_t(s) * ri;
        uint64_t mc = mf >= total_mcus ? 0 : std::min<uint64_t>(ri, total_mcus - mf);
        seg_blk_first.push_back(uint32_t(mf * bpm));
        seg_blk_count.push_back(uint32_t(mc * bpm));
        for (uint32_t j = 0; j < n; j++) {
            uint32_t remain = sg.nbytes - j * S;
            subs.push_back(SubInfo{uint32_t(s), j == 0, j + 1 == n, std::min<uint32_t>(remain, uint32_t(S)) * 8u, size_t(sg.offset) + size_t(j) * S});
        }
    }
    const uint32_t nsub = uint32_t(subs.size());
    // T = subsequences a CTA owns; its halo threads (K1Args::halo: 2 for large pictures, 768 bytes' worth for small
    // ones; K1_MODEL_HALO overrides) re-decode the subsequences before them
    uint32_t H = uint32_t(768 / S);
    if (const char* e = std::getenv("K1_MODEL_HALO")) H = uint32_t(std::atoi(e));
    // K1_MODEL_TOLERANT=1: damaged streams are handled as the device handles them instead of failing the model's self-checks
    const bool tolerant = std::getenv("K1_MODEL_TOLERANT") != nullptr && std::atoi(std::getenv("K1_MODEL_TOLERANT")) != 0;
    const uint32_t nctas = (nsub + T - 1) / T;
    std::vector<uint32_t> state(nsub, 0), used(nsub, 0), nnzv(nsub, 0);
    NullSink nsink;
    auto decode_from = [&](uint32_t g, uint32_t key, uint32_t* nnz_out) {
        uint32_t pb = StateOverflow(key), nb = 0, blk = 0, nnz = 0;
        int c = StateC(key), z = StateZ(key);
        HostLoader ld{clean + subs[g].start};
        DecodeSpan<false>(ld, &parser.lut(), sel, bpm, pb, subs[g].end_bit, c, z, nb, nnz, blk, 0xFFFFFFFFu, nsink);
        *nnz_out = (nnz + 7u) & ~7u;   // runs are padded to whole 32-byte stores
        uint32_t over = pb > subs[g].end_bit ? pb - subs[g].end_bit : 0;
        return PackState(over, c, z, nb > 0xFFFFu ? 0xFFFFu : nb);
    };
    if (stats) std::memset(stats, 0, sizeof(*stats));
    uint32_t round = 0;
    for (;; round++) {
        uint32_t boundary_changes = 0, ndec = 0;
        std::vector<uint32_t> prev_state = state;   // what other CTAs see during this round (worst case: old values)
        for (uint32_t cta = 0; cta < nctas; cta++) {
            const uint32_t g0 = cta * T, g1 = std::min(nsub, g0 + T);
            if (round > 0) {
                bool need0 = !subs[g0].first && StateKey(prev_state[g0 - 1]) != used[g0];
                if (!need0) continue;
            }
            // local window: round 0 includes the halo, later rounds start at the first owned subsequence
            const uint32_t lo = round == 0 ? (g0 >= H ? g0 - H : 0) : g0;
            std::vector<uint32_t> lstate(g1 - lo), lused(g1 - lo), lnnz(g1 - lo);
            for (uint32_t g = lo; g < g1; g++) {
                lstate[g - lo] = round == 0 ? 0 : state[g];
                lused[g - lo] = round == 0 ? 0 : used[g];
                lnnz[g - lo] = round == 0 ? 0 : nnzv[g];
            }
            std::vector<uint32_t> old(lstate);
            if (round == 0)
                for (uint32_t g = lo; g < g1; g++) {
                    lstate[g - lo] = decode_from(g, 0, &lnnz[g - lo]);
                    ndec++;
                }
            for (uint32_t iter = 0;; iter++) {
                std::vector<uint32_t> snap(lstate);
                bool any = false;
                for (uint32_t g = lo; g < g1; g++) {
                    if (subs[g].first) continue;
                    uint32_t in = lused[g - lo];
                    if (g > lo) in = StateKey(snap[g - 1 - lo]);
                    else if (round > 0) in = StateKey(prev_state[g - 1]);
                    if (in != lused[g - lo]) {
                        lstate[g - lo] = decode_from(g, in, &lnnz[g - lo]);
                        lused[g - lo] = in;
                        ndec++;
                        any = true;
                    }
                }
                if (stats) stats->max_local_iters = std::max(stats->max_local_iters, iter);
                if (!any) break;
            }
            for (uint32_t g = g0; g < g1; g++) {   // only owned results are stored
                state[g] = lstate[g - lo];
                used[g] = lused[g - lo];
                nnzv[g] = lnnz[g - lo];
            }
            const uint32_t gl = g0 + T - 1;
            if (gl < nsub && !subs[gl].last && (round == 0 || StateKey(old[gl - lo]) != StateKey(state[gl]))) boundary_changes++;
        }
        if (stats && round < 8) stats->decodes[round] = ndec;
        if (round > 0 && boundary_changes == 0) break;
        if (round > nctas + 2) return -6;
    }
    if (stats) {
        stats->rounds = round + 1;
        stats->nsub = nsub;
        stats->nctas = nctas;
    }
    // CTA partials + carry (as k1_scan / k1_write do)
    std::vector<uint32_t> cta_flag(nctas, 0), cta_tail(nctas, 0);
    for (uint32_t cta = 0; cta < nctas; cta++) {
        uint32_t tail = 0, flag = 0;
        for (uint32_t g = cta * T; g < std::min(nsub, (cta + 1) * T); g++) {
            if (subs[g].first) { flag = 1; tail = 0; }
            tail += StateBlocks(state[g]);
        }
        cta_flag[cta] = flag;
        cta_tail[cta] = tail;
    }
    const uint32_t nblocks = total_mcus * uint32_t(bpm);
    // the entry arena is NOT cleared on the device: start from garbage; per-block indices start
    // as "never decoded"
    const uint32_t cap = uint32_t(uint64_t(hscan.clean_bytes) * 8 / p.min_entry_bits + 7 * (uint64_t(hscan.clean_bytes) / 32 + hscan.segments.size() + 1) + 64);
    std::vector<uint32_t> entries(cap, 0x77777777u), blk_end(size_t(nblocks), 0xFFFFFFFFu);
    std::vector<int16_t> dcdiff(nblocks, int16_t(0x7777));
    uint32_t ent_run = 0;
    for (uint32_t cta = 0; cta < nctas; cta++) {
        uint32_t carry = 0;
        for (int64_t kk = int64_t(cta) - 1; kk >= 0; kk--) {
            carry += cta_tail[kk];
            if (cta_flag[kk]) break;
        }
        uint32_t run = carry;
        for (uint32_t g = cta * T; g < std::min(nsub, (cta + 1) * T); g++) {
            if (subs[g].first) run = 0;
            const uint32_t excl = run;
            run += StateBlocks(state[g]);
            const uint32_t n0 = ent_run;
            ent_run += nnzv[g];
            uint32_t key = subs[g].first ? 0 : StateKey(state[g - 1]);
            uint32_t pb = StateOverflow(key), cnt = 0, nnz = 0;
            int c = StateC(key), z = StateZ(key);
            uint32_t blk = seg_blk_first[subs[g].seg] + excl;
            const uint32_t limit = seg_blk_first[subs[g].seg] + seg_blk_count[subs[g].seg];
            HostSink sink{entries.data(), blk_end.data(), dcdiff.data(), n0, cap};
            HostLoader ld{clean + subs[g].start};
            DecodeSpan<true>(ld, &parser.lut(), sel, bpm, pb, subs[g].end_bit, c, z, cnt, nnz, blk, limit, sink);
            // counting pass and write pass must agree (the last thread of an interval may have counted padding)
            const uint32_t used_n = (sink.n - n0 + 7u) & ~7u;
            // (a damaged interval can hold more blocks than it should: the write pass stops at the interval's last block,
            // the counting pass did not - fewer entries than reserved is then legitimate, more never is)
            if (tolerant ? used_n > nnzv[g] : (subs[g].last ? used_n > nnzv[g] : used_n != nnzv[g])) return -7;
            // zero padding: the tail of the last group and the groups the counting pass reserved in vain
            for (uint32_t k2 = sink.n; k2 < n0 + nnzv[g] && k2 < cap; k2++) entries[k2] = kPadEntry;
        }
    }
    // densify (what K2 / the coefficient tap do)
    std::vector<int16_t> coef(size_t(nblocks) * 64, 0);
    for (uint32_t b = 0; b < nblocks; b++) {
        uint32_t e0 = b ? blk_end[b - 1] : 0u, e1 = blk_end[b];   // a block begins where its predecessor ends
        if (e0 == 0xFFFFFFFFu || e1 == 0xFFFFFFFFu || e1 < e0 || e1 - e0 > 0xFFFFu || e1 > cap) {
            if (!tolerant) return -8;
            if (e1 == 0xFFFFFFFFu) dcdiff[b] = 0;   // K2 / the DC kernels: a block no thread reached decodes as zero
            continue;
        }
        for (uint32_t k2 = e0; k2 < e1; k2++) coef[size_t(b) * 64 + kZig[CoefEntryPos(entries[k2])]] = int16_t(entries[k2] & 0xFFFFu);
    }
    // DC integration per component, reset at restart intervals
    {
        int pred[3] = {0, 0, 0};
        for (uint32_t m = 0; m < total_mcus; m++) {
            if (m % ri == 0) pred[0] = pred[1] = pred[2] = 0;
            for (int b = 0; b < bpm; b++) {
                pred[mcu_comp[b]] += dcdiff[size_t(m) * bpm + b];
                coef[(size_t(m) * bpm + b) * 64] = int16_t(pred[mcu_comp[b]]);
            }
        }
    }
    // reorder: decode order -> component-major raster
    size_t need = 0;
    for (int c = 0; c < p.ncomp; c++) need += size_t(p.blocks_w[c]) * p.blocks_h[c] * 64;
    if (count < need) return -2;
    size_t base = 0;
    for (int c = 0; c < p.ncomp; c++) {
        int H = p.ncomp == 1 ? 1 : p.hs[c], V = p.ncomp == 1 ? 1 : p.vs[c];
        for (int by = 0; by < p.blocks_h[c]; by++)
            for (int bx = 0; bx < p.blocks_w[c]; bx++) {
                size_t mcu = size_t(by / V) * p.mcus_x + size_t(bx / H);
                size_t kb = size_t(comp_first[c] + (by % V) * H + (bx % H));
                std::memcpy(out + base + (size_t(by) * p.blocks_w[c] + bx) * 64, &coef[(mcu * bpm + kb) * 64], 128);
            }
        base += size_t(p.blocks_w[c]) * p.blocks_h[c] * 64;
    }
    return 0;
}
