/* jpeg_oracle.c — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A scalar, single-threaded restatement of the decode path this repository
 * accelerates. Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load it; the product (librocjpeg.so) never does.
 *
 * What it follows:
 *   - marker parsing and accept/reject rules: the reference parser,
 *     src/rocjpeg_parser.cpp:43-124 (walk), :160-207 (SOF), :217-246 (DQT),
 *     :256-313 (DHT), :324-363 (SOS), :374-390 (DRI), :400-416 (scan slice up
 *     to the first FFD9), :432-470 (chroma-subsampling classification);
 *   - image info: src/rocjpeg_decoder.cpp:307-358;
 *   - entropy decode / dequantisation / IDCT: these run inside AMD VCN
 *     fixed-function hardware in the reference (submitted at
 *     src/rocjpeg_vaapi_decoder.cpp:677-689) and are absent from its source.
 *     They are restated from ITU-T T.81 (baseline Huffman, Annex F.2.2) and
 *     libjpeg's jidctint.c "islow" integer IDCT as BASELINE.json mandates, and
 *     are pinned against libjpeg-turbo 3.1.4.1 (jpeg_read_coefficients,
 *     jpeg_read_raw_data) by tests/test_oracle_pinning.py;
 *   - output assembly (ROI offsets, surface layouts, per-format copies):
 *     src/rocjpeg_decoder.cpp:143-180, :372-399, :450-494, :511-557, :576-636
 *     with the VCN surface layout per subsampling from
 *     src/rocjpeg_vaapi_decoder.cpp:612-637;
 *   - colour conversion and nearest-neighbour chroma upsampling:
 *     src/rocjpeg_hip_kernels.cpp:76-89 (444), :499-617 (440), :947-954 (YUYV),
 *     :1389-1433 (NV12), :1915-1927 (400); pinned against the reference's own
 *     kernels compiled for the CPU (oracle/_ref) by tests/test_oracle_pinning.py.
 *     The float->u8 pack (hipPack, :25-30, v_cvt_pk_u8_f32) has no stated
 *     rounding mode in the reference: this oracle fixes round-to-nearest-even
 *     with saturation (CUDA cvt.rni.sat.u8.f32). PARITY NOTE: that one
 *     convention is unpinned by the reference.
 *
 * Build: see oracle/build.py (gcc -O2 -ffp-contract=off -shared -fPIC).
 */
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>

#define ORC_OK 0
#define ORC_BAD_JPEG (-3)
#define ORC_NOT_SUPPORTED (-4)
#define ORC_INVALID (-2)

enum { CSS_444 = 0, CSS_440 = 1, CSS_422 = 2, CSS_420 = 3, CSS_411 = 4, CSS_400 = 5, CSS_UNKNOWN = -1 };
enum { FMT_NATIVE = 0, FMT_YUV_PLANAR = 1, FMT_Y = 2, FMT_RGB = 3, FMT_RGB_PLANAR = 4 };

typedef struct {
    int32_t width, height, ncomp, css;
    int32_t comp_id[3], hs[3], vs[3], tq[3];
    int32_t scan_ncomp, td[3], ta[3];
    int32_t hmax, vmax, mcus_x, mcus_y, blocks_per_mcu;
    int32_t blocks_w[3], blocks_h[3];     /* padded block grid of each component */
    int32_t restart_interval;
    uint32_t num_mcus_ref;                /* the reference's num_mcus (parser.cpp:197) */
    uint32_t scan_offset, scan_size;      /* entropy-coded slice: [offset, offset+size) */
    uint8_t qt[4][64];                    /* zig-zag order, as stored in the stream (low byte of a 16-bit table) */
    uint8_t qt_present[4];
    uint8_t dc_bits[4][16], dc_vals[4][12];
    uint8_t ac_bits[4][16], ac_vals[4][162];
    uint8_t dc_present[4], ac_present[4];
    uint32_t n_restart_markers;           /* RSTn markers found inside the slice */
    uint16_t qt16[4][64];                 /* the quantiser steps at full width, zig-zag order */
    int32_t features;                     /* ORC_FEAT_*: what the stream uses beyond what the reference's parser accepts */
} OrcInfo;

/* Widening beyond the reference's parser: the sequential Huffman process with 8-bit samples
 * is the same decode whether the frame header says SOF0 or SOF1 (T.81 Table B.1, F.1.1: "extended" only permits what
 * follows); quantiser tables with 16-bit steps (the reference rejects them, parser.cpp:230) and Huffman table ids 2 and 3
 * (rejected at parser.cpp:274) only change table storage. libjpeg-turbo writes SOF1 as soon as a quantiser step exceeds 255. */
#define ORC_FEAT_SOF1 1
#define ORC_FEAT_DQT16 2
#define ORC_FEAT_HUFF_ID23 4

static const uint8_t kZigzag[64] = {
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

const uint8_t *orc_zigzag_table(void) { return kZigzag; }

/* src/rocjpeg_parser.cpp:432-470 */
static int classify_css(const int h[3], const int v[3]) {
    if ((h[0] == 1 && h[1] == 1 && h[2] == 1 && v[0] == 1 && v[1] == 1 && v[2] == 1) ||
        (h[0] == 2 && h[1] == 2 && h[2] == 2 && v[0] == 2 && v[1] == 2 && v[2] == 2) ||
        (h[0] == 4 && h[1] == 4 && h[2] == 4 && v[0] == 4 && v[1] == 4 && v[2] == 4))
        return CSS_444;
    if (h[0] == 1 && h[1] == 1 && h[2] == 1 && v[0] == 2 && v[1] == 1 && v[2] == 1) return CSS_440;
    if ((h[0] == 2 && h[1] == 1 && h[2] == 1 && v[0] == 1 && v[1] == 1 && v[2] == 1) ||
        (h[0] == 2 && h[1] == 1 && h[2] == 1 && v[0] == 2 && v[1] == 2 && v[2] == 2) ||
        (h[0] == 2 && h[1] == 2 && h[2] == 2 && v[0] == 2 && v[1] == 1 && v[2] == 1))
        return CSS_422;
    if (h[0] == 2 && h[1] == 1 && h[2] == 1 && v[0] == 2 && v[1] == 1 && v[2] == 1) return CSS_420;
    if (h[0] == 4 && h[1] == 1 && h[2] == 1 && v[0] == 1 && v[1] == 1 && v[2] == 1) return CSS_411;
    if ((h[0] == 1 && h[1] == 0 && h[2] == 0 && v[0] == 1 && v[1] == 0 && v[2] == 0) ||
        (h[0] == 4 && h[1] == 0 && h[2] == 0 && v[0] == 4 && v[1] == 0 && v[2] == 0))
        return CSS_400;
    return CSS_UNKNOWN;
}

#define RD16(p) (((uint32_t)(p)[0] << 8) | (p)[1])

/* Marker walk. Same accept/reject decisions as the reference parser, plus
 * bounds checks (the reference reads past the end of truncated input). */
int orc_parse(const uint8_t *d, size_t len, OrcInfo *o) {
    memset(o, 0, sizeof(*o));
    o->css = CSS_UNKNOWN;
    if (!d || len < 4) return ORC_BAD_JPEG;
    if (d[0] != 0xFF || d[1] != 0xD8) return ORC_BAD_JPEG;       /* parser.cpp:64 */
    size_t p = 2;
    int seen_dht = 0, seen_dqt = 0, seen_sos = 0, seen_sof = 0;
    while (!seen_sos) {
        if (p + 4 > len) return ORC_BAD_JPEG;
        while (p < len && d[p] == 0xFF) p++;                     /* parser.cpp:75 */
        if (p + 3 > len) return ORC_BAD_JPEG;
        uint8_t m = d[p++];
        uint32_t seglen = RD16(d + p);
        size_t next = p + seglen;
        if (seglen < 2 || next > len) return ORC_BAD_JPEG;
        const uint8_t *s = d + p;
        switch (m) {
        case 0xC1:   /* SOF1: extended sequential, Huffman - the same decode when the samples have 8 bits */
            if (seglen < 8 || s[2] != 8) return ORC_BAD_JPEG;
            o->features |= ORC_FEAT_SOF1;
            /* fall through */
        case 0xC0: { /* SOF0, parser.cpp:160-207 */
            if (seglen < 8) return ORC_BAD_JPEG;
            o->height = (int)RD16(s + 3);
            o->width = (int)RD16(s + 5);
            o->ncomp = s[7];
            if (o->ncomp > 3) return ORC_BAD_JPEG;              /* parser.cpp:172 */
            if (seglen < 8u + 3u * (uint32_t)o->ncomp) return ORC_BAD_JPEG;
            for (int i = 0; i < o->ncomp; i++) {
                o->comp_id[i] = s[8 + 3 * i];
                uint8_t sf = s[9 + 3 * i];
                uint8_t tq = s[10 + 3 * i];
                if (tq >= 4) return ORC_BAD_JPEG;                /* parser.cpp:185 */
                o->hs[i] = sf >> 4;
                o->vs[i] = sf & 15;
                o->tq[i] = tq;
            }
            int hmaxf = o->hs[0] ? o->hs[0] : 1, vmaxf = o->vs[0] ? o->vs[0] : 1;
            o->num_mcus_ref = (uint32_t)((o->width + hmaxf * 8 - 1) / (hmaxf * 8)) *
                              (uint32_t)((o->height + vmaxf * 8 - 1) / (vmaxf * 8));  /* parser.cpp:197 */
            o->css = classify_css(o->hs, o->vs);
            seen_sof = 1;
            break;
        }
        case 0xC4: { /* DHT, parser.cpp:256-313 */
            int32_t rem = (int32_t)seglen - 2;
            const uint8_t *q = s + 2;
            while (rem > 0) {
                if (rem < 17) return ORC_BAD_JPEG;
                uint8_t idx = *q++;
                int is_ac = idx & 0xF0, id = idx & 0x0F;
                if (id >= 4) return ORC_BAD_JPEG;                /* T.81 B.2.4.2: Th 0..3 (the reference stops at 1, parser.cpp:274) */
                if (id >= 2) o->features |= ORC_FEAT_HUFF_ID23;
                uint32_t count = 0;
                for (int i = 0; i < 16; i++) count += q[i];
                if (is_ac) {
                    if (count > 162) return ORC_BAD_JPEG;        /* parser.cpp:291 */
                } else {
                    if (count > 12) return ORC_BAD_JPEG;         /* parser.cpp:298 */
                }
                if ((int32_t)(17 + count) > rem) return ORC_BAD_JPEG;
                if (is_ac) {
                    memcpy(o->ac_bits[id], q, 16);
                    memcpy(o->ac_vals[id], q + 16, count);
                    o->ac_present[id] = 1;
                } else {
                    memcpy(o->dc_bits[id], q, 16);
                    memcpy(o->dc_vals[id], q + 16, count);
                    o->dc_present[id] = 1;
                }
                q += 16 + count;
                rem -= 17 + (int32_t)count;
            }
            seen_dht = 1;
            break;
        }
        case 0xDB: { /* DQT, parser.cpp:217-246 */
            const uint8_t *q = s + 2, *end = s + seglen;
            while (q < end) {
                uint8_t idx = *q++;
                int wide = idx >> 4;                             /* Pq: 0 = 8-bit steps, 1 = 16-bit (rejected by the reference, parser.cpp:230) */
                idx &= 15;
                if (wide > 1) return ORC_BAD_JPEG;
                if (idx >= 4) return ORC_BAD_JPEG;               /* parser.cpp:234 */
                if (q + (wide ? 128 : 64) > end) return ORC_BAD_JPEG;
                for (int k = 0; k < 64; k++) {
                    o->qt16[idx][k] = wide ? (uint16_t)RD16(q + 2 * k) : q[k];
                    o->qt[idx][k] = (uint8_t)o->qt16[idx][k];
                }
                if (wide) o->features |= ORC_FEAT_DQT16;
                o->qt_present[idx] = 1;
                q += wide ? 128 : 64;
            }
            seen_dqt = 1;
            break;
        }
        case 0xDD: /* DRI, parser.cpp:374-390 */
            if (seglen != 4) return ORC_BAD_JPEG;
            o->restart_interval = (int)RD16(s + 2);
            break;
        case 0xDA: { /* SOS, parser.cpp:324-363 */
            if (seglen < 3) return ORC_BAD_JPEG;
            int n = s[2];
            if (n > 3) return ORC_BAD_JPEG;                      /* parser.cpp:333 */
            if (seglen < 6u + 2u * (uint32_t)n) return ORC_BAD_JPEG;
            o->scan_ncomp = n;
            for (int i = 0; i < n; i++) {
                uint8_t cid = s[3 + 2 * i], t = s[4 + 2 * i];
                if ((t & 15) >= 4 || (t >> 4) >= 4) return ORC_BAD_JPEG;      /* parser.cpp:347-354 */
                if (cid != o->comp_id[i]) return ORC_BAD_JPEG;                /* parser.cpp:355 */
                o->td[i] = t >> 4;
                o->ta[i] = t & 15;
            }
            seen_sos = 1;
            break;
        }
        default: /* APPn, COM, SOF2, ... skipped by length (parser.cpp:105-108) */
            break;
        }
        p = next;
    }
    if (!seen_dht || !seen_dqt) return ORC_BAD_JPEG;            /* parser.cpp:111-118 */
    /* Slice = bytes up to (not including) the first FF D9 (parser.cpp:400-416);
     * a stream without EOI runs to the end of the buffer. */
    size_t e = p;
    uint32_t nrst = 0;
    while (e < len) {
        if (d[e] == 0xFF && e + 1 < len) {
            if (d[e + 1] == 0xD9) break;
            if ((d[e + 1] & 0xF8) == 0xD0) nrst++;
        }
        e++;
    }
    o->scan_offset = (uint32_t)p;
    o->scan_size = (uint32_t)(e - p);
    o->n_restart_markers = nrst;
    (void)seen_sof;

    /* Derived geometry (T.81 A.1.1, A.2). */
    if (o->ncomp >= 1 && o->width > 0 && o->height > 0) {
        int hmax = 1, vmax = 1;
        for (int i = 0; i < o->ncomp; i++) {
            if (o->hs[i] > hmax) hmax = o->hs[i];
            if (o->vs[i] > vmax) vmax = o->vs[i];
        }
        o->hmax = hmax;
        o->vmax = vmax;
        if (o->ncomp == 1) {
            /* non-interleaved: MCU = one 8x8 block, sampling factors ignored */
            o->mcus_x = (o->width + 7) / 8;
            o->mcus_y = (o->height + 7) / 8;
            o->blocks_per_mcu = 1;
            o->blocks_w[0] = o->mcus_x;
            o->blocks_h[0] = o->mcus_y;
        } else {
            o->mcus_x = (o->width + 8 * hmax - 1) / (8 * hmax);
            o->mcus_y = (o->height + 8 * vmax - 1) / (8 * vmax);
            o->blocks_per_mcu = 0;
            for (int i = 0; i < o->ncomp; i++) {
                o->blocks_w[i] = o->mcus_x * o->hs[i];
                o->blocks_h[i] = o->mcus_y * o->vs[i];
                o->blocks_per_mcu += o->hs[i] * o->vs[i];
            }
        }
    }
    return ORC_OK;
}

/* Is this stream inside what the decode path supports? (baseline, 8-bit,
 * one interleaved scan over all components, css in {444,440,422,420,400}) */
int orc_supported(const OrcInfo *o) {
    if (o->width <= 0 || o->height <= 0) return ORC_NOT_SUPPORTED;
    if (o->ncomp != 1 && o->ncomp != 3) return ORC_NOT_SUPPORTED;
    if (o->scan_ncomp != o->ncomp) return ORC_NOT_SUPPORTED;
    if (o->css == CSS_UNKNOWN) return ORC_NOT_SUPPORTED;   /* 4:1:1 (api/rocjpeg.h:91, rejected by the reference's decoder) is decoded: section 8 f4 */
    if (o->blocks_per_mcu > 10) return ORC_NOT_SUPPORTED;
    if (o->css == CSS_422 && o->ncomp == 3 && o->hs[1] == o->hs[0]) return ORC_NOT_SUPPORTED; /* (2,2,2|2,1,1) */
    for (int i = 0; i < o->ncomp; i++) {
        if (!o->qt_present[o->tq[i]]) return ORC_BAD_JPEG;
        if (!o->dc_present[o->td[i]] || !o->ac_present[o->ta[i]]) return ORC_BAD_JPEG;
    }
    return ORC_OK;
}

/* ---------------------------------------------------------------- Huffman */

typedef struct {
    int32_t mincode[17], maxcode[18], valptr[17];
    const uint8_t *vals;
} HuffTab;

/* T.81 Annex C (code generation) + F.2.2.3 (decoder tables). */
static void build_hufftab(const uint8_t bits[16], const uint8_t *vals, HuffTab *t) {
    int32_t code = 0, k = 0;
    for (int l = 1; l <= 16; l++) {
        t->valptr[l] = k;
        t->mincode[l] = code;
        code += bits[l - 1];
        k += bits[l - 1];
        t->maxcode[l] = bits[l - 1] ? code - 1 : -1;
        code <<= 1;
    }
    t->maxcode[17] = 0x7fffffff;
    t->vals = vals;
}

typedef struct {
    const uint8_t *d;
    size_t pos, end;
    uint32_t acc;
    int nbits;
    int hit_marker; /* 0, or the marker byte met */
} BitRd;

static void br_fill(BitRd *b) {
    while (b->nbits <= 24) {
        uint32_t byte = 0;
        if (!b->hit_marker && b->pos < b->end) {
            byte = b->d[b->pos];
            if (byte == 0xFF) {
                uint8_t nx = (b->pos + 1 < b->end) ? b->d[b->pos + 1] : 0xD9;
                if (nx == 0x00) {
                    b->pos += 2;
                } else {
                    b->hit_marker = nx; /* feed zeros from here on */
                    byte = 0;
                }
            } else {
                b->pos++;
            }
        }
        b->acc |= byte << (24 - b->nbits);
        b->nbits += 8;
    }
}

static inline uint32_t br_peek(BitRd *b, int n) {
    br_fill(b);
    return b->acc >> (32 - n);
}
static inline void br_skip(BitRd *b, int n) {
    b->acc <<= n;
    b->nbits -= n;
}
static inline int32_t br_get(BitRd *b, int n) {
    if (n == 0) return 0;
    uint32_t v = br_peek(b, n);
    br_skip(b, n);
    return (int32_t)v;
}

static int huff_decode(BitRd *b, const HuffTab *t) {
    uint32_t w = br_peek(b, 16);
    for (int l = 1; l <= 16; l++) {
        int32_t code = (int32_t)(w >> (16 - l));
        if (t->maxcode[l] >= 0 && code <= t->maxcode[l] && code >= t->mincode[l]) {
            br_skip(b, l);
            return t->vals[t->valptr[l] + code - t->mincode[l]];
        }
    }
    br_skip(b, 16);
    return 0; /* invalid code: corrupt data */
}

static inline int32_t extend(int32_t v, int s) { /* T.81 F.2.2.1 EXTEND */
    return (s && v < (1 << (s - 1))) ? v - (1 << s) + 1 : v;
}

size_t orc_coef_count(const OrcInfo *o) {
    size_t n = 0;
    for (int i = 0; i < o->ncomp; i++) n += (size_t)o->blocks_w[i] * o->blocks_h[i] * 64;
    return n;
}

/* Decode the entropy-coded slice into quantised coefficients, natural
 * (de-zig-zagged) order, int16. Layout: component-major; within a component
 * the padded block grid in raster order, 64 coefficients per block. */
int orc_decode_coefficients(const uint8_t *data, size_t len, const OrcInfo *o, int16_t *coefs) {
    int rc = orc_supported(o);
    if (rc != ORC_OK) return rc;
    if ((size_t)o->scan_offset + o->scan_size > len) return ORC_BAD_JPEG;
    HuffTab dc[4], ac[4];
    for (int i = 0; i < 4; i++) {
        if (o->dc_present[i]) build_hufftab(o->dc_bits[i], o->dc_vals[i], &dc[i]);
        if (o->ac_present[i]) build_hufftab(o->ac_bits[i], o->ac_vals[i], &ac[i]);
    }
    size_t comp_base[3], acc = 0;
    for (int i = 0; i < o->ncomp; i++) {
        comp_base[i] = acc;
        acc += (size_t)o->blocks_w[i] * o->blocks_h[i] * 64;
    }
    memset(coefs, 0, acc * sizeof(int16_t));
    BitRd b = {data + o->scan_offset, 0, o->scan_size, 0, 0, 0};
    int32_t pred[3] = {0, 0, 0};
    int total_mcus = o->mcus_x * o->mcus_y;
    int ri = o->restart_interval;
    for (int m = 0; m < total_mcus; m++) {
        if (ri && m && (m % ri) == 0) {
            /* restart: drop padding bits, expect RSTn, reset predictors */
            b.acc = 0;
            b.nbits = 0;
            if (!b.hit_marker) {
                /* marker not yet met by the bit reader: it must be next */
                while (b.pos + 1 < b.end && !(b.d[b.pos] == 0xFF && (b.d[b.pos + 1] & 0xF8) == 0xD0)) b.pos++;
            }
            if (b.pos + 1 < b.end && b.d[b.pos] == 0xFF && (b.d[b.pos + 1] & 0xF8) == 0xD0) b.pos += 2;
            b.hit_marker = 0;
            pred[0] = pred[1] = pred[2] = 0;
        }
        int mx = m % o->mcus_x, my = m / o->mcus_x;
        for (int c = 0; c < o->ncomp; c++) {
            int H = (o->ncomp == 1) ? 1 : o->hs[c], V = (o->ncomp == 1) ? 1 : o->vs[c];
            for (int v = 0; v < V; v++)
                for (int h = 0; h < H; h++) {
                    int bx = mx * H + h, by = my * V + v;
                    int16_t *blk = coefs + comp_base[c] + ((size_t)by * o->blocks_w[c] + bx) * 64;
                    int t = huff_decode(&b, &dc[o->td[c]]);
                    int32_t diff = extend(br_get(&b, t & 15), t & 15);
                    pred[c] += diff;
                    blk[0] = (int16_t)pred[c];
                    for (int k = 1; k < 64;) {
                        int rs = huff_decode(&b, &ac[o->ta[c]]);
                        int r = rs >> 4, s = rs & 15;
                        if (s == 0) {
                            if (r == 15) { k += 16; continue; }
                            break; /* EOB */
                        }
                        k += r;
                        int32_t val = extend(br_get(&b, s), s);
                        if (k < 64) blk[kZigzag[k]] = (int16_t)val;
                        k++;
                    }
                }
        }
    }
    return ORC_OK;
}

/* ------------------------------------------------------------------- IDCT */
/* libjpeg jidctint.c "islow" (CONST_BITS 13, PASS1_BITS 2), branch-free form. */
#define FIX_0_298631336 2446
#define FIX_0_390180644 3196
#define FIX_0_541196100 4433
#define FIX_0_765366865 6270
#define FIX_0_899976223 7373
#define FIX_1_175875602 9633
#define FIX_1_501321110 12299
#define FIX_1_847759065 15137
#define FIX_1_961570560 16069
#define FIX_2_053119869 16819
#define FIX_2_562915447 20995
#define FIX_3_072711026 25172

static inline void islow_1d(const int32_t in[8], int32_t out[8], int shift) {
    int32_t z1, z2, z3, z4, z5, t0, t1, t2, t3, t10, t11, t12, t13;
    const int32_t rnd = 1 << (shift - 1);
    z1 = (in[2] + in[6]) * FIX_0_541196100;
    t2 = z1 + in[6] * (-FIX_1_847759065);
    t3 = z1 + in[2] * FIX_0_765366865;
    t0 = (in[0] + in[4]) * 8192;
    t1 = (in[0] - in[4]) * 8192;
    t10 = t0 + t3; t13 = t0 - t3; t11 = t1 + t2; t12 = t1 - t2;
    int32_t a = in[7], b = in[5], c = in[3], d = in[1];
    z1 = a + d; z2 = b + c; z3 = a + c; z4 = b + d;
    z5 = (z3 + z4) * FIX_1_175875602;
    a *= FIX_0_298631336; b *= FIX_2_053119869; c *= FIX_3_072711026; d *= FIX_1_501321110;
    z1 *= -FIX_0_899976223; z2 *= -FIX_2_562915447;
    z3 = z3 * (-FIX_1_961570560) + z5;
    z4 = z4 * (-FIX_0_390180644) + z5;
    a += z1 + z3; b += z2 + z4; c += z2 + z3; d += z1 + z4;
    out[0] = (t10 + d + rnd) >> shift; out[7] = (t10 - d + rnd) >> shift;
    out[1] = (t11 + c + rnd) >> shift; out[6] = (t11 - c + rnd) >> shift;
    out[2] = (t12 + b + rnd) >> shift; out[5] = (t12 - b + rnd) >> shift;
    out[3] = (t13 + a + rnd) >> shift; out[4] = (t13 - a + rnd) >> shift;
}

/* coef: 64 natural-order quantised coefficients; q_nat: 64 natural-order
 * quantiser steps; out: 8 rows at `pitch`. */
void orc_idct_islow_block(const int16_t *coef, const uint16_t *q_nat, uint8_t *out, int pitch) {
    int32_t ws[64], col[8], res[8];
    for (int c = 0; c < 8; c++) {
        for (int r = 0; r < 8; r++) col[r] = (int32_t)coef[r * 8 + c] * (int32_t)q_nat[r * 8 + c];
        islow_1d(col, res, 13 - 2);
        for (int r = 0; r < 8; r++) ws[r * 8 + c] = res[r];
    }
    for (int r = 0; r < 8; r++) {
        islow_1d(ws + r * 8, res, 13 + 2 + 3);
        for (int c = 0; c < 8; c++) {
            int32_t v = res[c] + 128;
            out[r * pitch + c] = (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
        }
    }
}

size_t orc_plane_bytes(const OrcInfo *o) {
    size_t n = 0;
    for (int i = 0; i < o->ncomp; i++) n += (size_t)o->blocks_w[i] * o->blocks_h[i] * 64;
    return n;
}

/* Dequantise + IDCT every block. planes: component-major, each plane is
 * (blocks_w*8) x (blocks_h*8) bytes, pitch = blocks_w*8. */
int orc_idct_planes(const OrcInfo *o, const int16_t *coefs, uint8_t *planes) {
    size_t cb = 0;
    for (int c = 0; c < o->ncomp; c++) {
        uint16_t qn[64];
        for (int k = 0; k < 64; k++) qn[kZigzag[k]] = o->qt16[o->tq[c]][k];
        int pitch = o->blocks_w[c] * 8;
        for (int by = 0; by < o->blocks_h[c]; by++)
            for (int bx = 0; bx < o->blocks_w[c]; bx++) {
                const int16_t *blk = coefs + cb + ((size_t)by * o->blocks_w[c] + bx) * 64;
                orc_idct_islow_block(blk, qn, planes + cb + (size_t)by * 8 * pitch + bx * 8, pitch);
            }
        cb += (size_t)o->blocks_w[c] * o->blocks_h[c] * 64;
    }
    return ORC_OK;
}

/* ---------------------------------------------------------- output stage */

/* hipPack convention: saturating round-to-nearest-even float -> u8. */
static inline uint8_t pack_u8(float f) {
    if (!(f > 0.0f)) return 0;
    if (f >= 255.0f) return 255;
    return (uint8_t)lrintf(f); /* default rounding mode = nearest-even */
}

typedef struct {
    const OrcInfo *o;
    const uint8_t *pl[3];
    int pitch[3];
} Surf;

/* Virtual VCN surface accessors (layouts: src/rocjpeg_vaapi_decoder.cpp:612-637).
 * Rows/columns beyond the picture come from the MCU-padded decoded planes. */
static inline uint8_t pl_at(const Surf *s, int c, int x, int y) { return s->pl[c][(size_t)y * s->pitch[c] + x]; }
/* YUYV packed row y, byte k: Y0 U Y1 V */
static inline uint8_t yuyv_at(const Surf *s, int y, int k) {
    if ((k & 1) == 0) return pl_at(s, 0, k >> 1, y);
    return pl_at(s, (k & 2) ? 2 : 1, k >> 2, y);
}
/* NV12 chroma row y, byte k: U V interleaved */
static inline uint8_t nv12uv_at(const Surf *s, int y, int k) { return pl_at(s, 1 + (k & 1), k >> 1, y); }

static inline void rgb_from_yuv(uint8_t yy, uint8_t uu, uint8_t vv, uint8_t *r, uint8_t *g, uint8_t *b) {
    /* src/rocjpeg_hip_kernels.cpp:76-89 */
    float y = (float)yy, u = (float)uu - 128.0f, v = (float)vv - 128.0f;
    *r = pack_u8(fmaf(1.5748f, v, y));
    *g = pack_u8(fmaf(-0.4681f, v, fmaf(-0.1873f, u, y)));
    *b = pack_u8(fmaf(1.8556f, u, y));
}

/* ROI validity rule: src/rocjpeg_decoder.cpp:126-131. Returns 1 if the crop
 * rectangle selects a region, 0 for "whole picture", <0 if it passes the
 * reference's test but lies outside the picture (the reference would read
 * out of bounds; this implementation rejects it). */
int orc_roi(const OrcInfo *o, const int16_t crop[4], int *x0, int *y0, int *w, int *h) {
    uint32_t rw = (uint32_t)((int)crop[2] - (int)crop[0]);
    uint32_t rh = (uint32_t)((int)crop[3] - (int)crop[1]);
    *x0 = 0; *y0 = 0; *w = o->width; *h = o->height;
    if (rw > 0 && rh > 0 && rw <= (uint32_t)o->width && rh <= (uint32_t)o->height) {
        if (crop[0] < 0 || crop[1] < 0 || crop[2] > o->width || crop[3] > o->height) return ORC_INVALID;
        *x0 = crop[0]; *y0 = crop[1]; *w = (int)rw; *h = (int)rh;
        return 1;
    }
    return 0;
}

/* Assemble the caller-visible output. dst[c]/dst_pitch[c] follow RocJpegImage;
 * a channel with a null pointer or zero pitch is skipped
 * (src/rocjpeg_decoder.cpp:373). Only the valid bytes of each row are written.
 *
 * Region of interest. Without a crop rectangle every format follows the
 * reference byte for byte. With one:
 *   - NATIVE is a raw copy of the VCN surface at the reference's byte offsets
 *     (CopyChannel, decoder.cpp:376-389): top*pitch + left, chroma rows from
 *     top>>1 for 4:2:0 / 4:4:0, left*2 for packed YUYV;
 *   - Y, YUV_PLANAR, RGB and RGB_PLANAR are the geometric crop of the
 *     full-picture result: out(x, y) = full(left + x, top + y), chroma sample
 *     ((left+x)>>sx, (top+y)>>sy); planar chroma starts at (left>>sx, top>>sy)
 *     and is (W>>sx) x (H>>sy). For even left/top this is what the reference
 *     computes for 4:2:2 / 4:2:0 / 4:0:0. It deliberately does NOT reproduce
 *     three reference defects, which read the wrong samples (or out of
 *     bounds): RGB from 4:4:4 adds the ROI offset to the chroma planes twice
 *     (decoder.cpp:464-466 with hip_kernels.cpp:66-68), RGB from 4:4:0 offsets
 *     chroma by `top` instead of `top>>1` rows (decoder.cpp:468-470), and an
 *     odd `left` swaps U and V for 4:2:0 / mis-phases YUYV (decoder.cpp:456-461).
 */
int orc_convert(const OrcInfo *o, const uint8_t *planes, int fmt, const int16_t crop[4],
                uint8_t *dst[4], const uint32_t dst_pitch[4]) {
    Surf s;
    s.o = o;
    size_t cb = 0;
    for (int c = 0; c < 3; c++) {
        s.pl[c] = NULL;
        s.pitch[c] = 0;
    }
    for (int c = 0; c < o->ncomp; c++) {
        s.pl[c] = planes + cb;
        s.pitch[c] = o->blocks_w[c] * 8;
        cb += (size_t)o->blocks_w[c] * o->blocks_h[c] * 64;
    }
    int x0, y0, W, H;
    int roi = orc_roi(o, crop, &x0, &y0, &W, &H);
    if (roi < 0) return ORC_INVALID;
    int css = o->css;
    int sx = (css == CSS_411) ? 2 : (css == CSS_422 || css == CSS_420) ? 1 : 0;   /* chroma shift, horizontal */
    int sy = (css == CSS_440 || css == CSS_420) ? 1 : 0;   /* chroma shift, vertical */
#define CH_OK(c) (dst[c] != NULL && dst_pitch[c] != 0)
    switch (fmt) {
    case FMT_NATIVE: {
        if (css == CSS_422) {
            /* packed YUYV, byte offset top*pitch + 2*left (decoder.cpp:384-388) */
            if (CH_OK(0))
                for (int y = 0; y < H; y++)
                    for (int j = 0; j < 2 * W; j++)
                        dst[0][(size_t)y * dst_pitch[0] + j] = yuyv_at(&s, y0 + y, 2 * x0 + j);
            break;
        }
        if (CH_OK(0))
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++) dst[0][(size_t)y * dst_pitch[0] + x] = pl_at(&s, 0, x0 + x, y0 + y);
        if (css == CSS_444 || css == CSS_440 || css == CSS_411) {
            /* decoder.cpp:157-158 with CopyChannel's roi rule :376-389. 4:1:1 has no VCN surface in the reference
             * (it is rejected there): its native layout here is three planes, chroma (W>>2) x H, like 4:4:4 / 4:4:0 */
            int ch = H >> sy, cy0 = y0 >> sy, cw = W >> sx, cx0 = x0 >> sx;
            for (int c = 1; c < 3; c++)
                if (CH_OK(c))
                    for (int y = 0; y < ch; y++)
                        for (int x = 0; x < cw; x++)
                            dst[c][(size_t)y * dst_pitch[c] + x] = pl_at(&s, c, cx0 + x, cy0 + y);
        } else if (css == CSS_420) {
            /* interleaved UV rows, byte offset (top>>1)*pitch + left (decoder.cpp:380-388) */
            if (CH_OK(1))
                for (int y = 0; y < (H >> 1); y++)
                    for (int j = 0; j < W; j++)
                        dst[1][(size_t)y * dst_pitch[1] + j] = nv12uv_at(&s, (y0 >> 1) + y, x0 + j);
        }
        break;
    }
    case FMT_YUV_PLANAR:
    case FMT_Y: {
        if (CH_OK(0))
            for (int y = 0; y < H; y++)
                for (int x = 0; x < W; x++) dst[0][(size_t)y * dst_pitch[0] + x] = pl_at(&s, 0, x0 + x, y0 + y);
        if (fmt == FMT_Y || css == CSS_400) break;
        /* chroma at the coded resolution. 4:4:4 / 4:4:0 honour each channel's own
         * pitch (CopyChannel, decoder.cpp:600-601); 4:2:2 / 4:2:0 write U and V
         * with pitch[1] (decoder.cpp:589-590, 596-597). */
        int cw = W >> sx, ch = H >> sy, cx0 = x0 >> sx, cy0 = y0 >> sy;
        for (int c = 1; c < 3; c++) {
            uint32_t pitch = (sx == 0) ? dst_pitch[c] : dst_pitch[1];
            if (dst[c] == NULL || pitch == 0) continue;
            for (int y = 0; y < ch; y++)
                for (int x = 0; x < cw; x++)
                    dst[c][(size_t)y * pitch + x] = pl_at(&s, c, cx0 + x, cy0 + y);
        }
        break;
    }
    case FMT_RGB:
    case FMT_RGB_PLANAR: {
        if (fmt == FMT_RGB && !CH_OK(0)) break;
        if (fmt == FMT_RGB_PLANAR && (!dst[0] || !dst[1] || !dst[2] || !dst_pitch[0])) break;
        for (int y = 0; y < H; y++)
            for (int x = 0; x < W; x++) {
                int X = x0 + x, Y = y0 + y;
                uint8_t yy = pl_at(&s, 0, X, Y), r, g, b;
                if (css == CSS_400) {
                    r = g = b = yy; /* hip_kernels.cpp:1915-1927 */
                } else {
                    /* nearest-neighbour chroma: hip_kernels.cpp:76-89 (444), :585-617 (440),
                     * :947-954 (422), :1389-1429 (420) */
                    uint8_t uu = pl_at(&s, 1, X >> sx, Y >> sy), vv = pl_at(&s, 2, X >> sx, Y >> sy);
                    rgb_from_yuv(yy, uu, vv, &r, &g, &b);
                }
                if (fmt == FMT_RGB) {
                    uint8_t *p = dst[0] + (size_t)y * dst_pitch[0] + 3 * x;
                    p[0] = r; p[1] = g; p[2] = b;
                } else { /* all three planes use pitch[0] (decoder.cpp:526-544) */
                    dst[0][(size_t)y * dst_pitch[0] + x] = r;
                    dst[1][(size_t)y * dst_pitch[0] + x] = g;
                    dst[2][(size_t)y * dst_pitch[0] + x] = b;
                }
            }
        break;
    }
    default:
        return ORC_INVALID;
    }
#undef CH_OK
    return ORC_OK;
}

/* src/rocjpeg_decoder.cpp:307-358 */
int orc_image_info(const OrcInfo *o, uint8_t *ncomp, int32_t *css, uint32_t widths[4], uint32_t heights[4]) {
    *ncomp = (uint8_t)o->ncomp;
    *css = o->css;
    widths[0] = (uint32_t)o->width;
    heights[0] = (uint32_t)o->height;
    widths[3] = heights[3] = 0;
    switch (o->css) {
    case CSS_444: widths[1] = widths[2] = widths[0]; heights[1] = heights[2] = heights[0]; break;
    case CSS_440: widths[1] = widths[2] = widths[0]; heights[1] = heights[2] = heights[0] >> 1; break;
    case CSS_422: widths[1] = widths[2] = widths[0] >> 1; heights[1] = heights[2] = heights[0]; break;
    case CSS_420: widths[1] = widths[2] = widths[0] >> 1; heights[1] = heights[2] = heights[0] >> 1; break;
    case CSS_400: widths[1] = widths[2] = 0; heights[1] = heights[2] = 0; break;
    case CSS_411: widths[1] = widths[2] = widths[0] >> 2; heights[1] = heights[2] = heights[0]; break;
    default: break;
    }
    return ORC_OK;
}

/* One-call full decode used by the CPU-baseline timing leg. */
int orc_decode(const uint8_t *data, size_t len, int fmt, const int16_t crop[4], uint8_t *dst[4],
               const uint32_t dst_pitch[4]) {
    OrcInfo o;
    int rc = orc_parse(data, len, &o);
    if (rc) return rc;
    rc = orc_supported(&o);
    if (rc) return rc;
    size_t n = orc_coef_count(&o);
    int16_t *coefs = (int16_t *)malloc(n * sizeof(int16_t));
    uint8_t *planes = (uint8_t *)malloc(n);
    if (!coefs || !planes) { free(coefs); free(planes); return -5; }
    rc = orc_decode_coefficients(data, len, &o, coefs);
    if (!rc) rc = orc_idct_planes(&o, coefs, planes);
    if (!rc) rc = orc_convert(&o, planes, fmt, crop, dst, dst_pitch);
    free(coefs);
    free(planes);
    return rc;
}

size_t orc_info_size(void) { return sizeof(OrcInfo); }
