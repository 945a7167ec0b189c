/* ljt_shim.c — ORACLE side-car (test infrastructure, NOT product code).
 *
 * Header-less access to libjpeg-turbo 3.1.4.1 as bundled with Pillow
 * (pillow.libs/libjpeg-*.so.62, libjpeg v62 ABI, LP64). The image has no
 * jpeglib.h, so the handful of struct offsets needed are hand-declared; they
 * were verified in this container and are re-checked at
 * run time by ljt_selfcheck(). Used for:
 *   - pinning the oracle: jpeg_read_coefficients / jpeg_read_raw_data,
 *   - the reported CPU baseline: multithreaded default decode of the same
 *     files on the host cores.
 */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <pthread.h>
#include <setjmp.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CINFO_SIZE 632
#define OFF_ERR 0
#define OFF_MEM 8
#define OFF_IMAGE_WIDTH 48
#define OFF_IMAGE_HEIGHT 52
#define OFF_NUM_COMPONENTS 56
#define OFF_JPEG_COLOR_SPACE 60
#define OFF_OUT_COLOR_SPACE 64
#define OFF_RAW_DATA_OUT 92
#define OFF_DCT_METHOD 96
#define OFF_DO_FANCY 100
#define OFF_OUTPUT_WIDTH 136
#define OFF_OUTPUT_HEIGHT 140
#define OFF_OUTPUT_SCANLINE 168
#define OFF_QUANT_TBL_PTRS 200
#define OFF_COMP_INFO 304
#define COMP_STRIDE 96
#define COMP_H 8
#define COMP_V 12
#define COMP_TQ 16
#define COMP_WIB 28
#define COMP_HIB 32
#define MEM_ACCESS_VIRT_BARRAY 64

typedef void *(*fn_std_error)(void *);
typedef void (*fn_create)(void *, int, size_t);
typedef void (*fn_mem_src)(void *, const unsigned char *, unsigned long);
typedef int (*fn_read_header)(void *, int);
typedef int (*fn_start)(void *);
typedef unsigned (*fn_read_scanlines)(void *, uint8_t **, unsigned);
typedef unsigned (*fn_read_raw)(void *, uint8_t ***, unsigned);
typedef void **(*fn_read_coefs)(void *);
typedef int (*fn_finish)(void *);
typedef void (*fn_destroy)(void *);
typedef int16_t (**(*fn_access)(void *, void *, unsigned, unsigned, int))[64];

static struct {
    void *h;
    fn_std_error std_error;
    fn_create create;
    fn_mem_src mem_src;
    fn_read_header read_header;
    fn_start start;
    fn_read_scanlines read_scanlines;
    fn_read_raw read_raw;
    fn_read_coefs read_coefs;
    fn_finish finish;
    fn_destroy destroy, abort_;
} L;

typedef struct {
    char mgr[512];
    jmp_buf jb;
} ErrCtx;

static void on_error_exit(void *cinfo) {
    ErrCtx *e = *(ErrCtx **)((char *)cinfo + OFF_ERR);
    longjmp(e->jb, 1);
}
static void on_output_message(void *cinfo) { (void)cinfo; }
static void on_emit_message(void *cinfo, int lvl) { (void)cinfo; (void)lvl; }

int ljt_open(const char *path) {
    if (L.h) return 0;
    L.h = dlopen(path, RTLD_NOW | RTLD_LOCAL);
    if (!L.h) return -1;
#define SYM(field, name) do { *(void **)(&L.field) = dlsym(L.h, name); if (!L.field) return -2; } while (0)
    SYM(std_error, "jpeg_std_error");
    SYM(create, "jpeg_CreateDecompress");
    SYM(mem_src, "jpeg_mem_src");
    SYM(read_header, "jpeg_read_header");
    SYM(start, "jpeg_start_decompress");
    SYM(read_scanlines, "jpeg_read_scanlines");
    SYM(read_raw, "jpeg_read_raw_data");
    SYM(read_coefs, "jpeg_read_coefficients");
    SYM(finish, "jpeg_finish_decompress");
    SYM(destroy, "jpeg_destroy_decompress");
    SYM(abort_, "jpeg_abort_decompress");
#undef SYM
    return 0;
}

#define I32(ci, off) (*(int32_t *)((char *)(ci) + (off)))
#define U32(ci, off) (*(uint32_t *)((char *)(ci) + (off)))
#define PTR(ci, off) (*(void **)((char *)(ci) + (off)))

static void setup(void *ci, ErrCtx *e) {
    memset(ci, 0, CINFO_SIZE + 64);
    memset(e, 0, sizeof(*e));
    PTR(ci, OFF_ERR) = L.std_error(e->mgr);
    ((void **)e->mgr)[0] = (void *)on_error_exit;
    ((void **)e->mgr)[1] = (void *)on_emit_message;
    ((void **)e->mgr)[2] = (void *)on_output_message;
    L.create(ci, 62, CINFO_SIZE);
}

typedef struct {
    int32_t width, height, ncomp, color_space;
    int32_t hs[4], vs[4], tq[4], wib[4], hib[4];
    uint16_t quant[4][64]; /* natural order; zeros if absent */
} LjtInfo;

static void fill_info(void *ci, LjtInfo *o) {
    memset(o, 0, sizeof(*o));
    o->width = (int32_t)U32(ci, OFF_IMAGE_WIDTH);
    o->height = (int32_t)U32(ci, OFF_IMAGE_HEIGHT);
    o->ncomp = I32(ci, OFF_NUM_COMPONENTS);
    o->color_space = I32(ci, OFF_JPEG_COLOR_SPACE);
    char *comp = (char *)PTR(ci, OFF_COMP_INFO);
    for (int c = 0; c < o->ncomp && c < 4; c++) {
        char *cp = comp + c * COMP_STRIDE;
        o->hs[c] = I32(cp, COMP_H);
        o->vs[c] = I32(cp, COMP_V);
        o->tq[c] = I32(cp, COMP_TQ);
        o->wib[c] = (int32_t)U32(cp, COMP_WIB);
        o->hib[c] = (int32_t)U32(cp, COMP_HIB);
    }
    for (int t = 0; t < 4; t++) {
        uint16_t *q = (uint16_t *)((void **)((char *)ci + OFF_QUANT_TBL_PTRS))[t];
        if (q) memcpy(o->quant[t], q, 128);
    }
}

int ljt_info(const uint8_t *data, size_t len, LjtInfo *o) {
    char ci[CINFO_SIZE + 64];
    ErrCtx e;
    setup(ci, &e);
    if (setjmp(e.jb)) { L.destroy(ci); return -1; }
    L.mem_src(ci, data, (unsigned long)len);
    L.read_header(ci, 1);
    fill_info(ci, o);
    L.destroy(ci);
    return 0;
}

/* Quantised coefficients, natural order, into a component-major array whose
 * per-component grid is bw[c] x bh[c] blocks (the MCU-padded grid). */
int ljt_read_coefficients(const uint8_t *data, size_t len, int16_t *out, const int32_t bw[3], const int32_t bh[3]) {
    char ci[CINFO_SIZE + 64];
    ErrCtx e;
    setup(ci, &e);
    if (setjmp(e.jb)) { L.destroy(ci); return -1; }
    L.mem_src(ci, data, (unsigned long)len);
    L.read_header(ci, 1);
    void **arrays = L.read_coefs(ci);
    if (!arrays) { L.destroy(ci); return -2; }
    int ncomp = I32(ci, OFF_NUM_COMPONENTS);
    fn_access access = *(fn_access *)((char *)PTR(ci, OFF_MEM) + MEM_ACCESS_VIRT_BARRAY);
    size_t base = 0;
    for (int c = 0; c < ncomp && c < 3; c++) {
        for (int r = 0; r < bh[c]; r++) {
            int16_t(**rows)[64] = access(ci, arrays[c], (unsigned)r, 1, 0);
            memcpy(out + base + (size_t)r * bw[c] * 64, rows[0], (size_t)bw[c] * 128);
        }
        base += (size_t)bw[c] * bh[c] * 64;
    }
    L.finish(ci);
    L.destroy(ci);
    return 0;
}

/* Raw (un-upsampled, un-colour-converted) component planes via
 * jpeg_read_raw_data with the default islow IDCT. planes: component-major,
 * plane c is (bw[c]*8) x (bh[c]*8) bytes; only the wib*8 x hib*8 region is
 * defined by libjpeg (dummy blocks are not inverse-transformed). */
int ljt_read_raw(const uint8_t *data, size_t len, uint8_t *planes, const int32_t bw[3], const int32_t bh[3]) {
    char ci[CINFO_SIZE + 64];
    ErrCtx e;
    uint8_t **rowptr[3] = {0, 0, 0};
    setup(ci, &e);
    if (setjmp(e.jb)) {
        L.destroy(ci);
        for (int c = 0; c < 3; c++) free(rowptr[c]);
        return -1;
    }
    L.mem_src(ci, data, (unsigned long)len);
    L.read_header(ci, 1);
    I32(ci, OFF_RAW_DATA_OUT) = 1;
    I32(ci, OFF_DCT_METHOD) = 0; /* JDCT_ISLOW */
    int ncomp = I32(ci, OFF_NUM_COMPONENTS);
    L.start(ci);
    char *comp = (char *)PTR(ci, OFF_COMP_INFO);
    int vmax = 1, vs[3] = {1, 1, 1};
    uint8_t *pbase[3];
    size_t base = 0;
    for (int c = 0; c < ncomp && c < 3; c++) {
        vs[c] = I32(comp + c * COMP_STRIDE, COMP_V);
        if (vs[c] > vmax) vmax = vs[c];
        pbase[c] = planes + base;
        base += (size_t)bw[c] * bh[c] * 64;
        rowptr[c] = (uint8_t **)malloc(sizeof(uint8_t *) * 8 * 4);
    }
    unsigned out_h = U32(ci, OFF_OUTPUT_HEIGHT);
    int imcu = 0;
    while (U32(ci, OFF_OUTPUT_SCANLINE) < out_h) {
        for (int c = 0; c < ncomp && c < 3; c++)
            for (int r = 0; r < vs[c] * 8; r++) {
                size_t row = (size_t)imcu * vs[c] * 8 + r;
                rowptr[c][r] = pbase[c] + row * (size_t)bw[c] * 8;
            }
        unsigned got = L.read_raw(ci, rowptr, (unsigned)(vmax * 8));
        if (!got) break;
        imcu++;
    }
    L.finish(ci);
    L.destroy(ci);
    for (int c = 0; c < 3; c++) free(rowptr[c]);
    return 0;
}

/* Default libjpeg-turbo decode (islow, fancy upsampling, JFIF colour).
 * mode 0: RGB interleaved into out (pitch bytes/row);
 * mode 1: grayscale/Y (out_color_space = JCS_GRAYSCALE);
 * mode 2: raw planes (as ljt_read_raw, out = scratch of sum(bw*bh*64)). */
static int decode_one(const uint8_t *data, size_t len, uint8_t *out, size_t pitch, int mode) {
    char ci[CINFO_SIZE + 64];
    ErrCtx e;
    setup(ci, &e);
    if (setjmp(e.jb)) { L.destroy(ci); return -1; }
    L.mem_src(ci, data, (unsigned long)len);
    L.read_header(ci, 1);
    if (mode == 2) {
        LjtInfo inf;
        fill_info(ci, &inf);
        L.destroy(ci);
        int32_t bw[3] = {0, 0, 0}, bh[3] = {0, 0, 0};
        int hmax = 1, vmax = 1;
        for (int c = 0; c < inf.ncomp && c < 3; c++) {
            if (inf.hs[c] > hmax) hmax = inf.hs[c];
            if (inf.vs[c] > vmax) vmax = inf.vs[c];
        }
        int mx = (inf.width + 8 * hmax - 1) / (8 * hmax), my = (inf.height + 8 * vmax - 1) / (8 * vmax);
        for (int c = 0; c < inf.ncomp && c < 3; c++) {
            bw[c] = (inf.ncomp == 1) ? (inf.width + 7) / 8 : mx * inf.hs[c];
            bh[c] = (inf.ncomp == 1) ? (inf.height + 7) / 8 : my * inf.vs[c];
        }
        return ljt_read_raw(data, len, out, bw, bh);
    }
    I32(ci, OFF_OUT_COLOR_SPACE) = (mode == 1) ? 1 /*JCS_GRAYSCALE*/ : 2 /*JCS_RGB*/;
    L.start(ci);
    unsigned out_h = U32(ci, OFF_OUTPUT_HEIGHT);
    while (U32(ci, OFF_OUTPUT_SCANLINE) < out_h) {
        uint8_t *rows[4];
        unsigned sl = U32(ci, OFF_OUTPUT_SCANLINE);
        for (int i = 0; i < 4; i++) rows[i] = out + (size_t)(sl + i < out_h ? sl + i : out_h - 1) * pitch;
        unsigned want = (out_h - sl) < 4 ? (out_h - sl) : 4;
        if (!L.read_scanlines(ci, rows, want)) break;
    }
    L.finish(ci);
    L.destroy(ci);
    return 0;
}

int ljt_decode(const uint8_t *data, size_t len, uint8_t *out, size_t pitch, int mode) {
    return decode_one(data, len, out, pitch, mode);
}

typedef struct {
    int n, mode, tid, nthreads;
    const uint8_t *const *datas;
    const size_t *lens;
    uint8_t *const *outs;
    const size_t *pitches;
    int rc;
} Job;

static void *worker(void *arg) {
    Job *j = (Job *)arg;
    for (int i = j->tid; i < j->n; i += j->nthreads) {
        int rc = decode_one(j->datas[i], j->lens[i], j->outs[i], j->pitches[i], j->mode);
        if (rc) j->rc = rc;
    }
    return NULL;
}

/* Decode n images with `nthreads` host threads (round-robin assignment). */
int ljt_decode_batch(int n, const uint8_t *const *datas, const size_t *lens, uint8_t *const *outs,
                     const size_t *pitches, int mode, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    if (nthreads > n) nthreads = n;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    Job *jobs = (Job *)malloc(sizeof(Job) * nthreads);
    int rc = 0;
    for (int t = 0; t < nthreads; t++) {
        jobs[t] = (Job){n, mode, t, nthreads, datas, lens, outs, pitches, 0};
        pthread_create(&th[t], NULL, worker, &jobs[t]);
    }
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], NULL);
        if (jobs[t].rc) rc = jobs[t].rc;
    }
    free(th);
    free(jobs);
    return rc;
}

size_t ljt_info_size(void) { return sizeof(LjtInfo); }
