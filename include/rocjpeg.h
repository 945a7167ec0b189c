/* rocjpeg.h — the drop-in C ABI of the B200-native JPEG decoder.
 *
 * This header declares, with identical names, values, struct layouts and
 * signatures, the public API of the reference (api/rocjpeg.h:46-343), so the
 * reference's jpegDecode / jpegDecodeBatched / jpegDecodePerf samples compile
 * and link against librocjpeg.so from this repository without modification.
 * Only the ABI is shared; the prose and the implementation behind it are new.
 *
 * Each declaration cites the reference line it replaces.
 */
#ifndef ROC_JPEG_H
#define ROC_JPEG_H

#define ROCJPEGAPI

#pragma once
#include "hip/hip_runtime.h"   /* api/rocjpeg.h:28 — resolved to this repo's CUDA shim */
#include "rocjpeg_version.h"

#if defined(__cplusplus)
extern "C" {
#endif

/* api/rocjpeg.h:46 */
#define ROCJPEG_MAX_COMPONENT 4

/* api/rocjpeg.h:53-67 — status codes returned by every entry point. */
typedef enum {
    ROCJPEG_STATUS_SUCCESS = 0,
    ROCJPEG_STATUS_NOT_INITIALIZED = -1,
    ROCJPEG_STATUS_INVALID_PARAMETER = -2,
    ROCJPEG_STATUS_BAD_JPEG = -3,
    ROCJPEG_STATUS_JPEG_NOT_SUPPORTED = -4,
    ROCJPEG_STATUS_OUTOF_MEMORY = -5,
    ROCJPEG_STATUS_EXECUTION_FAILED = -6,
    ROCJPEG_STATUS_ARCH_MISMATCH = -7,
    ROCJPEG_STATUS_INTERNAL_ERROR = -8,
    ROCJPEG_STATUS_IMPLEMENTATION_NOT_SUPPORTED = -9,
    ROCJPEG_STATUS_HW_JPEG_DECODER_NOT_SUPPORTED = -10,
    ROCJPEG_STATUS_RUNTIME_ERROR = -11,
    ROCJPEG_STATUS_NOT_IMPLEMENTED = -12,
} RocJpegStatus;

/* api/rocjpeg.h:86-94 — chroma layout of the coded image. */
typedef enum {
    ROCJPEG_CSS_444 = 0,
    ROCJPEG_CSS_440 = 1,
    ROCJPEG_CSS_422 = 2,
    ROCJPEG_CSS_420 = 3,
    ROCJPEG_CSS_411 = 4,
    ROCJPEG_CSS_400 = 5,
    ROCJPEG_CSS_UNKNOWN = -1
} RocJpegChromaSubsampling;

/* api/rocjpeg.h:104-107 — caller-owned DEVICE buffers; pitch is in bytes. */
typedef struct {
    uint8_t* channel[ROCJPEG_MAX_COMPONENT];
    uint32_t pitch[ROCJPEG_MAX_COMPONENT];
} RocJpegImage;

/* api/rocjpeg.h:124-141 — what rocJpegDecode writes into RocJpegImage.
 *   NATIVE      444/440: Y,U,V in channel 0,1,2; 422: packed YUYV in channel 0;
 *               420: Y in channel 0, interleaved UV in channel 1; 400: Y.
 *   YUV_PLANAR  Y,U,V in channel 0,1,2 at the coded chroma resolution.
 *   Y           luma only, channel 0.
 *   RGB         interleaved R,G,B bytes in channel 0.
 *   RGB_PLANAR  R,G,B planes in channel 0,1,2 (all with pitch[0]). */
typedef enum {
    ROCJPEG_OUTPUT_NATIVE = 0,
    ROCJPEG_OUTPUT_YUV_PLANAR = 1,
    ROCJPEG_OUTPUT_Y = 2,
    ROCJPEG_OUTPUT_RGB = 3,
    ROCJPEG_OUTPUT_RGB_PLANAR = 4,
    ROCJPEG_OUTPUT_FORMAT_MAX = 5
} RocJpegOutputFormat;

/* api/rocjpeg.h:153-166 — one set of parameters per decode call (shared by a
 * whole batch). A crop rectangle with positive extent no larger than the
 * picture selects a region of interest; target_dimension is unused, as in
 * the reference. */
typedef struct {
    RocJpegOutputFormat output_format;
    struct {
        int16_t left;
        int16_t top;
        int16_t right;
        int16_t bottom;
    } crop_rectangle;
    struct {
        uint32_t width;
        uint32_t height;
    } target_dimension;
} RocJpegDecodeParams;

/* api/rocjpeg.h:176-179 — HARDWARE (0, the samples' default) selects the
 * B200 CUDA pipeline; HYBRID returns NOT_IMPLEMENTED as in the reference. */
typedef enum {
    ROCJPEG_BACKEND_HARDWARE = 0,
    ROCJPEG_BACKEND_HYBRID = 1
} RocJpegBackend;

/* api/rocjpeg.h:187 — opaque parsed-stream handle. */
typedef void* RocJpegStreamHandle;

/* api/rocjpeg.h:204 */
RocJpegStatus ROCJPEGAPI rocJpegStreamCreate(RocJpegStreamHandle *jpeg_stream_handle);
/* api/rocjpeg.h:219 — host-side marker parse of one baseline JPEG. */
RocJpegStatus ROCJPEGAPI rocJpegStreamParse(const unsigned char *data, size_t length, RocJpegStreamHandle jpeg_stream_handle);
/* api/rocjpeg.h:234 */
RocJpegStatus ROCJPEGAPI rocJpegStreamDestroy(RocJpegStreamHandle jpeg_stream_handle);

/* api/rocjpeg.h:242 — opaque decoder handle (bound to one device). */
typedef void *RocJpegHandle;

/* api/rocjpeg.h:258 */
RocJpegStatus ROCJPEGAPI rocJpegCreate(RocJpegBackend backend, int device_id, RocJpegHandle *handle);
/* api/rocjpeg.h:273 */
RocJpegStatus ROCJPEGAPI rocJpegDestroy(RocJpegHandle handle);
/* api/rocjpeg.h:296 — widths/heights are 4-entry arrays. */
RocJpegStatus ROCJPEGAPI rocJpegGetImageInfo(RocJpegHandle handle, RocJpegStreamHandle jpeg_stream_handle, uint8_t *num_components, RocJpegChromaSubsampling *subsampling, uint32_t *widths, uint32_t *heights);
/* api/rocjpeg.h:314 — synchronous: pixels are in `destination` on return. */
RocJpegStatus ROCJPEGAPI rocJpegDecode(RocJpegHandle handle, RocJpegStreamHandle jpeg_stream_handle, const RocJpegDecodeParams *decode_params, RocJpegImage *destination);
/* api/rocjpeg.h:331 */
RocJpegStatus ROCJPEGAPI rocJpegDecodeBatched(RocJpegHandle handle, RocJpegStreamHandle *jpeg_stream_handles, int batch_size, const RocJpegDecodeParams *decode_params, RocJpegImage *destinations);
/* api/rocjpeg.h:343 */
extern const char* ROCJPEGAPI rocJpegGetErrorName(RocJpegStatus rocjpeg_status);

#if defined(__cplusplus)
}
#endif

#endif /* ROC_JPEG_H */
