/* C entry points over the reference's HIP launchers compiled for the CPU
 * (see oracle/build.py and oracle/ref_hip_emul/). ORACLE infrastructure only.
 * NOTE: the reference kernels process 8 pixels x 2 rows per thread with no
 * tail guard (src/rocjpeg_hip_kernels.cpp:59-221), so callers must hand them
 * padded buffers and compare only the valid region. */
#include "rocjpeg_hip_kernels.h"

extern "C" {
void ref_yuv444_to_rgb(uint32_t w, uint32_t h, uint8_t *dst, uint32_t dst_pitch, const uint8_t *src, uint32_t src_pitch,
                       uint32_t u_off, uint32_t v_off) {
    ColorConvertYUV444ToRGB(nullptr, w, h, dst, dst_pitch, src, src_pitch, u_off, v_off);
}
void ref_yuv440_to_rgb(uint32_t w, uint32_t h, uint8_t *dst, uint32_t dst_pitch, const uint8_t *src, uint32_t src_pitch,
                       uint32_t u_off, uint32_t v_off) {
    ColorConvertYUV440ToRGB(nullptr, w, h, dst, dst_pitch, src, src_pitch, u_off, v_off);
}
void ref_yuyv_to_rgb(uint32_t w, uint32_t h, uint8_t *dst, uint32_t dst_pitch, const uint8_t *src, uint32_t src_pitch) {
    ColorConvertYUYVToRGB(nullptr, w, h, dst, dst_pitch, src, src_pitch);
}
void ref_nv12_to_rgb(uint32_t w, uint32_t h, uint8_t *dst, uint32_t dst_pitch, const uint8_t *luma, uint32_t luma_pitch,
                     const uint8_t *chroma, uint32_t chroma_pitch) {
    ColorConvertNV12ToRGB(nullptr, w, h, dst, dst_pitch, luma, luma_pitch, chroma, chroma_pitch);
}
void ref_yuv400_to_rgb(uint32_t w, uint32_t h, uint8_t *dst, uint32_t dst_pitch, const uint8_t *luma, uint32_t luma_pitch) {
    ColorConvertYUV400ToRGB(nullptr, w, h, dst, dst_pitch, luma, luma_pitch);
}
void ref_yuv444_to_rgb_planar(uint32_t w, uint32_t h, uint8_t *r, uint8_t *g, uint8_t *b, uint32_t dst_pitch,
                              const uint8_t *src, uint32_t src_pitch, uint32_t u_off, uint32_t v_off) {
    ColorConvertYUV444ToRGBPlanar(nullptr, w, h, r, g, b, dst_pitch, src, src_pitch, u_off, v_off);
}
void ref_yuv440_to_rgb_planar(uint32_t w, uint32_t h, uint8_t *r, uint8_t *g, uint8_t *b, uint32_t dst_pitch,
                              const uint8_t *src, uint32_t src_pitch, uint32_t u_off, uint32_t v_off) {
    ColorConvertYUV440ToRGBPlanar(nullptr, w, h, r, g, b, dst_pitch, src, src_pitch, u_off, v_off);
}
void ref_yuyv_to_rgb_planar(uint32_t w, uint32_t h, uint8_t *r, uint8_t *g, uint8_t *b, uint32_t dst_pitch,
                            const uint8_t *src, uint32_t src_pitch) {
    ColorConvertYUYVToRGBPlanar(nullptr, w, h, r, g, b, dst_pitch, src, src_pitch);
}
void ref_nv12_to_rgb_planar(uint32_t w, uint32_t h, uint8_t *r, uint8_t *g, uint8_t *b, uint32_t dst_pitch,
                            const uint8_t *luma, uint32_t luma_pitch, const uint8_t *chroma, uint32_t chroma_pitch) {
    ColorConvertNV12ToRGBPlanar(nullptr, w, h, r, g, b, dst_pitch, luma, luma_pitch, chroma, chroma_pitch);
}
void ref_yuv400_to_rgb_planar(uint32_t w, uint32_t h, uint8_t *r, uint8_t *g, uint8_t *b, uint32_t dst_pitch,
                              const uint8_t *luma, uint32_t luma_pitch) {
    ColorConvertYUV400ToRGBPlanar(nullptr, w, h, r, g, b, dst_pitch, luma, luma_pitch);
}
void ref_uv_to_planar(uint32_t w, uint32_t h, uint8_t *u, uint8_t *v, uint32_t dst_pitch, const uint8_t *src,
                      uint32_t src_pitch) {
    ConvertInterleavedUVToPlanarUV(nullptr, w, h, u, v, dst_pitch, src, src_pitch);
}
void ref_yuyv_extract_y(uint32_t w, uint32_t h, uint8_t *y, uint32_t dst_pitch, const uint8_t *src, uint32_t src_pitch) {
    ExtractYFromPackedYUYV(nullptr, w, h, y, dst_pitch, src, src_pitch);
}
void ref_yuyv_to_planar(uint32_t w, uint32_t h, uint8_t *y, uint8_t *u, uint8_t *v, uint32_t luma_pitch,
                        uint32_t chroma_pitch, const uint8_t *src, uint32_t src_pitch) {
    ConvertPackedYUYVToPlanarYUV(nullptr, w, h, y, u, v, luma_pitch, chroma_pitch, src, src_pitch);
}
}
