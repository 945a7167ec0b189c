"""Output assembly through the REFERENCE's own kernels (oracle/_ref, CPU build).

Builds the VCN surface the reference would have received for a given chroma
subsampling (layouts: src/rocjpeg_vaapi_decoder.cpp:612-637) from decoded
component planes, then drives the reference's launchers exactly as
src/rocjpeg_decoder.cpp does (:450-494 RGB, :511-557 RGB planar, :576-605 YUV
planar, :620-636 Y), including its ROI pointer offsets. The result is what the
reference would have written, given those planes — the pin for the oracle's
orc_convert and, transitively, for the CUDA colour/layout kernels.

The reference kernels have no tail guards (8 px x 2 rows per thread), so every
buffer here is over-allocated and only the valid region is returned.
"""
import ctypes as C

import numpy as np

from oracle import CSS, RefKernels

PAD = 64


def _align(v, a):
    return (v + a - 1) // a * a


class Surface:
    """Pitched VCN-style surface in one flat allocation (hip_mapped_device_mem)."""

    def __init__(self, info, planes):
        css = CSS[info.css]
        W, H = info.width, info.height
        self.css = css
        pw = _align(max(p.shape[1] for p in planes), 256) + 256
        rows = _align(max(p.shape[0] for p in planes), 16) + 32
        self.pitch = [0, 0, 0]
        self.offset = [0, 0, 0]
        if css == "422":  # packed YUYV, one plane
            pitch = 2 * pw
            buf = np.zeros((rows, pitch), dtype=np.uint8)
            y, u, v = planes
            h = y.shape[0]
            buf[:h, 0:2 * y.shape[1]:2] = y
            buf[:h, 1:4 * u.shape[1]:4] = u[:h]
            buf[:h, 3:4 * v.shape[1]:4] = v[:h]
            self.mem = buf.reshape(-1)
            self.pitch[0] = pitch
        elif css == "420":  # NV12
            y, u, v = planes
            pitch = pw
            ybuf = np.zeros((rows, pitch), dtype=np.uint8)
            ybuf[:y.shape[0], :y.shape[1]] = y
            cbuf = np.zeros((rows, pitch), dtype=np.uint8)
            cbuf[:u.shape[0], 0:2 * u.shape[1]:2] = u
            cbuf[:v.shape[0], 1:2 * v.shape[1]:2] = v
            self.mem = np.concatenate([ybuf.reshape(-1), cbuf.reshape(-1)])
            self.pitch[0] = self.pitch[1] = pitch
            self.offset[1] = ybuf.size
        else:  # 444P, 422V (4:4:0), Y800: planar, same pitch for all planes
            pitch = pw
            bufs = []
            for i, p in enumerate(planes):
                b = np.zeros((rows, pitch), dtype=np.uint8)
                b[:p.shape[0], :p.shape[1]] = p
                self.offset[i] = i * b.size
                self.pitch[i] = pitch
                bufs.append(b.reshape(-1))
            self.mem = np.concatenate(bufs)
        self.mem = np.concatenate([self.mem, np.zeros(pw * 64, dtype=np.uint8)])
        self.W, self.H = W, H

    def ptr(self, off=0):
        return C.c_void_p(self.mem.ctypes.data + int(off))


def _roi(info, crop):
    """src/rocjpeg_decoder.cpp:126-134"""
    rw = (int(crop[2]) - int(crop[0])) & 0xFFFFFFFF
    rh = (int(crop[3]) - int(crop[1])) & 0xFFFFFFFF
    valid = rw > 0 and rh > 0 and rw <= info.width and rh <= info.height
    return valid, (rw if valid else info.width), (rh if valid else info.height)


def _out(rows, rowbytes):
    pitch = _align(rowbytes, 64) + 256
    return np.full((rows + 34, pitch), 0xCD, dtype=np.uint8), pitch


def reference_output(rk: RefKernels, info, planes, fmt, crop=(0, 0, 0, 0)):
    """Returns a list of (valid-region array) per channel, as the reference would write."""
    L = rk.lib
    s = Surface(info, planes)
    css = s.css
    valid, W, H = _roi(info, crop)
    left, top = (int(crop[0]), int(crop[1])) if valid else (0, 0)
    vp = lambda a: C.c_void_p(a.ctypes.data)

    def copy_channel(ch, height, rowbytes):
        """CopyChannel (decoder.cpp:372-399) restated: 2-D copy of `rowbytes` per row."""
        t, l = top, left
        if valid:
            if css in ("420", "440") and ch in (1, 2):
                t = top >> 1
            if css == "422":
                l = left * 2
        off = s.offset[ch] + (t * s.pitch[ch] + l if valid else 0)
        src = s.mem[off:off + height * s.pitch[ch]].reshape(height, s.pitch[ch]) if height else np.zeros((0, 0), np.uint8)
        return src[:, :rowbytes].copy()

    if fmt in ("rgb", "rgb_planar"):
        roi_off = roi_uv = 0
        l = left
        if valid:
            if css in ("440", "420"):
                roi_uv = (top >> 1) * s.pitch[1] + left
            elif css == "422":
                l = left * 2
            roi_off = top * s.pitch[0] + l
        if fmt == "rgb":
            dst, dp = _out(H, 3 * W)
            if css == "444":
                L.ref_yuv444_to_rgb(W, H, vp(dst), dp, s.ptr(roi_off), s.pitch[0], s.offset[1] + roi_off, s.offset[2] + roi_off)
            elif css == "440":  # chroma ROI offset commented out in the reference (decoder.cpp:470)
                L.ref_yuv440_to_rgb(W, H, vp(dst), dp, s.ptr(roi_off), s.pitch[0], s.offset[1], s.offset[2])
            elif css == "422":
                L.ref_yuyv_to_rgb(W, H, vp(dst), dp, s.ptr(roi_off), s.pitch[0])
            elif css == "420":
                L.ref_nv12_to_rgb(W, H, vp(dst), dp, s.ptr(roi_off), s.pitch[0], s.ptr(s.offset[1] + roi_uv), s.pitch[1])
            elif css == "400":
                L.ref_yuv400_to_rgb(W, H, vp(dst), dp, s.ptr(roi_off), s.pitch[0])
            return [dst[:H, :3 * W].copy()]
        r, dp = _out(H, W)
        g, _ = _out(H, W)
        b, _ = _out(H, W)
        if css == "444":
            L.ref_yuv444_to_rgb_planar(W, H, vp(r), vp(g), vp(b), dp, s.ptr(roi_off), s.pitch[0], s.offset[1] + roi_off, s.offset[2] + roi_off)
        elif css == "440":
            L.ref_yuv440_to_rgb_planar(W, H, vp(r), vp(g), vp(b), dp, s.ptr(roi_off), s.pitch[0], s.offset[1], s.offset[2])
        elif css == "422":
            L.ref_yuyv_to_rgb_planar(W, H, vp(r), vp(g), vp(b), dp, s.ptr(roi_off), s.pitch[0])
        elif css == "420":
            L.ref_nv12_to_rgb_planar(W, H, vp(r), vp(g), vp(b), dp, s.ptr(roi_off), s.pitch[0], s.ptr(s.offset[1] + roi_uv), s.pitch[1])
        elif css == "400":
            L.ref_yuv400_to_rgb_planar(W, H, vp(r), vp(g), vp(b), dp, s.ptr(roi_off), s.pitch[0])
        return [x[:H, :W].copy() for x in (r, g, b)]

    if fmt == "y":
        if css == "422":
            off = top * s.pitch[0] + left * 2 if valid else 0
            y, dp = _out(H, W)
            L.ref_yuyv_extract_y(W, H, vp(y), dp, s.ptr(off), s.pitch[0])
            return [y[:H, :W].copy()]
        return [copy_channel(0, H, W)]

    if fmt == "yuv_planar":
        if css == "400":
            return [copy_channel(0, H, W)]
        if css == "422":
            off = top * s.pitch[0] + left * 2 if valid else 0
            y, yp = _out(H, W)
            u, cp = _out(H, W)
            v, _ = _out(H, W)
            L.ref_yuyv_to_planar(W, H, vp(y), vp(u), vp(v), yp, cp, s.ptr(off), s.pitch[0])
            return [y[:H, :W].copy(), u[:H, :W >> 1].copy(), v[:H, :W >> 1].copy()]
        if css == "420":
            off = (top >> 1) * s.pitch[1] + left if valid else 0
            u, cp = _out(H >> 1, W)
            v, _ = _out(H >> 1, W)
            L.ref_uv_to_planar(W >> 1, H >> 1, vp(u), vp(v), cp, s.ptr(s.offset[1] + off), s.pitch[1])
            return [copy_channel(0, H, W), u[:H >> 1, :W >> 1].copy(), v[:H >> 1, :W >> 1].copy()]
        ch = H >> 1 if css == "440" else H
        return [copy_channel(0, H, W), copy_channel(1, ch, W), copy_channel(2, ch, W)]

    if fmt == "native":
        if css == "400":
            return [copy_channel(0, H, W)]
        if css == "422":
            return [copy_channel(0, H, 2 * W)]
        if css == "420":
            return [copy_channel(0, H, W), copy_channel(1, H >> 1, W)]
        ch = H >> 1 if css == "440" else H
        return [copy_channel(0, H, W), copy_channel(1, ch, W), copy_channel(2, ch, W)]
    raise ValueError(fmt)
