"""CPU oracle for the rocJPEG decode path — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package. The product (rocjpeg_b200 +
librocjpeg.so) never does, and fails loudly without its CUDA library.

Three layers:
  * `Oracle`      — our scalar C restatement (oracle/jpeg_oracle.c): parser,
                    Huffman, dequant + islow IDCT, reference output assembly.
  * `LibJpegTurbo`— header-less harness over Pillow's bundled libjpeg-turbo
                    (oracle/ljt_shim.c): pins coefficients / raw planes, and is
                    the reported multithreaded CPU baseline.
  * `RefParser`, `RefKernels` — the reference's own parser and HIP kernels
                    compiled for the CPU (oracle/_ref, built by oracle/build.py
                    from /root/reference where it lies): pins parser decisions
                    and the colour/upsample/layout arithmetic.
Parity status: entropy decode + IDCT pinned to libjpeg-turbo 3.1.4.1; parser
and colour kernels pinned to the reference's own code; the float->u8 pack
rounding (hipPack / v_cvt_pk_u8_f32) is a documented convention (RNE+sat),
unpinned by the reference.
"""
from __future__ import annotations

import ctypes as C
import glob
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CSS = {0: "444", 1: "440", 2: "422", 3: "420", 4: "411", 5: "400", -1: "unknown"}
FMT = {"native": 0, "yuv_planar": 1, "y": 2, "rgb": 3, "rgb_planar": 4}


class OrcInfo(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("ncomp", C.c_int32), ("css", C.c_int32),
        ("comp_id", C.c_int32 * 3), ("hs", C.c_int32 * 3), ("vs", C.c_int32 * 3), ("tq", C.c_int32 * 3),
        ("scan_ncomp", C.c_int32), ("td", C.c_int32 * 3), ("ta", C.c_int32 * 3),
        ("hmax", C.c_int32), ("vmax", C.c_int32), ("mcus_x", C.c_int32), ("mcus_y", C.c_int32),
        ("blocks_per_mcu", C.c_int32),
        ("blocks_w", C.c_int32 * 3), ("blocks_h", C.c_int32 * 3),
        ("restart_interval", C.c_int32),
        ("num_mcus_ref", C.c_uint32), ("scan_offset", C.c_uint32), ("scan_size", C.c_uint32),
        ("qt", (C.c_uint8 * 64) * 4), ("qt_present", C.c_uint8 * 4),
        ("dc_bits", (C.c_uint8 * 16) * 4), ("dc_vals", (C.c_uint8 * 12) * 4),
        ("ac_bits", (C.c_uint8 * 16) * 4), ("ac_vals", (C.c_uint8 * 162) * 4),
        ("dc_present", C.c_uint8 * 4), ("ac_present", C.c_uint8 * 4),
        ("n_restart_markers", C.c_uint32),
        ("qt16", (C.c_uint16 * 64) * 4), ("features", C.c_int32),
    ]


def _ensure_built():
    from . import build as _b

    _b.build()


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


class Oracle:
    """Scalar C restatement of the whole path (parse -> coefficients -> planes -> output)."""

    def __init__(self):
        _ensure_built()
        self.lib = C.CDLL(os.path.join(HERE, "_build", "liboracle.so"))
        L = self.lib
        L.orc_info_size.restype = C.c_size_t
        assert L.orc_info_size() == C.sizeof(OrcInfo), (L.orc_info_size(), C.sizeof(OrcInfo))
        L.orc_parse.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(OrcInfo)]
        L.orc_supported.argtypes = [C.POINTER(OrcInfo)]
        L.orc_coef_count.argtypes = [C.POINTER(OrcInfo)]
        L.orc_coef_count.restype = C.c_size_t
        L.orc_decode_coefficients.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(OrcInfo), C.c_void_p]
        L.orc_idct_planes.argtypes = [C.POINTER(OrcInfo), C.c_void_p, C.c_void_p]
        L.orc_idct_islow_block.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_convert.argtypes = [C.POINTER(OrcInfo), C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_roi.argtypes = [C.POINTER(OrcInfo), C.c_void_p] + [C.POINTER(C.c_int)] * 4
        L.orc_zigzag_table.restype = C.POINTER(C.c_uint8)
        self.zigzag = np.array([L.orc_zigzag_table()[i] for i in range(64)], dtype=np.int64)

    def parse(self, data: bytes):
        info = OrcInfo()
        rc = self.lib.orc_parse(data, len(data), C.byref(info))
        return rc, info

    def supported(self, info) -> int:
        return self.lib.orc_supported(C.byref(info))

    def coefficients(self, data: bytes, info=None) -> list[np.ndarray]:
        """Per component: int16 array (blocks_h, blocks_w, 64), natural order."""
        if info is None:
            rc, info = self.parse(data)
            assert rc == 0, rc
        n = self.lib.orc_coef_count(C.byref(info))
        flat = np.zeros(n, dtype=np.int16)
        rc = self.lib.orc_decode_coefficients(data, len(data), C.byref(info), flat.ctypes.data)
        if rc != 0:
            raise ValueError(f"oracle decode failed rc={rc}")
        return self.split(info, flat, 64)

    @staticmethod
    def split(info, flat, per_block):
        out, base = [], 0
        for c in range(info.ncomp):
            bw, bh = info.blocks_w[c], info.blocks_h[c]
            n = bw * bh * per_block
            out.append(flat[base:base + n].reshape(bh, bw, per_block))
            base += n
        return out

    def planes(self, data: bytes, info=None) -> list[np.ndarray]:
        """Per component: uint8 plane (blocks_h*8, blocks_w*8) — MCU-padded."""
        if info is None:
            rc, info = self.parse(data)
            assert rc == 0, rc
        coefs = np.concatenate([c.reshape(-1) for c in self.coefficients(data, info)])
        flat = np.zeros(coefs.size, dtype=np.uint8)
        self.lib.orc_idct_planes(C.byref(info), coefs.ctypes.data, flat.ctypes.data)
        out, base = [], 0
        for c in range(info.ncomp):
            bw, bh = info.blocks_w[c], info.blocks_h[c]
            out.append(flat[base:base + bw * bh * 64].reshape(bh * 8, bw * 8))
            base += bw * bh * 64
        return out

    def idct_block(self, coef_nat: np.ndarray, q_nat: np.ndarray) -> np.ndarray:
        coef = np.ascontiguousarray(coef_nat, dtype=np.int16).reshape(64)
        q = np.ascontiguousarray(q_nat, dtype=np.uint16).reshape(64)
        out = np.zeros((8, 8), dtype=np.uint8)
        self.lib.orc_idct_islow_block(coef.ctypes.data, q.ctypes.data, out.ctypes.data, 8)
        return out

    def roi(self, info, crop):
        c = np.array(crop, dtype=np.int16)
        x0, y0, w, h = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        rc = self.lib.orc_roi(C.byref(info), c.ctypes.data, C.byref(x0), C.byref(y0), C.byref(w), C.byref(h))
        return rc, x0.value, y0.value, w.value, h.value

    def convert(self, info, planes, fmt: str, crop, dst: list, pitches: list) -> int:
        flat = np.concatenate([p.reshape(-1) for p in planes])
        c = np.array(crop, dtype=np.int16)
        ptrs = (C.c_void_p * 4)(*[d.ctypes.data if d is not None else None for d in dst] + [None] * (4 - len(dst)))
        pit = (C.c_uint32 * 4)(*(list(pitches) + [0] * (4 - len(pitches))))
        return self.lib.orc_convert(C.byref(info), flat.ctypes.data, FMT[fmt], c.ctypes.data, ptrs, pit)

    def decode(self, data: bytes, fmt: str, crop=(0, 0, 0, 0), pitches=None, fill=0xCD):
        """Full decode. Returns (info, [channel arrays shaped (rows, pitch)])."""
        rc, info = self.parse(data)
        if rc != 0:
            raise ValueError(f"parse rc={rc}")
        rc = self.supported(info)
        if rc != 0:
            raise ValueError(f"unsupported rc={rc}")
        planes = self.planes(data, info)
        shapes = output_shapes(info, fmt, crop, self)
        dst, pit = [], []
        for i, (rows, rowbytes) in enumerate(shapes):
            p = rowbytes if pitches is None else pitches[i]
            pit.append(p)
            dst.append(np.full((rows, p), fill, dtype=np.uint8) if rows and p else None)
        rc = self.convert(info, planes, fmt, crop, dst, pit)
        if rc != 0:
            raise ValueError(f"convert rc={rc}")
        return info, dst


def output_shapes(info, fmt: str, crop=(0, 0, 0, 0), oracle: Oracle | None = None):
    """(rows, valid bytes per row) of every output channel, mirroring the samples'
    sizing rules (samples/rocjpeg_samples_utils.h:318-399)."""
    W, H = info.width, info.height
    if oracle is not None:
        rc, _, _, w, h = oracle.roi(info, crop)
        if rc >= 0:
            W, H = w, h
    css = CSS[info.css]
    if fmt == "rgb":
        return [(H, 3 * W)]
    if fmt == "rgb_planar":
        return [(H, W)] * 3
    if fmt == "y" or css == "400":
        return [(H, W)]
    if fmt == "native":
        if css == "444":
            return [(H, W)] * 3
        if css == "440":
            return [(H, W), (H >> 1, W), (H >> 1, W)]
        if css == "422":
            return [(H, 2 * W)]
        if css == "420":
            return [(H, W), (H >> 1, W)]
        if css == "411":
            return [(H, W), (H, W >> 2), (H, W >> 2)]
    if fmt == "yuv_planar":
        if css == "444":
            return [(H, W)] * 3
        if css == "440":
            return [(H, W), (H >> 1, W), (H >> 1, W)]
        if css == "422":
            return [(H, W), (H, W >> 1), (H, W >> 1)]
        if css == "420":
            return [(H, W), (H >> 1, W >> 1), (H >> 1, W >> 1)]
        if css == "411":
            return [(H, W), (H, W >> 2), (H, W >> 2)]
    raise ValueError((fmt, css))


class LjtInfo(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("ncomp", C.c_int32), ("color_space", C.c_int32),
        ("hs", C.c_int32 * 4), ("vs", C.c_int32 * 4), ("tq", C.c_int32 * 4),
        ("wib", C.c_int32 * 4), ("hib", C.c_int32 * 4),
        ("quant", (C.c_uint16 * 64) * 4),
    ]


def find_libjpeg_turbo() -> str:
    import PIL

    cands = glob.glob(os.path.join(os.path.dirname(PIL.__file__), "..", "pillow.libs", "libjpeg*.so*"))
    if not cands:
        raise RuntimeError("Pillow's bundled libjpeg-turbo not found")
    return os.path.realpath(cands[0])


class LibJpegTurbo:
    def __init__(self):
        _ensure_built()
        self.lib = C.CDLL(os.path.join(HERE, "_build", "libljt.so"))
        self.path = find_libjpeg_turbo()
        rc = self.lib.ljt_open(self.path.encode())
        if rc != 0:
            raise RuntimeError(f"ljt_open({self.path}) rc={rc}")
        L = self.lib
        L.ljt_info_size.restype = C.c_size_t
        assert L.ljt_info_size() == C.sizeof(LjtInfo)
        L.ljt_info.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(LjtInfo)]
        L.ljt_read_coefficients.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ljt_read_raw.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ljt_decode.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
        L.ljt_decode_batch.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]

    def info(self, data: bytes) -> LjtInfo:
        o = LjtInfo()
        rc = self.lib.ljt_info(data, len(data), C.byref(o))
        if rc:
            raise ValueError("libjpeg-turbo rejected the stream")
        return o

    def coefficients(self, data: bytes, oinfo) -> list[np.ndarray]:
        bw = (C.c_int32 * 3)(*oinfo.blocks_w)
        bh = (C.c_int32 * 3)(*oinfo.blocks_h)
        n = sum(oinfo.blocks_w[c] * oinfo.blocks_h[c] * 64 for c in range(oinfo.ncomp))
        flat = np.zeros(n, dtype=np.int16)
        rc = self.lib.ljt_read_coefficients(data, len(data), flat.ctypes.data, bw, bh)
        if rc:
            raise ValueError(f"ljt_read_coefficients rc={rc}")
        return Oracle.split(oinfo, flat, 64)

    def raw_planes(self, data: bytes, oinfo) -> list[np.ndarray]:
        bw = (C.c_int32 * 3)(*oinfo.blocks_w)
        bh = (C.c_int32 * 3)(*oinfo.blocks_h)
        n = sum(oinfo.blocks_w[c] * oinfo.blocks_h[c] * 64 for c in range(oinfo.ncomp))
        flat = np.zeros(n, dtype=np.uint8)
        rc = self.lib.ljt_read_raw(data, len(data), flat.ctypes.data, bw, bh)
        if rc:
            raise ValueError(f"ljt_read_raw rc={rc}")
        out, base = [], 0
        for c in range(oinfo.ncomp):
            w, h = oinfo.blocks_w[c] * 8, oinfo.blocks_h[c] * 8
            out.append(flat[base:base + w * h].reshape(h, w))
            base += w * h
        return out

    def decode_rgb(self, data: bytes, width: int, height: int) -> np.ndarray:
        out = np.zeros((height, width, 3), dtype=np.uint8)
        rc = self.lib.ljt_decode(data, len(data), out.ctypes.data, width * 3, 0)
        if rc:
            raise ValueError(f"ljt_decode rc={rc}")
        return out

    def decode_batch(self, datas: list[bytes], outs: list[np.ndarray], pitches: list[int], mode: int, nthreads: int):
        """mode 0 RGB, 1 gray, 2 raw planes; releases the GIL for the whole batch."""
        n = len(datas)
        dptr = (C.c_char_p * n)(*datas)
        lens = (C.c_size_t * n)(*[len(d) for d in datas])
        optr = (C.c_void_p * n)(*[o.ctypes.data for o in outs])
        pit = (C.c_size_t * n)(*pitches)
        rc = self.lib.ljt_decode_batch(n, dptr, lens, optr, pit, mode, nthreads)
        if rc:
            raise ValueError(f"ljt_decode_batch rc={rc}")


class RefParsed(C.Structure):
    _fields_ = [
        ("ok", C.c_int32),
        ("width", C.c_int32), ("height", C.c_int32), ("ncomp", C.c_int32), ("css", C.c_int32),
        ("comp_id", C.c_int32 * 3), ("hs", C.c_int32 * 3), ("vs", C.c_int32 * 3), ("tq", C.c_int32 * 3),
        ("scan_ncomp", C.c_int32), ("td", C.c_int32 * 3), ("ta", C.c_int32 * 3),
        ("restart_interval", C.c_int32),
        ("num_mcus", C.c_uint32), ("scan_offset", C.c_uint32), ("scan_size", C.c_uint32),
        ("qt", (C.c_uint8 * 64) * 4), ("qt_present", C.c_uint8 * 4),
        ("dc_bits", (C.c_uint8 * 16) * 2), ("dc_vals", (C.c_uint8 * 12) * 2),
        ("ac_bits", (C.c_uint8 * 16) * 2), ("ac_vals", (C.c_uint8 * 162) * 2),
        ("huff_present", C.c_uint8 * 2),
    ]


def ref_available() -> bool:
    _ensure_built()
    return os.path.exists(os.path.join(HERE, "_ref", "libref_parser.so")) and os.path.exists(
        os.path.join(HERE, "_ref", "libref_kernels.so"))


class RefParser:
    """The reference's RocJpegStreamParser (src/rocjpeg_parser.cpp) run as-is."""

    def __init__(self):
        _ensure_built()
        self.lib = C.CDLL(os.path.join(HERE, "_ref", "libref_parser.so"))
        self.lib.ref_parse.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(RefParsed)]

    def parse(self, data: bytes) -> RefParsed:
        # the reference parser reads a few bytes past the end on truncated
        # input (src/rocjpeg_parser.cpp:74,137-143,407): give it slack
        buf = C.create_string_buffer(data + b"\x00" * 64, len(data) + 64)
        o = RefParsed()
        self.lib.ref_parse(buf, len(data), C.byref(o))
        return o


class RefKernels:
    """The reference's HIP kernels (src/rocjpeg_hip_kernels.cpp) executed on the CPU."""

    def __init__(self):
        _ensure_built()
        self.lib = C.CDLL(os.path.join(HERE, "_ref", "libref_kernels.so"))
