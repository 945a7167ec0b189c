"""Baseline-JPEG writer from quantised coefficients — TEST TOOL.

Entropy-codes given coefficient blocks (T.81 Annex F.1.2) with arbitrary
sampling factors, Huffman tables, quantisation tables and restart interval, so
tests can build bitstreams no stock encoder in the image will produce:
lossless MCU-aligned crops of the reference's data/images fixtures (keeping
their tables and the (2x2,1x2,1x2) 4:2:2 layout), restart intervals of one MCU,
custom/optimised Huffman tables with 16-bit codes, extreme coefficient values.
Pure Python: small images only.
"""
from __future__ import annotations

import struct

import numpy as np

ZIGZAG = np.array([
    0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6, 7, 14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63])

# T.81 Annex K.3 typical tables
STD_DC_LUMA = ([0, 1, 5, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0], list(range(12)))
STD_DC_CHROMA = ([0, 3, 1, 1, 1, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0], list(range(12)))
STD_AC_LUMA = ([0, 2, 1, 3, 3, 2, 4, 3, 5, 5, 4, 4, 0, 0, 1, 0x7d], [
    0x01, 0x02, 0x03, 0x00, 0x04, 0x11, 0x05, 0x12, 0x21, 0x31, 0x41, 0x06, 0x13, 0x51, 0x61, 0x07,
    0x22, 0x71, 0x14, 0x32, 0x81, 0x91, 0xa1, 0x08, 0x23, 0x42, 0xb1, 0xc1, 0x15, 0x52, 0xd1, 0xf0,
    0x24, 0x33, 0x62, 0x72, 0x82, 0x09, 0x0a, 0x16, 0x17, 0x18, 0x19, 0x1a, 0x25, 0x26, 0x27, 0x28,
    0x29, 0x2a, 0x34, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48, 0x49,
    0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68, 0x69,
    0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x83, 0x84, 0x85, 0x86, 0x87, 0x88, 0x89,
    0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5, 0xa6, 0xa7,
    0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3, 0xc4, 0xc5,
    0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda, 0xe1, 0xe2,
    0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf1, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa])
STD_AC_CHROMA = ([0, 2, 1, 2, 4, 4, 3, 4, 7, 5, 4, 4, 0, 1, 2, 0x77], [
    0x00, 0x01, 0x02, 0x03, 0x11, 0x04, 0x05, 0x21, 0x31, 0x06, 0x12, 0x41, 0x51, 0x07, 0x61, 0x71,
    0x13, 0x22, 0x32, 0x81, 0x08, 0x14, 0x42, 0x91, 0xa1, 0xb1, 0xc1, 0x09, 0x23, 0x33, 0x52, 0xf0,
    0x15, 0x62, 0x72, 0xd1, 0x0a, 0x16, 0x24, 0x34, 0xe1, 0x25, 0xf1, 0x17, 0x18, 0x19, 0x1a, 0x26,
    0x27, 0x28, 0x29, 0x2a, 0x35, 0x36, 0x37, 0x38, 0x39, 0x3a, 0x43, 0x44, 0x45, 0x46, 0x47, 0x48,
    0x49, 0x4a, 0x53, 0x54, 0x55, 0x56, 0x57, 0x58, 0x59, 0x5a, 0x63, 0x64, 0x65, 0x66, 0x67, 0x68,
    0x69, 0x6a, 0x73, 0x74, 0x75, 0x76, 0x77, 0x78, 0x79, 0x7a, 0x82, 0x83, 0x84, 0x85, 0x86, 0x87,
    0x88, 0x89, 0x8a, 0x92, 0x93, 0x94, 0x95, 0x96, 0x97, 0x98, 0x99, 0x9a, 0xa2, 0xa3, 0xa4, 0xa5,
    0xa6, 0xa7, 0xa8, 0xa9, 0xaa, 0xb2, 0xb3, 0xb4, 0xb5, 0xb6, 0xb7, 0xb8, 0xb9, 0xba, 0xc2, 0xc3,
    0xc4, 0xc5, 0xc6, 0xc7, 0xc8, 0xc9, 0xca, 0xd2, 0xd3, 0xd4, 0xd5, 0xd6, 0xd7, 0xd8, 0xd9, 0xda,
    0xe2, 0xe3, 0xe4, 0xe5, 0xe6, 0xe7, 0xe8, 0xe9, 0xea, 0xf2, 0xf3, 0xf4, 0xf5, 0xf6, 0xf7, 0xf8,
    0xf9, 0xfa])


def build_codes(bits, vals):
    """T.81 Annex C: symbol -> (code, length)."""
    out, code, k = {}, 0, 0
    for length in range(1, 17):
        for _ in range(bits[length - 1]):
            out[vals[k]] = (code, length)
            code += 1
            k += 1
        code <<= 1
    return out


def flat_huffman(nsyms_by_len, symbols):
    """Helper: a table spec (bits, vals) from explicit per-length counts."""
    assert sum(nsyms_by_len) == len(symbols)
    return (list(nsyms_by_len), list(symbols))


class _BitWriter:
    def __init__(self):
        self.out = bytearray()
        self.acc = 0
        self.n = 0

    def put(self, code, length):
        if length == 0:
            return
        self.acc = (self.acc << length) | (code & ((1 << length) - 1))
        self.n += length
        while self.n >= 8:
            b = (self.acc >> (self.n - 8)) & 0xFF
            self.out.append(b)
            if b == 0xFF:
                self.out.append(0)
            self.n -= 8
        self.acc &= (1 << self.n) - 1

    def flush(self):
        if self.n:
            self.put((1 << (8 - self.n)) - 1, 8 - self.n)  # pad with 1-bits


def _mag(v):
    v = int(v)
    if v == 0:
        return 0, 0
    a = -v if v < 0 else v
    s = a.bit_length()
    return s, (v if v > 0 else v + (1 << s) - 1)


def encode_scan(coefs, hs, vs, mcus_x, mcus_y, dc_tabs, ac_tabs, td, ta, restart_interval=0):
    """coefs[c]: int array (blocks_h, blocks_w, 64) natural order (MCU-padded grid)."""
    ncomp = len(coefs)
    dcc = [build_codes(*dc_tabs[td[c]]) for c in range(ncomp)]
    acc = [build_codes(*ac_tabs[ta[c]]) for c in range(ncomp)]
    bw = _BitWriter()
    pred = [0] * ncomp
    total = mcus_x * mcus_y
    rst = 0
    for m in range(total):
        if restart_interval and m and m % restart_interval == 0:
            bw.flush()
            bw.out += bytes([0xFF, 0xD0 + (rst & 7)])
            rst += 1
            pred = [0] * ncomp
        mx, my = m % mcus_x, m // mcus_x
        for c in range(ncomp):
            H, V = (1, 1) if ncomp == 1 else (hs[c], vs[c])
            for v in range(V):
                for h in range(H):
                    blk = coefs[c][my * V + v, mx * H + h]
                    zz = blk[ZIGZAG]
                    dcv = int(zz[0])
                    s, bitsv = _mag(dcv - pred[c])
                    pred[c] = dcv
                    code, ln = dcc[c][s]
                    bw.put(code, ln)
                    bw.put(bitsv, s)
                    run = 0
                    nz = np.nonzero(zz[1:])[0]
                    last = int(nz[-1]) + 1 if nz.size else 0
                    for k in range(1, last + 1):
                        val = int(zz[k])
                        if val == 0:
                            run += 1
                            continue
                        while run > 15:
                            code, ln = acc[c][0xF0]
                            bw.put(code, ln)
                            run -= 16
                        s, bitsv = _mag(val)
                        code, ln = acc[c][(run << 4) | s]
                        bw.put(code, ln)
                        bw.put(bitsv, s)
                        run = 0
                    if last < 63:
                        code, ln = acc[c][0x00]
                        bw.put(code, ln)
    bw.flush()
    return bytes(bw.out)


def _seg(marker, payload):
    return bytes([0xFF, marker]) + struct.pack(">H", len(payload) + 2) + payload


def write_jpeg(width, height, coefs, hs, vs, tq, qts_zigzag, dc_tabs=None, ac_tabs=None, td=None, ta=None,
               restart_interval=0, extra_segments=b"", comp_ids=None, eoi=True, sof_marker=0xC0):
    """Assemble a complete sequential Huffman JPEG (SOF0, or SOF1 with sof_marker=0xC1).
    qts_zigzag: dict id -> 64 quantiser steps in zig-zag (stream) order; a table with a step above 255 is written with
    16-bit precision. dc_tabs/ac_tabs: dict id (0..3) -> (bits, vals)."""
    ncomp = len(coefs)
    dc_tabs = dc_tabs or {0: STD_DC_LUMA, 1: STD_DC_CHROMA}
    ac_tabs = ac_tabs or {0: STD_AC_LUMA, 1: STD_AC_CHROMA}
    td = td or [0, 1, 1][:ncomp]
    ta = ta or [0, 1, 1][:ncomp]
    comp_ids = comp_ids or [1, 2, 3][:ncomp]
    if ncomp == 1:
        mcus_x, mcus_y = (width + 7) // 8, (height + 7) // 8
    else:
        hmax, vmax = max(hs), max(vs)
        mcus_x, mcus_y = (width + 8 * hmax - 1) // (8 * hmax), (height + 8 * vmax - 1) // (8 * vmax)
    out = bytearray(b"\xFF\xD8")
    out += extra_segments
    for tid, q in sorted(qts_zigzag.items()):
        q = [int(x) for x in q]
        if max(q) > 255:
            out += _seg(0xDB, bytes([0x10 | tid]) + b"".join(struct.pack(">H", x) for x in q))
        else:
            out += _seg(0xDB, bytes([tid]) + bytes(q))
    sof = struct.pack(">BHHB", 8, height, width, ncomp)
    for c in range(ncomp):
        sof += bytes([comp_ids[c], (hs[c] << 4) | vs[c], tq[c]])
    out += _seg(sof_marker, sof)
    used_dc = sorted(set(td))
    used_ac = sorted(set(ta))
    dht = b""
    for t in used_dc:
        bits, vals = dc_tabs[t]
        dht += bytes([t]) + bytes(bits) + bytes(vals)
    for t in used_ac:
        bits, vals = ac_tabs[t]
        dht += bytes([0x10 | t]) + bytes(bits) + bytes(vals)
    out += _seg(0xC4, dht)
    if restart_interval:
        out += _seg(0xDD, struct.pack(">H", restart_interval))
    sos = bytes([ncomp])
    for c in range(ncomp):
        sos += bytes([comp_ids[c], (td[c] << 4) | ta[c]])
    sos += bytes([0, 63, 0])
    out += _seg(0xDA, sos)
    out += encode_scan(coefs, hs, vs, mcus_x, mcus_y, dc_tabs, ac_tabs, td, ta, restart_interval)
    if eoi:
        out += b"\xFF\xD9"
    return bytes(out)


def lossless_crop(orc, data, mcu_x0, mcu_y0, n_mcu_x, n_mcu_y, restart_interval=0):
    """MCU-aligned lossless crop of an existing JPEG, keeping its quantisation and
    Huffman tables and sampling factors (uses the oracle only to read coefficients)."""
    rc, info = orc.parse(data)
    assert rc == 0
    coefs = orc.coefficients(data, info)
    ncomp = info.ncomp
    hs = [info.hs[c] for c in range(ncomp)]
    vs = [info.vs[c] for c in range(ncomp)]
    sub = []
    for c in range(ncomp):
        H, V = (1, 1) if ncomp == 1 else (hs[c], vs[c])
        sub.append(np.ascontiguousarray(
            coefs[c][mcu_y0 * V:(mcu_y0 + n_mcu_y) * V, mcu_x0 * H:(mcu_x0 + n_mcu_x) * H]))
    if ncomp == 1:
        w, h = n_mcu_x * 8, n_mcu_y * 8
    else:
        w, h = n_mcu_x * 8 * max(hs), n_mcu_y * 8 * max(vs)
    qts = {t: bytes(info.qt[t]) for t in range(4) if info.qt_present[t]}
    dc_tabs = {t: (list(info.dc_bits[t]), list(info.dc_vals[t])[:sum(info.dc_bits[t])]) for t in range(2)
               if info.dc_present[t]}
    ac_tabs = {t: (list(info.ac_bits[t]), list(info.ac_vals[t])[:sum(info.ac_bits[t])]) for t in range(2)
               if info.ac_present[t]}
    return write_jpeg(w, h, sub, hs, vs, [info.tq[c] for c in range(ncomp)], qts, dc_tabs, ac_tabs,
                      [info.td[c] for c in range(ncomp)], [info.ta[c] for c in range(ncomp)], restart_interval,
                      comp_ids=[info.comp_id[c] for c in range(ncomp)])
