"""Regenerate tests/golden/ — run in the BUILD container (needs /root/reference).

  python tests/golden/make_golden.py

Writes small JPEG fixtures (*.jpg) and golden.json (SHA-256 of every stage's
output for every fixture). At generation time each stage is cross-checked
against the strongest available reference before its hash is recorded:

  coefficients : oracle == libjpeg-turbo 3.1.4.1 jpeg_read_coefficients
  planes       : oracle == libjpeg-turbo jpeg_read_raw_data (islow) on the
                 region libjpeg defines (width_in_blocks x height_in_blocks)
  parser       : oracle fields == the reference's RocJpegStreamParser
                 (src/rocjpeg_parser.cpp compiled as-is into oracle/_ref)
  outputs      : oracle == the reference's own HIP kernels + decoder assembly
                 executed on the CPU (oracle/_ref, tests/ref_assembly.py), for
                 all five output formats, uncropped, and with an even-aligned
                 crop wherever the reference is self-consistent (see the ROI
                 note above orc_convert in oracle/jpeg_oracle.c).

The fixtures mug_4xx_crop.jpg are lossless MCU-aligned crops of the reference's
data/images/mug_{400,420,422}.jpg (same quantisation/Huffman tables, sampling
factors incl. the (2x2,1x2,1x2) 4:2:2 layout), re-entropy-coded by
tests/jpeg_writer.py; the GPU box has no /root/reference, so these travel.
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import oracle  # noqa: E402
from rocjpeg_b200 import datagen  # noqa: E402
import jpeg_writer as jw  # noqa: E402
from ref_assembly import reference_output  # noqa: E402

REF_IMAGES = "/root/reference/data/images"
FORMATS = ["native", "yuv_planar", "y", "rgb", "rgb_planar"]
CROP = (16, 8, 80, 56)


def sha(arrs):
    h = hashlib.sha256()
    for a in arrs:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def custom_huffman_stream(orc):
    """4:2:0 picture re-coded with non-standard tables whose AC codes reach 16 bits
    and whose symbol order is permuted, restart interval = 1 MCU."""
    base = datagen.make_jpeg(96, 80, "420", seed=11)
    rc, info = orc.parse(base)
    coefs = orc.coefficients(base, info)
    rng = np.random.default_rng(5)
    ac_syms = list(jw.STD_AC_LUMA[1])
    rng.shuffle(ac_syms)
    # keep EOB and ZRL cheap so the stream stays small; the rest permuted
    for s in (0x00, 0xF0, 0x01, 0x02, 0x11):
        ac_syms.remove(s)
    ac_syms = [0x00, 0x01, 0x02, 0x11, 0xF0] + ac_syms
    ac_bits = [0, 1, 2, 2, 3, 4, 6, 8, 10, 12, 14, 16, 18, 20, 22, 24]
    assert sum(ac_bits) == 162
    dc_bits = [0, 0, 2, 2, 2, 2, 2, 1, 1, 0, 0, 0, 0, 0, 0, 0]
    dc_syms = [3, 4, 2, 5, 1, 6, 0, 7, 8, 9, 10, 11]
    tabs_dc = {0: (dc_bits, dc_syms), 1: jw.STD_DC_CHROMA}
    tabs_ac = {0: (ac_bits, ac_syms), 1: jw.STD_AC_CHROMA}
    qts = {t: bytes(info.qt[t]) for t in range(4) if info.qt_present[t]}
    return jw.write_jpeg(96, 80, coefs, [2, 1, 1], [2, 1, 1], [info.tq[c] for c in range(3)], qts, tabs_dc, tabs_ac,
                         [0, 1, 1], [0, 1, 1], restart_interval=1)


def extreme_coefficient_stream():
    """4:4:4, all-ones quantiser, large-magnitude coefficients incl. a full block of
    non-zeros, ZRL runs and category-10 AC values (within the islow int32 domain)."""
    rng = np.random.default_rng(9)
    bw, bh = 4, 3
    coefs = []
    for c in range(3):
        a = np.zeros((bh, bw, 64), dtype=np.int16)
        for by in range(bh):
            for bx in range(bw):
                kind = (by * bw + bx + c) % 4
                if kind == 0:
                    a[by, bx] = rng.integers(-40, 41, 64)
                elif kind == 1:
                    a[by, bx, 0] = rng.integers(-1000, 1000)
                    a[by, bx, 63] = 1
                elif kind == 2:
                    a[by, bx, 0] = rng.integers(-1000, 1000)
                    a[by, bx, jw.ZIGZAG[40]] = -700
                    a[by, bx, jw.ZIGZAG[17]] = 1023
                else:
                    a[by, bx, 0] = (-1) ** bx * 1000
        coefs.append(a)
    q1 = {0: bytes([1] * 64)}
    return jw.write_jpeg(31, 23, coefs, [1, 1, 1], [1, 1, 1], [0, 0, 0], q1)


def build_cases(orc):
    cases = {}
    for css in datagen.CSS_NAMES:
        cases[f"synth_{css}_123x77"] = datagen.make_jpeg(123, 77, css, seed=1)
        cases[f"synth_{css}_123x77_dri"] = datagen.make_jpeg(123, 77, css, seed=2, restart_rows=1)
    cases["synth_420_64x64"] = datagen.make_jpeg(64, 64, "420", seed=3)
    cases["synth_444_500x375"] = datagen.make_jpeg(500, 375, "444", seed=100)
    cases["synth_422_500x375"] = datagen.make_jpeg(500, 375, "422", seed=101)
    cases["synth_420_500x375_dri7"] = datagen.make_jpeg(500, 375, "420", seed=102, restart_mcus=7)
    cases["synth_400_333x211"] = datagen.make_jpeg(333, 211, "400", seed=4)
    if os.path.isdir(REF_IMAGES):
        mug = {n: open(os.path.join(REF_IMAGES, f"mug_{n}.jpg"), "rb").read() for n in ("400", "420", "422")}
        cases["mug_420_crop"] = jw.lossless_crop(orc, mug["420"], 100, 60, 12, 6)
        cases["mug_422_crop"] = jw.lossless_crop(orc, mug["422"], 100, 60, 12, 6)
        cases["mug_400_crop"] = jw.lossless_crop(orc, mug["400"], 200, 120, 24, 12)
        cases["mug_422_crop_dri1"] = jw.lossless_crop(orc, mug["422"], 90, 50, 9, 5, restart_interval=1)
    cases["custom_huffman_420_dri1"] = custom_huffman_stream(orc)
    cases["extreme_coefs_444"] = extreme_coefficient_stream()
    return cases


def main():
    orc = oracle.Oracle()
    ljt = oracle.LibJpegTurbo()
    have_ref = oracle.ref_available()
    rp = oracle.RefParser() if have_ref else None
    rk = oracle.RefKernels() if have_ref else None
    cases = build_cases(orc)
    golden = {}
    for name, data in sorted(cases.items()):
        rc, info = orc.parse(data)
        assert rc == 0 and orc.supported(info) == 0, name
        coefs = orc.coefficients(data, info)
        lc = ljt.coefficients(data, info)
        assert all((a == b).all() for a, b in zip(coefs, lc)), f"{name}: coefficients != libjpeg-turbo"
        planes = orc.planes(data, info)
        lp = ljt.raw_planes(data, info)
        li = ljt.info(data)
        for c in range(info.ncomp):
            hh, ww = li.hib[c] * 8, li.wib[c] * 8
            assert (planes[c][:hh, :ww] == lp[c][:hh, :ww]).all(), f"{name}: plane {c} != libjpeg-turbo"
        if rp is not None:
            r = rp.parse(data)
            assert r.ok and (r.width, r.height, r.css, r.restart_interval) == (
                info.width, info.height, info.css, info.restart_interval), name
            assert (r.scan_offset, r.scan_size) == (info.scan_offset, info.scan_size), name
        entry = {"bytes": len(data), "width": info.width, "height": info.height, "css": oracle.CSS[info.css],
                 "restart_interval": info.restart_interval, "coefficients": sha(coefs), "planes": sha(planes),
                 "outputs": {}}
        for fmt in FORMATS:
            crops = [(0, 0, 0, 0)]
            if info.width >= 96 and info.height >= 64:
                crops.append(CROP)
            for crop in crops:
                _, dst = orc.decode(data, fmt, crop)
                shapes = oracle.output_shapes(info, fmt, crop, orc)
                valid = [d[:rows, :rb] for d, (rows, rb) in zip(dst, shapes) if d is not None]
                cropped = crop != (0, 0, 0, 0)
                ref_defect = cropped and fmt in ("rgb", "rgb_planar") and oracle.CSS[info.css] in ("444", "440")
                if rk is not None and not ref_defect:
                    ref = reference_output(rk, info, planes, fmt, crop)
                    for a, b in zip(valid, ref):
                        assert a.shape == b.shape and (a == b).all(), f"{name} {fmt} {crop}: != reference kernels"
                entry["outputs"][f"{fmt}|{','.join(map(str, crop))}"] = sha(valid)
        golden[name] = entry
        with open(os.path.join(HERE, name + ".jpg"), "wb") as f:
            f.write(data)
        print(f"{name:32s} {len(data):7d} B  {info.width}x{info.height} css={oracle.CSS[info.css]} ri={info.restart_interval}")
    meta = {"libjpeg_turbo": os.path.basename(ljt.path), "reference_pinned": bool(have_ref), "crop": list(CROP)}
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump({"meta": meta, "cases": golden}, f, indent=1, sort_keys=True)
    print("wrote", len(golden), "cases; reference pinned:", have_ref)


if __name__ == "__main__":
    main()
