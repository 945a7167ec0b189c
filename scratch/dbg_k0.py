import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
from rocjpeg_b200 import api, datagen
torch.cuda.set_device(0)
def run(w, h, css, rows, n, lanes):
    os.environ["ROCJPEG_B200_LANES"] = str(lanes)
    dec = api.Decoder(0, 0)
    d = datagen.make_jpeg(w, h, css, seed=400, restart_rows=rows)
    datas = [d] * n
    streams, dests, keep = [], [], []
    for d in datas:
        s = api.JpegStream(); assert s.parse(d) == 0
        inf = s.info()
        buf = torch.zeros(inf.width * inf.height * 3 + 64, dtype=torch.uint8, device="cuda")
        streams.append(s); keep.append(buf)
        dests.append([(buf.data_ptr(), inf.width), (buf.data_ptr() + inf.width * inf.height, inf.width), (buf.data_ptr() + 2 * inf.width * inf.height, inf.width)])
    for it in range(6):
        st = dec.decode_batched(streams, api.make_params("y"), dests)
        if st != 0:
            print("FAIL", w, h, css, rows, n, lanes, "iter", it, "status", st); return False
    bad = 0
    for i, s in enumerate(streams[:2]):
        inf = s.info(); hs = s.host_scan(); ds = dec.scan_status(i); print("reserved flags", hex(ds.reserved))
        if ds.scan_size != hs.scan_size or ds.segments_seen != hs.restart_markers_seen + 1:
            print("status mismatch", i, ds.scan_size, hs.scan_size, ds.segments_seen, hs.restart_markers_seen + 1); bad += 1
        for k in range(inf.num_segments):
            if dec.device_segment(i, k) != s.segment(k):
                a, b = dec.device_segment(i, k), s.segment(k)
                first = next((j for j in range(min(len(a), len(b))) if a[j] != b[j]), None)
                diffs = [j for j in range(min(len(a), len(b))) if a[j] != b[j]]
                print("segment mismatch img", i, "seg", k, len(a), len(b), "first diff at", first, "ndiff", len(diffs), "last diff", diffs[-1] if diffs else None); bad += 1
                print("  dev ", a[max(0,first-16):first+48].hex())
                print("  host", b[max(0,first-16):first+48].hex())
                raw = d[inf.scan_offset:]
                key = b[first-16:first]
                pos = raw.find(key)
                print("  raw pos of the 16 bytes before:", pos, "mod 64:", pos % 64 if pos>=0 else None, "raw:", raw[pos:pos+80].hex() if pos>=0 else None)
                if bad > 5: break
    print("ok" if not bad else "BAD", w, h, css, rows, n, lanes, len(d))
    dec.close()
    return not bad
w, h, css, rows, n, lanes = sys.argv[1:7]
run(int(w), int(h), css, int(rows), int(n), int(lanes))
