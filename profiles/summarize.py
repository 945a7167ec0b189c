"""Turn ncu outputs brought back in gpurun_out/ into the small, tracked summaries under profiles/.

  python profiles/summarize.py <tag> <launches.csv> [<report.ncu-rep> [<workload>]]

With <workload> (bench.py --workload name) the DRAM bytes of one step's launches are also recorded per
stage in profiles/traffic.json, which bench.py reports as roofline.traffic.

Writes profiles/<tag>_launches.md (per-kernel launch count, mean duration, share of the step) and,
when a full report is given, profiles/<tag>_kernels.md (DRAM bytes, throughput %, issue %, stalls).
"""
import collections
import csv
import subprocess
import sys

tag, launches = sys.argv[1], sys.argv[2]
rep = sys.argv[3] if len(sys.argv) > 3 else None
rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
hdr, data = rows[hi], rows[hi + 1:]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= vi:
        continue
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] == "ns" else v * 1e3 if r[ui] == "ms" else v
    agg.setdefault(r[ki].split("(")[0].replace("rjb::<unnamed>::", "").replace("void ", ""), []).append(v)
tot = sum(sum(v) for v in agg.values())
with open(f"profiles/{tag}_launches.md", "w") as f:
    f.write(f"# {tag}: ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare shares)\n\n")
    f.write("| kernel | launches | mean us | share |\n|---|---|---|---|\n")
    for k, v in agg.items():
        f.write(f"| {k[:60]} | {len(v)} | {sum(v) / len(v):.1f} | {100 * sum(v) / tot:.1f}% |\n")
if rep:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rr = list(csv.reader(raw.splitlines()))
    h, u, d = rr[0], rr[1], rr[2:]
    cols = [("Kernel Name", "kernel"), ("gpu__time_duration.sum", "us"), ("launch__grid_size", "grid"),
            ("launch__registers_per_thread", "regs"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
            ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
            ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm %"),
            ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ %"),
            ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
            ("smsp__inst_executed.sum", "warp inst"),
            ("smsp__thread_inst_executed_per_inst_executed.ratio", "thr/inst"),
            ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st long_sb"),
            ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st short_sb"),
            ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st barrier"),
            ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem conflicts")]
    idx = [(h.index(c), n) for c, n in cols if c in h]
    with open(f"profiles/{tag}_kernels.md", "w") as f:
        f.write(f"# {tag}: ncu --set full --clock-control none, one row per captured launch\n\n")
        f.write("| " + " | ".join(n for _, n in idx) + " |\n|" + "---|" * len(idx) + "\n")
        for r in d:
            cells = []
            for i, n in idx:
                v = r[i]
                if n == "kernel":
                    v = v.split("(")[0].replace("unnamed>::", "").replace("void ", "")
                else:
                    try:
                        v = f"{float(v.replace(',', '')):.4g} {u[i]}".replace(" inst", "").replace(" register/thread", "")
                    except ValueError:
                        pass
                cells.append(v)
            f.write("| " + " | ".join(cells) + " |\n")
if rep and len(sys.argv) > 4:
    import json
    import os

    stage_of = {"k0_reduce": "destuff", "k0_scan": "destuff", "k0_apply": "destuff", "k1_sync": "huffman_sync", "k1_fused": "huffman_sync", "k1_scan": "huffman_write",
                "k1_write": "huffman_write", "dc_sums": "dc", "dc_scan": "dc", "dc_apply": "dc", "dc_image": "dc", "k2_idct": "idct",
                "k23_fused": "output", "k23_warp": "output", "k3_output": "output"}
    ki, ri, wi = h.index("Kernel Name"), h.index("dram__bytes_read.sum"), h.index("dram__bytes_write.sum")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    per_stage, seen = {}, {}
    per_step = {"k1_sync": 2}   # launches of a kernel in one step (default 1)
    for r in d:   # launches in order: one step's worth, wherever in the step the capture window began
        name = r[ki].split("(")[0].split("<")[0].replace("void ", "").split("::")[-1]
        st = stage_of.get(name)
        if st is None:
            continue
        if seen.get(name, 0) >= per_step.get(name, 1):
            break
        seen[name] = seen.get(name, 0) + 1
        b = float(r[ri].replace(",", "")) * scale.get(u[ri], 1.0) + float(r[wi].replace(",", "")) * scale.get(u[wi], 1.0)
        per_stage[st] = per_stage.get(st, 0.0) + b
    path = "profiles/traffic.json"
    t = json.load(open(path)) if os.path.exists(path) else {}
    t[sys.argv[4]] = {k: int(v) for k, v in per_stage.items()}
    # issue-slot utilisation of the entropy stage (what bench.py reports as roofline_k1.issue_utilisation)
    ii = h.index("smsp__issue_active.avg.pct_of_peak_sustained_active") if "smsp__issue_active.avg.pct_of_peak_sustained_active" in h else None
    if ii is not None:
        iss = {}
        for r in d:
            name = r[ki].split("(")[0].split("<")[0].replace("void ", "").split("::")[-1]
            if name in ("k1_sync", "k1_write", "k1_fused", "k2_idct", "k23_warp", "k23_fused", "k3_output", "k0_apply") and name not in iss:
                iss[name] = round(float(r[ii].replace(",", "")) / 100.0, 3)
        if iss:
            t[sys.argv[4]]["k1_issue"] = {k: v for k, v in iss.items() if k.startswith("k1_")}
            t[sys.argv[4]]["issue"] = iss
    t.setdefault("_note", "dram__bytes_read.sum + dram__bytes_write.sum per stage of ONE step, from the ncu --set full capture named by "
                          "the tag in profiles/<tag>_kernels.md (cold cache, one pipeline lane)")
    json.dump(t, open(path, "w"), indent=1, sort_keys=True)
print("ok")
