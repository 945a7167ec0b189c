"""Top instructions by warp-stall samples for one kernel of an ncu report (needs -lineinfo / --import-source).

  python profiles/hot_instructions.py <report.ncu-rep> <kernel regex> [N]
"""
import csv
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{pat}"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
# several launches may match: sections start with a "Kernel Name" row followed by a header row
sections, i = [], 0
while i < len(rows):
    if rows[i] and rows[i][0] == "Kernel Name":
        name, hdr, j = rows[i][1], rows[i + 1], i + 2
        while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
            j += 1
        sections.append((name, hdr, rows[i + 2:j]))
        i = j
    else:
        i += 1
name, h, d = sections[0]
si, src, ie = h.index("# Samples"), h.index("Source"), h.index("Instructions Executed")
cols = ["stall_short_sb", "stall_long_sb", "stall_wait", "stall_math", "stall_mio", "stall_lg", "stall_barrier", "stall_not_selected",
        "stall_branch_resolving", "stall_dispatch", "stall_no_inst"]
ci = [h.index(c) for c in cols]
d = [r for r in d if len(r) > max(ci) and r[si].isdigit()]
tot = sum(int(r[si]) for r in d)
print(f"# {name}: {tot} samples, {sum(int(r[ie]) for r in d)} warp instructions")
for c, i in zip(cols, ci):
    print(f"#   {c}: {sum(int(r[i]) for r in d)}")
for r in sorted(d, key=lambda r: -int(r[si]))[:topn]:
    print(r[si].rjust(6), r[ie].rjust(9), r[src].strip()[:64].ljust(64), " ".join(f"{c[6:]}={r[i]}" for c, i in zip(cols, ci) if int(r[i]) > 0))
