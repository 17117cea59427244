#!/usr/bin/env python
"""Per-source-line summary of an ncu report (needs -lineinfo and --import-source on):
   python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top]
Prints, per CUDA source line, warp-stall samples (% of kernel), executed instructions and the dominant stall reasons."""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
if not hi:
    sys.exit("no source rows (was the report captured with --import-source on and built with -lineinfo?)")
per_all = {}
for h in hi:
    H = rows[h]
    fname = ""
    for r in rows[max(0, h - 3):h]:
        if r and r[0] == "File Path":
            fname = r[1].split("/")[-1]
    c = {name: H.index(name) for name in ("Line No", "# Samples", "Instructions Executed") if name in H}
    stalls = [(i, n) for i, n in enumerate(H) if n.startswith("stall_") and "Not Issued" not in n]
    src_col = 1
    per = defaultdict(lambda: [0, 0, defaultdict(int), ""])
    per_all[fname] = per
    cur_line, cur_src = None, ""
    for r in rows[h + 1:]:
        if not r or r[0] in ("File Path", "Function Name", "Line No"):
            if r and r[0] == "File Path":
                break
            continue
        if r[0].strip():
            cur_line, cur_src = r[0], r[src_col]
        try:
            s = int(r[c["# Samples"]] or 0)
            ie = int(r[c["Instructions Executed"]] or 0)
        except (ValueError, IndexError):
            continue
        e = per[cur_line]
        e[0] += s
        e[1] += ie
        e[3] = cur_src
        for i, n in stalls:
            try:
                e[2][n] += int(r[i] or 0)
            except ValueError:
                pass
flat = [(f, l, e) for f, per in per_all.items() for l, e in per.items()]
tot = sum(e[0] for _, _, e in flat) or 1
toti = sum(e[1] for _, _, e in flat) or 1
print(f"total samples {tot}, total warp instructions {toti}")
for f, line, e in sorted(flat, key=lambda t: -t[2][0])[:top]:
    st = ", ".join(f"{n[6:]}={v * 100 // max(e[0], 1)}%" for n, v in sorted(e[2].items(), key=lambda kv: -kv[1])[:3] if v)
    print(f"{f[:16]:16s} L{line:>4s} {100 * e[0] / tot:5.1f}% smp {100 * e[1] / toti:5.1f}% inst | {e[3].strip()[:80]} | {st}")
