#!/usr/bin/env python
"""profiles/r1_traffic.json from the committed ncu --set full captures: DRAM bytes (read + write) and duration of the
profiled launch, per query / per leaf.  Run from the repo root after refreshing profiles/r1_*.ncu-rep."""
import csv
import io
import json
import subprocess

UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
# capture -> (items in the profiled launch, unit name)
CAPS = {"k_locate_only": (4194304, "query"), "k_eval_binned": (4194304, "query"), "k_scatter": (4194304, "query"),
        "k_unpermute": (4194304, "query"), "k_fit": (100006, "leaf"), "k_fit_chain": (100006, "leaf"),
        "k_fit_C3": (1000021, "leaf"), "k_eval_dmma": (20000000, "query"),
        "k_eval_binned_C3": (40000000, "query"), "k_locate_wide_C4": (20000000, "query")}
out = {}
for key, (items, unit) in CAPS.items():
    txt = subprocess.run(["ncu", "-i", f"profiles/r1_{key}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    H, U, V = rows[0], rows[1], rows[2]
    def get(name):
        i = H.index(name)
        return float(V[i].replace(",", "")) * UNITS[U[i]]
    rd, wr, us = get("dram__bytes_read.sum"), get("dram__bytes_write.sum"), get("gpu__time_duration.sum")
    plural = "queries" if unit == "query" else "leaves"
    out[key] = {"dram_bytes_read": rd, "dram_bytes_write": wr, "duration_us_under_ncu": us, "kernel": V[H.index("Kernel Name")],
                f"{plural}_in_profiled_launch": items, f"dram_bytes_per_{unit}": (rd + wr) / items}
out["_how"] = ("ncu --set full --clock-control none, one launch per kernel (tools/evidence_profiles.sh: bench.py --queries 20000000 = "
               "4 194 304-query chunks; tools/prof_fit.sh C2 / C3; tools/prof_dense.sh = C4 with 2e7 queries; tools/prof_c3_eval.sh = one 4e7-query chunk of C3; bench_c4.py = 2e7 queries); dram__bytes_read.sum + dram__bytes_write.sum")
json.dump(out, open("profiles/r1_traffic.json", "w"), indent=1)
for k, v in out.items():
    if k != "_how":
        print(k, {a: (round(b, 2) if isinstance(b, float) else b) for a, b in v.items() if a != "kernel"})
