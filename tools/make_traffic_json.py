#!/usr/bin/env python
"""profiles/r2_traffic.json from the committed ncu --set full captures of this round (tools/evidence_r2.sh): DRAM bytes (read +
write) and duration of the profiled launch per query, keyed by configuration and by the kernel class bench.py's roofline uses.
Run from the repo root after copying gpurun_out/r2_*.ncu-rep to profiles/."""
import csv
import io
import json
import os
import subprocess

UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
# config -> kernel class -> [(capture, queries in the profiled launch)]
CAPS = {"C3": {"locate": [("r2_k_locate_ring_C3", 25_000_000)], "eval_binned": [("r2_k_eval_binned_mma_C3", 25_000_000)],
               "scatter": [("r2_k_scatter_C3", 25_000_000), ("r2_k_unpermute_C3", 25_000_000)]},
        "C2": {"locate": [("r2_k_locate_ring_C2", 4_194_304)], "eval_binned": [("r2_k_eval_binned_exact_C2", 4_194_304)],
               "scatter": [("r2_k_scatter_C2", 4_194_304), ("r2_k_unpermute_C2", 4_194_304)]}}


def read(rep):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    H, U, V = rows[0], rows[1], rows[2]

    def get(name):
        i = H.index(name)
        return float(V[i].replace(",", "")) * UNITS[U[i]]
    return get("dram__bytes_read.sum"), get("dram__bytes_write.sum"), get("gpu__time_duration.sum"), V[H.index("Kernel Name")]


out = {}
for cfg, classes in CAPS.items():
    out[cfg] = {}
    for cls, caps in classes.items():
        per_q, parts = 0.0, []
        for cap, items in caps:
            rep = f"profiles/{cap}.ncu-rep"
            if not os.path.exists(rep):
                parts = None
                break
            rd, wr, us, name = read(rep)
            per_q += (rd + wr) / items
            parts.append({"capture": rep, "kernel": name[:60], "dram_bytes_read": rd, "dram_bytes_write": wr, "duration_us_under_ncu": us,
                          "queries_in_profiled_launch": items})
        if parts:
            out[cfg][cls] = {"dram_bytes_per_query": per_q, "parts": parts,
                             "source": " + ".join(p["capture"] for p in parts) + ": dram__bytes_read.sum + dram__bytes_write.sum per query x queries per launch"}
out["_how"] = "ncu --set full --clock-control none, one launch per kernel (tools/evidence_r2.sh: tools/bench_pipeline.py --queries 50000000 --steps 1)"
json.dump(out, open("profiles/r2_traffic.json", "w"), indent=1)
for cfg in CAPS:
    for cls, e in out[cfg].items():
        print(cfg, cls, round(e["dram_bytes_per_query"], 1), "B/query")
