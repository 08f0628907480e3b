#!/usr/bin/env python
"""ncu report -> profiles/<name>.json: the per-launch counters bench.py quotes (warp instructions, DRAM bytes,
duration, issue utilisation) of the k_round / k_push_frames launches in the report, stamped with the hash of the
sources the library was built from (manette_b200.build.source_hash) so that stale numbers are refused.

usage: tools/ncu_counters.py <report.ncu-rep> <out.json> --next-calls N [--note TEXT]
  --next-calls: next() calls served by ONE captured launch (envs on its work list)."""
import argparse
import csv
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from manette_b200 import build as mb_build  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("report")
ap.add_argument("out")
ap.add_argument("--next-calls", type=float, required=True)
ap.add_argument("--note", default="")
a = ap.parse_args()
out = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
H = rows[0]


def col(r, name):
    return float(r[H.index(name)].replace(",", "")) if name in H and r[H.index(name)] not in ("", "n/a") else None


launches = []
for r in rows[2:]:
    U = rows[1]
    def unit_scale(name):
        u = U[H.index(name)] if name in H else ""
        return {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}.get(u, 1.0)
    launches.append({
        "kernel": r[H.index("Kernel Name")],
        "duration_s": col(r, "gpu__time_duration.sum") * unit_scale("gpu__time_duration.sum"),
        "warp_inst": col(r, "smsp__inst_executed.sum"),
        "thread_inst_per_warp_inst": col(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
        "issue_active_pct": col(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "warps_active_pct": col(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
        "dram_bytes": col(r, "dram__bytes_read.sum") * unit_scale("dram__bytes_read.sum") +
                      col(r, "dram__bytes_write.sum") * unit_scale("dram__bytes_write.sum"),
        "registers": col(r, "launch__registers_per_thread"),
        "grid": col(r, "launch__grid_size"),
    })
n = len(launches)
doc = {"source_hash": mb_build.source_hash(), "report": os.path.basename(a.report), "note": a.note,
       "next_calls_per_launch": a.next_calls, "launches": launches,
       "warp_inst_per_next": sum(l["warp_inst"] for l in launches) / n / a.next_calls,
       "dram_bytes_per_next": sum(l["dram_bytes"] for l in launches) / n / a.next_calls}
with open(a.out, "w") as f:
    json.dump(doc, f, indent=1)
print(json.dumps({k: doc[k] for k in ("source_hash", "warp_inst_per_next", "dram_bytes_per_next")}))
