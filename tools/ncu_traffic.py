#!/usr/bin/env python
"""DRAM traffic per source iteration and per entry point from one `ncu --set full` capture of tools/gpu_probe.py
(one iteration's worth of launches).  Writes profiles/ncu_traffic.json, which bench.py reports as roofline.traffic.
Usage: python tools/ncu_traffic.py file.ncu-rep workload:layout [out.json] [commit]"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ENTRY = [("sweep_", "SweepSolver"), ("p2p_", "SweepSolver_sync"), ("moments_mma_kernel<3", "LTimes"), ("moments_mma_kernel<4", "LPlusTimes"),
         ("ltimes_slab", "LTimes"), ("moments_rowmma_stream", "LTimes"), ("moments_rowmma_resident", "LPlusTimes"), ("scatter_slab", "scattering"), ("scatter_mma", "scattering"),
         ("scatter_fractions", "scattering_fractions"), ("slab_fractions", "scattering_fractions"), ("slab_matrices", "scattering_matrices"), ("moments_transpose", "scattering_transpose"), ("population", "population"), ("source", "source")]


def main():
    rep, key = sys.argv[1], sys.argv[2]
    out_path = sys.argv[3] if len(sys.argv) > 3 else None
    commit = sys.argv[4] if len(sys.argv) > 4 else None
    out = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(out.splitlines()))
    h = rows[0]
    per = {}
    for r in rows[2:]:
        d = dict(zip(h, r))
        name = d["Kernel Name"]
        b = (float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"]))
        unit = rows[1][h.index("dram__bytes_read.sum")]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[unit]
        for pat, entry in ENTRY:
            if pat in name:
                e = per.setdefault(entry, {"dram_bytes": 0.0, "launches": 0, "gpu_time_ms": 0.0})
                e["dram_bytes"] += b * scale
                e["launches"] += 1
                e["gpu_time_ms"] += float(d["gpu__time_duration.sum"]) * (1.0 if rows[1][h.index("gpu__time_duration.sum")] == "ms" else 1e-3)
                break
    path = out_path or os.path.join(ROOT, "profiles", "ncu_traffic.json")
    allk = json.load(open(path)) if os.path.exists(path) else {}
    for e in per.values():  # the roofline is per launch: average over the launches of the capture
        e["dram_bytes_per_launch"] = e["dram_bytes"] / e["launches"]
        e["gpu_time_ms_per_launch"] = e["gpu_time_ms"] / e["launches"]
    allk[key] = {"source": os.path.basename(rep), "commit": commit, "per_entry_point": per}
    with open(path, "w") as f:
        json.dump(allk, f, indent=1, sort_keys=True)
    print(json.dumps(allk[key], indent=1))


if __name__ == "__main__":
    main()
