#!/usr/bin/env python
"""ncu CSV with dram__bytes_read.sum / dram__bytes_write.sum / gpu__time_duration.sum per launch -> per-kernel totals (JSON).

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc ... --csv --log-file x.csv <cmd>
    python tools/traffic_summary.py x.csv > profiles/rNN_traffic.json
"""
import collections
import csv
import json
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}


def main(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    rows = rows[rows.index(hdr) + 1:]
    ki, mi, vi, ui, ii = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "Metric Unit", "ID"))
    per = collections.OrderedDict()
    for r in rows:
        try:
            v = float(r[vi].replace(",", "")) * UNIT.get(r[ui], 1.0)
        except ValueError:
            continue
        k = r[ki].split("(")[0]
        d = per.setdefault(k, {"launches": set(), "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "time_us": 0.0})
        d["launches"].add(r[ii])
        if r[mi] == "dram__bytes_read.sum":
            d["dram_read_bytes"] += v
        elif r[mi] == "dram__bytes_write.sum":
            d["dram_write_bytes"] += v
        elif r[mi] == "gpu__time_duration.sum":
            d["time_us"] += v
    out = {}
    for k, d in per.items():
        n = len(d["launches"])
        out[k] = {"launches": n, "dram_read_bytes": d["dram_read_bytes"], "dram_write_bytes": d["dram_write_bytes"],
                  "dram_bytes_per_launch": (d["dram_read_bytes"] + d["dram_write_bytes"]) / max(1, n), "time_us": d["time_us"]}
    return out


if __name__ == "__main__":
    res = main(sys.argv[1])
    if len(sys.argv) > 2 and sys.argv[2] == "--bench-json":
        # the file bench.py reads (profiles/conv_tc_traffic.json): the dominant kernel's per-launch DRAM bytes, tagged with the mode, batch and
        # the hash of the kernel sources it was captured from:  traffic_summary.py x.csv --bench-json <mode name> <batch> "<command>"
        import os
        sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        from glare_b200.build import source_hash
        k = next(k for k in res if "conv_tc_kernel" in k)
        tot = {"launches": 0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "time_us": 0.0}
        for kk, d in res.items():
            if "conv_tc_kernel" in kk:
                for f in tot:
                    tot[f] += d[f]
        json.dump({"kernel": "conv_tc_kernel", "mode": sys.argv[3], "batch": int(sys.argv[4]), "command": sys.argv[5] if len(sys.argv) > 5 else None,
                   "csrc_sha16": source_hash(), "launches": tot["launches"],
                   "dram_bytes_per_launch": (tot["dram_read_bytes"] + tot["dram_write_bytes"]) / max(1, tot["launches"]),
                   "dram_bytes_per_step": tot["dram_read_bytes"] + tot["dram_write_bytes"], "time_us_under_ncu": tot["time_us"]}, sys.stdout, indent=1)
    else:
        json.dump(res, sys.stdout, indent=1)
