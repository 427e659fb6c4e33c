#!/usr/bin/env python
"""Summarise ncu artefacts brought back from the GPU box into small text files for profiles/.

    python tools/ncu_summary.py launches gpurun_out/x_launches.csv  > profiles/rNN_launches.txt
    python tools/ncu_summary.py full     gpurun_out/x_prof.ncu-rep  > profiles/rNN_full.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tensor.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "lts__t_sector_hit_rate.pct", "lts__t_bytes.sum", "smsp__cycles_active.avg"]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    rows = rows[rows.index(hdr) + 1:]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(r[ui], v / 1e3)
        agg.setdefault(r[ki].split("(")[0][:70], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print("# %s: gpu__time_duration.sum per launch (ncu, cold-cache, serialised) -- compare SHARES" % path)
    print("%-72s %6s %12s %12s %7s" % ("kernel", "n", "total_us", "avg_us", "share"))
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print("%-72s %6d %12.1f %12.1f %6.1f%%" % (k, len(v), sum(v), sum(v) / len(v), 100 * sum(v) / tot))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    seen = collections.OrderedDict()
    for r in rows[2:]:
        seen.setdefault(r[ki], []).append(r)
    print("# %s: ncu --set full --clock-control none, one block per distinct kernel (first captured launch; n = launches captured)" % path)
    for k, rs in seen.items():
        print("\n== %s   (n=%d)" % (k, len(rs)))
        r = rs[0]
        for key in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print("  %-70s %s %s" % (key, r[i], units[i]))


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
