#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/ (run in the build container, no GPU needed).

  python scripts/ncu_summary.py launches gpurun_out/r1_launches_c4.csv > profiles/r1_launches_c4.txt
  python scripts/ncu_summary.py full gpurun_out/r1_matvec.ncu-rep   > profiles/r1_matvec_full.txt
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg", "lts__t_bytes.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
        "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sectors_op_write.sum", "lts__t_sectors_op_read.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
        "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct",
        "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]


def launches(path):
    rows = []
    with open(path) as f:
        txt = f.read()
    start = txt.find('"ID"')
    rd = csv.DictReader(io.StringIO(txt[start:]))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        scale = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(unit, 1e-6)
        rows.append((re.sub(r"\(.*", "", r["Kernel Name"]).strip(), v * scale))
    agg = OrderedDict()
    for name, ms in rows:
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ms
    total = sum(a[1] for a in agg.values())
    print("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)")
    print("# %d launches, %.3f ms summed" % (len(rows), total))
    print("%-60s %8s %12s %10s %7s" % ("kernel", "launches", "total_ms", "avg_ms", "share"))
    for name, (cnt, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-60s %8d %12.3f %10.4f %6.1f%%" % (name[:60], cnt, ms, ms / cnt, 100 * ms / total))


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr, units = rd[0], rd[1]
    print("# ncu --set full --clock-control none: %s" % path)
    for row in rd[2:]:
        d = dict(zip(hdr, row))
        u = dict(zip(hdr, units))
        print("== %s  grid=%s block=%s" % (d.get("Kernel Name", "?")[:80], d.get("Grid Size"), d.get("Block Size")))
        for k in KEYS:
            if k in d and d[k] != "":
                print("   %-75s %18s %s" % (k, d[k], u.get(k, "")))
        try:
            rdB = float(d["dram__bytes_read.sum"].replace(",", ""))
            wrB = float(d["dram__bytes_write.sum"].replace(",", ""))
            mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = rdB * mult.get(u["dram__bytes_read.sum"], 1) + wrB * mult.get(u["dram__bytes_write.sum"], 1)
            t = float(d["gpu__time_duration.sum"].replace(",", "")) * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(u["gpu__time_duration.sum"], 1e-9)
            print("   -> dram traffic per launch = %.6e B, %.1f GB/s under ncu" % (tot, tot / t / 1e9))
        except Exception as e:
            print("   (traffic: %s)" % e)


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
