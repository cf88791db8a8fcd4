#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from an ncu report (source page), grouped with a little context.
   python scripts/ncu_hot.py gpurun_out/x.ncu-rep [topN]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
tot = sum(int(r["# Samples"] or 0) for r in rd)
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
idx = sorted(range(len(rd)), key=lambda i: -int(rd[i]["# Samples"] or 0))[:top]
print("total samples", tot, "instructions", len(rd))
for i in sorted(idx):
    r = rd[i]
    print("%5d %6.2f%%  %s" % (i, 100.0 * int(r["# Samples"]) / max(1, tot), r["Source"].strip()[:110]))
