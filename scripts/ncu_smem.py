#!/usr/bin/env python
"""Per-SASS-instruction shared-memory wavefronts of an ncu report: python scripts/ncu_smem.py rep [min_wavefronts]"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1
tot = sum(float(r["L1 Wavefronts Shared"] or 0) for r in rd)
print("total shared wavefronts", tot)
for i, r in enumerate(rd):
    w = float(r["L1 Wavefronts Shared"] or 0)
    if w >= thr:
        print("%5d exec %10s wf %12.0f ideal %12s (%.1f/inst) samples %6s  %s" % (i, r["Instructions Executed"], w, r["L1 Wavefronts Shared Ideal"], w / max(1.0, float(r["Instructions Executed"])), r["# Samples"], r["Source"].strip()[:90]))
