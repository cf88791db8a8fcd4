#!/usr/bin/env python
"""Dump a SASS index range with samples: python scripts/ncu_sass.py rep lo hi"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
lo, hi = int(sys.argv[2]), int(sys.argv[3])
for i in range(lo, min(hi, len(rd))):
    r = rd[i]
    print("%5d %6s %8s  %s" % (i, r["# Samples"], r["Instructions Executed"], r["Source"].strip()[:100]))
