#!/usr/bin/env python
"""Split the SASS of an ncu report at BAR.SYNC instructions and print, per section, executed warp instructions,
stall samples, DMMA/DFMA counts and shared wavefronts: python scripts/ncu_phases.py rep"""
import csv, io, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rd = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
stall_keys = [k for k in rd[0].keys() if k.startswith("stall_") and "Not Issued" not in k]
sec = []
cur = dict(first=0, inst=0.0, samples=0.0, wf=0.0, dmma=0.0, dfma=0.0, lds=0.0, sts=0.0, n=0, stalls={k: 0.0 for k in stall_keys})
for i, r in enumerate(rd):
    ex = float(r["Instructions Executed"] or 0)
    cur["inst"] += ex
    cur["samples"] += float(r["# Samples"] or 0)
    cur["wf"] += float(r["L1 Wavefronts Shared"] or 0)
    src = r["Source"]
    if "DMMA" in src: cur["dmma"] += ex
    if "DFMA" in src or "DMUL" in src or "DADD" in src: cur["dfma"] += ex
    if "LDS" in src: cur["lds"] += ex
    if "STS" in src: cur["sts"] += ex
    for k in stall_keys:
        cur["stalls"][k] += float(r[k] or 0)
    cur["n"] += 1
    if "BAR.SYNC" in src or i == len(rd) - 1:
        cur["last"] = i
        sec.append(cur)
        cur = dict(first=i + 1, inst=0.0, samples=0.0, wf=0.0, dmma=0.0, dfma=0.0, lds=0.0, sts=0.0, n=0, stalls={k: 0.0 for k in stall_keys})
tot = sum(s["inst"] for s in sec); tots = sum(s["samples"] for s in sec)
print("total warp instructions %.3e, samples %d" % (tot, tots))
for s in sec:
    top = sorted(s["stalls"].items(), key=lambda kv: -kv[1])[:4]
    print("sass %4d-%4d: inst %5.1f%% samples %5.1f%% wf %.2e dmma %.2e fp64 %.2e lds %.2e sts %.2e | %s" % (
        s["first"], s["last"], 100 * s["inst"] / tot, 100 * s["samples"] / max(1, tots), s["wf"], s["dmma"], s["dfma"], s["lds"], s["sts"],
        ", ".join("%s %.0f%%" % (k[6:], 100 * v / max(1, s["samples"])) for k, v in top)))
