#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of liboptimet_b200.so (static, from cuobjdump -sass; no GPU needed):
   python scripts/sass_summary.py > profiles/sass_summary.txt
UBLKCP = cp.async.bulk (TMA bulk copy engine), SYNCS = mbarrier operations, DMMA = FP64 tensor-core MMA (8x8x4),
DFMA/DMUL/DADD = FP64 pipe, LDS/STS = shared-memory accesses, LDG/STG = global."""
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "optimet_b200", "liboptimet_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
keys = ["UBLKCP", "SYNCS", "DMMA", "DFMA", "DMUL", "DADD", "LDS", "STS", "LDG", "STG", "SHFL", "BAR", "UTMALDG", "UTCHMMA"]
kern = OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        kern[cur] = dict.fromkeys(keys, 0)
        kern[cur]["total"] = 0
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and cur:
        op = m.group(1)
        kern[cur]["total"] += 1
        for k in keys:
            if op.startswith(k):
                kern[cur][k] += 1
demangled = subprocess.run(["c++filt"], input="\n".join(kern.keys()), capture_output=True, text=True).stdout.splitlines()
print("# static SASS counts per kernel, %s (sm_100a)" % os.path.basename(lib))
print("%-78s %6s " % ("kernel", "total") + " ".join("%7s" % k for k in keys))
for (name, c), dn in zip(kern.items(), demangled):
    short = re.sub(r"\(.*", "", dn)[:78]
    print("%-78s %6d " % (short, c["total"]) + " ".join("%7d" % c[k] for k in keys))
