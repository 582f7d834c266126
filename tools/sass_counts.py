#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of libacf_b200.so (what proves the Blackwell-native paths: UBLKCP / UTMALDG / SYNCS).
    python tools/sass_counts.py > profiles/r2_sass_counts.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "acfcalculator_b200", "libacf_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
WANT = ["DFMA", "DADD", "DMUL", "LDS", "STS", "LDG", "STG", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "BAR", "SHFL", "MUFU",
        "LDL", "STL", "HMMA", "UTC"]
kern, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("acfb200::", "").replace("void ", "")
        kern[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and name:
        op = m.group(1)
        kern[name]["total"] += 1
        for w in WANT:
            if op.startswith(w):
                kern[name][w] += 1
                break
print("%-58s %6s " % ("kernel (sm_100a SASS, %s)" % os.path.basename(lib), "instr") + " ".join("%7s" % w for w in WANT))
for k, c in kern.items():
    print("%-58s %6d " % (k[:58], c["total"]) + " ".join("%7s" % (c[w] or ".") for w in WANT))
