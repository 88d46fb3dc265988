#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of libminiamr_b200.so (cuobjdump -sass; no GPU needed):
the Blackwell data-movement instructions (UBLKCP = 1-D bulk TMA, SYNCS = mbarrier, LDGSTS =
cp.async, ...) and the FP64 arithmetic of every kernel.  usage: sass_histogram.py [lib.so]"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "miniamr_b200", "libminiamr_b200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = kern.replace("(anonymous namespace)::", "").replace("mamr::", "")
        kern = re.sub(r"^void ", "", re.sub(r"\(.*", "", kern))
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
KEYS = ["UBLKCP", "SYNCS", "LDGSTS", "LDS", "STS", "LDG", "STG", "DADD", "DFMA", "DMUL", "SHFL", "BAR", "ATOM", "RED",
        "MEMBAR", "FENCE"]
print("arch:", re.findall(r"arch = (\S+)", out)[:1])
print(f"{'kernel':58s} {'total':>6s} " + " ".join(f"{k:>6s}" for k in KEYS))
for k, h in hist.items():
    print(f"{k[:58]:58s} {sum(h.values()):6d} " + " ".join(f"{h.get(x, 0):6d}" for x in KEYS))
