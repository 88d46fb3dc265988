#!/usr/bin/env python
"""Compare the --report_diffusion lines ("<ts> var <v> sum <s> old ...", driver.c:92-95) of two
runs of the same command line: the drop-in and the reference.  usage: a.txt b.txt [rtol]"""
import re
import sys

pat = re.compile(r"^(\d+) var (\d+) sum (\S+) old")


def sums(path):
    out = []
    for line in open(path, errors="replace"):
        m = pat.match(line.strip())
        if m:
            out.append((int(m.group(1)), int(m.group(2)), float(m.group(3))))
    return out


a, b = sums(sys.argv[1]), sums(sys.argv[2])
rtol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-12
assert len(a) == len(b) and len(a) > 0, (len(a), len(b))
worst = 0.0
for (ta, va, sa), (tb, vb, sb) in zip(a, b):
    assert (ta, va) == (tb, vb), ((ta, va), (tb, vb))
    worst = max(worst, abs(sa - sb)/abs(sb))
print(f"{len(a)} check sums compared, largest relative difference {worst:.3g} (printed with 6 decimals)")
assert worst <= rtol, worst
