#!/usr/bin/env python
"""SASS instruction mix of one kernel from `ncu --page source --csv` (executed warp instructions
and stall samples per opcode).  python profiles/instmix.py <src.csv>"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
ops, samp, tot = collections.Counter(), collections.Counter(), 0
for r in rows:
    if "Source" in r and "Instructions Executed" in r:
        hdr = r
        ia, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ie:
        continue
    m = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia].strip())
    if not m:
        continue
    try:
        n = int(r[ie] or 0)
    except ValueError:
        continue
    op = m.group(2).split(".")[0]
    ops[op] += n
    tot += n
    samp[op] += int(r[isamp] or 0)
print("total warp instructions", tot)
for op, n in ops.most_common(24):
    print(f"{op:10s} {n:12d} {100 * n / tot:5.1f}%  stall samples {samp[op]}")
