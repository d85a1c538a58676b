#!/usr/bin/env python
"""Stall-reason totals and the hottest SASS lines of one kernel from `ncu --page source --csv`.
python profiles/stalls.py <src.csv> [n_lines]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
nl = int(sys.argv[2]) if len(sys.argv) > 2 else 12
hdr, tot, lines = None, {}, []
for r in rows:
    if "Source" in r and "# Samples" in r:
        hdr = r
        cols = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
        continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(hdr, r))
    try:
        n = int(d["# Samples"])
    except ValueError:
        continue
    best = ("", 0)
    for c in cols:
        try:
            v = int(d[c])
        except ValueError:
            continue
        tot[c] = tot.get(c, 0) + v
        if v > best[1]:
            best = (c, v)
    lines.append((n, d["Source"].strip()[:80], best[0]))
s = sum(tot.values())
print("samples", s, {k[6:]: round(100 * v / s, 1) for k, v in sorted(tot.items(), key=lambda x: -x[1])[:9]})
for n, src, why in sorted(lines, reverse=True)[:nl]:
    print(f"{100 * n / s:5.1f}%  {why[6:]:18s} {src}")
