#!/usr/bin/env python
"""Summarise `ncu --page source --csv --print-source sass` output: top stalled SASS instructions per kernel."""
import csv
import sys

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
sections = []
hdr = None
for r in csv.reader(open(path)):
    if not r:
        continue
    if r[0] == "Kernel Name":
        sections.append({"name": r[1], "rows": []})
    elif r[0] == "Address":
        hdr = r
    elif sections and hdr and len(r) >= len(hdr) - 2:
        sections[-1]["rows"].append(r)
isrc, isamp, iex = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
for sec in sections:
    data = [(int(r[isamp] or 0), n, r) for n, r in enumerate(sec["rows"])]
    tot = sum(d[0] for d in data) or 1
    print("=====", sec["name"][:100], "samples", tot, "sass lines", len(data))
    for s, n, r in sorted(data, reverse=True)[:top]:
        st = sorted(((int(r[i] or 0), hdr[i]) for i in stall_cols if i < len(r)), reverse=True)[:2]
        print(f"{100*s/tot:5.1f}%  #{n:4d} exec={r[iex]:>9s} {r[isrc].strip()[:64]:64s} {st}")
