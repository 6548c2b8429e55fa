#!/usr/bin/env python
"""Dynamic instructions / stall samples per source file and per line range of an .ncu-rep.
usage: python tools/ncu_regions.py <rep> <nstates> [file:lo-hi:name ...]"""
import csv, io, subprocess, collections, sys
rep, nstates = sys.argv[1], float(sys.argv[2])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur, h = None, None
tot, samp, lines = collections.Counter(), collections.Counter(), collections.defaultdict(list)
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        h = r
        iS, iI = h.index("# Samples"), h.index("Instructions Executed")
        continue
    if h and len(r) > 10 and r[2] == "-":
        try:
            i, s = int(r[iI] or 0), int(r[iS] or 0)
        except ValueError:
            continue
        tot[cur] += i
        samp[cur] += s
        lines[cur].append((i, s, int(r[0])))
T, S = sum(tot.values()), sum(samp.values())
print(f"total {T/nstates:.0f} warp-instructions per state, {S} samples")
for f in tot:
    print(f"{f:28s} {tot[f]/nstates:9.0f} inst/state {100*tot[f]/T:5.1f}%   samples {100*samp[f]/S:5.1f}%")
for spec in sys.argv[3:]:
    f, rg, name = spec.split(":")
    a, b = map(int, rg.split("-"))
    ii = sum(i for i, s, l in lines[f] if a <= l <= b)
    ss = sum(s for i, s, l in lines[f] if a <= l <= b)
    print(f"{f:16s} {name:24s} {ii/nstates:8.0f} inst/state {100*ss/S:5.1f}% samples")
