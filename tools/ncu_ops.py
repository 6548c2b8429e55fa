#!/usr/bin/env python
"""Executed-instruction histogram by opcode, and the source lines that execute the most of given opcodes.
usage: python tools/ncu_ops.py <rep> <nstates> [OPCODE ...]"""
import csv, io, re, subprocess, sys, collections
rep, nstates = sys.argv[1], float(sys.argv[2])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
cur, h, curline = None, None, None
per = collections.defaultdict(collections.Counter)
ops, samp = collections.Counter(), collections.Counter()
for r in rows:
    if r and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if r and r[0] == "Line No":
        h = r; iI, iS = h.index("Instructions Executed"), h.index("# Samples"); continue
    if not h or len(r) <= iI:
        continue
    if r[0] != "":
        curline = (cur, int(r[0])); continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3])
    if not m or not curline:
        continue
    try:
        i, s = int(r[iI] or 0), int(r[iS] or 0)
    except ValueError:
        continue
    op = m.group(2).split(".")[0]
    if m.group(2).startswith("IMAD.MOV"):
        op = "IMAD.MOV"
    ops[op] += i; samp[op] += s; per[op][curline] += i
tot, ts = sum(ops.values()), sum(samp.values()) or 1
print(f"total {tot / nstates:.0f} warp-instructions per state")
for op, c in ops.most_common(28):
    print(f"{op:10s} {c / nstates:8.0f}/state {100 * c / tot:5.1f}%  samples {100 * samp[op] / ts:5.1f}%")
for op in sys.argv[3:]:
    print(op, [(f"{k[0].replace('wbc_', '').replace('.cuh', '')}:{k[1]}", round(v / nstates)) for k, v in per[op].most_common(16)])
