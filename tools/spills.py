#!/usr/bin/env python
"""Static spill (LDL/STL) instructions per source line of one kernel.  usage: python tools/spills.py <lib.so> <kernel-substring>"""
import collections, os, re, subprocess, sys, tempfile
lib, kern = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
cur, infn, seq, n = None, False, [], 0
for line in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m:
        infn = kern in m.group(1); continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)).replace("wbc_", "").replace(".cuh", ""), int(m.group(2))); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        n += 1
        if "LDL" in line or "STL" in line:
            seq.append((cur, "L" if "LDL" in line else "S"))
print("instructions", n, "spill instrs", len(seq))
out, prev = [], None
for c, k in seq:
    key = (c, k)
    if key == prev: out[-1][2] += 1
    else: out.append([c, k, 1])
    prev = key
print(" ".join(f"{c[0]}:{c[1]}{k}x{m}" for c, k, m in out))
