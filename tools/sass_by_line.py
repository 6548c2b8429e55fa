#!/usr/bin/env python
"""Static SASS instruction count per source line for one kernel (needs -lineinfo).
usage: python tools/sass_by_line.py <lib.so> <kernel-substring> [topN]"""
import collections, re, subprocess, sys, tempfile, os
lib, kern = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
cur, infn, cnt, total = None, False, collections.Counter(), 0
for line in txt.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", line)
    if m:
        infn = kern in m.group(1)
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line) and cur:
        cnt[cur] += 1
        total += 1
byfile = collections.Counter()
for (f, l), c in cnt.items():
    byfile[f] += c
print("total", total, dict(byfile.most_common(10)))
for (f, l), c in cnt.most_common(top):
    print(f"{c:6d}  {f}:{l}")
