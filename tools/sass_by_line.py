#!/usr/bin/env python3
"""SASS instruction count per source line of one kernel (needs -lineinfo):  sass_by_line.py file.o <mangled-substring>"""
import collections
import re
import subprocess
import sys
import tempfile
import os

obj, key = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
cub = [os.path.join(d, f) for f in os.listdir(d) if f.endswith(".cubin")][0]
txt = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout
on, cur = False, None
cnt, ops = collections.Counter(), collections.defaultdict(collections.Counter)
for line in txt.splitlines():
    if line.startswith("//---") and ".text." in line:
        on = key in line
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", line)
    if m and cur:
        cnt[cur] += 1
        ops[cur][m.group(1) + (".WIDE" if ".WIDE" in m.group(2) else "")] += 1
tot = sum(cnt.values())
print("total", tot)
for k, v in sorted(cnt.items(), key=lambda kv: -kv[1])[:45]:
    print("%-18s %4d  %5.1f%%  %s" % ("%s:%d" % k, v, 100.0 * v / tot, dict(ops[k])))
