#!/usr/bin/env python
"""Static SASS instruction count per source line / per source region of one kernel.

    python tools/sass_by_line.py LIB.so KERNEL_SUBSTRING [--regions]
Needs the library built with -lineinfo.  Uses cuobjdump -xelf + nvdisasm -g.
"""
import collections
import os
import re
import subprocess
import sys
import tempfile

lib, want = os.path.abspath(sys.argv[1]), sys.argv[2]
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
out = subprocess.run(["nvdisasm", "-g", cubin], cwd=tmp, capture_output=True, text=True).stdout
per_line = collections.Counter()
cur, on, total = None, False, 0
for line in out.splitlines():
    if line.startswith(".text."):
        on = want in line
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        per_line[cur] += 1
        total += 1
print("total", total)
if "--regions" in sys.argv:
    # bucket by 25 source lines
    reg = collections.Counter()
    for (f, l), n in per_line.items():
        reg[(f, l // 25 * 25)] += n
    for (f, l), n in sorted(reg.items()):
        print(f"{f}:{l:5d}-{l+24:5d} {n:6d}")
else:
    for (f, l), n in sorted(per_line.items(), key=lambda kv: -kv[1])[:60]:
        print(f"{f}:{l:5d} {n:6d}")
