#!/usr/bin/env python
"""Instruction count of every kernel in a shared library (cuobjdump -sass).

    python tools/sass_size.py kaldi-decoder_b200/lib/libkd_b200.so [substring]
"""
import re
import subprocess
import sys

lib = sys.argv[1]
want = sys.argv[2] if len(sys.argv) > 2 else ""
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
name, n = None, 0
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        if name and want in name:
            print(f"{n:6d}  {name}")
        name, n = m.group(1), 0
    elif re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", line):
        n += 1
if name and want in name:
    print(f"{n:6d}  {name}")
