#!/usr/bin/env python
"""SASS opcode histogram per kernel of the built library (cuobjdump -sass, no GPU needed).
usage: sass_opcodes.py [lib.so] > profiles/rN_sass_opcodes.txt
Lists, per kernel, the instruction count and the opcodes that prove what the kernel is built from:
UTCHMMA / UTCBAR / LDTM (tcgen05.mma, its commit barrier, tcgen05.ld from TMEM), UBLKCP (cp.async.bulk = TMA 1-D),
SYNCS (mbarrier), LDGSTS (cp.async), REDUX (redux.sync), DADD/DMUL/DFMA (binary64), MUFU, ATOMS/ATOMG/RED."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "nav-slam_b200", "_build", "libnavslam_b200.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEY = ["UTCHMMA", "UTCBAR", "LDTM", "UTCCP", "UBLKCP", "SYNCS", "LDGSTS", "REDUX", "DADD", "DMUL", "DFMA", "DSETP",
       "MUFU", "F2F", "ATOMS", "ATOMG", "RED", "LDG", "STG", "LDS", "STS", "SHFL", "VOTE", "BAR", "ACQBULK", "ELECT"]
kern, hist = None, collections.OrderedDict()
for line in txt.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        kern = re.sub(r"\(.*", "", kern)
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
print(f"# {os.path.relpath(lib, ROOT)}: cuobjdump -sass, opcode counts per kernel (static instructions)")
for k, h in hist.items():
    tot = sum(h.values())
    sel = " ".join(f"{o}={h[o]}" for o in KEY if h.get(o))
    print(f"{k}\n    total={tot}  {sel}")
