#!/usr/bin/env python
"""Count SASS instructions per kernel, split by issue pipe (FMA-pipe integer ops vs ALU-pipe ops).
usage: sasscount.py file.cubin|file.so [name-substring]"""
import re, subprocess, sys, collections
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
sel = sys.argv[2] if len(sys.argv) > 2 else ""
fn = None; counts = collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m: fn = m.group(1); counts[fn] = collections.Counter(); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and fn: counts[fn][m.group(2)] += 1
for fn, c in counts.items():
    if sel not in fn: continue
    fma = sum(v for k, v in c.items() if k.startswith(("IMAD", "FFMA", "FMUL", "FADD", "HFMA")))
    wide = sum(v for k, v in c.items() if k.startswith(("IMAD.WIDE", "IMAD.HI")))
    alu = sum(v for k, v in c.items() if k.startswith(("IADD3", "LOP3", "SHF", "ISETP", "SEL", "VIADD", "MOV", "PRMT", "LEA", "IABS", "FSEL", "VIMNMX", "P2R", "R2P", "PLOP3", "CS2R", "FMNMX")))
    tot = sum(c.values())
    print(f"{fn[:70]:70s} total={tot:5d} fma={fma:5d} (wide/hi={wide}) alu={alu:5d} other={tot-fma-alu}")
    if "-v" in sys.argv: print("   ", dict(c.most_common(14)))
