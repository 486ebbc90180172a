#!/usr/bin/env python
"""Per-instruction view of an ncu source page CSV (ncu -i X.ncu-rep --page source --csv):
aggregates stall samples by opcode and by contiguous code region (between branch targets)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
col = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot_samples = sum(int(r[col["# Samples"]] or 0) for r in data)
by_op = collections.Counter(); by_op_exec = collections.Counter()
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
stall_tot = collections.Counter()
for r in data:
    src = r[col["Source"]].strip()
    op = src.split()[0] if not src.startswith("@") else src.split()[1]
    op = op.split(".")[0] + ("." + op.split(".")[1] if op.startswith("IMAD.") and len(op.split(".")) > 1 else "")
    n = int(r[col["# Samples"]] or 0)
    by_op[op] += n
    by_op_exec[op] += int(r[col["Instructions Executed"]] or 0)
    for s in stall_cols:
        stall_tot[s] += int(r[col[s]] or 0)
print("total samples", tot_samples)
print("stalls:", {k: v for k, v in stall_tot.most_common(8)})
tot_exec = sum(by_op_exec.values())
print(f"{'op':14s} {'samples%':>8s} {'exec%':>7s} {'samples/exec(rel)':>18s}")
for op, n in by_op.most_common(18):
    e = by_op_exec[op]
    print(f"{op:14s} {100*n/tot_samples:8.1f} {100*e/tot_exec:7.1f} {(n/tot_samples)/(e/tot_exec) if e else 0:18.2f}")
# regions: split at every 64 instructions for a coarse heat map
if "-regions" in sys.argv:
    step = 64
    for i in range(0, len(data), step):
        blk = data[i:i+step]
        n = sum(int(r[col["# Samples"]] or 0) for r in blk); e = sum(int(r[col["Instructions Executed"]] or 0) for r in blk)
        print(f"instr {i:5d}-{i+len(blk):5d}: samples {100*n/tot_samples:5.1f}%  exec {100*e/tot_exec:5.1f}%  first: {blk[0][col['Source']].strip()[:50]}")
