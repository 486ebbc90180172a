#!/usr/bin/env python
"""Digest ncu exports into the small text/JSON summaries committed under profiles/.

  profile_digest.py launches <launches.csv> <out.md>        per-kernel totals/shares of an ncu launch list
  profile_digest.py kernel <raw.csv> <source.csv> <out.md> [perms_per_launch]
                                                             pipe/stall/opcode digest of one `ncu --set full` capture
"""
import collections
import csv
import json
import sys


def launches(path, out):
    rows = list(csv.reader(open(path)))
    for i, r in enumerate(rows):
        if "Kernel Name" in r:
            h, start = r, i + 1
            break
    ki, vi, gi, bi = h.index("Kernel Name"), h.index("Metric Value"), h.index("Grid Size"), h.index("Block Size")
    seq = [(r[ki].split("(")[0], float(r[vi].replace(",", "")), r[gi], r[bi]) for r in rows[start:] if len(r) > vi]
    agg = collections.OrderedDict()
    for k, v, _, _ in seq:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v for _, v, _, _ in seq)
    with open(out, "w") as f:
        f.write(f"# ncu launch list digest ({path})\n\n")
        f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` — cold-cache, serialised per-launch times;\n"
                "use the SHARES, not the absolutes.\n\n")
        f.write(f"{len(seq)} launches, {tot / 1e6:.3f} ms total\n\n| kernel | launches | total ms | avg us | share |\n|---|---|---|---|---|\n")
        for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write(f"| `{k}` | {n} | {t / 1e6:.3f} | {t / n / 1e3:.1f} | {100 * t / tot:.1f}% |\n")
        # one commit = the launches between two consecutive leaf_hash launches
        idx = [i for i, s in enumerate(seq) if "leaf_hash" in s[0]]
        if len(idx) >= 3:
            a, b = idx[-2], idx[-1]
            # commit boundaries: from the first NTT launch after the previous tree's last level kernel
            j = a
            while j > 0 and "level_hash" not in seq[j - 1][0] and "fused_top" not in seq[j - 1][0]:
                j -= 1
            k2 = b
            while k2 > 0 and "level_hash" not in seq[k2 - 1][0] and "fused_top" not in seq[k2 - 1][0]:
                k2 -= 1
            f.write("\n## one commit (launch order)\n\n| kernel | grid | block | us |\n|---|---|---|---|\n")
            for s in seq[j:k2]:
                f.write(f"| `{s[0]}` | {s[2]} | {s[3]} | {s[1] / 1e3:.1f} |\n")
            f.write(f"\ncommit total under ncu: {sum(s[1] for s in seq[j:k2]) / 1e6:.3f} ms\n")


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__cycles_elapsed.avg", "l1tex__t_sector_hit_rate.pct",
        "lts__t_sector_hit_rate.pct", "sass__inst_executed_register_spilling", "smsp__sass_inst_executed_op_local_ld.sum",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]


def kernel(raw, src, out, perms=None):
    rows = list(csv.reader(open(raw)))
    h, u = rows[0], rows[1]
    with open(out, "w") as f:
        for r in rows[2:]:
            name = r[h.index("Kernel Name")]
            f.write(f"# ncu --set full digest: `{name[:120]}`\n\n| metric | unit | value |\n|---|---|---|\n")
            vals = {}
            for i, k in enumerate(h):
                if k in KEYS or k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio"):
                    if r[i] not in ("", "n/a"):
                        f.write(f"| {k} | {u[i]} | {r[i]} |\n")
                        vals[k] = r[i]
            try:
                rd, wr = float(vals["dram__bytes_read.sum"]), float(vals["dram__bytes_write.sum"])
                mult = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1}
                rd *= mult[u[h.index("dram__bytes_read.sum")]]
                wr *= mult[u[h.index("dram__bytes_write.sum")]]
                f.write(f"\nDRAM traffic per launch: {rd + wr:.0f} bytes\n")
                json.dump({"kernel": name, "dram_bytes_per_launch": rd + wr, "source": raw},
                          open(out.replace(".md", "_traffic.json"), "w"))
            except Exception:
                pass
        if src:
            srows = list(csv.reader(open(src)))
            sh = srows[1]
            si, ei = sh.index("Source"), sh.index("Instructions Executed")
            cnt = collections.Counter()
            for r in srows[2:]:
                if len(r) <= ei:
                    continue
                t = r[si].split()
                op = t[0] if not t[0].startswith("@") else t[1]
                cnt[op] += int(r[ei])
            tot = sum(cnt.values())
            f.write(f"\n## executed SASS opcode histogram ({len(srows) - 2} static instructions, {tot} warp instructions)\n\n")
            f.write("| opcode | share |" + (" per warp-permutation |" if perms else "") + "\n|---|---|" + ("---|" if perms else "") + "\n")
            for op, e in cnt.most_common(24):
                f.write(f"| {op} | {100 * e / tot:.2f}% |" + (f" {e / (perms / 32):.0f} |" if perms else "") + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3])
    else:
        kernel(sys.argv[2], sys.argv[3] if sys.argv[3] != "-" else None, sys.argv[4],
               float(sys.argv[5]) if len(sys.argv) > 5 else None)
