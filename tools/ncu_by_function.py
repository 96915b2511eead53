#!/usr/bin/env python
"""Aggregate an ncu source-page export of the step kernel by device function (development tool).

    python tools/ncu_by_function.py <report.ncu-rep> <object.o> [top]
Uses the symbol table of the object file (cuobjdump -elf) to attribute SASS addresses to the out-of-line device functions."""
import collections
import csv
import io
import subprocess
import os
import sys

rep, obj = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 32
elf = subprocess.run(["cuobjdump", "-elf", obj], capture_output=True, text=True).stdout
syms, on = [], False
for l in elf.splitlines():
    if l.startswith(".section .symtab"):
        on = True
    elif l.startswith(".section") and on:
        break
    elif on:
        p = l.split()
        # out-of-line device functions are named $<kernel>$<function>; keep those of the single-record kernel k_hru_step_nn*
        if len(p) >= 7 and p[3] == "0x2" and p[6].startswith("$") and os.environ.get("NCU_KERNEL", "k_hru_step_nn3ILb1") in p[6].split("$")[1]:
            syms.append((int(p[1], 16), int(p[2], 16), p[6]))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:k_hru_step", "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
for k, r in enumerate(data):  # a report with several launches repeats the two header rows: keep the first launch
    if r and r[0] == "Kernel Name":
        data = data[:k]
        break
ia = hdr.index("Address")
a0 = int(data[0][ia], 16)
cols = ["# Samples", "Instructions Executed", "Thread Instructions Executed", "stall_long_sb", "stall_no_inst", "stall_wait", "stall_branch_resolving",
        "stall_short_sb", "stall_math", "L2 Theoretical Sectors Local", "L2 Theoretical Sectors Global"]
ci = [hdr.index(c) for c in cols]


def find(off):
    for v, sz, n in syms:
        if v <= off < v + sz and "$" in n:
            return n.split("$")[-1]
    return "(kernel body)"


agg = collections.defaultdict(collections.Counter)
for r in data:
    fn = find(int(r[ia], 16) - a0)
    for c, i in zip(cols, ci):
        try:
            agg[fn][c] += float(r[i] or 0)
        except ValueError:
            pass
tot = collections.Counter()
for fn in agg:
    tot.update(agg[fn])
dem = subprocess.run(["c++filt"], input="\n".join(agg.keys()), capture_output=True, text=True).stdout.splitlines()
names = {k: d.split("(")[0][:46] for k, d in zip(agg.keys(), dem)}
print("TOTAL", {c: int(tot[c]) for c in cols})
print("%-46s %6s %8s %5s %6s %6s %6s %6s %6s %8s" % ("function", "samp%", "winst(M)", "thr/w", "longsb", "noinst", "wait", "brres", "shsb", "L2loc(M)"))
for fn, c in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    print("%-46s %6.2f %8.2f %5.1f %6.0f %6.0f %6.0f %6.0f %6.0f %8.2f" % (
        names[fn], 100 * c["# Samples"] / tot["# Samples"], c["Instructions Executed"] / 1e6,
        c["Thread Instructions Executed"] / max(c["Instructions Executed"], 1), c["stall_long_sb"], c["stall_no_inst"], c["stall_wait"],
        c["stall_branch_resolving"], c["stall_short_sb"], c["L2 Theoretical Sectors Local"] / 1e6))
