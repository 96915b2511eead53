#!/usr/bin/env python
"""Aggregate an ncu source-page export of the step kernel by CUDA source line (development tool).

    python tools/ncu_by_line.py <report.ncu-rep> <object.o> [top] [--fn NAME] [--mem]
Joins ncu's per-SASS-instruction samples with `nvdisasm -gi` line info of the same object (by instruction offset inside the kernel's
text section) and prints the source lines that collect the most stall samples, with the outermost non-inlined function each belongs
to.  --fn restricts to lines whose inline chain / file mentions NAME; --mem lists only local-memory instructions."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

args = [a for a in sys.argv[1:] if not a.startswith("--")]
rep, obj = args[0], args[1]
top = int(args[2]) if len(args) > 2 else 40
fn_filter = None
if "--fn" in sys.argv:
    fn_filter = sys.argv[sys.argv.index("--fn") + 1]
    args = [a for a in args if a != fn_filter]
    top = int(args[2]) if len(args) > 2 else 40
only_mem = "--mem" in sys.argv
kern = os.environ.get("NCU_KERNEL", "k_hru_step_nn3ILb1")

tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
insn = {}  # offset -> (file, line, chain, sass, symbol)
on, cur, sym = False, ("?", 0, ""), ""
for l in dis:
    if l.lstrip().startswith(".section"):
        on = (".text." in l) and (kern in l)
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        chain = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
        cur = (os.path.basename(m.group(1)), int(m.group(2)), " < ".join(f"{os.path.basename(f)}:{n}" for f, n in chain))
        continue
    m = re.match(r"^(\$?[_A-Za-z][^:\s]*):\s*$", l)
    if m and not l.startswith(".L_"):
        sym = m.group(1)
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        insn[int(m.group(1), 16)] = cur + (m.group(2).strip(), sym)

src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + os.environ.get("NCU_KREGEX", "k_hru_step"), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
for k, r in enumerate(data):
    if r and r[0] == "Kernel Name":
        data = data[:k]
        break
ia, isrc = hdr.index("Address"), hdr.index("Source")
a0 = int(data[0][ia], 16)
cols = ["# Samples", "Instructions Executed", "Thread Instructions Executed", "stall_long_sb", "stall_wait", "stall_no_inst", "L2 Theoretical Sectors Local"]
ci = [hdr.index(c) for c in cols]
agg = collections.defaultdict(collections.Counter)
mism = 0
dem_cache = {}
for r in data:
    off = int(r[ia], 16) - a0
    info = insn.get(off)
    if info is None:
        mism += 1
        continue
    f, ln, chain, sass, sym = info
    if sass.split()[0].lstrip("@!P0123456789T ") != r[isrc].split(";")[0].strip().split()[0].lstrip("@!P0123456789T ") and mism < 5:
        pass
    if only_mem and not re.search(r"\b(LDL|STL)", sass):
        continue
    fnname = sym.split("$")[-1]
    key = (f, ln, fnname, chain)
    if fn_filter and fn_filter not in (f + chain + fnname):
        continue
    for c, i in zip(cols, ci):
        try:
            agg[key][c] += float(r[i] or 0)
        except ValueError:
            pass
names = list({k[2] for k in agg})
dem = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
dm = {n: d.split("(")[0].replace("vic::", "")[:28] for n, d in zip(names, dem)}
tot = collections.Counter()
for k in agg:
    tot.update(agg[k])
print(f"instructions without line info: {mism}; TOTAL", {c: int(tot[c]) for c in cols})
print("%-22s %-28s %6s %8s %5s %6s %6s %8s  %s" % ("file:line", "function", "samp%", "winst(k)", "thr/w", "longsb", "wait", "L2loc(k)", "inlined at"))
for key, c in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    f, ln, fnname, chain = key
    print("%-22s %-28s %6.2f %8.1f %5.1f %6.0f %6.0f %8.1f  %s" % (
        f"{f}:{ln}", dm.get(fnname, fnname), 100 * c["# Samples"] / max(tot["# Samples"], 1), c["Instructions Executed"] / 1e3,
        c["Thread Instructions Executed"] / max(c["Instructions Executed"], 1), c["stall_long_sb"], c["stall_wait"],
        c["L2 Theoretical Sectors Local"] / 1e3, chain[:90]))
