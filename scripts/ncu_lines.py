#!/usr/bin/env python
"""Per-source-line dynamic instruction counts and stall samples of one kernel: joins the SASS page of an .ncu-rep (executed
instructions and samples per address) with `nvdisasm -g` of the cubin (address -> file:line).
usage: ncu_lines.py <rep.csv from `ncu -i rep --page source --csv`> <nvdisasm -g -c output> <mangled-name substring> <units: launches*steps>"""
import collections
import csv
import re
import sys

csvf, disf, key, units = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
amap, fil, line, on = {}, None, None, False
for ln in open(disf):
    if ln.startswith("//----"):
        on = key in ln
        continue
    if not on:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        fil, line = m.group(1).split("/")[-1], int(m.group(2))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        amap[int(m.group(1), 16)] = (fil, line, m.group(2).strip())
rows = list(csv.reader(open(csvf)))
hi = [i for i, r in enumerate(rows) if "# Samples" in r][0]
ci = {n: i for i, n in enumerate(rows[hi])}
agg = collections.defaultdict(lambda: [0.0, 0.0])
ops = collections.Counter()
base = None
for r in rows[hi + 1:]:
    try:
        a = int(r[ci["Address"]], 16); ex = float(r[ci["Instructions Executed"]] or 0); sm = float(r[ci["# Samples"]] or 0)
    except Exception:
        continue
    base = a if base is None else base
    if a - base not in amap:
        continue
    f, l, ins = amap[a - base]
    agg[(f, l)][0] += ex / units
    agg[(f, l)][1] += sm
    t = ins.split()
    ops[(t[1] if t[0].startswith("@") else t[0]).split(".")[0]] += ex / units
tot = sum(v[0] for v in agg.values()); tots = sum(v[1] for v in agg.values())
print(f"instructions per unit: {tot:.1f}")
print("opcode mix per unit:", ", ".join(f"{k} {v:.0f}" for k, v in ops.most_common(18)))
src = {}
import os
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "differentialdynamicprogramming.jl_b200", "csrc")
for (f, l), v in sorted(agg.items(), key=lambda t: -t[1][0])[:top]:
    if f not in src:
        try:
            src[f] = open(os.path.join(root, f)).read().splitlines()
        except Exception:
            src[f] = []
    text = src[f][l - 1].strip()[:100] if 0 < l <= len(src[f]) else ""
    print(f"{f[:16]:16s}:{l:4d} {v[0]:7.1f} instr {v[1] / tots * 100:5.1f}% samples | {text}")
