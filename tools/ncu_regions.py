#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` SASS listing into loop regions: instructions executed, FP64-pipe share and stall
samples per innermost loop (loops = backward branches), so that the cost of the hot loop, the near band, the edges and the
chunk prologues can be read off one capture.   python tools/ncu_regions.py listing.csv [min_share]"""
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
# a listing may hold several kernels, each introduced by a "Kernel Name" row: take the one whose name contains argv[3] (default:
# the first)
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
want = sys.argv[3] if len(sys.argv) > 3 else ""
pick = next((i for i in starts if want in rows[i][1]), starts[0])
end = next((i for i in starts if i > pick), len(rows))
print("kernel:", rows[pick][1])
rows = rows[pick:end]
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
ins = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    ins.append(dict(addr=int(r[ix["Address"]], 16), sass=r[ix["Source"]].strip(), n=int(r[ix["Instructions Executed"]] or 0),
                    samp=int(r[ix["# Samples"]] or 0)))
base = ins[0]["addr"]
pos = {d["addr"]: i for i, d in enumerate(ins)}
loops = []
for i, d in enumerate(ins):
    m = re.search(r"\bBRA\b.*?(0x[0-9a-f]+)", d["sass"])
    if m:
        t = int(m.group(1), 16)
        if t in pos and pos[t] <= i:
            loops.append((pos[t], i))
# innermost loop of every instruction
owner = [None] * len(ins)
for lo, hi in sorted(loops, key=lambda x: x[1] - x[0], reverse=True):
    for k in range(lo, hi + 1):
        owner[k] = (lo, hi)
tot_n = sum(d["n"] for d in ins)
tot_s = sum(d["samp"] for d in ins)
isfp = lambda s: re.match(r"(@!?U?P\d+\s+)?D(FMA|MUL|ADD|SETP|MNMX)", s) is not None
agg = {}
for d, o in zip(ins, owner):
    a = agg.setdefault(o, dict(n=0, fp=0, samp=0, mufu=0, lds=0, cnt=0))
    a["n"] += d["n"]; a["samp"] += d["samp"]; a["cnt"] += 1
    if isfp(d["sass"]): a["fp"] += d["n"]
    if "MUFU" in d["sass"]: a["mufu"] += d["n"]
    if "LDS" in d["sass"]: a["lds"] += d["n"]
print(f"total warp instructions {tot_n:.3e}, FP64 {sum(a['fp'] for a in agg.values()):.3e}, samples {tot_s}")
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.3
print(f"{'region (instr idx)':>22s} {'static':>6s} {'inst %':>7s} {'fp64 %':>7s} {'issue-cyc %':>11s} {'samples %':>9s}  fp64/inst")
cyc_tot = sum(a["n"] + a["fp"] for a in agg.values())
for o, a in sorted(agg.items(), key=lambda kv: -(kv[1]["n"] + kv[1]["fp"])):
    share = 100.0 * (a["n"] + a["fp"]) / cyc_tot
    if share < thr:
        continue
    name = "straight-line" if o is None else f"{o[0]}..{o[1]} @{ins[o[0]]['addr'] - base:#x}"
    print(f"{name:>22s} {a['cnt']:6d} {100.0 * a['n'] / tot_n:7.2f} {100.0 * a['fp'] / max(1, sum(x['fp'] for x in agg.values())):7.2f} "
          f"{share:11.2f} {100.0 * a['samp'] / max(1, tot_s):9.2f}  {a['fp'] / max(1, a['n']):.3f}")
