"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line.
usage: python tools/ncu_lines.py dump.csv [topN]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr = None, None
lines = {}
tot_samples = 0
tot_inst = 0
with open(path, newline="") as f:
    for row in csv.reader(f):
        if not row:
            continue
        if row[0] == "File Path":
            cur_file = row[1].split("/")[-1]
            continue
        if row[0] == "Function Name":
            continue
        if row[0] == "Line No":
            hdr = row
            idx = {n: i for i, n in enumerate(hdr)}
            continue
        if row[0] == "" or hdr is None:
            continue
        try:
            ln = int(row[0])
        except ValueError:
            continue

        def g(name):
            try:
                return float(row[idx[name]])
            except (ValueError, KeyError, IndexError):
                return 0.0
        s = g("# Samples")
        inst = g("Instructions Executed")
        key = (cur_file, ln)
        d = lines.setdefault(key, dict(src=row[1].strip()[:110], s=0, inst=0, bar=0, lsb=0, ssb=0, wait=0, br=0))
        d["s"] += s
        d["inst"] += inst
        d["bar"] += g("stall_barrier")
        d["lsb"] += g("stall_long_sb")
        d["ssb"] += g("stall_short_sb")
        d["wait"] += g("stall_wait")
        d["br"] += g("stall_branch_resolving")
        tot_samples += s
        tot_inst += inst
print(f"total samples {tot_samples:.0f}  total warp-instructions {tot_inst:.0f}")
print("by file:")
byf = defaultdict(lambda: [0, 0])
for (fn, ln), d in lines.items():
    byf[fn][0] += d["s"]
    byf[fn][1] += d["inst"]
for fn, (s, i) in sorted(byf.items(), key=lambda kv: -kv[1][0]):
    print(f"  {fn:28s} samples {100*s/tot_samples:5.1f}%  inst {100*i/max(tot_inst,1):5.1f}%")
print(f"top {top} lines by samples:  (%samples | %inst | barrier longsb shortsb wait branch)")
for (fn, ln), d in sorted(lines.items(), key=lambda kv: -kv[1]["s"])[:top]:
    print(f"{100*d['s']/tot_samples:5.1f}% {100*d['inst']/max(tot_inst,1):5.1f}% | {d['bar']:6.0f} {d['lsb']:6.0f} {d['ssb']:5.0f} "
          f"{d['wait']:5.0f} {d['br']:5.0f} | {fn}:{ln}  {d['src']}")
