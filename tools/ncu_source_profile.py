#!/usr/bin/env python3
"""`ncu --page source --csv --print-source cuda,sass` export -> executed warp instructions and stall samples per CUDA source line
(and per function the line belongs to). python tools/ncu_source_profile.py file.csv.gz [top]"""
import collections, csv, gzip, io, sys
rows = list(csv.reader(io.TextIOWrapper(gzip.open(sys.argv[1]))))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname, lines, cur_file = None, [], ""
tot_i = tot_s = 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Function Name":
        fname = r[1]
        continue
    if len(r) >= 8 and r[0].isdigit():  # a CUDA source line with its aggregated metrics
        try:
            inst, smp = int(r[7]), int(r[6])
        except ValueError:
            continue
        lines.append((cur_file, int(r[0]), r[1].strip(), inst, smp))
        tot_i += inst
        tot_s += smp
print((fname or "")[:110])
print(f"warp instructions {tot_i}, stall samples {tot_s}")
print("--- top lines by executed warp instructions")
for f, ln, src, inst, smp in sorted(lines, key=lambda x: -x[3])[:top]:
    print(f"{100 * inst / tot_i:5.1f}% inst {100 * smp / max(tot_s, 1):5.1f}% stall  {f}:{ln:<5d} {src[:110]}")
