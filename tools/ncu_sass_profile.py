#!/usr/bin/env python3
"""SASS-level `ncu --page source --csv` export (tools/ncu_export.sh) -> where a kernel's warp instructions go:
share and lanes-per-instruction by opcode class and by position in the instruction stream.
    python tools/ncu_sass_profile.py gpurun_out/<name>_source_k<N>.csv.gz [block]"""
import collections, csv, gzip, io, re, sys
rows = list(csv.reader(io.TextIOWrapper(gzip.open(sys.argv[1]))))
B = int(sys.argv[2]) if len(sys.argv) > 2 else 40
print(rows[0][1][:100])
h = rows[1]
ix = {n: i for i, n in enumerate(h)}
data = [r for r in rows[2:] if len(r) > ix['Thread Instructions Executed'] and r[ix['Instructions Executed']].isdigit()]
I = lambda r, k: int(r[ix[k]])
tot = sum(I(r, 'Instructions Executed') for r in data)
thr = sum(I(r, 'Thread Instructions Executed') for r in data)
smp = sum(I(r, '# Samples') for r in data)
print(f"warp instructions {tot}, lanes/inst {thr / tot:.2f}, SASS lines {len(data)}, stall samples {smp}")
ops, opl = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[ix['Source']])
    op = m.group(2).split('.')[0] if m else '?'
    ops[op] += I(r, 'Instructions Executed')
    opl[op] += I(r, 'Thread Instructions Executed')
print("opcode     share  lanes")
for op, n in ops.most_common(20):
    print(f"{op:10s} {100 * n / tot:5.1f}% {opl[op] / max(n, 1):5.1f}")
print(f"--- by position (blocks of {B} SASS instructions with >= 1 % of the executed instructions): share, lanes, stall-sample share")
for b in range(0, len(data), B):
    blk = data[b:b + B]
    n = sum(I(r, 'Instructions Executed') for r in blk)
    t = sum(I(r, 'Thread Instructions Executed') for r in blk)
    s = sum(I(r, '# Samples') for r in blk)
    if n / tot > 0.01:
        print(f"{b:5d} {100 * n / tot:5.1f}% lanes {t / max(n, 1):5.1f} samples {100 * s / max(smp, 1):5.1f}%  {blk[0][ix['Source']].strip()[:60]}")
