#!/usr/bin/env python3
"""ptxas -v log (rust-pathtracer_b200/csrc/build.log) -> one line per kernel: registers, spill bytes, stack, static smem.
    python tools/ptxas_table.py [build.log]"""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
path = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rust-pathtracer_b200", "csrc", "build.log")
rows, cur = [], None
for line in open(path):
    m = re.search(r"Compiling entry function '([^']+)'", line)
    if m:
        cur = {"name": m.group(1), "regs": None, "spill_st": 0, "spill_ld": 0, "stack": 0, "smem": 0}
        rows.append(cur)
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m and "seen" not in cur:  # the first one belongs to the entry function (later ones: its non-inlined callees)
        cur["seen"] = True
        cur["stack"], cur["spill_st"], cur["spill_ld"] = map(int, m.groups())
    m = re.search(r"Used (\d+) registers", line)
    if m:
        cur["regs"] = int(m.group(1))
        m2 = re.search(r"(\d+) bytes smem", line)
        cur["smem"] = int(m2.group(1)) if m2 else 0
names = subprocess.run(["c++filt"], input="\n".join(r["name"] for r in rows), capture_output=True, text=True).stdout.splitlines()
print("| kernel | regs | spill st/ld (B) | stack (B) | static smem (B) |\n|---|---|---|---|---|")
for r, n in zip(rows, names):
    n = re.sub(r"\(anonymous namespace\)::", "", n)
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\(.*$", "", n)
    print(f"| `{n}` | {r['regs']} | {r['spill_st']}/{r['spill_ld']} | {r['stack']} | {r['smem']} |")
