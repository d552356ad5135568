#!/usr/bin/env python3
"""ncu launch list (gpu__time_duration.sum, dram__bytes_read.sum, dram__bytes_write.sum; --csv --log-file) -> per-kernel
averages per launch, written as JSON for bench.py's roofline.traffic. Usage: ncu_traffic.py launches.csv out.json [passes]"""
import csv, json, re, sys
from collections import defaultdict

src, dst = sys.argv[1], sys.argv[2]
passes = int(sys.argv[3]) if len(sys.argv) > 3 else 2
rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
hdr, rows = rows[0], rows[1:]
ix = {n: i for i, n in enumerate(hdr)}
per = defaultdict(dict)
names = {}
for r in rows:
    per[int(r[ix["ID"]])][r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    names[int(r[ix["ID"]])] = r[ix["Kernel Name"]]
units = {r[ix["Metric Name"]]: r[ix["Metric Unit"]] for r in rows}


def short(n):
    m = re.search(r"(k_[a-z_]+)(<[^>]*>)?", n)
    base, t = m.group(1), m.group(2) or ""
    if base in ("k_shade_surface", "k_shade_vertex", "k_nee"):
        first = re.search(r"(\d+)", t.split(",")[0])  # the material class is the first template argument (2 = diffuse, 3 = GGX)
        return base + ("<diffuse>" if first and first.group(1) == "2" else "<ggx>")
    return base


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]


def to_ms(v, unit):
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[unit]


agg = defaultdict(lambda: {"launches": 0, "ms": 0.0, "dram_read": 0.0, "dram_write": 0.0})
for i, m in per.items():
    a = agg[short(names[i])]
    a["launches"] += 1
    a["ms"] += to_ms(m["gpu__time_duration.sum"], units["gpu__time_duration.sum"])
    a["dram_read"] += to_bytes(m["dram__bytes_read.sum"], units["dram__bytes_read.sum"])
    a["dram_write"] += to_bytes(m["dram__bytes_write.sum"], units["dram__bytes_write.sum"])
total_ms = sum(a["ms"] for a in agg.values())
out = {"source": src, "passes": passes, "note": "ncu per-launch values are cold-cache and serialised: compare shares, not absolutes", "kernels": {}}
for k, a in sorted(agg.items()):
    out["kernels"][k] = {"launches_per_pass": a["launches"] / passes, "ms_per_pass": a["ms"] / passes, "share": a["ms"] / total_ms,
                         "dram_bytes_per_launch": (a["dram_read"] + a["dram_write"]) / a["launches"],
                         "dram_read_bytes_per_launch": a["dram_read"] / a["launches"], "dram_write_bytes_per_launch": a["dram_write"] / a["launches"]}
json.dump(out, open(dst, "w"), indent=1)
for k, v in out["kernels"].items():
    print(f"{k:28s} {v['launches_per_pass']:5.1f} launches/pass {v['ms_per_pass']:8.3f} ms/pass share {v['share']:.3f} dram/launch {v['dram_bytes_per_launch']/1e6:9.1f} MB")
