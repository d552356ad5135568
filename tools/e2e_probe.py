#!/usr/bin/env python3
"""Where an end-to-end step (scene upload + render + film download, bench.py's e2e leg) spends its host time, step by step:
python tools/e2e_probe.py [steps]. Prints every step slower than 1.3 x the median and the per-phase medians."""
import gc, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import parity, torch
p = parity.pkg()
world, st, flat0 = parity.load_scene("cornell")
r = p.CudaRenderer(device=0)
pinned = torch.empty((st.height, st.width, 4), dtype=torch.float32).pin_memory()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
mode = sys.argv[2] if len(sys.argv) > 2 else "plain"  # "head": a second, long-lived scene exists (bench.py's headline job); "head+instr": and it has run instrumented steps
if mode != "plain":
    head = r.make_scene(world, st.wavelength_bounds)
    for i in range(3):
        head.render_pt_device(st.params(seed=i, spp_total=0, flags=3 if mode == "head+instr" else 0))
    torch.cuda.synchronize()
print("mode", mode)
rows = []
gc.collect(); gc.disable()
for it in range(-2, n):
    t0 = time.perf_counter()
    sc = r.make_scene(world, st.wavelength_bounds)
    t1 = time.perf_counter()
    cnt = sc.render_pt_into(st.params(seed=it + 5), pinned.data_ptr())
    t2 = time.perf_counter()
    sc.close()
    t3 = time.perf_counter()
    if it >= 0:
        rows.append((1e3 * (t3 - t0), 1e3 * (t1 - t0), 1e3 * (t2 - t1), cnt.device_ms, 1e3 * (t3 - t2)))
a = np.array(rows)
med = np.median(a, axis=0)
print("medians: step %.2f ms = make_scene %.2f + render+D2H %.2f (device %.2f) + destroy %.2f" % tuple(med))
for i, row in enumerate(rows):
    if row[0] > 1.3 * med[0]:
        print("slow step %3d: step %.2f ms = make_scene %.2f + render+D2H %.2f (device %.2f) + destroy %.2f" % ((i,) + row))
print("steps over 1.3 x median:", int((a[:, 0] > 1.3 * med[0]).sum()), "of", n)
