#!/usr/bin/env python3
"""Minimal driver for ncu: N full-frame passes of the bench workload (scenes/cornell.npz, 1080p @ 16 spp)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity
name = sys.argv[1] if len(sys.argv) > 1 else "cornell"
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 2
world, st, flat = parity.load_scene(name)
sc = parity.cuda_scene(flat)
for i in range(passes):
    ptr, cnt = sc.render_pt_device(st.params(seed=i, spp_total=0, flags=1))
    print(f"pass {i}: {cnt.device_ms:.2f} ms, {cnt.segments / cnt.device_ms / 1e6:.3f} Gseg/s, launches {cnt.kernel_launches}")
    for k in sc.kernel_times():
        print(f"    {k['name']:28s} {k['launches']:3d} launches {k['ms']:8.3f} ms")
sc.close()
