#!/usr/bin/env python3
"""Times library variants (librpt_var_*.so, `make -C rust-pathtracer_b200/csrc variant NAME=.. EXTRA=..`) on a scene's configured
workload, best of 3 instrumented passes: python tools/variant_bench.py <scene> <so>..."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity
p = parity.pkg()
name = sys.argv[1]
world, st, flat = parity.load_scene(name)
for so in sys.argv[2:]:
    lib = p.ffi.load_library(os.path.join(p.ffi.PKG_DIR, so))
    sc = parity._bake_unbaked_importance_map(p.ffi.Scene(lib, flat, 0), flat)
    best = None
    for i in range(4):
        ptr, cnt = sc.render_pt_device(st.params(seed=i, spp_total=0, flags=1))
        kt = {k["name"]: k["ms"] for k in sc.kernel_times()}
        if i > 0 and (best is None or cnt.device_ms < best[0]):
            best = (cnt.device_ms, kt, cnt.segments)
    ks = "  ".join(f"{k.replace('k_', '')} {v:7.2f}" for k, v in sorted(best[1].items(), key=lambda kv: -kv[1])[:6])
    print(f"{name:18s} {so:28s} {best[0]:8.2f} ms  {best[2] / best[0] / 1e6:6.3f} Gseg/s   {ks}")
    sc.close()
