#!/bin/bash
# Round 2, GPU call AC: one-level walk for scenes without transformed mesh instances (TRAV_BVH_FLAT, default there) against the
# two-level walk (RPT_FLAT=0), and the lazily rebuilt watertight constants against the eager variant build, same session.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r02ac_tests_all.log 2>&1
tail -3 gpurun_out/r02ac_tests_all.log
timeout 900 python - > gpurun_out/r02ac_flat_walk.txt 2> gpurun_out/r02ac.err <<'PY'
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import parity
p = parity.pkg()
for rep in range(2):
    for name, kw in (("cornell", {}), ("furnace", {}), ("hdri2", {"spp": 16}), ("hdri", {"spp": 16}), ("sun_test", {}), ("instanced_monkeys", {}), ("gem", {"spp": 64})):
        world, st, flat = parity.load_scene(name, **kw)
        for so, flatmode in (("librpt_var_eager.so", "0"), ("librpt_b200.so", "0"), ("librpt_b200.so", "1")):
            os.environ["RPT_FLAT"] = flatmode
            lib = p.ffi.load_library(os.path.join(p.ffi.PKG_DIR, so))
            sc = parity._bake_unbaked_importance_map(p.ffi.Scene(lib, flat, 0), flat)
            best = None
            for i in range(5):
                ptr, c = sc.render_pt_device(st.params(seed=i, spp_total=0, flags=1))
                kt = {k["name"]: k["ms"] for k in sc.kernel_times()}
                if i and (best is None or c.device_ms < best[0]):
                    best = (c.device_ms, kt, c)
            ms, kt, c = best
            print(f"{name:18s} {so:22s} FLAT={flatmode} {ms:9.3f} ms {c.segments / ms / 1e6:6.3f} Gseg/s  trace {kt.get('k_trace', 0):8.3f}  shadow {kt.get('k_shadow', 0):8.3f}", flush=True)
            sc.close()
PY
cat gpurun_out/r02ac_flat_walk.txt; tail -3 gpurun_out/r02ac.err
