#!/bin/bash
# Round 2, GPU call AD: register caps of the one-level walk: 8 (shipped: 64 registers, no spills), 9 (56) and 10 (48) CTAs per SM.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python - > gpurun_out/r02ad_flat_caps.txt 2> gpurun_out/r02ad.err <<'PY'
import os, sys
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import parity
p = parity.pkg()
for rep in range(2):
    for name, kw in (("cornell", {}), ("furnace", {}), ("hdri2", {"spp": 16}), ("sun_test", {})):
        world, st, flat = parity.load_scene(name, **kw)
        for so in ("librpt_b200.so", "librpt_var_flat9.so", "librpt_var_flat10.so"):
            lib = p.ffi.load_library(os.path.join(p.ffi.PKG_DIR, so))
            sc = parity._bake_unbaked_importance_map(p.ffi.Scene(lib, flat, 0), flat)
            best = None
            for i in range(5):
                ptr, c = sc.render_pt_device(st.params(seed=i, spp_total=0, flags=1))
                kt = {k["name"]: k["ms"] for k in sc.kernel_times()}
                if i and (best is None or c.device_ms < best[0]):
                    best = (c.device_ms, kt, c)
            ms, kt, c = best
            print(f"{name:18s} {so:22s} {ms:9.3f} ms {c.segments / ms / 1e6:6.3f} Gseg/s  trace {kt.get('k_trace', 0):8.3f}  shadow {kt.get('k_shadow', 0):8.3f}", flush=True)
            sc.close()
PY
cat gpurun_out/r02ad_flat_caps.txt; tail -3 gpurun_out/r02ad.err
